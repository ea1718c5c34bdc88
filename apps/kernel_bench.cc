/**
 * \file kernel_bench.cc
 * \brief Device-timed microbenchmarks + self-checks of the sm_100a data-plane kernels.
 *
 * For each kernel: warm up, then time `iters` launches with CUDA events on the
 * launching stream, rotating through buffers whose total footprint exceeds the
 * 126 MB L2 (so every iteration streams from HBM), and report the algorithmic
 * bandwidth next to the measured copy roofline (MEASURED_PEAKS.json hbm_gbs,
 * passed as argv[1], default 6571). With two visible GPUs and `--peer` the copy
 * destinations live on device 1 (NVLink). Prints one JSON object per line.
 *   usage: kernel_bench [hbm_gbs] [--peer] [--quick]
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels/ps_kernels.h"

#define CK(expr)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__,   \
              cudaGetErrorString(e_));                                             \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

namespace {

double g_hbm = 6571.0;

struct Timer {
  cudaEvent_t a, b;
  Timer() {
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
  }
  void Start(cudaStream_t s) { CK(cudaEventRecord(a, s)); }
  float StopMs(cudaStream_t s) {
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
};

void Report(const char* name, size_t bytes_per_iter, double algo_bytes, float ms_per_iter,
            const char* extra = "") {
  const double gbs = algo_bytes / (ms_per_iter * 1e-3) / 1e9;
  printf("{\"kernel\":\"%s\",\"bytes\":%zu,\"us\":%.2f,\"algo_GBps\":%.1f,\"frac_of_measured_hbm\":%.3f%s}\n",
         name, bytes_per_iter, ms_per_iter * 1e3, gbs, gbs / g_hbm, extra);
  fflush(stdout);
}

float HalfToFloatBf16(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int SelfCheck(cudaStream_t st) {
  int bad = 0;
  const size_t n = 100003;  // ragged on purpose
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = std::sin(0.001f * i) * (1.f + (i % 97)) * 0.01f;
  float *d_src, *d_out;
  void* d_wire;
  CK(cudaMalloc(&d_src, n * 4));
  CK(cudaMalloc(&d_out, n * 4));
  CK(cudaMalloc(&d_wire, n * 4 + 4096));
  CK(cudaMemcpy(d_src, h.data(), n * 4, cudaMemcpyHostToDevice));
  std::vector<float> back(n);
  // f32 -> bf16 (scale 0.5) -> f32
  CK((cudaError_t)ps_launch_copy(d_wire, d_src, n * 4, PS_CODEC_F32_TO_BF16, 0.5f, 0, (ps_stream_t)st));
  CK((cudaError_t)ps_launch_decode(d_out, d_wire, n, PS_GRAD_BF16, (ps_stream_t)st));
  CK(cudaMemcpyAsync(back.data(), d_out, n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double worst = 0;
  for (size_t i = 0; i < n; ++i) {
    const double ref = 0.5 * h[i];
    worst = std::fmax(worst, std::fabs(back[i] - ref) / (std::fabs(ref) + 1e-6));
  }
  if (worst > 1.0 / 128) { ++bad; }
  printf("{\"check\":\"f32_to_bf16\",\"max_rel_err\":%.5f,\"ok\":%d}\n", worst, worst <= 1.0 / 128);
  // f32 -> fp8 block -> f32
  CK((cudaError_t)ps_launch_copy(d_wire, d_src, n * 4, PS_CODEC_F32_TO_FP8BLOCK, 1.f, 0, (ps_stream_t)st));
  CK((cudaError_t)ps_launch_decode(d_out, d_wire, n, PS_GRAD_FP8BLOCK, (ps_stream_t)st));
  CK(cudaMemcpyAsync(back.data(), d_out, n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  worst = 0;
  for (size_t b = 0; b < n; b += 32) {
    float amax = 0;
    for (size_t i = b; i < b + 32 && i < n; ++i) amax = std::fmax(amax, std::fabs(h[i]));
    for (size_t i = b; i < b + 32 && i < n; ++i) {
      // e4m3 has 3 mantissa bits: error <= amax/16 per block after power-of-two scaling
      worst = std::fmax(worst, std::fabs(back[i] - h[i]) / (amax + 1e-12));
    }
  }
  if (worst > 0.0725) ++bad;
  printf("{\"check\":\"f32_to_fp8block\",\"max_err_over_block_amax\":%.5f,\"ok\":%d}\n", worst,
         worst <= 0.0725);
  // raw copy, both flavours, odd size
  std::vector<unsigned char> hb(n * 4 + 3), hb2(n * 4 + 3);
  for (size_t i = 0; i < hb.size(); ++i) hb[i] = static_cast<unsigned char>(i * 131 + 7);
  unsigned char *d_a, *d_b;
  CK(cudaMalloc(&d_a, hb.size()));
  CK(cudaMalloc(&d_b, hb.size()));
  CK(cudaMemcpy(d_a, hb.data(), hb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_b, 0, hb.size()));
  CK((cudaError_t)ps_launch_copy(d_b, d_a, hb.size(), PS_CODEC_RAW, 1.f, 0, (ps_stream_t)st));
  CK(cudaMemcpyAsync(hb2.data(), d_b, hb.size(), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int raw_ok = memcmp(hb.data(), hb2.data(), hb.size()) == 0;
  if (!raw_ok) ++bad;
  printf("{\"check\":\"raw_copy\",\"tma\":%d,\"ok\":%d}\n", getenv("PS_COPY_TMA") ? atoi(getenv("PS_COPY_TMA")) : 0, raw_ok);
  // fused update vs host AdamW (W=2, bf16 grads)
  {
    const size_t m = 4099;
    std::vector<float> p(m), g0(m), g1(m);
    for (size_t i = 0; i < m; ++i) {
      p[i] = std::cos(0.01f * i);
      g0[i] = std::sin(0.02f * i) * 0.1f;
      g1[i] = std::cos(0.03f * i) * 0.1f;
    }
    float *dp, *dm, *dv, *dg0, *dg1;
    void *w0, *w1, *outb;
    CK(cudaMalloc(&dp, m * 4)); CK(cudaMalloc(&dm, m * 4)); CK(cudaMalloc(&dv, m * 4));
    CK(cudaMalloc(&dg0, m * 4)); CK(cudaMalloc(&dg1, m * 4));
    CK(cudaMalloc(&w0, m * 2 + 64)); CK(cudaMalloc(&w1, m * 2 + 64)); CK(cudaMalloc(&outb, m * 2 + 64));
    CK(cudaMemcpy(dp, p.data(), m * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dm, 0, m * 4)); CK(cudaMemset(dv, 0, m * 4));
    CK(cudaMemcpy(dg0, g0.data(), m * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg1, g1.data(), m * 4, cudaMemcpyHostToDevice));
    CK((cudaError_t)ps_launch_copy(w0, dg0, m * 4, PS_CODEC_F32_TO_BF16, 1.f, 0, (ps_stream_t)st));
    CK((cudaError_t)ps_launch_copy(w1, dg1, m * 4, PS_CODEC_F32_TO_BF16, 1.f, 0, (ps_stream_t)st));
    std::vector<uint16_t> hw0(m), hw1(m);
    CK(cudaMemcpyAsync(hw0.data(), w0, m * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hw1.data(), w1, m * 2, cudaMemcpyDeviceToHost, st));
    ps_update_args a;
    memset(&a, 0, sizeof(a));
    a.n = m; a.num_grads = 2; a.grad_format = PS_GRAD_BF16;
    a.grads[0] = w0; a.grads[1] = w1;
    a.master = dp; a.m = dm; a.v = dv; a.num_outs = 1; a.outs[0] = outb; a.body_outs = 1;
    ps_opt_params o;
    o.optimizer = PS_OPT_ADAMW; o.lr = 1e-2f; o.beta1 = 0.9f; o.beta2 = 0.95f; o.eps = 1e-8f;
    o.weight_decay = 0.1f; o.bias_corr1 = 1.f - 0.9f; o.bias_corr2 = 1.f - 0.95f; o.grad_scale = 0.5f;
    CK((cudaError_t)ps_launch_update(&a, &o, 0, (ps_stream_t)st));
    std::vector<float> pn(m);
    CK(cudaMemcpyAsync(pn.data(), dp, m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double werr = 0;
    for (size_t i = 0; i < m; ++i) {
      const float g = 0.5f * (HalfToFloatBf16(hw0[i]) + HalfToFloatBf16(hw1[i]));
      const float mm = 0.1f * g, vv = 0.05f * g * g;
      const float ref = p[i] - 1e-2f * ((mm / 0.1f) / (std::sqrt(vv / 0.05f) + 1e-8f) + 0.1f * p[i]);
      werr = std::fmax(werr, std::fabs(ref - pn[i]));
    }
    if (werr > 1e-5) ++bad;
    printf("{\"check\":\"fused_adamw_update\",\"max_abs_err\":%.3g,\"ok\":%d}\n", werr, werr <= 1e-5);
  }
  fflush(stdout);
  return bad;
}

}  // namespace

int main(int argc, char** argv) {
  bool peer = false, quick = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--peer")) peer = true;
    else if (!strcmp(argv[i], "--quick")) quick = true;
    else g_hbm = atof(argv[i]);
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  CK(cudaSetDevice(0));
  if (peer && ndev < 2) {
    fprintf(stderr, "--peer needs 2 GPUs\n");
    peer = false;
  }
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const int bad = SelfCheck(st);

  // ---- raw copy sweep ----
  const size_t kRot = 4;  // 4 x (src+dst) x 256 MB >> L2
  const size_t big = quick ? (64u << 20) : (256u << 20);
  std::vector<char*> src(kRot), dst(kRot);
  for (size_t i = 0; i < kRot; ++i) {
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&src[i], big));
    CK(cudaMemset(src[i], 1, big));
    if (peer) {
      CK(cudaSetDevice(1));
      CK(cudaMalloc(&dst[i], big));
      CK(cudaSetDevice(0));
    } else {
      CK(cudaMalloc(&dst[i], big));
    }
  }
  if (peer) {
    cudaError_t e = cudaDeviceEnablePeerAccess(1, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
    cudaGetLastError();
  }
  Timer t;
  const char* flavour = (getenv("PS_COPY_TMA") && atoi(getenv("PS_COPY_TMA"))) ? "tma" : "ldg";
  std::vector<size_t> sizes = {4096, 65536, 1u << 20, 4u << 20, 16u << 20, 64u << 20, big};
  std::vector<int> ctas = {0, 16, 32, 74, 148, 296};
  for (size_t sz : sizes) {
    for (int c : ctas) {
      if (sz < (16u << 20) && c != 0) continue;
      const int iters = sz >= (16u << 20) ? 20 : 200;
      for (int w = 0; w < 3; ++w)
        CK((cudaError_t)ps_launch_copy(dst[w % kRot], src[w % kRot], sz, PS_CODEC_RAW, 1.f, c, (ps_stream_t)st));
      t.Start(st);
      for (int i = 0; i < iters; ++i)
        CK((cudaError_t)ps_launch_copy(dst[i % kRot], src[i % kRot], sz, PS_CODEC_RAW, 1.f, c, (ps_stream_t)st));
      const float ms = t.StopMs(st) / iters;
      char extra[128];
      snprintf(extra, sizeof(extra), ",\"flavour\":\"%s\",\"ctas\":%d,\"peer\":%d", flavour, c, peer ? 1 : 0);
      // a copy moves 2 bytes of traffic per payload byte locally, 1 over NVLink per direction
      Report(peer ? "copy_raw_peer" : "copy_raw", sz, (peer ? 1.0 : 2.0) * sz, ms, extra);
    }
  }
  // ---- codec kernels (local) ----
  if (!peer) {
    const size_t n = big / 4;  // fp32 elements
    for (int codec : {PS_CODEC_F32_TO_BF16, PS_CODEC_F32_TO_FP8BLOCK, PS_CODEC_BF16_TO_FP8BLOCK}) {
      const size_t src_bytes = codec == PS_CODEC_BF16_TO_FP8BLOCK ? n * 2 : n * 4;
      const size_t out_bytes = codec == PS_CODEC_F32_TO_BF16 ? n * 2 : n + n / 32;
      for (int w = 0; w < 3; ++w)
        CK((cudaError_t)ps_launch_copy(dst[w % kRot], src[w % kRot], src_bytes, codec, 0.25f, 0, (ps_stream_t)st));
      t.Start(st);
      const int iters = 20;
      for (int i = 0; i < iters; ++i)
        CK((cudaError_t)ps_launch_copy(dst[i % kRot], src[i % kRot], src_bytes, codec, 0.25f, 0, (ps_stream_t)st));
      const float ms = t.StopMs(st) / iters;
      const char* nm = codec == PS_CODEC_F32_TO_BF16 ? "push_f32_to_bf16"
                       : codec == PS_CODEC_F32_TO_FP8BLOCK ? "push_f32_to_fp8block" : "push_bf16_to_fp8block";
      Report(nm, src_bytes, static_cast<double>(src_bytes + out_bytes), ms);
    }
    // ---- fused update: W in {1,2,4}, bf16 and fp8 grads, 1 and 1+W outputs ----
    const size_t ne = quick ? (16u << 20) : (64u << 20);  // elements per shard
    float *master, *m, *v;
    CK(cudaMalloc(&master, ne * 4)); CK(cudaMalloc(&m, ne * 4)); CK(cudaMalloc(&v, ne * 4));
    CK(cudaMemset(master, 0, ne * 4)); CK(cudaMemset(m, 0, ne * 4)); CK(cudaMemset(v, 0, ne * 4));
    std::vector<void*> slots(4), outs(5);
    for (auto& s : slots) { CK(cudaMalloc(&s, ne * 2 + 4096)); CK(cudaMemset(s, 0, ne * 2 + 4096)); }
    for (auto& o : outs) CK(cudaMalloc(&o, ne * 2 + 64));
    for (int fmt : {PS_GRAD_BF16, PS_GRAD_FP8BLOCK}) {
      for (int W : {1, 2, 4}) {
        for (int fan : {1, 1 + W}) {
          ps_update_args a;
          memset(&a, 0, sizeof(a));
          a.n = ne; a.num_grads = W; a.grad_format = fmt;
          for (int w = 0; w < W; ++w) a.grads[w] = slots[w];
          a.master = master; a.m = m; a.v = v;
          a.num_outs = fan;
          a.body_outs = fan;
          for (int k = 0; k < fan; ++k) a.outs[k] = outs[k];
          ps_opt_params o;
          o.optimizer = PS_OPT_ADAMW; o.lr = 1e-3f; o.beta1 = 0.9f; o.beta2 = 0.95f; o.eps = 1e-8f;
          o.weight_decay = 0.f; o.bias_corr1 = 0.1f; o.bias_corr2 = 0.05f; o.grad_scale = 1.f / W;
          for (int w = 0; w < 3; ++w) CK((cudaError_t)ps_launch_update(&a, &o, 0, (ps_stream_t)st));
          t.Start(st);
          const int iters = 10;
          for (int i = 0; i < iters; ++i) CK((cudaError_t)ps_launch_update(&a, &o, 0, (ps_stream_t)st));
          const float ms = t.StopMs(st) / iters;
          const double gbytes = fmt == PS_GRAD_BF16 ? 2.0 : (1.0 + 1.0 / 32);
          const double bytes = ne * (W * gbytes + 12.0 + 12.0 + 2.0 * fan);
          char extra[128];
          snprintf(extra, sizeof(extra), ",\"W\":%d,\"fanout\":%d,\"grad\":\"%s\",\"elems\":%zu", W, fan,
                   fmt == PS_GRAD_BF16 ? "bf16" : "fp8block", ne);
          Report("update_adamw_fused", static_cast<size_t>(bytes), bytes, ms, extra);
        }
      }
    }
  }
  printf("{\"kernel_launches\":%llu,\"selfcheck_failures\":%d}\n", ps_kernel_launch_count(), bad);
  return bad ? 1 : 0;
}
