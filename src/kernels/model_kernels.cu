/**
 * \file model_kernels.cu
 * \brief Fused elementwise kernels of the worker-side Llama block (sm_100a).
 *
 * Not part of the reference (it has no model code); they exist so that the flagship
 * PS-training step does not spend HBM bandwidth on PyTorch's unfused elementwise chains:
 *   rope_split   qkv[T,(H+2KV)*D] -> q[T,H,D], k[T,KV,D] rotated (half-split RoPE), one pass;
 *                the backward merges dq, dk, dv back into dqkv with the inverse rotation.
 *   swiglu       gu[T,2F] -> silu(g)*u, and its backward into dgu[T,2F].
 * All are 16-byte vectorised, bf16 in/out, fp32 math, grid-stride over 8-element packets.
 */
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels/model_kernels.h"

namespace ps_kernels_internal {
void CountLaunch(int n);
}

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int4 ld16(const void* p) { return *reinterpret_cast<const int4*>(p); }
__device__ __forceinline__ void st16(void* p, const int4& v) { *reinterpret_cast<int4*>(p) = v; }
__device__ __forceinline__ void unpack8(const int4& q, float* f) {
  const uint32_t w[4] = {static_cast<uint32_t>(q.x), static_cast<uint32_t>(q.y),
                         static_cast<uint32_t>(q.z), static_cast<uint32_t>(q.w)};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ int4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_int4(static_cast<int>(w[0]), static_cast<int>(w[1]), static_cast<int>(w[2]),
                   static_cast<int>(w[3]));
}

/*!
 * One packet = 8 elements of the first half of a head and the 8 matching elements of the
 * second half. FORWARD: src = qkv rows (stride in_stride), dst = contiguous [T, heads, D].
 * BACKWARD (INVERSE): src = contiguous grads, dst = dqkv rows; rotation by -angle.
 */
template <bool INVERSE>
__global__ void __launch_bounds__(kThreads)
k_rope(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
       const float* __restrict__ cos_t, const float* __restrict__ sin_t, size_t tokens, int seq,
       int heads, int hd, size_t strided_row, size_t strided_off) {
  const int half = hd / 2, packs = half / 8;
  const size_t total = tokens * heads * packs;
  const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
    const int pk = static_cast<int>(i % packs);
    const int h = static_cast<int>((i / packs) % heads);
    const size_t t = i / (static_cast<size_t>(packs) * heads);
    const int s = static_cast<int>(t % seq);
    const size_t strided = t * strided_row + strided_off + static_cast<size_t>(h) * hd + pk * 8;
    const size_t dense = (t * heads + h) * hd + pk * 8;
    const size_t in = INVERSE ? dense : strided, out = INVERSE ? strided : dense;
    float a[8], b[8], c[8], sn[8], ra[8], rb[8];
    unpack8(ld16(src + in), a);
    unpack8(ld16(src + in + half), b);
    const float4 c0 = *reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(s) * half + pk * 8);
    const float4 c1 = *reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(s) * half + pk * 8 + 4);
    const float4 s0 = *reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(s) * half + pk * 8);
    const float4 s1 = *reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(s) * half + pk * 8 + 4);
    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
    sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sj = INVERSE ? -sn[j] : sn[j];
      ra[j] = a[j] * c[j] - b[j] * sj;
      rb[j] = b[j] * c[j] + a[j] * sj;
    }
    st16(dst + out, pack8(ra));
    st16(dst + out + half, pack8(rb));
  }
}

/*! copy the V part of dqkv from the dense dv gradient (backward of the split) */
__global__ void __launch_bounds__(kThreads)
k_scatter_rows(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t tokens,
               int width, size_t dst_row, size_t dst_off) {
  const int packs = width / 8;
  const size_t total = tokens * packs;
  const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
    const size_t t = i / packs;
    const int pk = static_cast<int>(i % packs);
    st16(dst + t * dst_row + dst_off + pk * 8, ld16(src + t * width + pk * 8));
  }
}

/*! dense[T,width] <- rows of a strided matrix (the V part of qkv) */
__global__ void __launch_bounds__(kThreads)
k_gather_rows(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t tokens,
              int width, size_t src_row, size_t src_off) {
  const int packs = width / 8;
  const size_t total = tokens * packs;
  const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
    const size_t t = i / packs;
    const int pk = static_cast<int>(i % packs);
    st16(dst + t * width + pk * 8, ld16(src + t * src_row + src_off + pk * 8));
  }
}

__global__ void __launch_bounds__(kThreads)
k_swiglu_fwd(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ out, size_t tokens, int f) {
  const int packs = f / 8;
  const size_t total = tokens * packs;
  const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
    const size_t t = i / packs;
    const int pk = static_cast<int>(i % packs);
    float g[8], u[8], o[8];
    unpack8(ld16(gu + t * 2 * f + pk * 8), g);
    unpack8(ld16(gu + t * 2 * f + f + pk * 8), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = __fdividef(1.f, 1.f + __expf(-g[j]));
      o[j] = g[j] * sg * u[j];
    }
    st16(out + t * f + pk * 8, pack8(o));
  }
}

__global__ void __launch_bounds__(kThreads)
k_swiglu_bwd(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dout,
             __nv_bfloat16* __restrict__ dgu, size_t tokens, int f) {
  const int packs = f / 8;
  const size_t total = tokens * packs;
  const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
    const size_t t = i / packs;
    const int pk = static_cast<int>(i % packs);
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(ld16(gu + t * 2 * f + pk * 8), g);
    unpack8(ld16(gu + t * 2 * f + f + pk * 8), u);
    unpack8(ld16(dout + t * f + pk * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = __fdividef(1.f, 1.f + __expf(-g[j]));
      const float silu = g[j] * sg;
      du[j] = d[j] * silu;
      dg[j] = d[j] * u[j] * sg * (1.f + g[j] * (1.f - sg));
    }
    st16(dgu + t * 2 * f + pk * 8, pack8(dg));
    st16(dgu + t * 2 * f + f + pk * 8, pack8(du));
  }
}

int GridFor(size_t items) {
  size_t want = (items + kThreads - 1) / kThreads;
  const size_t cap = 148 * 8;
  if (want > cap) want = cap;
  return want < 1 ? 1 : static_cast<int>(want);
}

}  // namespace

extern "C" int ps_launch_rope_split(const void* qkv, void* q, void* k, void* v, const float* cos_t,
                                    const float* sin_t, size_t tokens, int seq, int n_heads,
                                    int n_kv, int hd, ps_stream_t stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (hd % 16 != 0) return cudaErrorInvalidValue;
  const size_t row = static_cast<size_t>(n_heads + 2 * n_kv) * hd;
  const auto* src = static_cast<const __nv_bfloat16*>(qkv);
  k_rope<false><<<GridFor(tokens * n_heads * (hd / 16)), kThreads, 0, st>>>(
      src, static_cast<__nv_bfloat16*>(q), cos_t, sin_t, tokens, seq, n_heads, hd, row, 0);
  k_rope<false><<<GridFor(tokens * n_kv * (hd / 16)), kThreads, 0, st>>>(
      src, static_cast<__nv_bfloat16*>(k), cos_t, sin_t, tokens, seq, n_kv, hd, row,
      static_cast<size_t>(n_heads) * hd);
  k_gather_rows<<<GridFor(tokens * (n_kv * hd / 8)), kThreads, 0, st>>>(
      src, static_cast<__nv_bfloat16*>(v), tokens, n_kv * hd, row,
      static_cast<size_t>(n_heads + n_kv) * hd);
  ps_kernels_internal::CountLaunch(3);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ps_launch_rope_merge_bwd(const void* dq, const void* dk, const void* dv, void* dqkv,
                                        const float* cos_t, const float* sin_t, size_t tokens,
                                        int seq, int n_heads, int n_kv, int hd, ps_stream_t stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (hd % 16 != 0) return cudaErrorInvalidValue;
  const size_t row = static_cast<size_t>(n_heads + 2 * n_kv) * hd;
  auto* dst = static_cast<__nv_bfloat16*>(dqkv);
  k_rope<true><<<GridFor(tokens * n_heads * (hd / 16)), kThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(dq), dst, cos_t, sin_t, tokens, seq, n_heads, hd, row, 0);
  k_rope<true><<<GridFor(tokens * n_kv * (hd / 16)), kThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(dk), dst, cos_t, sin_t, tokens, seq, n_kv, hd, row,
      static_cast<size_t>(n_heads) * hd);
  k_scatter_rows<<<GridFor(tokens * (n_kv * hd / 8)), kThreads, 0, st>>>(
      static_cast<const __nv_bfloat16*>(dv), dst, tokens, n_kv * hd, row,
      static_cast<size_t>(n_heads + n_kv) * hd);
  ps_kernels_internal::CountLaunch(3);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ps_launch_swiglu_fwd(const void* gu, void* out, size_t tokens, int f,
                                    ps_stream_t stream_) {
  if (f % 8 != 0) return cudaErrorInvalidValue;
  k_swiglu_fwd<<<GridFor(tokens * (f / 8)), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(gu), static_cast<__nv_bfloat16*>(out), tokens, f);
  ps_kernels_internal::CountLaunch(1);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ps_launch_swiglu_bwd(const void* gu, const void* dout, void* dgu, size_t tokens, int f,
                                    ps_stream_t stream_) {
  if (f % 8 != 0) return cudaErrorInvalidValue;
  k_swiglu_bwd<<<GridFor(tokens * (f / 8)), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(gu), static_cast<const __nv_bfloat16*>(dout),
      static_cast<__nv_bfloat16*>(dgu), tokens, f);
  ps_kernels_internal::CountLaunch(1);
  return static_cast<int>(cudaGetLastError());
}
