/**
 * \file engine_bench.cc
 * \brief The copy engine (src/kernels/engine_kernels.cu) against one kernel launch per copy, in
 *        isolation from the van: message rate, per-message latency and bandwidth as a host thread
 *        that only POLLS a completion word sees them (host clock around the completion of the
 *        last item — exactly what the van's receive thread observes).
 *
 *   usage: engine_bench [--peer] [--ctas N] [--idle-us N]
 *
 * For each message size: `burst` messages are posted back to back (flag values 1..burst on one
 * word, like the gate word of a descriptor ring), the thread waits for the last value; then the
 * same with ps_launch_copy_signal on a stream. `latency`: one message at a time. Prints JSON lines.
 * The data is verified after every configuration (each message has its own pattern).
 */
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels/ps_kernels.h"

#define CK(expr)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                   \
    }                                                                                            \
  } while (0)

namespace {

double NowUs() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool WaitFlag(volatile unsigned long long* flag, unsigned long long value, double timeout_s) {
  const double t0 = NowUs();
  while (__atomic_load_n(const_cast<unsigned long long*>(flag), __ATOMIC_ACQUIRE) < value) {
    if (NowUs() - t0 > timeout_s * 1e6) return false;
  }
  return true;
}

}  // namespace

int main(int argc, char** argv) {
  bool peer = false;
  int ctas = 0, idle_us = 200;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--peer")) peer = true;
    if (!strcmp(argv[i], "--ctas") && i + 1 < argc) ctas = atoi(argv[++i]);
    if (!strcmp(argv[i], "--idle-us") && i + 1 < argc) idle_us = atoi(argv[++i]);
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (peer && ndev < 2) {
    fprintf(stderr, "--peer needs two GPUs\n");
    return 2;
  }
  CK(cudaSetDevice(0));
  if (peer) {
    cudaError_t e = cudaDeviceEnablePeerAccess(1, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
    cudaGetLastError();
  }
  const size_t kMaxMsg = 16u << 20;
  const int kSlots = 24;  // 24 x 16 MB source + destination: far larger than the 126 MB L2
  unsigned char *src = nullptr, *dst = nullptr;
  CK(cudaMalloc(&src, kMaxMsg * kSlots));
  if (peer) CK(cudaSetDevice(1));
  CK(cudaMalloc(&dst, kMaxMsg * kSlots));
  CK(cudaSetDevice(0));
  CK(ps_launch_fill_u32(src, kMaxMsg * kSlots / 4, 12345u, nullptr) == 0 ? cudaSuccess : cudaErrorUnknown);
  CK(cudaDeviceSynchronize());
  unsigned long long* flag = nullptr;
  CK(cudaHostAlloc(reinterpret_cast<void**>(&flag), 64, cudaHostAllocMapped | cudaHostAllocPortable));
  unsigned long long* flag_dev = nullptr;
  CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&flag_dev), flag, 0));
  unsigned* counter = nullptr;
  CK(cudaMalloc(&counter, 256));
  CK(cudaMemset(counter, 0, 256));
  cudaStream_t stream;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  ps_engine* eng = ps_engine_create(0, ctas, idle_us);
  if (!eng) {
    fprintf(stderr, "engine creation failed\n");
    return 2;
  }
  unsigned long long* sums = nullptr;
  CK(cudaMalloc(&sums, 16));
  auto checksum = [&](const unsigned char* p, size_t bytes) {
    unsigned long long h = 0;
    ps_launch_checksum_u32(p, bytes / 4, sums, nullptr);
    CK(cudaMemcpy(&h, sums, 8, cudaMemcpyDeviceToHost));
    return h;
  };
  int failures = 0;
  const size_t sizes[] = {0, 1024, 65536, 1u << 20, 4096000, 16u << 20};
  for (size_t sz : sizes) {
    const int burst = sz >= (1u << 20) ? 240 : 2000;
    for (int mode = 0; mode < 2; ++mode) {  // 0 = engine, 1 = one launch per message
      // warm-up + timed
      double best_us = 1e30, lat_us = 0;
      for (int rep = 0; rep < 4; ++rep) {
        *flag = 0;
        CK(cudaMemsetAsync(dst, 0, sz ? sz * kSlots : 4, stream));
        CK(cudaStreamSynchronize(stream));
        const double t0 = NowUs();
        for (int i = 0; i < burst; ++i) {
          unsigned char* d = dst + static_cast<size_t>(i % kSlots) * kMaxMsg;
          const unsigned char* s = src + static_cast<size_t>(i % kSlots) * kMaxMsg;
          if (mode == 0) {
            if (ps_engine_post(eng, d, s, sz, flag_dev, static_cast<unsigned long long>(i + 1), nullptr) != 0) return 3;
          } else {
            ps_signal sig = {counter, flag_dev, static_cast<unsigned long long>(i + 1)};
            if (ps_launch_copy_signal(d, s, sz, PS_CODEC_RAW, 1.f, 0, &sig,
                                      reinterpret_cast<ps_stream_t>(stream)) != 0) return 3;
          }
        }
        const double t_posted = NowUs();
        if (!WaitFlag(flag, static_cast<unsigned long long>(burst), 20)) {
          fprintf(stderr, "timeout: flag=%llu of %d (mode %d size %zu)\n", *flag, burst, mode, sz);
          return 4;
        }
        const double t1 = NowUs();
        if (rep > 0 && t1 - t0 < best_us) {
          best_us = t1 - t0;
          lat_us = t_posted - t0;
        }
      }
      if (sz >= 4) {  // every slot must hold its source's bytes
        CK(cudaDeviceSynchronize());
        for (int i = 0; i < kSlots && i < burst; ++i) {
          if (checksum(dst + static_cast<size_t>(i) * kMaxMsg, sz & ~size_t(3)) !=
              checksum(src + static_cast<size_t>(i) * kMaxMsg, sz & ~size_t(3))) {
            ++failures;
            fprintf(stderr, "DATA MISMATCH mode %d size %zu slot %d\n", mode, sz, i);
            break;
          }
        }
      }
      // one at a time: post, wait, post, ...
      double one_us = 0;
      const int n1 = 200;
      *flag = 0;
      for (int i = 0; i < n1 + 20; ++i) {
        const double t0 = NowUs();
        if (mode == 0) {
          ps_engine_post(eng, dst, src, sz, flag_dev, static_cast<unsigned long long>(i + 1), nullptr);
        } else {
          ps_signal sig = {counter, flag_dev, static_cast<unsigned long long>(i + 1)};
          ps_launch_copy_signal(dst, src, sz, PS_CODEC_RAW, 1.f, 0, &sig, reinterpret_cast<ps_stream_t>(stream));
        }
        if (!WaitFlag(flag, static_cast<unsigned long long>(i + 1), 20)) return 4;
        if (i >= 20) one_us += NowUs() - t0;
      }
      printf("{\"bench\":\"engine_bench\",\"path\":\"%s\",\"peer\":%d,\"bytes\":%zu,\"burst\":%d,"
             "\"us_per_msg\":%.3f,\"host_us_per_post\":%.3f,\"GBps\":%.1f,\"one_at_a_time_us\":%.2f}\n",
             mode == 0 ? "engine" : "launch", peer ? 1 : 0, sz, burst, best_us / burst, lat_us / burst,
             sz * static_cast<double>(burst) / best_us / 1e3, one_us / n1);
      fflush(stdout);
    }
  }
  if (peer) {
    // both directions at once: an engine on each GPU writes into the other one's memory (what every
    // GPU of a joint job does: its worker pushes out while its server answers pulls)
    CK(cudaSetDevice(1));
    cudaError_t pe = cudaDeviceEnablePeerAccess(0, 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CK(pe);
    cudaGetLastError();
    unsigned char *src1 = nullptr, *dst0 = nullptr;
    CK(cudaMalloc(&src1, kMaxMsg * kSlots));
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&dst0, kMaxMsg * kSlots));
    unsigned long long* flag1 = nullptr;
    CK(cudaHostAlloc(reinterpret_cast<void**>(&flag1), 64, cudaHostAllocMapped | cudaHostAllocPortable));
    unsigned long long* flag1_dev = nullptr;
    CK(cudaSetDevice(1));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&flag1_dev), flag1, 0));
    ps_engine* eng1 = ps_engine_create(1, ctas, idle_us);
    if (!eng1) return 2;
    for (size_t sz : {static_cast<size_t>(4096000), static_cast<size_t>(16u << 20)}) {
      for (int both = 0; both < 2; ++both) {
        const int burst = 240;
        double best_us = 1e30;
        for (int rep = 0; rep < 4; ++rep) {
          *flag = 0;
          *flag1 = 0;
          const double t0 = NowUs();
          for (int i = 0; i < burst; ++i) {
            const size_t off = static_cast<size_t>(i % kSlots) * kMaxMsg;
            ps_engine_post(eng, dst + off, src + off, sz, flag_dev, static_cast<unsigned long long>(i + 1), nullptr);
            if (both) {
              ps_engine_post(eng1, dst0 + off, src1 + off, sz, flag1_dev, static_cast<unsigned long long>(i + 1), nullptr);
            }
          }
          if (!WaitFlag(flag, burst, 20) || (both && !WaitFlag(flag1, burst, 20))) return 4;
          const double t1 = NowUs();
          if (rep > 0 && t1 - t0 < best_us) best_us = t1 - t0;
        }
        printf("{\"bench\":\"engine_bench\",\"path\":\"engine\",\"pattern\":\"%s\",\"bytes\":%zu,"
               "\"GBps_per_direction\":%.1f}\n", both ? "0->1 and 1->0 at once" : "0->1 only", sz,
               sz * static_cast<double>(burst) / best_us / 1e3);
        fflush(stdout);
      }
    }
    ps_engine_destroy(eng1);
    CK(cudaSetDevice(0));
    // local and NVLink copies alternating in ONE queue (a joint worker's pushes: every other key lives
    // on the local server): the slow ones must not hold up the fast ones
    {
      const size_t sz = 4096000;
      const int burst = 240;
      double best_us = 1e30;
      for (int rep = 0; rep < 4; ++rep) {
        *flag = 0;
        const double t0 = NowUs();
        for (int i = 0; i < burst; ++i) {
          const size_t off = static_cast<size_t>(i % kSlots) * kMaxMsg;
          unsigned char* d = (i & 1) ? dst + off : dst0 + off;  // odd: GPU 1 (NVLink), even: GPU 0 (local)
          ps_engine_post(eng, d, src + off, sz, flag_dev, static_cast<unsigned long long>(i + 1), nullptr);
        }
        if (!WaitFlag(flag, burst, 20)) return 4;
        const double t1 = NowUs();
        if (rep > 0 && t1 - t0 < best_us) best_us = t1 - t0;
      }
      printf("{\"bench\":\"engine_bench\",\"path\":\"engine\",\"pattern\":\"local and 0->1 alternating\",\"bytes\":%zu,"
             "\"GBps_per_direction\":%.1f,\"us_per_msg\":%.2f}\n", sz, sz * static_cast<double>(burst) / best_us / 1e3,
             best_us / burst);
    }
  }
  unsigned long long launches = 0, items = 0;
  ps_engine_stats(eng, &launches, &items);
  printf("{\"engine_launches\":%llu,\"engine_items\":%llu,\"data_failures\":%d}\n", launches, items, failures);
  ps_engine_destroy(eng);
  return failures ? 1 : 0;
}
