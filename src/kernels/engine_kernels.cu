/**
 * \file engine_kernels.cu
 * \brief The copy engine: a work-queue driven, on-demand persistent kernel — the "NIC" of the
 *        NVLink van (sm_100a).
 *
 * An RDMA NIC takes work requests from a queue in host memory, moves the bytes, and writes a
 * completion (the reference drives ibv_post_send / ibv_poll_cq this way,
 * src/rdma_transport.h:211-231, src/rdma_van.h:609-709). A kernel launch per message is the
 * wrong shape for that: a 1 KB ... 4 MB push costs 1-6 us of copy time but 3-5 us of driver time
 * on the issuing thread, every time. Here the van POSTS a 64-byte descriptor into a ring in
 * mapped host memory (a few stores, no driver call) and a resident kernel executes it:
 *
 *   CTA 0 (dispatcher, one thread)   polls `posted` in host memory, copies new descriptors into
 *                                    a device-side ring and publishes its tail at gpu scope
 *   CTAs 1..G-1 (workers)            walk the device ring in order; an item is cut into 64 KB
 *                                    chunks and spread over as many workers as it has chunks
 *                                    (start CTA rotates, so small items land on different CTAs
 *                                    and many items are in flight at once); stores may target
 *                                    peer HBM over NVLink
 *   completion                       every participating CTA fences at system scope and counts
 *                                    itself in; the last one waits for its turn (completions
 *                                    are published in posting order, like a stream) and stores
 *                                    the item's value to its flag with st.release.sys — the
 *                                    gate word of the receiver's descriptor ring
 *
 * The kernel is "on-demand persistent": it exits after `idle_us` without work (so device-wide
 * synchronisation still terminates) and the next post relaunches it. The exit is a two-phase
 * handshake through mapped host memory (intent -> re-check `posted` -> final), which makes the
 * race "engine decides to leave while the host posts" lose-free without atomics across PCIe.
 *
 * The engine only takes work that needs no stream ordering (no producer event) — everything
 * else stays on the launch path; CudaDomain keeps the two paths from overtaking each other.
 */
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "kernels/ps_kernels.h"

namespace ps_kernels_internal {
void CountLaunch(int n);
int NumSMs();
}  // namespace ps_kernels_internal

namespace {

constexpr int kThreads = 256;
constexpr unsigned kHostRing = 4096;     // descriptors in mapped host memory (power of two)
constexpr unsigned kDevRing = 256;       // descriptors staged in device memory (power of two)
constexpr unsigned long long kChunk = 64 * 1024;

struct Item {  // 64 bytes
  unsigned char* dst;
  const unsigned char* src;
  unsigned long long bytes;
  unsigned long long* flag;
  unsigned long long flag_value;
  unsigned long long reserved[3];
};
static_assert(sizeof(Item) == 64, "descriptor is one 64-byte line");

/*! \brief host <-> device mailbox, in mapped pinned host memory; one writer per word */
struct HostCtl {
  alignas(64) volatile unsigned long long posted;       // host -> device: descriptors written so far
  alignas(64) volatile unsigned long long retired;      // device -> host: completions published so far
  alignas(64) volatile unsigned long long exit_intent;  // device -> host: head+1 while it considers leaving
  alignas(64) volatile unsigned long long exit_final;   // device -> host: epoch of the kernel that left
  alignas(64) volatile unsigned long long exit_head;    // device -> host: first descriptor it did not take
};

struct DevState {
  Item ring[kDevRing];
  unsigned arrive[kDevRing];
  unsigned long long tail;          // descriptors published to the workers
  unsigned long long signal_head;   // next completion to publish
  unsigned long long stop_at;       // ~0 while running; the dispatcher's final head when leaving
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const volatile unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(volatile unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

/*! \brief all threads of a CTA copy [off, off+len) of the item; 16-byte vectors when aligned */
__device__ __forceinline__ void copy_chunk(unsigned char* __restrict__ dst,
                                           const unsigned char* __restrict__ src,
                                           unsigned long long len, bool aligned) {
  unsigned long long done = 0;
  if (aligned) {
    const unsigned long long n16 = len / 16;
    int4* d4 = reinterpret_cast<int4*>(dst);
    const int4* s4 = reinterpret_cast<const int4*>(src);
    unsigned long long i = threadIdx.x;
    for (; i + 3 * kThreads < n16; i += 4 * kThreads) {
      int4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_stream(s4 + i + u * kThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u) st_stream(d4 + i + u * kThreads, v[u]);
    }
    for (; i < n16; i += kThreads) st_stream(d4 + i, ld_stream(s4 + i));
    done = n16 * 16;
  }
  for (unsigned long long i = done + threadIdx.x; i < len; i += kThreads) dst[i] = src[i];
}

__global__ void __launch_bounds__(kThreads)
k_copy_engine(HostCtl* ctl, const Item* host_ring, DevState* st, unsigned long long head0,
              unsigned long long epoch, unsigned long long idle_ns) {
  const unsigned W = gridDim.x - 1;  // worker CTAs
  if (blockIdx.x == 0) {
    // ---------------- dispatcher ----------------
    if (threadIdx.x != 0) return;
    unsigned long long head = head0, idle_since = 0;
    for (;;) {
      const unsigned long long posted = ld_acquire_sys(&ctl->posted);
      if (posted != head) {
        idle_since = 0;
        while (head != posted) {
          // a slot of the device ring is free again once its completion has been published
          while (head - ld_acquire_gpu(&st->signal_head) >= kDevRing) {
          }
          const int4* s = reinterpret_cast<const int4*>(host_ring + (head & (kHostRing - 1)));
          int4* d = reinterpret_cast<int4*>(&st->ring[head & (kDevRing - 1)]);
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = __ldcv(s + q);  // host memory: never from a stale cache line
          ++head;
          st_release_gpu(&st->tail, head);
        }
        continue;
      }
      const unsigned long long now = globaltimer_ns();
      if (idle_since == 0) {
        idle_since = now;
        continue;
      }
      if (now - idle_since < idle_ns) continue;
      // nothing for a while: leave — unless the host posts right now. Phase 1: say so ...
      st_release_sys(&ctl->exit_intent, head + 1);
      __threadfence_system();
      // ... phase 2: look again. The host does the mirror image (post, fence, read the intent),
      // so one of the two always notices the other.
      if (ld_acquire_sys(&ctl->posted) != head) {
        st_release_sys(&ctl->exit_intent, 0ull);
        idle_since = 0;
        continue;
      }
      st_release_gpu(&st->stop_at, head);
      while (ld_acquire_gpu(&st->signal_head) != head) {
      }
      ctl->exit_head = head;
      __threadfence_system();
      st_release_sys(&ctl->exit_final, epoch);
      return;
    }
  }
  // ---------------- workers ----------------
  __shared__ Item item;
  __shared__ int leave;
  const unsigned me = blockIdx.x - 1;
  unsigned rot = 0;  // where the participants of the current item start (same sequence in every CTA)
  for (unsigned long long k = head0;; ++k) {
    if (threadIdx.x == 0) {
      leave = 0;
      for (;;) {
        if (ld_acquire_gpu(&st->tail) > k) {
          item = st->ring[k & (kDevRing - 1)];
          break;
        }
        if (ld_acquire_gpu(&st->stop_at) <= k) {
          leave = 1;
          break;
        }
      }
    }
    __syncthreads();
    if (leave) return;
    const unsigned long long bytes = item.bytes;
    const unsigned long long nchunks = bytes ? (bytes + kChunk - 1) / kChunk : 0;
    const unsigned P = nchunks >= W ? W : (nchunks ? static_cast<unsigned>(nchunks) : 1u);
    const unsigned r = (me + W - rot) % W;
    rot = (rot + P) % W;
    if (r < P) {
      const bool aligned =
          ((reinterpret_cast<unsigned long long>(item.dst) | reinterpret_cast<unsigned long long>(item.src)) & 15) == 0;
      for (unsigned long long c = r; c < nchunks; c += P) {
        const unsigned long long off = c * kChunk;
        const unsigned long long len = bytes - off < kChunk ? bytes - off : kChunk;
        copy_chunk(item.dst + off, item.src + off, len, aligned);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&st->arrive[k & (kDevRing - 1)], 1u) == P - 1) {
          // every chunk of item k is in place; completions leave in posting order
          while (ld_acquire_gpu(&st->signal_head) != k) {
          }
          st->arrive[k & (kDevRing - 1)] = 0;
          __threadfence_system();
          if (item.flag) st_release_sys(item.flag, item.flag_value);
          st_release_sys(&ctl->retired, k + 1);
          st_release_gpu(&st->signal_head, k + 1);
        }
      }
    }
    __syncthreads();  // `item` is overwritten by the next iteration
  }
}

inline unsigned long long HostLoad(const volatile unsigned long long* p) {
  return __atomic_load_n(const_cast<const unsigned long long*>(p), __ATOMIC_ACQUIRE);
}

}  // namespace

struct ps_engine {
  int device = 0;
  int grid = 0;
  unsigned long long idle_ns = 0;
  cudaStream_t stream = nullptr;
  HostCtl* ctl = nullptr;        // host address
  HostCtl* ctl_dev = nullptr;    // the same memory as the device sees it
  Item* ring = nullptr;
  Item* ring_dev = nullptr;
  DevState* state = nullptr;
  std::mutex mu;
  unsigned long long posted = 0;
  unsigned long long epoch = 0;
  unsigned long long next_head = 0;  // where the next kernel starts
  bool running = false;
  std::atomic<unsigned long long> launches{0};
  std::atomic<unsigned long long> items{0};
};

namespace {

bool EngineLaunch(ps_engine* e) {
  e->ctl->exit_intent = 0;
  ++e->epoch;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  if (cudaMemsetAsync(&e->state->stop_at, 0xff, sizeof(unsigned long long), e->stream) != cudaSuccess) {  // "never"
    return false;
  }
  k_copy_engine<<<e->grid, kThreads, 0, e->stream>>>(e->ctl_dev, e->ring_dev, e->state, e->next_head, e->epoch,
                                                     e->idle_ns);
  if (cudaGetLastError() != cudaSuccess) return false;
  ps_kernels_internal::CountLaunch(1);
  ++e->launches;
  e->running = true;
  return true;
}

}  // namespace

extern "C" ps_engine* ps_engine_create(int device, int num_ctas, int idle_us) {
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  ps_engine* e = new ps_engine();
  e->device = device;
  // every CTA must be resident at once (workers wait for each other's completions)
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_copy_engine, kThreads, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    delete e;
    return nullptr;
  }
  const int sms = ps_kernels_internal::NumSMs();
  int grid = num_ctas > 0 ? num_ctas : sms + 1;  // default: one worker per SM + the dispatcher
  if (grid > sms * (per_sm > 2 ? 2 : per_sm)) grid = sms * (per_sm > 2 ? 2 : per_sm);
  if (grid < 2) grid = 2;
  e->grid = grid;
  e->idle_ns = static_cast<unsigned long long>(idle_us > 0 ? idle_us : 200) * 1000ull;
  bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&e->ctl), sizeof(HostCtl), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
  ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&e->ring), sizeof(Item) * kHostRing, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&e->state), sizeof(DevState)) == cudaSuccess;
  if (ok) {
    memset(const_cast<HostCtl*>(e->ctl), 0, sizeof(HostCtl));
    memset(e->ring, 0, sizeof(Item) * kHostRing);
    ok = cudaMemset(e->state, 0, sizeof(DevState)) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->ctl_dev), const_cast<HostCtl*>(e->ctl), 0) == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->ring_dev), e->ring, 0) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    ps_engine_destroy(e);
    return nullptr;
  }
  return e;
}

extern "C" void ps_engine_destroy(ps_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);  // the kernel leaves by itself once idle
  if (e->state) cudaFree(e->state);
  if (e->ring) cudaFreeHost(e->ring);
  if (e->ctl) cudaFreeHost(const_cast<HostCtl*>(e->ctl));
  if (e->stream) cudaStreamDestroy(e->stream);
  cudaGetLastError();
  delete e;
}

extern "C" int ps_engine_post(ps_engine* e, void* dst, const void* src, size_t bytes, unsigned long long* flag,
                              unsigned long long value) {
  std::lock_guard<std::mutex> lk(e->mu);
  // back-pressure: the host ring holds kHostRing descriptors that have not been retired
  while (e->posted - HostLoad(&e->ctl->retired) >= kHostRing - 1) std::this_thread::yield();
  Item& it = e->ring[e->posted & (kHostRing - 1)];
  it.dst = static_cast<unsigned char*>(dst);
  it.src = static_cast<const unsigned char*>(src);
  it.bytes = bytes;
  it.flag = flag;
  it.flag_value = value;
  ++e->posted;
  __atomic_store_n(const_cast<unsigned long long*>(&e->ctl->posted), e->posted, __ATOMIC_RELEASE);
  ++e->items;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  if (e->running) {
    if (HostLoad(&e->ctl->exit_final) != e->epoch) {
      const unsigned long long intent = HostLoad(&e->ctl->exit_intent);
      if (intent == 0) return 0;  // alive, and it will see this descriptor
      // it is thinking about leaving: it either notices the post and stays, or finishes leaving
      for (;;) {
        if (HostLoad(&e->ctl->exit_final) == e->epoch) break;
        if (HostLoad(&e->ctl->exit_intent) == 0) return 0;
      }
    }
    e->running = false;
    e->next_head = HostLoad(&e->ctl->exit_head);
  }
  if (cudaSetDevice(e->device) != cudaSuccess) return -1;
  return EngineLaunch(e) ? 0 : -1;
}

extern "C" int ps_engine_idle(ps_engine* e) {
  std::lock_guard<std::mutex> lk(e->mu);
  return HostLoad(&e->ctl->retired) == e->posted;
}

extern "C" void ps_engine_drain(ps_engine* e) {
  unsigned long long upto;
  {
    std::lock_guard<std::mutex> lk(e->mu);
    upto = e->posted;
  }
  while (HostLoad(&e->ctl->retired) < upto) std::this_thread::yield();
}

extern "C" void ps_engine_stats(ps_engine* e, unsigned long long* launches, unsigned long long* items) {
  if (launches) *launches = e->launches.load();
  if (items) *items = e->items.load();
}
