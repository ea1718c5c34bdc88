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
 *   CTA 0, warp 0 (dispatcher)       polls `posted` in host memory, copies new descriptors into
 *                                    a device-side ring (eight per PCIe round trip) and
 *                                    publishes its tail at gpu scope
 *   CTAs 1..G-1 (workers)            walk the device ring in order; an item is cut into 512 KB
 *                                    chunks which the workers TAKE one at a time (claim counter
 *                                    per item) and move with cp.async.bulk through shared memory
 *                                    (TMA) — many items are in flight at once, a slow chunk
 *                                    (NVLink) never holds up the rest; stores may target peer HBM
 *   completion (one thread of CTA 0) every participating CTA fences at system scope and counts
 *                                    itself in; the completer walks the items in posting order
 *                                    (completions are published like on a stream) and stores an
 *                                    item's value to its flag with st.release.sys — the gate
 *                                    word of the receiver's descriptor ring. Finished neighbours
 *                                    that signal the same word are folded into one store.
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
constexpr unsigned kDevRing = 1024;      // descriptors staged in device memory (power of two)
constexpr unsigned kMaxWorkers = 512;

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
  unsigned arrive[kDevRing];   // chunks of the item in this slot that are in place
  unsigned claim[kDevRing];    // chunks of the item in this slot that have been taken
  unsigned long long tail;          // descriptors published to the workers
  unsigned long long signal_head;   // next completion to publish
  unsigned long long stop_at;       // ~0 while running; the dispatcher's final head when leaving
  unsigned long long progress[kMaxWorkers];  // per worker CTA: first descriptor it has not read yet
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
    const unsigned T = blockDim.x;
    for (; i + 3 * T < n16; i += 4 * T) {
      int4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_stream(s4 + i + u * T);
#pragma unroll
      for (int u = 0; u < 4; ++u) st_stream(d4 + i + u * T, v[u]);
    }
    for (; i < n16; i += T) st_stream(d4 + i, ld_stream(s4 + i));
    done = n16 * 16;
  }
  for (unsigned long long i = done + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}

// ---- TMA flavour of the chunk copy: one elected thread streams the chunk global -> shared ->
// global with cp.async.bulk (UBLKCP), kTmaStages x 32 KB in flight per CTA and no register staging.
// A CTA sustains 2-3x the bytes in flight of the LDG/STG loop, so far fewer CTAs saturate NVLink.
constexpr int kTmaStages = 4;
constexpr unsigned kTmaSub = 32 * 1024;

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "ENG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra ENG_DONE;\n\t"
      "bra ENG_WAIT;\n\t"
      "ENG_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

/*!
 * \brief ONE thread copies [0, len16) (a multiple of 16, both addresses 16-byte aligned). `seq`
 *        counts the 32 KB pieces this CTA has ever moved: it picks the stage and the mbarrier
 *        parity, so the barriers keep their phase from chunk to chunk. All stores have completed
 *        (not just left shared memory) when this returns.
 */
__device__ __forceinline__ void tma_copy(unsigned char* dst, const unsigned char* src, unsigned long long len16,
                                         unsigned char* smem, unsigned long long* full, unsigned long long& seq) {
  const unsigned long long n = (len16 + kTmaSub - 1) / kTmaSub;
  constexpr int kLook = kTmaStages - 2;  // loads run this many pieces ahead of the stores
  auto piece = [&](unsigned long long j) -> unsigned {
    const unsigned long long left = len16 - j * kTmaSub;
    return static_cast<unsigned>(left < kTmaSub ? left : kTmaSub);
  };
  for (unsigned long long j = 0; j < n + kLook; ++j) {
    if (j < n) {
      const int sidx = static_cast<int>((seq + j) % kTmaStages);
      // the stage was last read by the store of piece j - kTmaStages (issued kLook steps after its
      // load): at most one younger store may still be reading shared memory
      if (j >= static_cast<unsigned long long>(kTmaStages)) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      const unsigned b = piece(j);
      mbar_expect_tx(&full[sidx], b);
      bulk_g2s(smem + static_cast<size_t>(sidx) * kTmaSub, src + j * kTmaSub, b, &full[sidx]);
    }
    if (j >= static_cast<unsigned long long>(kLook)) {
      const unsigned long long i = j - kLook;
      const int sidx = static_cast<int>((seq + i) % kTmaStages);
      mbar_wait(&full[sidx], static_cast<unsigned>(((seq + i) / kTmaStages) & 1));
      bulk_s2g(dst + i * kTmaSub, smem + static_cast<size_t>(sidx) * kTmaSub, piece(i));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");  // async-proxy stores before the generic-proxy signalling
  seq += n;
}

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(volatile unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/*! \brief pieces an item is cut into (an empty item — a pure signal — counts as one) */
__device__ __forceinline__ unsigned NumChunks(unsigned long long bytes, unsigned long long kChunk) {
  const unsigned long long n = (bytes + kChunk - 1) / kChunk;
  return n ? static_cast<unsigned>(n) : 1u;
}

template <bool TMA>
__global__ void __launch_bounds__(kThreads)
k_copy_engine(HostCtl* ctl, const Item* host_ring, DevState* st, unsigned long long head0,
              unsigned long long epoch, unsigned long long idle_ns, unsigned long long kChunk) {
  const unsigned W = gridDim.x - 1;  // worker CTAs
  if (blockIdx.x == 0) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
      // ---------------- dispatcher (one warp) ----------------
      // lane 0 polls `posted` in host memory; new descriptors cross PCIe eight at a time (every
      // lane moves 16 bytes), so a burst costs one round trip per eight messages, not one each
      unsigned long long head = head0, idle_since = 0, safe = head0;
      for (;;) {
        unsigned long long posted = 0;
        if (lane == 0) posted = ld_acquire_sys(&ctl->posted);
        posted = __shfl_sync(0xffffffffu, posted, 0);
        if (posted != head) {
          idle_since = 0;
          while (head != posted) {
            const unsigned n = posted - head < 8 ? static_cast<unsigned>(posted - head) : 8u;
            // a slot of the device ring is free again once its completion has been published AND
            // every worker has moved past it (a worker that found nothing left to take may still be
            // about to look at the slot's claim counter)
            if (head + n - safe > kDevRing) {
              for (;;) {
                unsigned long long m = ld_acquire_gpu(&st->signal_head);
                for (unsigned w = lane; w < W; w += 32) {
                  const unsigned long long p = ld_acquire_gpu(&st->progress[w]);
                  m = p < m ? p : m;
                }
                for (int o = 16; o > 0; o >>= 1) {
                  const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
                  m = other < m ? other : m;
                }
                safe = m;
                if (head + n - safe <= kDevRing) break;
              }
            }
            __syncwarp();
            if ((lane >> 2) < n) {
              const unsigned long long idx = head + (lane >> 2);
              const int4 v = __ldcv(reinterpret_cast<const int4*>(host_ring + (idx & (kHostRing - 1))) + (lane & 3));
              reinterpret_cast<int4*>(&st->ring[idx & (kDevRing - 1)])[lane & 3] = v;
              if ((lane & 3) == 0) {  // nobody looks at the slot's old counters any more (see `safe`)
                st->arrive[idx & (kDevRing - 1)] = 0;
                st->claim[idx & (kDevRing - 1)] = 0;
              }
            }
            __syncwarp();
            head += n;
            if (lane == 0) {
              __threadfence();
              st_release_gpu(&st->tail, head);
            }
          }
          continue;
        }
        int leave = 0;
        if (lane == 0) {
          const unsigned long long now = globaltimer_ns();
          if (ld_acquire_gpu(&st->signal_head) != head) {
            idle_since = 0;  // copies still running: not idle
          } else if (idle_since == 0) {
            idle_since = now;
          } else if (now - idle_since >= idle_ns) {
            // nothing for a while: leave — unless the host posts right now. Phase 1: say so ...
            st_release_sys(&ctl->exit_intent, head + 1);
            __threadfence_system();
            // ... phase 2: look again. The host does the mirror image (post, fence, read the
            // intent), so one of the two always notices the other.
            if (ld_acquire_sys(&ctl->posted) != head) {
              st_release_sys(&ctl->exit_intent, 0ull);
              idle_since = 0;
            } else {
              st_release_gpu(&st->stop_at, head);  // the completer finishes the handshake
              leave = 1;
            }
          }
        }
        if (__shfl_sync(0xffffffffu, leave, 0)) return;
      }
    }
    if (warp == 1 && lane == 0) {
      // ---------------- completer (one thread) ----------------
      // Publishes completions in posting order, like a stream: item k is complete when all of
      // its participants have counted themselves in (each after a system-scope fence behind its
      // stores). Consecutive finished items that signal the same word are folded into ONE
      // st.release.sys of the newest value: the receiver only compares "reached k yet?".
      unsigned long long k = head0;
      for (;;) {
        unsigned long long tail = ld_acquire_gpu(&st->tail);
        if (tail == k) {
          if (ld_acquire_gpu(&st->stop_at) == k) {
            ctl->exit_head = k;
            __threadfence_system();
            st_release_sys(&ctl->exit_final, epoch);
            return;
          }
          continue;
        }
        const Item* it = &st->ring[k & (kDevRing - 1)];
        const unsigned P = NumChunks(it->bytes, kChunk);
        while (ld_acquire_gpu_u32(&st->arrive[k & (kDevRing - 1)]) != P) {
        }
        unsigned long long* flag = it->flag;
        unsigned long long value = it->flag_value;
        ++k;
        // fold in whatever else has finished already
        while (k != tail) {
          const Item* nx = &st->ring[k & (kDevRing - 1)];
          if (nx->flag != flag) break;
          if (ld_acquire_gpu_u32(&st->arrive[k & (kDevRing - 1)]) != NumChunks(nx->bytes, kChunk)) break;
          value = nx->flag_value;
          ++k;
        }
        if (flag) {
          st_release_sys(flag, value);
          st_relaxed_sys(&ctl->retired, k);
        } else {
          st_release_sys(&ctl->retired, k);
        }
        st_release_gpu(&st->signal_head, k);
      }
    }
    return;
  }
  // ---------------- workers ----------------
  __shared__ Item item;
  __shared__ int leave;
  extern __shared__ __align__(128) unsigned char stage_mem[];  // TMA flavour: kTmaStages x 32 KB
  __shared__ __align__(8) unsigned long long full[kTmaStages];
  unsigned long long tma_seq = 0;
  if (TMA) {
    if (threadIdx.x == 0) {
      for (int q = 0; q < kTmaStages; ++q) mbar_init(&full[q], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const unsigned me = blockIdx.x - 1;
  __shared__ unsigned my_chunk;
  // Every worker walks the items in posting order and TAKES chunks (atomic claim counter per item)
  // until none is left: a worker stuck on a slow chunk (NVLink) never holds up chunks that were
  // statically "its own", the others simply take more — items of different cost mix freely.
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
    const unsigned nchunks = NumChunks(bytes, kChunk);
    const bool aligned =
        ((reinterpret_cast<unsigned long long>(item.dst) | reinterpret_cast<unsigned long long>(item.src)) & 15) == 0;
    unsigned* claim = &st->claim[k & (kDevRing - 1)];
    for (;;) {
      if (threadIdx.x == 0) {
        // look before taking: once an item is fully handed out the other ~150 workers pass with a load
        unsigned c = nchunks;
        if (ld_acquire_gpu_u32(claim) < nchunks) c = atomicAdd(claim, 1u);
        my_chunk = c;
      }
      __syncthreads();
      const unsigned c = my_chunk;
      if (c >= nchunks) break;
      if (bytes) {
        const unsigned long long off = static_cast<unsigned long long>(c) * kChunk;
        const unsigned long long len = bytes - off < kChunk ? bytes - off : kChunk;
        if (TMA && aligned) {
          const unsigned long long len16 = len & ~15ull;
          if (threadIdx.x == 0) {
            if (len16) tma_copy(item.dst + off, item.src + off, len16, stage_mem, full, tma_seq);
            for (unsigned long long t = len16; t < len; ++t) item.dst[off + t] = item.src[off + t];  // < 16 bytes
          }
        } else {
          copy_chunk(item.dst + off, item.src + off, len, aligned);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system();  // the chunk is visible system-wide before it is counted
        atomicAdd(&st->arrive[k & (kDevRing - 1)], 1u);
      }
    }
    if (threadIdx.x == 0) st_release_gpu(&st->progress[me], k + 1);  // done with this slot for good
    __syncthreads();  // `item` / `my_chunk` are overwritten by the next iteration
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
  // bytes one worker CTA moves per turn (PS_ENGINE_CHUNK_KB) and how it moves them (PS_ENGINE_TMA=1:
  // cp.async.bulk through shared memory, one worker per SM; 0: LDG/STG, two workers per SM). Measured,
  // 4 MB / 16 MB messages (profiles/r2/engine_bench_tma_sweep.txt):
  //   local HBM   LDG 256 KB 1300 / 1984 GB/s   TMA 256 KB 1637 / 2662   TMA 512 KB 1789 / 2967 (0.90 of the copy roofline)
  //   NVLink      LDG 256 KB  649 /  666 GB/s   TMA 256 KB  611 /  676   TMA 512 KB  649 /  700
  unsigned long long chunk = 512 * 1024;
  bool tma = true;
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
  if (e->tma) {
    k_copy_engine<true><<<e->grid, 64, kTmaStages * kTmaSub, e->stream>>>(e->ctl_dev, e->ring_dev, e->state,
                                                                         e->next_head, e->epoch, e->idle_ns, e->chunk);
  } else {
    k_copy_engine<false><<<e->grid, kThreads, 0, e->stream>>>(e->ctl_dev, e->ring_dev, e->state, e->next_head,
                                                             e->epoch, e->idle_ns, e->chunk);
  }
  if (cudaGetLastError() != cudaSuccess) return false;
  ps_kernels_internal::CountLaunch(1);
  ++e->launches;
  e->running = true;
  return true;
}

}  // namespace

namespace {
// One engine per device and process: two of them (the worker van's and the server van's of a joint
// process) would fight for the SMs' shared memory and take turns instead of running together.
std::mutex g_registry_mu;
ps_engine* g_engines[64] = {nullptr};
int g_refs[64] = {0};
ps_engine* EngineCreate(int device, int num_ctas, int idle_us);
void EngineDestroy(ps_engine* e);
}  // namespace

extern "C" ps_engine* ps_engine_create(int device, int num_ctas, int idle_us) {
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_registry_mu);
  if (!g_engines[device]) g_engines[device] = EngineCreate(device, num_ctas, idle_us);
  if (g_engines[device]) ++g_refs[device];
  return g_engines[device];
}

extern "C" void ps_engine_destroy(ps_engine* e) {
  if (!e) return;
  std::lock_guard<std::mutex> lk(g_registry_mu);
  const int device = e->device;
  if (device >= 0 && device < 64 && g_engines[device] == e) {
    if (--g_refs[device] > 0) return;
    g_engines[device] = nullptr;
  }
  EngineDestroy(e);
}

namespace {
ps_engine* EngineCreate(int device, int num_ctas, int idle_us) {
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  ps_engine* e = new ps_engine();
  e->device = device;
  if (const char* t = getenv("PS_ENGINE_TMA")) e->tma = atoi(t) != 0;
  // every CTA must be resident at once (the dispatcher waits for the slowest worker)
  int per_sm = 0;
  cudaError_t occ;
  if (e->tma) {
    cudaFuncSetAttribute(k_copy_engine<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaStages * kTmaSub);
    occ = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_copy_engine<true>, 64, kTmaStages * kTmaSub);
  } else {
    occ = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_copy_engine<false>, kThreads, 0);
  }
  if (occ != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    delete e;
    return nullptr;
  }
  const int sms = ps_kernels_internal::NumSMs();
  const int cap = per_sm > 2 ? 2 : per_sm;            // LDG: two workers per SM; TMA (128 KB each): one
  int grid = num_ctas > 0 ? num_ctas : cap * sms;     // (the dispatcher's CTA takes one of the slots)
  if (grid > sms * cap) grid = sms * cap;
  if (grid > static_cast<int>(kMaxWorkers)) grid = static_cast<int>(kMaxWorkers);
  if (grid < 2) grid = 2;
  e->grid = grid;
  e->idle_ns = static_cast<unsigned long long>(idle_us > 0 ? idle_us : 200) * 1000ull;
  if (const char* ck = getenv("PS_ENGINE_CHUNK_KB")) {
    const long kb = atol(ck);
    if (kb >= 4 && kb <= 16384) e->chunk = static_cast<unsigned long long>(kb) * 1024ull;
  }
  bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
  const unsigned host_flags = cudaHostAllocMapped | cudaHostAllocPortable;
  ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&e->ctl), sizeof(HostCtl), host_flags) == cudaSuccess;
  ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&e->ring), sizeof(Item) * kHostRing, host_flags) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&e->state), sizeof(DevState)) == cudaSuccess;
  if (ok) {
    memset(const_cast<HostCtl*>(e->ctl), 0, sizeof(HostCtl));
    memset(e->ring, 0, sizeof(Item) * kHostRing);
    ok = cudaMemset(e->state, 0, sizeof(DevState)) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->ctl_dev), const_cast<HostCtl*>(e->ctl), 0) ==
                   cudaSuccess;
    ok = ok && cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->ring_dev), e->ring, 0) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    EngineDestroy(e);
    return nullptr;
  }
  return e;
}

void EngineDestroy(ps_engine* e) {
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
}  // namespace

namespace {
/*!
 * \brief wait until `target` descriptors have retired. A kernel that faulted (a descriptor naming
 *        memory that is gone) never retires anything: after every second without progress the
 *        engine's stream is asked for its status, and a CUDA error ends the process with a message
 *        instead of a silent hang (the reference's counterpart: a failed work completion in PollCQ
 *        is fatal, src/rdma_van.h:609-616).
 */
void WaitRetired(ps_engine* e, unsigned long long target) {
  unsigned long long seen = HostLoad(&e->ctl->retired);
  if (seen >= target) return;
  auto since = std::chrono::steady_clock::now();
  for (unsigned spins = 0;; ++spins) {
    const unsigned long long now_retired = HostLoad(&e->ctl->retired);
    if (now_retired >= target) return;
    std::this_thread::yield();
    if ((spins & 1023u) != 1023u) continue;
    const auto now = std::chrono::steady_clock::now();
    if (now_retired != seen) {
      seen = now_retired;
      since = now;
    } else if (now - since > std::chrono::seconds(1)) {
      since = now;
      const cudaError_t st = cudaStreamQuery(e->stream);
      if (st != cudaSuccess && st != cudaErrorNotReady) {
        fprintf(stderr, "pslite copy engine (device %d): the engine kernel failed: %s; %llu of %llu descriptors retired\n",
                e->device, cudaGetErrorString(st), now_retired, target);
        abort();
      }
    }
  }
}
}  // namespace

extern "C" int ps_engine_post(ps_engine* e, void* dst, const void* src, size_t bytes, unsigned long long* flag,
                              unsigned long long value, unsigned long long* ticket) {
  std::lock_guard<std::mutex> lk(e->mu);
  // back-pressure: the host ring holds kHostRing descriptors that have not been retired
  if (e->posted - HostLoad(&e->ctl->retired) >= kHostRing - 1) WaitRetired(e, e->posted - (kHostRing - 2));
  Item& it = e->ring[e->posted & (kHostRing - 1)];
  it.dst = static_cast<unsigned char*>(dst);
  it.src = static_cast<const unsigned char*>(src);
  it.bytes = bytes;
  it.flag = flag;
  it.flag_value = value;
  ++e->posted;
  if (ticket) *ticket = e->posted;
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
  WaitRetired(e, upto);
}

extern "C" int ps_engine_done(ps_engine* e, unsigned long long ticket) {
  return HostLoad(&e->ctl->retired) >= ticket;
}

extern "C" void ps_engine_wait(ps_engine* e, unsigned long long ticket) {
  WaitRetired(e, ticket);
}

extern "C" void ps_engine_stats(ps_engine* e, unsigned long long* launches, unsigned long long* items) {
  if (launches) *launches = e->launches.load();
  if (items) *items = e->items.load();
}
