/**
 * \file cuda_domain.cc
 * \brief CudaDomain implementation (see cuda_domain.h).
 */
#include "van/cuda_domain.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <unordered_map>

#include "kernels/ps_kernels.h"
#include "van/fd_exchange.h"

namespace ps {

#define PS_CUDA_CHECK(expr)                                                       \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    CHECK(e_ == cudaSuccess) << "CUDA: " #expr " -> " << cudaGetErrorString(e_);  \
  } while (0)

int CudaDeviceCount() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

namespace {

/*!
 * \brief the driver entry points of virtual memory management and NVSwitch multicast, resolved
 *        through the runtime (no link-time dependency on libcuda.so, absent on build hosts).
 *        This replaces what round 1 borrowed from torch.distributed._symmetric_memory.
 */
struct VmmApi {
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  bool vmm = false, multicast = false;

  static const VmmApi& Get() {
    static const VmmApi api = [] {
      VmmApi a;
      auto load = [](const char* name, void* slot) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
          cudaGetLastError();
          fn = nullptr;
        }
        *static_cast<void**>(slot) = fn;
        return fn != nullptr;
      };
      bool v = true, m = true;
      v &= load("cuDeviceGet", &a.DeviceGet);
      v &= load("cuDeviceGetAttribute", &a.DeviceGetAttribute);
      v &= load("cuMemGetAllocationGranularity", &a.MemGetAllocationGranularity);
      v &= load("cuMemCreate", &a.MemCreate);
      v &= load("cuMemRelease", &a.MemRelease);
      v &= load("cuMemAddressReserve", &a.MemAddressReserve);
      v &= load("cuMemAddressFree", &a.MemAddressFree);
      v &= load("cuMemMap", &a.MemMap);
      v &= load("cuMemUnmap", &a.MemUnmap);
      v &= load("cuMemSetAccess", &a.MemSetAccess);
      v &= load("cuMemExportToShareableHandle", &a.MemExportToShareableHandle);
      v &= load("cuMemImportFromShareableHandle", &a.MemImportFromShareableHandle);
      m &= load("cuMulticastCreate", &a.MulticastCreate);
      m &= load("cuMulticastAddDevice", &a.MulticastAddDevice);
      m &= load("cuMulticastBindMem", &a.MulticastBindMem);
      m &= load("cuMulticastGetGranularity", &a.MulticastGetGranularity);
      a.vmm = v;
      a.multicast = v && m;
      return a;
    }();
    return api;
  }
};

class CudaDomain : public MemDomain {
 public:
  explicit CudaDomain(int dev) : dev_(dev) {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    int lo = 0, hi = 0;
    PS_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PS_CUDA_CHECK(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, hi));
    max_ctas_ = GetEnv("PS_COPY_CTAS", 0);
    // PS_COPY_ENGINE=1: raw copies that need no stream ordering are posted to the copy engine
    // (an on-demand persistent kernel fed through a ring in mapped host memory) instead of
    // costing a kernel launch each. Meant for processes that own their GPU: a resident kernel
    // and another process's kernels would only alternate by time slice.
    if (GetEnv("PS_COPY_ENGINE", 0) != 0) {
      engine_ = ps_engine_create(dev_, GetEnv("PS_ENGINE_CTAS", 0), GetEnv("PS_ENGINE_IDLE_US", 200));
      if (!engine_) LOG(WARNING) << "PS_COPY_ENGINE: the copy engine could not be created; copies are launched";
    }
  }
  ~CudaDomain() override {
    cudaSetDevice(dev_);
    if (engine_) ps_engine_destroy(engine_);
    cudaStreamSynchronize(stream_);
    for (cudaEvent_t e : free_events_) cudaEventDestroy(e);
    for (auto& kv : imported_) cudaIpcCloseMemHandle(kv.second);
    for (auto& a : arenas_) cudaFree(a->base);
    if (sig_counter_) cudaFree(sig_counter_);
    cudaStreamDestroy(stream_);
  }
  const char* name() const override { return "nvl"; }
  int device() const override { return dev_; }
  void* Stream() override {
    // whoever asks may enqueue work of their own here: from now on the engine path checks that
    // the stream is idle every time (see CopySignal)
    stream_shared_.store(true, std::memory_order_release);
    return stream_;
  }

  bool Handles(int device_type, const void* /*ptr*/) override { return device_type == GPU; }

  /*!
   * Landing slots come from large device arenas (PS_NVL_ARENA_MB each, default 1024) carved
   * by an offset allocator: one CUDA IPC handle / peer mapping per arena instead of one per
   * slot, and a slot can be returned and re-cut without invalidating any peer's mapping
   * (cudaFree of memory a peer still has open is undefined). The reference's counterpart is
   * the registered-memory MemoryAllocator, src/rdma_utils.h:75-140.
   */
  void* Alloc(size_t bytes) override {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    ReclaimRetired(false);
    const uint64_t want = AlignUp(bytes ? bytes : 1, 512);
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& a : arenas_) {
        const uint64_t off = a->alloc.Alloc(want);
        if (off != UINT64_MAX) return a->base + off;
      }
    }
    const uint64_t chunk = static_cast<uint64_t>(GetEnv("PS_NVL_ARENA_MB", 1024)) << 20;
    std::unique_ptr<DevArena> a(new DevArena());
    a->size = std::max<uint64_t>(chunk, AlignUp(want, 2u << 20));
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, a->size);
    if (e != cudaSuccess) {
      // memory may be sitting in retired slots: wait for their users, then look again
      cudaGetLastError();
      ReclaimRetired(true);
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& old : arenas_) {
        const uint64_t off = old->alloc.Alloc(want);
        if (off != UINT64_MAX) return old->base + off;
      }
      e = cudaMalloc(&base, a->size);
    }
    if (e != cudaSuccess && a->size > want) {  // little memory left: fall back to an exact fit
      cudaGetLastError();
      a->size = AlignUp(want, 2u << 20);
      e = cudaMalloc(&base, a->size);
    }
    PS_CUDA_CHECK(e);
    a->base = static_cast<char*>(base);
    a->alloc.Reset(a->size, 512);
    const uint64_t off = a->alloc.Alloc(want);
    std::lock_guard<std::mutex> lk(mu_);
    arenas_.push_back(std::move(a));
    return arenas_.back()->base + off;
  }
  /*!
   * A kernel already enqueued may still read or write the slot, so its bytes must not be re-cut
   * yet — but stalling the whole device for that (cudaDeviceSynchronize, round 1) also stalls a
   * training step that happens to re-size one slot. The slot is RETIRED instead: an event marks
   * "everything enqueued on the data stream so far" and the engine ticket "everything posted so
   * far"; the bytes return to the arena once both have passed (checked on later Alloc / Free).
   */
  void Free(void* p) override {
    cudaSetDevice(dev_);
    Retired r;
    r.ptr = static_cast<char*>(p);
    r.event = AcquireEvent();
    if (cudaEventRecord(r.event, stream_) != cudaSuccess) {
      cudaGetLastError();
      cudaStreamSynchronize(stream_);  // cannot track it: fall back to waiting now
    }
    r.engine_ticket = engine_ticket_.load(std::memory_order_acquire);
    {
      std::lock_guard<std::mutex> lk(mu_);
      retired_.push_back(r);
    }
    ReclaimRetired(false);
  }

  /*! \brief give retired slots whose last possible user has finished back to their arena */
  void ReclaimRetired(bool wait) {
    std::vector<Retired> ready;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto it = retired_.begin(); it != retired_.end();) {
        bool done = !engine_ || ps_engine_done(engine_, it->engine_ticket);
        if (done) {
          const cudaError_t e = wait ? cudaEventSynchronize(it->event) : cudaEventQuery(it->event);
          done = e == cudaSuccess;
          if (!done) cudaGetLastError();
        } else if (wait) {
          ps_engine_wait(engine_, it->engine_ticket);
          done = cudaEventSynchronize(it->event) == cudaSuccess;
        }
        if (done) {
          ready.push_back(*it);
          it = retired_.erase(it);
        } else {
          ++it;
        }
      }
      for (const Retired& r : ready) {
        free_events_.push_back(r.event);
        bool ours = false;
        for (auto& a : arenas_) {
          if (r.ptr >= a->base && r.ptr < a->base + a->size) {
            a->alloc.Free(static_cast<uint64_t>(r.ptr - a->base));
            ours = true;
            break;
          }
        }
        if (!ours) cudaFree(r.ptr);  // not from an arena (never happens for slots handed out by Alloc)
      }
    }
  }

  bool Export(const void* p, RegionDesc* out) override {
    {
      // fast path: the pointer lies in an allocation that was exported before (one map
      // lookup instead of two driver calls on every pull request)
      std::lock_guard<std::mutex> lk(mu_);
      auto it = ranges_.upper_bound(reinterpret_cast<uint64_t>(p));
      if (it != ranges_.begin()) {
        --it;
        if (reinterpret_cast<uint64_t>(p) < it->first + it->second.first) {
          memcpy(out->handle, &it->second.second, 64);
          out->pid = pid_;
          out->dev = dev_;
          out->base = it->first;
          out->size = it->second.first;
          return true;
        }
      }
    }
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    CUdeviceptr base = 0;
    size_t size = 0;
    // resolved through the runtime so that the library has no link-time dependency on
    // libcuda.so (absent on GPU-less build / CI hosts)
    typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static GetRangeFn get_range = [] {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) !=
              cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        fn = nullptr;
      }
      return reinterpret_cast<GetRangeFn>(fn);
    }();
    if (!get_range || get_range(&base, &size, reinterpret_cast<CUdeviceptr>(p)) != CUDA_SUCCESS) {
      return false;
    }
    std::lock_guard<std::mutex> lk(mu_);
    auto it = exported_.find(base);
    if (it == exported_.end()) {
      cudaIpcMemHandle_t h;
      cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
      if (e != cudaSuccess) {
        LOG(WARNING) << "cudaIpcGetMemHandle failed (" << cudaGetErrorString(e)
                     << "): memory from a VMM / expandable-segments allocator cannot be exported";
        cudaGetLastError();
        return false;
      }
      it = exported_.emplace(base, h).first;
      ranges_[static_cast<uint64_t>(base)] = std::make_pair(size, h);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle must fit RegionDesc::handle");
    memcpy(out->handle, &it->second, 64);
    out->pid = pid_;
    out->dev = dev_;
    out->base = static_cast<uint64_t>(base);
    out->size = size;
    return true;
  }

  void Unexport(uint64_t base) override {
    std::lock_guard<std::mutex> lk(mu_);
    ranges_.erase(base);
    exported_.erase(static_cast<CUdeviceptr>(base));
  }

  void* Import(const RegionDesc& d) override {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    if (d.pid == pid_) {
      // same address space: no IPC; only make sure the two devices can see each other
      if (d.dev >= 0 && d.dev != dev_) {
        cudaError_t e = cudaDeviceEnablePeerAccess(d.dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          LOG(FATAL) << "cudaDeviceEnablePeerAccess(" << d.dev << "): " << cudaGetErrorString(e);
        }
        cudaGetLastError();
      }
      return reinterpret_cast<void*>(d.base);
    }
    std::string key(d.handle, 64);
    std::lock_guard<std::mutex> lk(mu_);
    auto it = imported_.find(key);
    if (it != imported_.end()) return it->second;
    cudaIpcMemHandle_t h;
    memcpy(&h, d.handle, 64);
    void* base = nullptr;
    PS_CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    imported_[key] = base;
    return base;
  }

  Ticket CopyAsync(void* dst, const void* src, size_t n, int codec, float scale,
                   void* wait_event, int src_device_type = UNK) override {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    if (engine_) EngineQuiesce();
    if (wait_event) {
      PS_CUDA_CHECK(cudaStreamWaitEvent(stream_, static_cast<cudaEvent_t>(wait_event), 0));
    }
    if (n) {
      bool src_on_device = src_device_type == GPU;
      if (src_device_type == UNK) {  // untagged source: ask the driver (about 1 us)
        cudaPointerAttributes attr;
        src_on_device = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
        if (!src_on_device) cudaGetLastError();
      }
      if (!src_on_device) {
        CHECK_EQ(codec, (int)kCodecRaw) << "host-resident values can only be sent raw";
        PS_CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, stream_));
      } else {
        int rc = ps_launch_copy(dst, src, n, codec, scale, max_ctas_,
                                reinterpret_cast<ps_stream_t>(stream_));
        CHECK_EQ(rc, 0) << "copy kernel launch failed: "
                        << cudaGetErrorString(static_cast<cudaError_t>(rc));
      }
    }
    cudaEvent_t ev = AcquireEvent();
    PS_CUDA_CHECK(cudaEventRecord(ev, stream_));
    Ticket t;
    t.event = ev;
    return t;
  }

  // Peers on another host cannot map HBM: their payloads are staged through host memory and
  // travel in socket frames (the reference without GPUDirect does the same, src/ucx_van.h:1097-1115)
  bool NeedsStaging(int device_type, const void* /*ptr*/) override { return device_type == GPU; }
  void CopyToHost(void* host, const void* src, size_t n, void* wait_event) override {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    if (engine_) EngineQuiesce();
    if (wait_event) PS_CUDA_CHECK(cudaStreamWaitEvent(stream_, static_cast<cudaEvent_t>(wait_event), 0));
    PS_CUDA_CHECK(cudaMemcpyAsync(host, src, n, cudaMemcpyDeviceToHost, stream_));
    PS_CUDA_CHECK(cudaStreamSynchronize(stream_));
  }
  void CopyFromHost(void* dst, const void* host, size_t n) override {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    if (engine_) EngineQuiesce();
    PS_CUDA_CHECK(cudaMemcpyAsync(dst, host, n, cudaMemcpyHostToDevice, stream_));
    PS_CUDA_CHECK(cudaStreamSynchronize(stream_));
  }

  /*! \brief raw device-to-device items share launches (ps_launch_copy_multi), one event in all */
  Ticket CopyBatchAsync(const std::vector<CopyItem>& items) override {
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    if (engine_) EngineQuiesce();
    for (const CopyItem& it : items) {
      if (it.wait_event) {
        PS_CUDA_CHECK(cudaStreamWaitEvent(stream_, static_cast<cudaEvent_t>(it.wait_event), 0));
      }
    }
    std::vector<ps_copy_seg> raw;
    raw.reserve(items.size());
    for (const CopyItem& it : items) {
      if (it.n_src_bytes == 0) continue;
      if (it.codec == kCodecRaw && it.src_device_type == GPU) {
        raw.push_back(ps_copy_seg{it.dst, it.src, it.n_src_bytes});
        continue;
      }
      // transforms and host-resident sources keep their own launch (still no event of their own)
      Ticket t = CopyAsync(it.dst, it.src, it.n_src_bytes, it.codec, it.scale, nullptr, it.src_device_type);
      std::lock_guard<std::mutex> lk(mu_);
      free_events_.push_back(static_cast<cudaEvent_t>(t.event));  // recorded, never waited on
    }
    if (!raw.empty()) {
      const int rc = ps_launch_copy_multi(raw.data(), static_cast<int>(raw.size()), max_ctas_,
                                          reinterpret_cast<ps_stream_t>(stream_));
      CHECK_EQ(rc, 0) << "multi-copy launch failed: " << cudaGetErrorString(static_cast<cudaError_t>(rc));
    }
    cudaEvent_t ev = AcquireEvent();
    PS_CUDA_CHECK(cudaEventRecord(ev, stream_));
    Ticket t;
    t.event = ev;
    return t;
  }

  /*!
   * The completion word lives in host shared memory (the control block of a descriptor ring the
   * receiving van polls). Page-lock and map that block so that a copy kernel can store to it:
   * the GPU, not a host thread, announces that a payload has landed.
   */
  void* MapSignalWord(void* page, size_t bytes, void* host_word) override {
    if (cudaSetDevice(dev_) != cudaSuccess) return nullptr;
    cudaError_t e = cudaHostRegister(page, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e != cudaSuccess && e != cudaErrorHostMemoryAlreadyRegistered) {
      LOG(WARNING) << "cudaHostRegister of a descriptor ring failed (" << cudaGetErrorString(e)
                   << "): completions of this connection go through events and a host thread";
      cudaGetLastError();
      return nullptr;
    }
    cudaGetLastError();
    void* dptr = nullptr;
    if (cudaHostGetDevicePointer(&dptr, host_word, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaHostUnregister(page);
      return nullptr;
    }
    std::lock_guard<std::mutex> lk(mu_);
    if (!sig_counter_) {
      PS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&sig_counter_), 256));
      PS_CUDA_CHECK(cudaMemset(sig_counter_, 0, 256));
    }
    return dptr;
  }
  void UnmapSignalWord(void* page) override {
    cudaSetDevice(dev_);
    cudaStreamSynchronize(stream_);  // no kernel may still want to store to the word
    cudaHostUnregister(page);
    cudaGetLastError();
  }

  bool CopySignal(const CopyItem& item, void* word, uint64_t value) override {
    if (!sig_counter_ || !word) return false;
    PS_CUDA_CHECK(cudaSetDevice(dev_));
    if (engine_) {
      // The two paths must not overtake each other (completions on one word are ordered): the
      // engine takes a copy only while nothing is pending on the stream, and the stream gets work
      // only once the engine has retired everything it was given.
      const bool eligible = item.wait_event == nullptr && item.codec == kCodecRaw &&
                            (item.n_src_bytes == 0 || item.src_device_type == GPU);
      if (eligible && StreamIdle()) {
        unsigned long long ticket = 0;
        CHECK_EQ(ps_engine_post(engine_, item.dst, item.src, item.n_src_bytes,
                                static_cast<unsigned long long*>(word), value, &ticket), 0)
            << "copy engine: post failed";
        engine_ticket_.store(ticket, std::memory_order_release);
        return true;
      }
      EngineQuiesce();
    }
    if (item.wait_event) {
      PS_CUDA_CHECK(cudaStreamWaitEvent(stream_, static_cast<cudaEvent_t>(item.wait_event), 0));
    }
    ps_signal sig;
    sig.counter = sig_counter_;
    sig.flag = static_cast<unsigned long long*>(word);
    sig.value = value;
    const ps_stream_t st = reinterpret_cast<ps_stream_t>(stream_);
    int rc = 0;
    if (item.n_src_bytes == 0) {
      rc = ps_launch_signal(&sig, st);
    } else {
      bool src_on_device = item.src_device_type == GPU;
      if (item.src_device_type == UNK) {
        cudaPointerAttributes attr;
        src_on_device = cudaPointerGetAttributes(&attr, item.src) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
        if (!src_on_device) cudaGetLastError();
      }
      if (!src_on_device) {
        CHECK_EQ(item.codec, (int)kCodecRaw) << "host-resident values can only be sent raw";
        PS_CUDA_CHECK(cudaMemcpyAsync(item.dst, item.src, item.n_src_bytes, cudaMemcpyDefault, stream_));
        rc = ps_launch_signal(&sig, st);
      } else {
        rc = ps_launch_copy_signal(item.dst, item.src, item.n_src_bytes, item.codec, item.scale,
                                   max_ctas_, &sig, st);
      }
    }
    CHECK_EQ(rc, 0) << "copy kernel launch failed: " << cudaGetErrorString(static_cast<cudaError_t>(rc));
    return true;
  }

  bool Ready(Ticket t) override {
    if (!t.event) return true;
    cudaError_t e = cudaEventQuery(static_cast<cudaEvent_t>(t.event));
    if (e == cudaErrorNotReady) {
      cudaGetLastError();
      return false;
    }
    return e == cudaSuccess;
  }

  void Wait(Ticket t) override {
    if (!t.event) return;
    cudaEvent_t ev = static_cast<cudaEvent_t>(t.event);
    PS_CUDA_CHECK(cudaEventSynchronize(ev));
    std::lock_guard<std::mutex> lk(mu_);
    free_events_.push_back(ev);
  }

  /*!
   * Symmetric memory, natively: every member cuMemCreate()s a block with a shareable POSIX
   * handle, publishes the descriptor through FdExchange, imports and maps every other member's
   * block; on an NVSwitch box member 0 creates a multicast object, everybody adds its device and
   * binds its block, and maps the multicast address (NVLS: multimem.st / multimem.ld_reduce).
   * Counterpart of the reference's registered-memory exchange (src/rdma_utils.h:75-140,
   * src/rdma_transport.h:469-633), for NVLink + NVSwitch.
   */
  bool SymmetricAlloc(const SymmetricGroup& g, const std::string& tag, size_t bytes, SymmetricBuffer* out) override {
    const VmmApi& api = VmmApi::Get();
    FdExchange* fx = FdExchange::Get(g.job_port);
    if (!api.vmm || !fx || g.index < 0 || g.pids.empty()) return false;
    if (cudaSetDevice(dev_) != cudaSuccess) return false;
    cudaFree(nullptr);  // make sure the primary context exists and is current
    const int n = static_cast<int>(g.pids.size());
    CUdevice cudev;
    if (api.DeviceGet(&cudev, dev_) != CUDA_SUCCESS) return false;
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev_;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (api.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) {
      return false;
    }
    int mc_cap = 0;
    if (api.multicast && n > 1) api.DeviceGetAttribute(&mc_cap, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev);
    if (GetEnv("PS_DISABLE_MULTICAST", 0) != 0) mc_cap = 0;
    CUmulticastObjectProp mcprop;
    memset(&mcprop, 0, sizeof(mcprop));
    mcprop.numDevices = static_cast<unsigned>(n);
    mcprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t size = AlignUp(bytes ? bytes : 1, gran);
    if (mc_cap) {
      size_t mcgran = 0;
      mcprop.size = size;
      if (api.MulticastGetGranularity(&mcgran, &mcprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mcgran) {
        size = (size + mcgran - 1) / mcgran * mcgran;
      } else {
        mc_cap = 0;
      }
    }
    mcprop.size = size;
    auto map_rw = [&](CUmemGenericAllocationHandle h, CUdeviceptr* va) -> bool {
      if (api.MemAddressReserve(va, size, gran, 0, 0) != CUDA_SUCCESS) return false;
      if (api.MemMap(*va, size, 0, h, 0) != CUDA_SUCCESS) return false;
      CUmemAccessDesc acc;
      memset(&acc, 0, sizeof(acc));
      acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      acc.location.id = dev_;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      return api.MemSetAccess(*va, size, &acc, 1) == CUDA_SUCCESS;
    };
    CUmemGenericAllocationHandle mine;
    if (api.MemCreate(&mine, size, &prop, 0) != CUDA_SUCCESS) return false;
    CUdeviceptr va = 0;
    if (!map_rw(mine, &va)) return false;
    if (cudaMemset(reinterpret_cast<void*>(va), 0, size) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    int fd = -1;
    if (api.MemExportToShareableHandle(&fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) return false;
    out->local = reinterpret_cast<void*>(va);
    out->bytes = size;
    out->index = g.index;
    out->count = n;
    out->peers.assign(static_cast<size_t>(n), nullptr);
    out->peers[static_cast<size_t>(g.index)] = out->local;
    out->mc = nullptr;
    auto endpoint = [&](int i) { return FdExchange::EndpointName(g.job_port, g.pids[static_cast<size_t>(i)]); };
    // value of the "/mem" token: block size, and in bit 63 "my device can do multicast"
    fx->Publish(tag + "/mem", fd, static_cast<uint64_t>(size) | (mc_cap ? (1ull << 63) : 0ull));
    bool ok = true, all_mc = mc_cap != 0;
    for (int i = 0; i < n && ok; ++i) {
      if (i == g.index) continue;
      int pfd = -1;
      uint64_t v = 0;
      ok = FdExchange::Fetch(endpoint(i), tag + "/mem", &pfd, &v) && pfd >= 0 && (v & ~(1ull << 63)) == size;
      if (!ok) {
        if (pfd >= 0) close(pfd);
        break;
      }
      all_mc = all_mc && (v >> 63) != 0;
      CUmemGenericAllocationHandle ph;
      ok = api.MemImportFromShareableHandle(&ph, reinterpret_cast<void*>(static_cast<uintptr_t>(pfd)),
                                            CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
      close(pfd);
      CUdeviceptr pva = 0;
      ok = ok && map_rw(ph, &pva);
      if (ok) out->peers[static_cast<size_t>(i)] = reinterpret_cast<void*>(pva);
    }
    // a token every member publishes and fetches from every other member = a barrier
    auto barrier = [&](const std::string& what, bool good) -> bool {
      fx->Publish(tag + what, -1, good ? 1 : 0);
      bool all = good;
      for (int i = 0; i < n; ++i) {
        if (i == g.index) continue;
        uint64_t v = 0;
        all = FdExchange::Fetch(endpoint(i), tag + what, nullptr, &v) && v == 1 && all;
      }
      return all;
    };
    ok = barrier("/mapped", ok);
    if (ok && all_mc) {
      // NVLS: one multicast object over all members' devices
      CUmemGenericAllocationHandle mch = 0;
      bool mc_ok = true;
      int mcfd = -1;
      if (g.index == 0) {
        mc_ok = api.MulticastCreate(&mch, &mcprop) == CUDA_SUCCESS &&
                api.MemExportToShareableHandle(&mcfd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
        fx->Publish(tag + "/mc", mc_ok ? mcfd : -1, mc_ok ? 1 : 0);
      } else {
        uint64_t v = 0;
        mc_ok = FdExchange::Fetch(endpoint(0), tag + "/mc", &mcfd, &v) && v == 1 && mcfd >= 0 &&
                api.MemImportFromShareableHandle(&mch, reinterpret_cast<void*>(static_cast<uintptr_t>(mcfd)),
                                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
      }
      mc_ok = mc_ok && api.MulticastAddDevice(mch, cudev) == CUDA_SUCCESS;
      mc_ok = barrier("/mc_added", mc_ok);  // every device is in the team before anybody binds
      mc_ok = mc_ok && api.MulticastBindMem(mch, 0, mine, 0, size, 0) == CUDA_SUCCESS;
      mc_ok = barrier("/mc_bound", mc_ok);
      CUdeviceptr mcva = 0;
      if (mc_ok && map_rw(mch, &mcva)) out->mc = reinterpret_cast<void*>(mcva);
      const bool everybody = barrier("/mc_mapped", out->mc != nullptr);
      if (!everybody) out->mc = nullptr;  // all or nobody: the members must agree on the path
      if (mcfd >= 0) close(mcfd);
      if (!out->mc) {
        LOG(WARNING) << "symmetric buffer '" << tag << "': NVSwitch multicast could not be set up; "
                     << "peers are reachable one by one";
      }
    }
    ok = barrier("/done", ok);
    fx->Retract(tag + "/mem");
    close(fd);
    return ok;
  }

  /*! \brief everything this domain was given (stream and engine) has completed */
  void Quiesce() {
    cudaSetDevice(dev_);
    if (engine_) ps_engine_wait(engine_, engine_ticket_.load(std::memory_order_acquire));
    cudaStreamSynchronize(stream_);
  }
  /*! \brief is `p` inside one of this device's landing-slot arenas? */
  bool Owns(const void* p) {
    std::lock_guard<std::mutex> lk(mu_);
    const char* c = static_cast<const char*>(p);
    for (auto& a : arenas_) {
      if (c >= a->base && c < a->base + a->size) return true;
    }
    return false;
  }
  /*! \brief let kernels of THIS device store into memory that lives on local ordinal `owner` */
  void EnablePeerTo(int owner) {
    if (owner < 0 || owner == dev_) return;
    if (cudaSetDevice(dev_) != cudaSuccess) return;
    const cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      LOG(WARNING) << "cudaDeviceEnablePeerAccess(" << owner << ") from device " << dev_ << ": " << cudaGetErrorString(e);
    }
    cudaGetLastError();
  }

  void EngineStats(uint64_t* launches, uint64_t* items) override {
    unsigned long long l = 0, i = 0;
    if (engine_) ps_engine_stats(engine_, &l, &i);
    *launches = l;
    *items = i;
  }

 private:
  /*! \brief before the stream gets work: everything THIS domain gave the engine must have completed */
  void EngineQuiesce() {
    ps_engine_wait(engine_, engine_ticket_.load(std::memory_order_acquire));
    launch_pending_.store(true, std::memory_order_release);
  }
  /*! \brief has everything ever enqueued on the data stream completed? */
  bool StreamIdle() {
    if (!launch_pending_.load(std::memory_order_acquire) && !stream_shared_.load(std::memory_order_acquire)) {
      return true;
    }
    const cudaError_t e = cudaStreamQuery(stream_);
    if (e == cudaSuccess) {
      launch_pending_.store(false, std::memory_order_release);
      return true;
    }
    cudaGetLastError();
    return false;
  }
  cudaEvent_t AcquireEvent() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!free_events_.empty()) {
        cudaEvent_t e = free_events_.back();
        free_events_.pop_back();
        return e;
      }
    }
    cudaEvent_t e;
    PS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return e;
  }

  struct Retired {
    char* ptr = nullptr;
    cudaEvent_t event = nullptr;
    unsigned long long engine_ticket = 0;
  };
  std::vector<Retired> retired_;  // freed slots whose last possible reader has not finished yet
  struct DevArena {
    char* base = nullptr;
    uint64_t size = 0;
    ArenaAllocator alloc;
  };
  std::vector<std::unique_ptr<DevArena>> arenas_;
  int dev_;
  const int32_t pid_ = static_cast<int32_t>(getpid());  // (a system call each time otherwise)
  int max_ctas_ = 0;
  ps_engine* engine_ = nullptr;
  std::atomic<unsigned long long> engine_ticket_{0};  // this domain's newest post to the (shared) engine
  std::atomic<bool> launch_pending_{false};  // the launch path has been used since the stream was last seen idle
  std::atomic<bool> stream_shared_{false};   // Stream() was handed out: foreign work may be on it
  unsigned* sig_counter_ = nullptr;  // CTA arrival counter of the signalling kernels (self-resetting)
  cudaStream_t stream_ = nullptr;
  std::mutex mu_;
  std::vector<cudaEvent_t> free_events_;
  std::unordered_map<CUdeviceptr, cudaIpcMemHandle_t> exported_;
  std::map<uint64_t, std::pair<size_t, cudaIpcMemHandle_t>> ranges_;  // base -> (size, handle)
  std::unordered_map<std::string, void*> imported_;
};

/*!
 * \brief several GPUs driven by ONE process (DMLC_NUM_GPU_DEV / PS_NUM_GPU_DEV > 1): one CudaDomain
 *        per device behind the MemDomain interface. A copy runs on the device its SOURCE lives on
 *        (that device's stream, engine and completion counter), a landing slot is cut on the device
 *        the sender names, and a peer's region is opened once and made reachable from every local
 *        device. The reference does this with one UCX context per device and routes by
 *        src / dst device id (src/ucx_van.h:662-682, 938-1006; tests/test_benchmark.cc:58-90 puts
 *        key k on device k % local_size).
 */
class MultiCudaDomain : public MemDomain {
 public:
  MultiCudaDomain(int first, int count) {
    for (int i = 0; i < count; ++i) subs_.emplace_back(new CudaDomain(first + i));
  }
  const char* name() const override { return "nvl"; }
  int device() const override { return subs_[0]->device(); }
  int num_devices() const override { return static_cast<int>(subs_.size()); }
  void* Stream() override { return subs_[0]->Stream(); }
  bool Handles(int device_type, const void* /*ptr*/) override { return device_type == GPU; }
  void* Alloc(size_t bytes) override { return subs_[0]->Alloc(bytes); }
  void* AllocOn(size_t bytes, int device) override { return Sub(device)->Alloc(bytes); }
  void Free(void* p) override { OwnerOf(p, -1)->Free(p); }
  bool Export(const void* p, RegionDesc* out) override { return OwnerOf(p, -1)->Export(p, out); }
  void Unexport(uint64_t base) override {
    for (auto& d : subs_) d->Unexport(base);
  }
  void* Import(const RegionDesc& d) override {
    void* base = subs_[0]->Import(d);
    // which local ordinal owns the mapped memory? every other device of this process needs peer access to it
    int owner = d.dev;
    if (d.pid != static_cast<int32_t>(getpid())) {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, base) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
        owner = attr.device;
      } else {
        cudaGetLastError();
      }
    }
    for (size_t i = 1; i < subs_.size(); ++i) subs_[i]->EnablePeerTo(owner);
    return base;
  }
  Ticket CopyAsync(void* dst, const void* src, size_t n, int codec, float scale, void* wait_event,
                   int src_device_type = UNK) override {
    return OwnerOf(src, -1)->CopyAsync(dst, src, n, codec, scale, wait_event, src_device_type);
  }
  bool NeedsStaging(int device_type, const void* /*ptr*/) override { return device_type == GPU; }
  void CopyToHost(void* host, const void* src, size_t n, void* wait_event) override {
    OwnerOf(src, -1)->CopyToHost(host, src, n, wait_event);
  }
  void CopyFromHost(void* dst, const void* host, size_t n) override { OwnerOf(dst, -1)->CopyFromHost(dst, host, n); }
  void* MapSignalWord(void* page, size_t bytes, void* host_word) override {
    void* word = nullptr;
    for (auto& d : subs_) {
      void* w = d->MapSignalWord(page, bytes, host_word);
      if (!w) return nullptr;
      if (word && w != word) {
        LOG(WARNING) << "the devices of this process see a mapped host word at different addresses: "
                     << "completions of this connection go through events and a host thread";
        return nullptr;
      }
      word = w;
    }
    return word;
  }
  void UnmapSignalWord(void* page) override {
    for (auto& d : subs_) d->UnmapSignalWord(page);
  }
  bool CopySignal(const CopyItem& item, void* word, uint64_t value) override {
    // completions on one word must keep their order: everything for one connection runs on the
    // device of its first copy unless the source says otherwise AND nothing else is in flight there.
    // (Per-connection traffic of the benchmark and the trainer comes from one device at a time.)
    CudaDomain* d = item.n_src_bytes ? OwnerOf(item.src, item.src_device_id) : LastFor(word);
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = last_for_word_.find(word);
      // (after Quiesce the previous device has signalled everything it was given)
      if (it != last_for_word_.end() && it->second != d) it->second->Quiesce();
      last_for_word_[word] = d;
    }
    return d->CopySignal(item, word, value);
  }
  bool Ready(Ticket t) override { return subs_[0]->Ready(t); }
  void Wait(Ticket t) override { subs_[0]->Wait(t); }
  bool SymmetricAlloc(const SymmetricGroup& g, const std::string& tag, size_t bytes, SymmetricBuffer* out) override {
    return subs_[0]->SymmetricAlloc(g, tag, bytes, out);
  }
  void EngineStats(uint64_t* launches, uint64_t* items) override {
    *launches = *items = 0;
    for (auto& d : subs_) {
      uint64_t l = 0, i = 0;
      d->EngineStats(&l, &i);
      *launches += l;
      *items += i;
    }
  }

 private:
  CudaDomain* Sub(int device) {
    for (auto& d : subs_) {
      if (d->device() == device) return d.get();
    }
    return subs_[0].get();
  }
  CudaDomain* OwnerOf(const void* p, int hint) {
    if (hint >= 0) return Sub(hint);
    for (auto& d : subs_) {
      if (d->Owns(p)) return d.get();
    }
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess &&
        (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
      return Sub(attr.device);
    }
    cudaGetLastError();
    return subs_[0].get();
  }
  CudaDomain* LastFor(void* word) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = last_for_word_.find(word);
    return it != last_for_word_.end() ? it->second : subs_[0].get();
  }
  std::vector<std::unique_ptr<CudaDomain>> subs_;
  std::mutex mu_;
  std::unordered_map<void*, CudaDomain*> last_for_word_;
};

}  // namespace

MemDomain* CreateCudaDomain(int instance_idx) {
  const int n = CudaDeviceCount();
  if (n <= 0) {
    LOG(ERROR) << "the nvl van needs a CUDA device and none is visible";
    return nullptr;
  }
  int dev = -1;
  if (const char* v = Environment::Get()->find("PS_CUDA_DEVICE")) {
    dev = atoi(v);
  } else if (const char* lr = Environment::Get()->find("LOCAL_RANK")) {
    dev = atoi(lr) % n;
  } else {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  }
  // DMLC_NUM_GPU_DEV (the reference's name, src/ucx_van.h:941-942) / PS_NUM_GPU_DEV: this process
  // drives that many consecutive devices starting at `dev`
  const int count = GetEnv("PS_NUM_GPU_DEV", GetEnv("DMLC_NUM_GPU_DEV", 1));
  // DMLC_GROUP_SIZE instances in one process (the reference gives each its own rail / NIC port,
  // src/postoffice.cc:38-47): PS_INSTANCE_GPU_STRIDE=s moves instance i to device dev + i * s * count,
  // so every instance gets its own GPU(s), NVLink ports and copy engine. 0 (default): all instances
  // of the process share its device.
  dev += instance_idx * GetEnv("PS_INSTANCE_GPU_STRIDE", 0) * std::max(1, count);
  CHECK(dev >= 0 && dev < n) << "CUDA device " << dev << " out of range (" << n << " visible)";
  if (count > 1) {
    CHECK_LE(dev + count, n) << "devices " << dev << ".." << dev + count - 1 << " requested, " << n << " visible";
    return new MultiCudaDomain(dev, count);
  }
  return new CudaDomain(dev);
}

}  // namespace ps
