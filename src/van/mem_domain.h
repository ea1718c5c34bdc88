/**
 * \file mem_domain.h
 * \brief Peer-mappable memory: the abstraction the one-sided van is written against.
 *
 * A MemDomain is "a kind of memory that another endpoint can map and write":
 *   - CudaDomain (cuda_domain.h): B200 HBM, exported with CUDA IPC handles and
 *     written by sm_100a copy kernels over NVLink peer mappings;
 *   - ShmDomain (here): POSIX shared memory written by memcpy — the GPU-less
 *     twin used by CPU-only CI and by same-host CPU workers (the role the
 *     reference's IPCTransport plays, src/rdma_transport.h:469-633).
 * It replaces the reference's ibverbs machinery: RegionDesc stands in for
 * (ibv_mr, rkey), Import for the rendezvous address exchange, CopyAsync for
 * RDMA WRITE, and the Ticket for the completion-queue entry
 * (src/rdma_utils.h:75-140, src/rdma_transport.h:211-231).
 *
 * Also here: ArenaAllocator (offset first-fit with coalescing; the reference's
 * MemoryAllocator, src/rdma_utils.h:75-140) and IndexPool (dense index <->
 * pointer table; the reference's AddressPool, src/van_common.h:72-122).
 */
#ifndef PS_VAN_MEM_DOMAIN_H_
#define PS_VAN_MEM_DOMAIN_H_
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ps/internal/message.h"
#include "ps/internal/symmetric.h"
#include "ps/internal/utils.h"
#include "kernels/host_kernels.h"
#include "ps/sarray.h"
#include "van/shm_util.h"

namespace ps {

/*! \brief round `v` up / down to a multiple of `a` (a power of two) */
inline uint64_t AlignUp(uint64_t v, uint64_t a) { return (v + a - 1) & ~(a - 1); }
inline uint64_t AlignDown(uint64_t v, uint64_t a) { return v & ~(a - 1); }

/*! \brief everything a peer needs to map one exported allocation */
struct RegionDesc {
  int32_t region = -1;     // id, unique within the exporting node
  int32_t owner = -1;      // exporting node id
  int32_t pid = 0;         // exporting process
  int32_t dev = -1;        // CUDA ordinal, -1 for host memory
  uint64_t base = 0;       // virtual address in the exporter
  uint64_t size = 0;
  char handle[64] = {0};   // cudaIpcMemHandle_t, or a shm object name
  char host[64] = {0};     // exporter hostname (same-process detection)

  std::string Serialize() const { return std::string(reinterpret_cast<const char*>(this), sizeof(*this)); }
  static bool Parse(const std::string& s, size_t at, RegionDesc* out) {
    if (s.size() < at + sizeof(RegionDesc)) return false;
    memcpy(out, s.data() + at, sizeof(RegionDesc));
    return true;
  }
};

/*! \brief completion token of an asynchronous copy */
struct Ticket {
  void* event = nullptr;  // domain-specific; nullptr = already complete
};

class MemDomain {
 public:
  virtual ~MemDomain() {}
  /*!
   * \brief collective over `group`: allocate `bytes` of zero-filled symmetric memory under the
   *        job-wide name `tag`. Handles travel as file descriptors through FdExchange (unix
   *        sockets); every step that needs all members doubles as a barrier. False if this
   *        domain cannot share memory that way.
   */
  virtual bool SymmetricAlloc(const SymmetricGroup& /*group*/, const std::string& /*tag*/, size_t /*bytes*/,
                              SymmetricBuffer* /*out*/) {
    return false;
  }
  virtual const char* name() const = 0;
  /*! \brief true if values tagged (type, ptr) should travel one-sided through this domain */
  virtual bool Handles(int device_type, const void* ptr) = 0;
  /*!
   * \brief peers on ANOTHER host cannot map this domain's memory: their payloads travel in socket
   *        frames. Must bytes at `ptr` be staged through host memory for that (device memory)?
   */
  virtual bool NeedsStaging(int /*device_type*/, const void* /*ptr*/) { return false; }
  /*! \brief blocking copy of this domain's memory to plain host memory, after `wait_event` (may be null) */
  virtual void CopyToHost(void* host, const void* src, size_t n, void* /*wait_event*/) { memcpy(host, src, n); }
  /*! \brief blocking copy of plain host memory into this domain's memory */
  virtual void CopyFromHost(void* dst, const void* host, size_t n) { memcpy(dst, host, n); }
  /*! \brief CUDA ordinal this domain is bound to, -1 for host domains */
  virtual int device() const { return -1; }
  /*! \brief allocate exportable memory (landing slots) */
  virtual void* Alloc(size_t bytes) = 0;
  /*! \brief ... on a given device of a domain that drives several (else: Alloc) */
  virtual void* AllocOn(size_t bytes, int /*device*/) { return Alloc(bytes); }
  /*! \brief how many devices this domain drives (1 for host domains) */
  virtual int num_devices() const { return 1; }
  virtual void Free(void* p) = 0;
  /*! \brief describe the exportable allocation that contains `p` (region id left unset) */
  virtual bool Export(const void* p, RegionDesc* out) = 0;
  /*! \brief drop cached export state of the allocation that starts at `base` (it is about to be freed) */
  virtual void Unexport(uint64_t /*base*/) {}
  /*! \brief map a peer's region; returns the local address of its base */
  virtual void* Import(const RegionDesc& d) = 0;
  /*!
   * \brief enqueue dst <- codec(src, n_src_bytes). `dst` may be a peer mapping.
   *        `wait_event` (domain specific, may be null) gates the copy on the
   *        producer of `src`.
   */
  virtual Ticket CopyAsync(void* dst, const void* src, size_t n_src_bytes, int codec, float scale,
                           void* wait_event, int src_device_type = UNK) = 0;
  /*! \brief one element of a coalesced batch (see CopyBatchAsync) */
  struct CopyItem {
    void* dst = nullptr;
    const void* src = nullptr;
    size_t n_src_bytes = 0;
    int codec = 0;
    float scale = 1.f;
    void* wait_event = nullptr;
    int src_device_type = UNK;
    int src_device_id = -1;  // device the source lives on, when the sender knows (SArray placement)
  };
  /*!
   * \brief enqueue all items and return ONE ticket that completes after the last of them
   *        (and after everything enqueued on the domain's stream before this call). Domains
   *        with a launch cost override this to merge the items into few launches.
   */
  virtual Ticket CopyBatchAsync(const std::vector<CopyItem>& items) {
    Ticket last;
    for (const CopyItem& it : items) {
      if (last.event) Wait(last);  // stream order makes the newest ticket cover the older ones
      last = CopyAsync(it.dst, it.src, it.n_src_bytes, it.codec, it.scale, it.wait_event, it.src_device_type);
    }
    if (items.empty()) last = CopyAsync(nullptr, nullptr, 0, 0, 1.f, nullptr);
    return last;
  }
  /*!
   * \brief make the 8-byte completion word at `host_word` (inside the page-aligned host range
   *        [page, page+bytes), e.g. the control block of a shared-memory ring) storable by this
   *        domain's copy engine; returns the address that engine must use, or null if the
   *        domain cannot signal completion by itself (the van then falls back to tickets).
   */
  virtual void* MapSignalWord(void* /*page*/, size_t /*bytes*/, void* /*host_word*/) { return nullptr; }
  virtual void UnmapSignalWord(void* /*page*/) {}
  /*!
   * \brief CopyAsync without a ticket: when the bytes of `item` are globally visible (and all
   *        work enqueued before it has completed), the copy engine ITSELF stores `value` to
   *        `word` (an address from MapSignalWord) with release semantics at system scope.
   *        `item.n_src_bytes == 0` only signals (after `item.wait_event`). Values passed for
   *        one word must be increasing. False: not supported for this item (use CopyAsync).
   */
  virtual bool CopySignal(const CopyItem& /*item*/, void* /*word*/, uint64_t /*value*/) { return false; }
  /*! \brief block until the copy behind `t` is globally visible; recycles the ticket */
  virtual void Wait(Ticket t) = 0;
  /*! \brief non-blocking: has the copy behind `t` completed? (does not recycle the ticket) */
  virtual bool Ready(Ticket t) { return t.event == nullptr; }
  /*! \brief copy-engine accounting (kernel launches it needed, descriptors it executed) */
  virtual void EngineStats(uint64_t* launches, uint64_t* items) {
    *launches = 0;
    *items = 0;
  }
  /*! \brief stream-like handle applications may enqueue their own work on (may be null) */
  virtual void* Stream() { return nullptr; }
  /*! \brief the van stopped: drop global names (mappings stay valid until the process ends) */
  virtual void ReleaseNames() {}
};

/*! \brief first-fit offset allocator with coalescing; thread-safe */
class ArenaAllocator {
 public:
  ArenaAllocator() {}
  ArenaAllocator(uint64_t size, uint64_t align) { Reset(size, align); }
  void Reset(uint64_t size, uint64_t align) {
    std::lock_guard<std::mutex> lk(mu_);
    size_ = size;
    align_ = align;
    free_.clear();
    used_.clear();
    free_[0] = size;
  }
  /*! \brief offset of a block of >= bytes, or UINT64_MAX */
  uint64_t Alloc(uint64_t bytes) {
    bytes = AlignUp(bytes ? bytes : 1, align_);
    std::lock_guard<std::mutex> lk(mu_);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second < bytes) continue;
      const uint64_t off = it->first, len = it->second;
      free_.erase(it);
      if (len > bytes) free_[off + bytes] = len - bytes;
      used_[off] = bytes;
      return off;
    }
    return UINT64_MAX;
  }
  bool Free(uint64_t off) {
    std::lock_guard<std::mutex> lk(mu_);
    auto u = used_.find(off);
    if (u == used_.end()) return false;
    uint64_t len = u->second;
    used_.erase(u);
    auto next = free_.lower_bound(off);
    if (next != free_.end() && off + len == next->first) {
      len += next->second;
      next = free_.erase(next);
    }
    if (next != free_.begin()) {
      auto prev = std::prev(next);
      if (prev->first + prev->second == off) {
        prev->second += len;
        return true;
      }
    }
    free_[off] = len;
    return true;
  }
  uint64_t BytesInUse() {
    std::lock_guard<std::mutex> lk(mu_);
    uint64_t n = 0;
    for (auto& kv : used_) n += kv.second;
    return n;
  }
  uint64_t size() const { return size_; }

 private:
  std::mutex mu_;
  uint64_t size_ = 0, align_ = 256;
  std::map<uint64_t, uint64_t> free_;  // offset -> length
  std::map<uint64_t, uint64_t> used_;
};

/*! \brief dense index <-> pointer table (32-bit immediates for descriptors) */
template <typename T>
class IndexPool {
 public:
  explicit IndexPool(size_t capacity = 0) {
    if (capacity == 0) capacity = static_cast<size_t>(GetEnv("BYTEPS_ADDRESS_POOL_SIZE", 10240));
    table_.assign(capacity, nullptr);
  }
  /*! \brief store `p`, return its index (grows when full) */
  uint32_t Store(T* p) {
    std::lock_guard<std::mutex> lk(mu_);
    for (size_t probe = 0; probe < table_.size(); ++probe) {
      size_t i = (cursor_ + probe) % table_.size();
      if (table_[i] == nullptr) {
        table_[i] = p;
        cursor_ = i + 1;
        return static_cast<uint32_t>(i);
      }
    }
    table_.push_back(p);
    cursor_ = table_.size();
    return static_cast<uint32_t>(table_.size() - 1);
  }
  T* Get(uint32_t i) {
    std::lock_guard<std::mutex> lk(mu_);
    return i < table_.size() ? table_[i] : nullptr;
  }
  T* Release(uint32_t i) {
    std::lock_guard<std::mutex> lk(mu_);
    if (i >= table_.size()) return nullptr;
    T* p = table_[i];
    table_[i] = nullptr;
    return p;
  }

 private:
  std::mutex mu_;
  size_t cursor_ = 0;
  std::vector<T*> table_;
};

/*!
 * \brief host shared memory domain. Memory comes from named POSIX shm arenas
 *        (PS_SHM_ARENA_MB each, default 256) so a peer process can map it.
 */
class ShmDomain : public MemDomain {
 public:
  ShmDomain() {
    arena_bytes_ = static_cast<uint64_t>(GetEnv("PS_SHM_ARENA_MB", 256)) << 20;
    static const int swept = SweepStaleShm("pslite_b200_");  // arenas of processes that were killed
    (void)swept;
    // PS_SHM_ASYNC=1: copies run on a background "stream" thread and complete later, like
    // kernels on a CUDA stream: the van's completion / batching logic gets exercised on CPU
    if (GetEnv("PS_SHM_ASYNC", 0) != 0) copier_.reset(new std::thread(&ShmDomain::CopierLoop, this));
  }
  ~ShmDomain() override {
    if (copier_) {
      {
        std::lock_guard<std::mutex> lk(q_mu_);
        q_stop_ = true;
      }
      q_cv_.notify_all();
      copier_->join();
    }
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& a : arenas_) {
      munmap(a->base, a->size);
      if (!a->retired) shm_unlink(a->name.c_str());
    }
    for (auto& m : imported_) munmap(m.second.first, m.second.second);
  }
  const char* name() const override { return "shm"; }
  void ReleaseNames() override {
    // the process-wide Postoffice (and with it this domain) is never destroyed: unlink here
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& a : arenas_) {
      if (!a->retired) shm_unlink(a->name.c_str());
      a->retired = true;  // still mapped (old pointers stay valid) but never handed out again
    }
  }

  bool Handles(int /*device_type*/, const void* ptr) override { return FindArena(ptr) != nullptr; }
  /*! \brief PS_TEST_STAGE_ARENA=1: treat arena memory like device memory (exercises the staging of
   *         the nvl van for peers on other hosts without a GPU) */
  bool NeedsStaging(int /*device_type*/, const void* ptr) override {
    static const bool pretend = GetEnv("PS_TEST_STAGE_ARENA", 0) != 0;
    return pretend && FindArena(ptr) != nullptr;
  }

  void* Alloc(size_t bytes) override {
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& a : arenas_) {
        if (a->retired) continue;
        uint64_t off = a->alloc.Alloc(bytes);
        if (off != UINT64_MAX) return a->base + off;
      }
    }
    Arena* a = NewArena(std::max<uint64_t>(arena_bytes_, AlignUp(bytes, 4096)));
    uint64_t off = a->alloc.Alloc(bytes);
    CHECK_NE(off, UINT64_MAX);
    return a->base + off;
  }
  void Free(void* p) override {
    Arena* a = FindArena(p);
    if (a) a->alloc.Free(static_cast<uint64_t>(static_cast<char*>(p) - a->base));
  }
  bool Export(const void* p, RegionDesc* out) override {
    Arena* a = FindArena(p);
    if (!a || a->retired) return false;
    out->pid = pid_;
    out->dev = -1;
    out->base = reinterpret_cast<uint64_t>(a->base);
    out->size = a->size;
    memset(out->handle, 0, sizeof(out->handle));
    strncpy(out->handle, a->name.c_str(), sizeof(out->handle) - 1);
    return true;
  }
  void* Import(const RegionDesc& d) override {
    if (d.pid == pid_) return reinterpret_cast<void*>(d.base);
    std::lock_guard<std::mutex> lk(mu_);
    std::string key(d.handle);
    auto it = imported_.find(key);
    if (it != imported_.end()) return it->second.first;
    int fd = shm_open(d.handle, O_RDWR, 0600);
    CHECK_GE(fd, 0) << "shm_open(" << d.handle << "): " << strerror(errno);
    void* base = mmap(nullptr, d.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    CHECK(base != MAP_FAILED) << strerror(errno);
    imported_[key] = std::make_pair(base, static_cast<size_t>(d.size));
    return base;
  }
  Ticket CopyAsync(void* dst, const void* src, size_t n, int codec, float scale,
                   void* /*wait_event*/, int /*src_device_type*/ = UNK) override {
    if (copier_) {
      AsyncOp* op = new AsyncOp();
      op->dst = dst;
      op->src = src;
      op->n = (dst != src || codec != kCodecRaw) ? n : 0;
      op->codec = codec;
      op->scale = scale;
      {
        std::lock_guard<std::mutex> lk(q_mu_);
        q_.push_back(op);
        q_len_.fetch_add(1, std::memory_order_release);
      }
      q_cv_.notify_one();
      Ticket t;
      t.event = op;
      return t;
    }
    if (n) CHECK_EQ(ps_host_copy(dst, src, n, codec, scale), 0) << "unknown wire codec " << codec;
    // make the payload visible before the descriptor that announces it
    std::atomic_thread_fence(std::memory_order_release);
    return Ticket();
  }
  /*! \brief CPU twin: memfd-backed blocks mapped by every member; no multicast (mc stays null) */
  bool SymmetricAlloc(const SymmetricGroup& g, const std::string& tag, size_t bytes, SymmetricBuffer* out) override;
  void* MapSignalWord(void* /*page*/, size_t /*bytes*/, void* host_word) override { return host_word; }
  bool CopySignal(const CopyItem& item, void* word, uint64_t value) override {
    auto* flag = static_cast<std::atomic<uint64_t>*>(word);
    const size_t n = (item.dst != item.src || item.codec != kCodecRaw) ? item.n_src_bytes : 0;
    if (copier_) {
      // the stand-in for a stream: the copier thread stores the flag after the bytes, in order
      AsyncOp* op = new AsyncOp();
      op->dst = item.dst;
      op->src = item.src;
      op->n = n;
      op->codec = item.codec;
      op->scale = item.scale;
      op->flag = flag;
      op->flag_value = value;
      {
        std::lock_guard<std::mutex> lk(q_mu_);
        q_.push_back(op);
        q_len_.fetch_add(1, std::memory_order_release);
      }
      q_cv_.notify_one();
      return true;
    }
    if (n) CHECK_EQ(ps_host_copy(item.dst, item.src, n, item.codec, item.scale), 0) << "unknown wire codec " << item.codec;
    flag->store(value, std::memory_order_release);
    return true;
  }
  bool Ready(Ticket t) override {
    return !t.event || static_cast<AsyncOp*>(t.event)->done.load(std::memory_order_acquire);
  }
  void Wait(Ticket t) override {
    if (!t.event) return;
    AsyncOp* op = static_cast<AsyncOp*>(t.event);
    while (!op->done.load(std::memory_order_acquire)) std::this_thread::yield();
    delete op;
  }

 private:
  const int32_t pid_ = static_cast<int32_t>(getpid());  // (a system call each time otherwise: glibc does not cache it)
  struct AsyncOp {
    void* dst = nullptr;
    const void* src = nullptr;
    size_t n = 0;
    int codec = 0;
    float scale = 1.f;
    std::atomic<bool> done{false};
    std::atomic<uint64_t>* flag = nullptr;  // CopySignal: store flag_value here, then self-delete
    uint64_t flag_value = 0;
  };
  void CopierLoop() {
    std::unique_lock<std::mutex> lk(q_mu_);
    for (;;) {
      if (q_.empty() && !q_stop_) {
        // a GPU stream does not go to sleep between kernels: poll for a while so that this
        // stand-in adds copy latency, not the wake-up latency of an idle host thread
        lk.unlock();
        const auto until = std::chrono::steady_clock::now() + std::chrono::milliseconds(2);
        while (q_len_.load(std::memory_order_acquire) == 0 && std::chrono::steady_clock::now() < until) {
          std::this_thread::yield();  // the tests run more stand-in threads than the host has cores
        }
        lk.lock();
      }
      q_cv_.wait(lk, [this] { return q_stop_ || !q_.empty(); });
      if (q_.empty()) return;
      AsyncOp* op = q_.front();
      q_.erase(q_.begin());
      q_len_.fetch_sub(1, std::memory_order_release);
      lk.unlock();
      if (op->n) ps_host_copy(op->dst, op->src, op->n, op->codec, op->scale);
      if (op->flag) {
        op->flag->store(op->flag_value, std::memory_order_release);
        delete op;  // nobody holds a ticket for a signalled copy
      } else {
        op->done.store(true, std::memory_order_release);
      }
      lk.lock();
    }
  }
  struct Arena {
    std::string name;
    bool retired = false;
    char* base = nullptr;
    uint64_t size = 0;
    ArenaAllocator alloc;
  };
  std::unique_ptr<std::thread> copier_;
  std::mutex q_mu_;
  std::condition_variable q_cv_;
  std::vector<AsyncOp*> q_;
  std::atomic<int> q_len_{0};
  bool q_stop_ = false;
  Arena* NewArena(uint64_t bytes) {
    static std::atomic<int> counter{0};
    std::unique_ptr<Arena> a(new Arena());
    a->name = "/" + ShmScopedPrefix("pslite_b200_") + std::to_string(getpid()) + "_" + std::to_string(counter++);
    a->size = AlignUp(bytes, 4096);
    shm_unlink(a->name.c_str());
    int fd = shm_open(a->name.c_str(), O_CREAT | O_RDWR | O_EXCL, 0600);
    CHECK_GE(fd, 0) << "shm_open(" << a->name << "): " << strerror(errno);
    CHECK_EQ(ftruncate(fd, static_cast<off_t>(a->size)), 0) << strerror(errno);
    void* base = mmap(nullptr, a->size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    CHECK(base != MAP_FAILED) << strerror(errno);
    a->base = static_cast<char*>(base);
    a->alloc.Reset(a->size, 256);
    std::lock_guard<std::mutex> lk(mu_);
    arenas_.push_back(std::move(a));
    return arenas_.back().get();
  }
  Arena* FindArena(const void* p) {
    std::lock_guard<std::mutex> lk(mu_);
    const char* c = static_cast<const char*>(p);
    for (auto& a : arenas_) {
      if (c >= a->base && c < a->base + a->size) return a.get();
    }
    return nullptr;
  }
  std::mutex mu_;
  uint64_t arena_bytes_;
  std::vector<std::unique_ptr<Arena>> arenas_;
  std::map<std::string, std::pair<void*, size_t>> imported_;
};

}  // namespace ps

#include "van/fd_exchange.h"

namespace ps {
inline bool ShmDomain::SymmetricAlloc(const SymmetricGroup& g, const std::string& tag, size_t bytes,
                                      SymmetricBuffer* out) {
  FdExchange* fx = FdExchange::Get(g.job_port);
  if (!fx || g.index < 0 || g.pids.empty()) return false;
  const size_t size = AlignUp(bytes ? bytes : 1, 4096);
  const std::string name = "/" + ShmScopedPrefix("pslite_b200_") + std::to_string(getpid()) + "_sym_" + tag;
  shm_unlink(name.c_str());
  const int fd = shm_open(name.c_str(), O_CREAT | O_RDWR | O_EXCL, 0600);
  if (fd < 0) return false;
  shm_unlink(name.c_str());  // the descriptor is all anybody needs
  if (ftruncate(fd, static_cast<off_t>(size)) != 0) {
    close(fd);
    return false;
  }
  void* mine = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (mine == MAP_FAILED) {
    close(fd);
    return false;
  }
  memset(mine, 0, size);
  out->local = mine;
  out->bytes = size;
  out->index = g.index;
  out->count = static_cast<int>(g.pids.size());
  out->peers.assign(g.pids.size(), nullptr);
  out->peers[g.index] = mine;
  out->mc = nullptr;
  fx->Publish(tag + "/mem", fd, size);
  bool ok = true;
  for (size_t i = 0; i < g.pids.size() && ok; ++i) {
    if (static_cast<int>(i) == g.index) continue;
    int pfd = -1;
    uint64_t psize = 0;
    ok = FdExchange::Fetch(FdExchange::EndpointName(g.job_port, g.pids[i]), tag + "/mem", &pfd, &psize) &&
         pfd >= 0 && psize == size;
    if (ok) {
      void* m = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, pfd, 0);
      ok = m != MAP_FAILED;
      if (ok) out->peers[i] = m;
    }
    if (pfd >= 0) close(pfd);
  }
  // nobody retracts (or exits) before everybody has fetched
  fx->Publish(tag + "/done", -1, ok ? 1 : 0);
  for (size_t i = 0; i < g.pids.size(); ++i) {
    if (static_cast<int>(i) == g.index) continue;
    uint64_t v = 0;
    ok = FdExchange::Fetch(FdExchange::EndpointName(g.job_port, g.pids[i]), tag + "/done", nullptr, &v) && v == 1 && ok;
  }
  fx->Retract(tag + "/mem");  // (the "/done" token stays: a slower member may not have asked for it yet)
  close(fd);
  return ok;
}
}  // namespace ps
#endif  // PS_VAN_MEM_DOMAIN_H_
