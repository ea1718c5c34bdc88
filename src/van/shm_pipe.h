/**
 * \file shm_pipe.h
 * \brief ShmPipe: a unidirectional byte stream through a POSIX shared-memory ring.
 *
 * On one box every peer is "same host", so message descriptors do not need the
 * kernel's socket path (2 system calls and 2 copies per message). The sending
 * van writes its frames into this ring; the receiving van's thread polls it and
 * only falls back to sleeping in epoll — woken by a 1-byte doorbell on the TCP
 * connection — when it has been idle for a while. This is the host-side mailbox
 * of the SURVEY design (§5.8 "descriptor in a per-pair SPSC mailbox ring").
 *
 * Layout: [Ctl | data[capacity]]; head/tail are free-running byte counters on
 * separate cache-line pairs; capacity is a power of two. A frame larger than the
 * ring streams through it (the writer publishes as it goes).
 */
#ifndef PS_VAN_SHM_PIPE_H_
#define PS_VAN_SHM_PIPE_H_
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/uio.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>

namespace ps {

class ShmPipe {
 public:
  ~ShmPipe() {
    if (base_) munmap(base_, map_bytes_);
    if (owner_ && !name_.empty()) shm_unlink(name_.c_str());
  }

  /*! \brief producer side: create and map a fresh ring */
  static std::unique_ptr<ShmPipe> Create(const std::string& name, size_t capacity) {
    size_t cap = 4096;
    while (cap < capacity) cap <<= 1;
    shm_unlink(name.c_str());
    int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return nullptr;
    const size_t bytes = sizeof(Ctl) + cap;
    // reserve the pages now: on a small /dev/shm (containers default to 64 MB) a ring that
    // only ftruncate()d would SIGBUS on first touch instead of failing here
    if (ftruncate(fd, static_cast<off_t>(bytes)) != 0 ||
        posix_fallocate(fd, 0, static_cast<off_t>(bytes)) != 0) {
      close(fd);
      shm_unlink(name.c_str());
      return nullptr;
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
      shm_unlink(name.c_str());
      return nullptr;
    }
    std::unique_ptr<ShmPipe> pipe(new ShmPipe());
    pipe->Adopt(p, bytes, name, true);
    Ctl* c = pipe->ctl_;
    new (&c->tail) std::atomic<uint64_t>(0);
    new (&c->head) std::atomic<uint64_t>(0);
    new (&c->sleeping) std::atomic<uint32_t>(1);  // the first message rings the doorbell
    new (&c->gate_done) std::atomic<uint64_t>(0);
    c->capacity = static_cast<uint32_t>(cap);
    std::atomic_thread_fence(std::memory_order_seq_cst);
    return pipe;
  }

  /*! \brief consumer side: map an existing ring (the name can be unlinked afterwards) */
  static std::unique_ptr<ShmPipe> Attach(const std::string& name) {
    int fd = shm_open(name.c_str(), O_RDWR, 0600);
    if (fd < 0) return nullptr;
    struct stat st;
    if (fstat(fd, &st) != 0 || static_cast<size_t>(st.st_size) <= sizeof(Ctl)) {
      close(fd);
      return nullptr;
    }
    void* p = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return nullptr;
    std::unique_ptr<ShmPipe> pipe(new ShmPipe());
    pipe->Adopt(p, static_cast<size_t>(st.st_size), name, false);
    return pipe;
  }

  // ---- producer ---------------------------------------------------------------
  /*! \brief append n bytes (streams when n exceeds the free space); false if the reader is gone */
  bool Write(const void* src, size_t n) {
    const char* p = static_cast<const char*>(src);
    const uint64_t cap = ctl_->capacity, mask = cap - 1;
    uint64_t tail = ctl_->tail.load(std::memory_order_relaxed);
    auto idle_since = std::chrono::steady_clock::time_point();
    while (n) {
      const uint64_t head = ctl_->head.load(std::memory_order_acquire);
      const uint64_t space = cap - (tail - head);
      if (space == 0) {
        // ring full: the reader may be asleep — it must be woken by the caller's doorbell,
        // so report "needs doorbell" through the flag and keep waiting politely
        if (idle_since == std::chrono::steady_clock::time_point()) {
          idle_since = std::chrono::steady_clock::now();
        } else if (std::chrono::steady_clock::now() - idle_since > std::chrono::seconds(60)) {
          return false;
        }
        if (full_hook_) full_hook_();
        std::this_thread::yield();
        continue;
      }
      idle_since = std::chrono::steady_clock::time_point();
      const size_t chunk = static_cast<size_t>(std::min<uint64_t>(space, n));
      const size_t at = static_cast<size_t>(tail & mask);
      const size_t first = std::min(chunk, static_cast<size_t>(cap - at));
      memcpy(data_ + at, p, first);
      if (chunk > first) memcpy(data_, p + first, chunk - first);
      tail += chunk;
      p += chunk;
      n -= chunk;
      ctl_->tail.store(tail, std::memory_order_release);
    }
    return true;
  }
  /*!
   * \brief append all pieces of one frame. A frame of at most half the ring is published with ONE
   *        store of `tail` once it fits (the reader never sees a partial frame, and the cache line
   *        of `tail` changes hands once per frame instead of once per piece); larger frames stream.
   */
  bool WriteV(const struct iovec* iov, int niov) {
    size_t total = 0;
    for (int i = 0; i < niov; ++i) total += iov[i].iov_len;
    const uint64_t cap = ctl_->capacity, mask = cap - 1;
    if (total > cap / 2) {
      for (int i = 0; i < niov; ++i) {
        if (!Write(iov[i].iov_base, iov[i].iov_len)) return false;
      }
      return true;
    }
    uint64_t tail = ctl_->tail.load(std::memory_order_relaxed);
    auto idle_since = std::chrono::steady_clock::time_point();
    while (cap - (tail - ctl_->head.load(std::memory_order_acquire)) < total) {
      if (idle_since == std::chrono::steady_clock::time_point()) {
        idle_since = std::chrono::steady_clock::now();
      } else if (std::chrono::steady_clock::now() - idle_since > std::chrono::seconds(60)) {
        return false;
      }
      if (full_hook_) full_hook_();
      std::this_thread::yield();
    }
    for (int i = 0; i < niov; ++i) {
      const char* p = static_cast<const char*>(iov[i].iov_base);
      const size_t n = iov[i].iov_len;
      const size_t at = static_cast<size_t>(tail & mask);
      const size_t first = std::min(n, static_cast<size_t>(cap - at));
      memcpy(data_ + at, p, first);
      if (n > first) memcpy(data_, p + first, n - first);
      tail += n;
    }
    ctl_->tail.store(tail, std::memory_order_release);
    return true;
  }
  /*! \brief bytes that can be written right now without waiting for the reader */
  size_t FreeSpace() const {
    const uint64_t used = ctl_->tail.load(std::memory_order_relaxed) - ctl_->head.load(std::memory_order_acquire);
    return static_cast<size_t>(ctl_->capacity - used);
  }
  size_t capacity() const { return ctl_->capacity; }
  /*! \brief after a frame: true if the reader declared itself asleep (ring its doorbell) */
  bool ReaderNeedsDoorbell() {
    std::atomic_thread_fence(std::memory_order_seq_cst);
    if (ctl_->sleeping.load(std::memory_order_seq_cst) == 0) return false;
    return ctl_->sleeping.exchange(0, std::memory_order_seq_cst) != 0;
  }
  /*! \brief called while the ring is full (lets the owner ring the doorbell) */
  void set_full_hook(std::function<void()> f) { full_hook_ = std::move(f); }

  // ---- completion gate ---------------------------------------------------------
  // A frame may announce a payload that a device kernel is still writing into the reader's
  // memory. Such a frame carries gate = k (the k-th gated frame of this ring); whoever moves
  // the payload stores k to `gate_done` once it is globally visible (the copy kernel itself,
  // with st.release.sys through a device mapping of this word, or the CPU twin). The reader
  // leaves the frame in the ring until gate_done >= k: the role of the reference's
  // RDMA WRITE_WITH_IMM completion (src/rdma_transport.h:211-231), with no host thread in
  // between the copy and the descriptor.
  /*! \brief host address of the 8-byte completion word (its page can be mapped into a device) */
  void* gate_word() { return &ctl_->gate_done; }
  uint64_t gate_done() const { return ctl_->gate_done.load(std::memory_order_acquire); }
  /*! \brief producer side, CPU twin of the kernel's store */
  void SignalGate(uint64_t k) { ctl_->gate_done.store(k, std::memory_order_release); }
  /*! \brief start of the mapping and the bytes in front of the data area (page-lockable) */
  void* map_base() { return base_; }

  // ---- consumer ---------------------------------------------------------------
  size_t Readable() const {
    return static_cast<size_t>(ctl_->tail.load(std::memory_order_acquire) - rcur_);
  }
  /*! \brief copy the next n bytes without consuming them; false if fewer have been published */
  bool Peek(void* dst, size_t n) const {
    if (Readable() < n) return false;
    const uint64_t cap = ctl_->capacity, mask = cap - 1;
    const size_t at = static_cast<size_t>(rcur_ & mask);
    const size_t first = std::min(n, static_cast<size_t>(cap - at));
    memcpy(dst, data_ + at, first);
    if (n > first) memcpy(static_cast<char*>(dst) + first, data_, n - first);
    return true;
  }
  /*!
   * \brief consume exactly n bytes, waiting for the writer if needed; false on a dead writer.
   *        The space is handed back to the writer (store of `head`) only every quarter ring or
   *        at Commit(): one store per frame instead of one per piece.
   */
  bool Read(void* dst, size_t n) {
    char* p = static_cast<char*>(dst);
    const uint64_t cap = ctl_->capacity, mask = cap - 1;
    auto idle_since = std::chrono::steady_clock::time_point();
    while (n) {
      const uint64_t avail = ctl_->tail.load(std::memory_order_acquire) - rcur_;
      if (avail == 0) {
        Commit();  // the writer may be waiting for exactly the space we have not returned yet
        if (idle_since == std::chrono::steady_clock::time_point()) {
          idle_since = std::chrono::steady_clock::now();
        } else if (std::chrono::steady_clock::now() - idle_since > std::chrono::seconds(60)) {
          return false;
        }
        std::this_thread::yield();
        continue;
      }
      idle_since = std::chrono::steady_clock::time_point();
      const size_t chunk = static_cast<size_t>(std::min<uint64_t>(avail, n));
      const size_t at = static_cast<size_t>(rcur_ & mask);
      const size_t first = std::min(chunk, static_cast<size_t>(cap - at));
      memcpy(p, data_ + at, first);
      if (chunk > first) memcpy(p + first, data_, chunk - first);
      rcur_ += chunk;
      p += chunk;
      n -= chunk;
      if (rcur_ - published_ >= cap / 4) Commit();
    }
    return true;
  }
  /*! \brief end of a frame: return everything consumed so far to the writer */
  void Commit() {
    if (published_ != rcur_) {
      published_ = rcur_;
      ctl_->head.store(rcur_, std::memory_order_release);
    }
  }
  /*! \brief announce "about to sleep"; returns false (and cancels) if data is already there */
  bool PrepareSleep() {
    ctl_->sleeping.store(1, std::memory_order_seq_cst);
    std::atomic_thread_fence(std::memory_order_seq_cst);
    if (Readable() > 0) {
      ctl_->sleeping.store(0, std::memory_order_seq_cst);
      return false;
    }
    return true;
  }
  void CancelSleep() { ctl_->sleeping.store(0, std::memory_order_seq_cst); }

  const std::string& name() const { return name_; }
  void Unlink() {
    if (!name_.empty()) shm_unlink(name_.c_str());
  }

 private:
  struct Ctl {
    alignas(128) std::atomic<uint64_t> tail;
    alignas(128) std::atomic<uint64_t> head;
    alignas(128) std::atomic<uint32_t> sleeping;
    uint32_t capacity;
    char pad[128 - sizeof(std::atomic<uint32_t>) - sizeof(uint32_t)];
    alignas(128) std::atomic<uint64_t> gate_done;  // written by a device kernel or the CPU twin
    char pad2[128 - sizeof(std::atomic<uint64_t>)];
  };
  ShmPipe() {}
  void Adopt(void* p, size_t bytes, const std::string& name, bool owner) {
    base_ = p;
    map_bytes_ = bytes;
    ctl_ = static_cast<Ctl*>(p);
    data_ = static_cast<char*>(p) + sizeof(Ctl);
    name_ = name;
    owner_ = owner;
    if (!owner) rcur_ = published_ = ctl_->head.load(std::memory_order_relaxed);
  }
  void* base_ = nullptr;
  size_t map_bytes_ = 0;
  Ctl* ctl_ = nullptr;
  char* data_ = nullptr;
  std::string name_;
  bool owner_ = false;
  std::function<void()> full_hook_;
  uint64_t rcur_ = 0;       // consumer only: bytes consumed
  uint64_t published_ = 0;  // consumer only: value of `head` the writer can see
};

}  // namespace ps
#endif  // PS_VAN_SHM_PIPE_H_
