/**
 * \file spsc_queue.h
 * \brief Bounded wait-free single-producer / single-consumer ring.
 *
 * Written for this project (the reference vendors a third-party ring,
 * include/ps/internal/spsc_queue.h:32-175). Layout choices: power-of-two
 * capacity so wrap is a mask; head and tail live on separate 128-byte lines (the
 * adjacent-line prefetcher on current x86 pulls pairs of 64-byte lines); each
 * side caches the other side's index so the common case touches one shared line.
 * The same ring layout is reused for the cross-process mailbox in shared memory
 * (src/van/shm_ring.h), which is why it is trivially relocatable (indices only).
 */
#ifndef PS_INTERNAL_SPSC_QUEUE_H_
#define PS_INTERNAL_SPSC_QUEUE_H_
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <new>
#include <utility>

namespace ps {

template <typename T>
class SPSCQueue {
 public:
  /*! \param min_capacity rounded up to a power of two (>= 2) */
  explicit SPSCQueue(size_t min_capacity) {
    size_t cap = 2;
    while (cap < min_capacity) cap <<= 1;
    mask_ = cap - 1;
    slots_ = static_cast<T*>(::operator new[](cap * sizeof(T), std::align_val_t(alignof(T) > 64 ? alignof(T) : 64)));
  }
  ~SPSCQueue() {
    T tmp;
    while (try_pop(&tmp)) {
    }
    ::operator delete[](slots_, std::align_val_t(alignof(T) > 64 ? alignof(T) : 64));
  }
  SPSCQueue(const SPSCQueue&) = delete;
  SPSCQueue& operator=(const SPSCQueue&) = delete;

  /*! \brief producer side; false if full */
  template <typename U>
  bool try_push(U&& v) {
    const uint64_t t = tail_.load(std::memory_order_relaxed);
    if (t - head_cache_ > mask_) {
      head_cache_ = head_.load(std::memory_order_acquire);
      if (t - head_cache_ > mask_) return false;
    }
    new (&slots_[t & mask_]) T(std::forward<U>(v));
    tail_.store(t + 1, std::memory_order_release);
    return true;
  }
  /*! \brief consumer side; false if empty */
  bool try_pop(T* out) {
    const uint64_t h = head_.load(std::memory_order_relaxed);
    if (h == tail_cache_) {
      tail_cache_ = tail_.load(std::memory_order_acquire);
      if (h == tail_cache_) return false;
    }
    T* slot = &slots_[h & mask_];
    *out = std::move(*slot);
    slot->~T();
    head_.store(h + 1, std::memory_order_release);
    return true;
  }
  size_t size() const {
    return static_cast<size_t>(tail_.load(std::memory_order_acquire) -
                               head_.load(std::memory_order_acquire));
  }
  bool empty() const { return size() == 0; }
  size_t capacity() const { return mask_ + 1; }

 private:
  static constexpr size_t kLine = 128;
  T* slots_ = nullptr;
  uint64_t mask_ = 0;
  alignas(kLine) std::atomic<uint64_t> tail_{0};
  uint64_t head_cache_ = 0;  // producer-private
  alignas(kLine) std::atomic<uint64_t> head_{0};
  uint64_t tail_cache_ = 0;  // consumer-private
  char pad_[kLine - sizeof(std::atomic<uint64_t>) - sizeof(uint64_t)];
};

}  // namespace ps
#endif  // PS_INTERNAL_SPSC_QUEUE_H_
