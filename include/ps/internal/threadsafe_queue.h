/**
 * \file threadsafe_queue.h
 * \brief Blocking MPMC queue used between the van receive thread and customers.
 *
 * Two modes (parity: reference include/ps/internal/threadsafe_queue.h:18-118):
 *  - default: mutex + condition variable over a deque.
 *  - DMLC_LOCKLESS_QUEUE=1: an SPSC ring; producers serialise on a tiny spinlock,
 *    the consumer spins for DMLC_POLLING_IN_NANOSECOND ns, then yields, then naps.
 *    This removes the futex wake from the 1 KB push/pull latency path.
 */
#ifndef PS_INTERNAL_THREADSAFE_QUEUE_H_
#define PS_INTERNAL_THREADSAFE_QUEUE_H_
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include "ps/internal/env.h"
#include "ps/internal/utils.h"
#include "ps/internal/spin_mutex.h"
#include "ps/internal/spsc_queue.h"

namespace ps {

template <typename T>
class ThreadsafeQueue {
 public:
  ThreadsafeQueue() {
    const char* v = Environment::Get()->find("DMLC_LOCKLESS_QUEUE");
    lockless_ = v && atoi(v) != 0;
    if (lockless_) {
      const char* p = Environment::Get()->find("DMLC_POLLING_IN_NANOSECOND");
      spin_ns_ = p ? atoll(p) : 1000;
      ring_.reset(new SPSCQueue<T>(kRingCapacity));
    }
  }
  ~ThreadsafeQueue() {}

  void Push(T v) {
    if (lockless_) {
      while (push_lock_.test_and_set(std::memory_order_acquire)) {
      }
      while (!ring_->try_push(std::move(v))) std::this_thread::yield();
      push_lock_.clear(std::memory_order_release);
      return;
    }
    {
      std::lock_guard<SpinMutex> lk(mu_);
      items_.push_back(std::move(v));
      count_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_one();
  }

  /*! \brief block until an item is available */
  void WaitAndPop(T* out) {
    if (lockless_) {
      auto t0 = std::chrono::steady_clock::now();
      int naps = 0;
      while (!ring_->try_pop(out)) {
        auto waited = std::chrono::duration_cast<std::chrono::nanoseconds>(
                          std::chrono::steady_clock::now() - t0).count();
        if (waited < spin_ns_) continue;
        if (++naps < 64) {
          std::this_thread::yield();
        } else {
          std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
      }
      return;
    }
    // stay hot for a while before sleeping: a producer that finds no sleeper skips the futex
    // wake (a system call per message otherwise), and the consumer skips the context switch.
    // How long is decided by the gaps seen lately (SpinBudget).
    if (count_.load(std::memory_order_acquire) == 0) {
      const auto t0 = std::chrono::steady_clock::now();
      SpinPoll([this] { return count_.load(std::memory_order_acquire) != 0; }, budget_.floor_us(),
               budget_.window_us());
      std::unique_lock<SpinMutex> lk(mu_);
      cv_.wait(lk, [this] { return !items_.empty(); });
      budget_.Observe(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0)
                          .count());
      *out = std::move(items_.front());
      items_.pop_front();
      count_.fetch_sub(1, std::memory_order_release);
      return;
    }
    std::unique_lock<SpinMutex> lk(mu_);
    cv_.wait(lk, [this] { return !items_.empty(); });
    *out = std::move(items_.front());
    items_.pop_front();
    count_.fetch_sub(1, std::memory_order_release);
  }

  /*! \brief non-blocking pop */
  bool TryPop(T* out) {
    if (lockless_) return ring_->try_pop(out);
    if (count_.load(std::memory_order_acquire) == 0) return false;
    std::lock_guard<SpinMutex> lk(mu_);
    if (items_.empty()) return false;
    *out = std::move(items_.front());
    items_.pop_front();
    count_.fetch_sub(1, std::memory_order_release);
    return true;
  }

  size_t Size() {
    if (lockless_) return ring_->size();
    std::lock_guard<SpinMutex> lk(mu_);
    return items_.size();
  }

  /*! \brief one polite busy-wait step */
  static void CpuRelax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }

 private:
  static constexpr size_t kRingCapacity = 32768;
  bool lockless_ = false;
  long long spin_ns_ = 1000;
  SpinMutex mu_;  // never held for longer than a deque operation
  std::condition_variable_any cv_;
  std::deque<T> items_;
  std::atomic<size_t> count_{0};  // == items_.size(), readable without the lock
  SpinBudget budget_{GetEnv("PS_QUEUE_SPIN_US", 20), GetEnv("PS_SPIN_MAX_US", 1000)};
  std::unique_ptr<SPSCQueue<T>> ring_;
  std::atomic_flag push_lock_ = ATOMIC_FLAG_INIT;
};

}  // namespace ps
#endif  // PS_INTERNAL_THREADSAFE_QUEUE_H_
