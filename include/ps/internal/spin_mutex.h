/**
 * \file spin_mutex.h
 * \brief Lock for the short critical sections of the message path.
 *
 * A contended std::mutex parks the loser in the kernel; on the virtual machines this code runs
 * on, getting a parked thread back costs on the order of 100 µs — more than ten whole push/pull
 * round trips — while the sections guarded here (a deque push, a tracker slot update, a map
 * lookup) last well under a microsecond. SpinMutex therefore waits in user space: test-and-
 * test-and-set with `pause`, and only after a few thousand polls (a holder that lost its CPU)
 * starts yielding the time slice. It satisfies Lockable, so std::lock_guard / std::unique_lock
 * work, and std::condition_variable_any can sleep on it where a thread really has to wait.
 */
#ifndef PS_INTERNAL_SPIN_MUTEX_H_
#define PS_INTERNAL_SPIN_MUTEX_H_
#include <atomic>
#include <thread>

namespace ps {

class SpinMutex {
 public:
  SpinMutex() = default;
  SpinMutex(const SpinMutex&) = delete;
  SpinMutex& operator=(const SpinMutex&) = delete;

  void lock() {
    int polls = 0;
    for (;;) {
      if (!held_.exchange(true, std::memory_order_acquire)) return;
      while (held_.load(std::memory_order_relaxed)) {
        if (++polls < 4096) {
#if defined(__x86_64__) || defined(__i386__)
          __builtin_ia32_pause();
#endif
        } else {
          std::this_thread::yield();
        }
      }
    }
  }
  bool try_lock() {
    return !held_.load(std::memory_order_relaxed) && !held_.exchange(true, std::memory_order_acquire);
  }
  void unlock() { held_.store(false, std::memory_order_release); }

 private:
  std::atomic<bool> held_{false};
};

}  // namespace ps
#endif  // PS_INTERNAL_SPIN_MUTEX_H_
