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
#include <chrono>
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

/*!
 * \brief poll `ready()` for up to `window_us`: the first `hot_us` with `pause` (lowest latency,
 *        owns the CPU), the rest with sched_yield between polls — on an oversubscribed host a
 *        polling thread must not keep a thread with real work off the CPU. True if `ready()`.
 */
template <typename Pred>
inline bool SpinPoll(Pred ready, int hot_us, int window_us) {
  if (ready()) return true;
  if (window_us <= 0) return false;
  const auto t0 = std::chrono::steady_clock::now();
  const auto hot_end = t0 + std::chrono::microseconds(hot_us < window_us ? hot_us : window_us);
  const auto end = t0 + std::chrono::microseconds(window_us);
  int polls = 0;
  for (;;) {  // hot phase
    if (ready()) return true;
    if ((++polls & 63) == 0 && std::chrono::steady_clock::now() >= hot_end) break;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  while (std::chrono::steady_clock::now() < end) {  // polite phase
    if (ready()) return true;
    std::this_thread::yield();
  }
  return ready();
}

/*!
 * \brief how long a consumer thread should busy-poll before it goes to sleep.
 *
 * Traffic comes in bursts (the keys of one round, the layers of one backward pass): inside a
 * burst the next message is microseconds away, between bursts the consumer idles for a few
 * hundred microseconds — and sleeping through such a gap costs more than the gap (a wake-up is
 * 100-300 us on a virtual machine). SpinBudget remembers how long the *long* gaps have been
 * lately (those that outlasted the floor window): if they are shorter than `cap`, polling
 * through them is worth it and the window grows to twice their typical length; if they are
 * longer (an idle job, a compute phase), the window falls back to `floor` and the thread sleeps.
 * One instance per consumer; the counters are relaxed atomics only to keep sanitizers quiet.
 */
class SpinBudget {
 public:
  SpinBudget(int floor_us, int cap_us) : floor_(floor_us), cap_(cap_us), long_gap_us_(cap_us / 2) {}
  int window_us() const {
    if (cap_ <= floor_) return floor_;
    const int typical = long_gap_us_.load(std::memory_order_relaxed);
    if (typical > cap_) return floor_;
    const int w = 2 * typical;
    return w < floor_ ? floor_ : (w > cap_ ? cap_ : w);
  }
  /*! \brief the consumer waited `gap_us` for its next item (polling or asleep) */
  void Observe(long long gap_us) {
    if (cap_ <= floor_ || gap_us < floor_) return;  // inside a burst: says nothing about the gaps between
    const long long clipped = gap_us > 4LL * cap_ ? 4LL * cap_ : gap_us;
    const int old = long_gap_us_.load(std::memory_order_relaxed);
    long_gap_us_.store(static_cast<int>((3LL * old + clipped) / 4), std::memory_order_relaxed);
  }
  int floor_us() const { return floor_; }

 private:
  const int floor_;
  const int cap_;
  std::atomic<int> long_gap_us_;
};

}  // namespace ps
#endif  // PS_INTERNAL_SPIN_MUTEX_H_
