/**
 * \file event_trace.h
 * \brief PS_EVENT_TRACE=<file>: a timeline of the message path inside one process.
 *
 * Every instrumented point appends (nanoseconds, thread, tag, a, b) to a preallocated array —
 * one relaxed fetch_add and five stores, cheap enough to leave the timing of a 10 µs round trip
 * intact — and the array is written to <file>.<pid> at exit. Merging the files of a worker and a
 * server (CLOCK_MONOTONIC is shared on one host) shows where a request spends its time between
 * the application thread, the rings, the receive threads and the customer threads:
 * `scripts/event_timeline.py`. Off (one predictable branch per point) unless the variable is set.
 */
#ifndef PS_CORE_EVENT_TRACE_H_
#define PS_CORE_EVENT_TRACE_H_
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace ps {

class EventTrace {
 public:
  static inline bool enabled() { return State().on; }
  static inline void Mark(const char* tag, long a = 0, long b = 0) {
    Store& s = State();
    if (!s.on) return;
    const int i = s.next.fetch_add(1, std::memory_order_relaxed);
    if (i >= kMax) return;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    static thread_local const int tid = static_cast<int>(syscall(SYS_gettid));
    Event& e = s.events[i];
    e.ns = static_cast<long long>(ts.tv_sec) * 1000000000LL + ts.tv_nsec;
    e.tid = tid;
    e.tag = tag;
    e.a = a;
    e.b = b;
  }

 private:
  static constexpr int kMax = 1 << 20;
  struct Event {
    long long ns;
    int tid;
    const char* tag;
    long a, b;
  };
  struct Store {
    bool on = false;
    std::string path;
    std::atomic<int> next{0};
    Event* events = nullptr;
  };
  static Store& State() {
    static Store* s = [] {
      Store* st = new Store();
      const char* p = getenv("PS_EVENT_TRACE");
      if (p && *p) {
        st->path = std::string(p) + "." + std::to_string(getpid());
        st->events = static_cast<Event*>(malloc(sizeof(Event) * kMax));
        // touch every page now: a first-touch fault inside Mark() would cost more than the event it records
        if (st->events) memset(st->events, 0, sizeof(Event) * kMax);
        st->on = st->events != nullptr;
        if (st->on) atexit(&Dump);
      }
      return st;
    }();
    return *s;
  }
  static void Dump() {
    Store& s = State();
    FILE* f = fopen(s.path.c_str(), "w");
    if (!f) return;
    const int n = s.next.load() < kMax ? s.next.load() : kMax;
    for (int i = 0; i < n; ++i) {
      const Event& e = s.events[i];
      fprintf(f, "%lld\t%d\t%s\t%ld\t%ld\n", e.ns, e.tid, e.tag ? e.tag : "?", e.a, e.b);
    }
    fclose(f);
  }
};

}  // namespace ps
#endif  // PS_CORE_EVENT_TRACE_H_
