/**
 * \file sampler.h
 * \brief A tiny in-process sampling profiler (PS_SAMPLE_PROFILE=<file>): SIGPROF every 1 ms of
 *        consumed CPU time, the interrupted thread's call stack is recorded; at exit the raw
 *        addresses are written to <file>.<pid> for `scripts/symbolize_samples.py`.
 *
 * This image has no perf / gdb / valgrind; the host path (descriptors, queues, thread hops) is
 * what bounds the push/pull benchmark, so it has to be measurable somehow. Not async-signal-safe
 * by the letter (backtrace), fine in practice after the warm-up call in Start().
 */
#ifndef PS_CORE_SAMPLER_H_
#define PS_CORE_SAMPLER_H_
#include <execinfo.h>
#include <signal.h>
#include <sys/syscall.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace ps {

class SampleProfiler {
 public:
  static void StartIfRequested() {
    const char* path = getenv("PS_SAMPLE_PROFILE");
    if (!path || !*path) return;
    static bool started = false;
    if (started) return;
    started = true;
    State().path = std::string(path) + "." + std::to_string(getpid());
    void* warm[4];
    backtrace(warm, 4);  // loads libgcc now, not inside the signal handler
    struct sigaction sa;
    sa.sa_sigaction = &OnSignal;
    sigemptyset(&sa.sa_mask);
    sa.sa_flags = SA_SIGINFO | SA_RESTART;
    sigaction(SIGPROF, &sa, nullptr);
    struct itimerval tv;
    tv.it_interval.tv_sec = 0;
    tv.it_interval.tv_usec = 1000;
    tv.it_value = tv.it_interval;
    setitimer(ITIMER_PROF, &tv, nullptr);
    atexit(&Dump);
  }

 private:
  static constexpr int kDepth = 12;
  static constexpr int kMaxSamples = 200000;
  struct Store {
    std::string path;
    std::atomic<int> next{0};
    void* frames[kMaxSamples][kDepth];
    int depth[kMaxSamples];
    int tid[kMaxSamples];
  };
  static Store& State() {
    static Store* s = new Store();
    return *s;
  }
  static void OnSignal(int, siginfo_t*, void*) {
    Store& s = State();
    const int i = s.next.fetch_add(1);
    if (i >= kMaxSamples) return;
    s.depth[i] = backtrace(s.frames[i], kDepth);
    s.tid[i] = static_cast<int>(syscall(SYS_gettid));
  }
  static void Dump() {
    struct itimerval off = {};
    setitimer(ITIMER_PROF, &off, nullptr);
    Store& s = State();
    FILE* f = fopen(s.path.c_str(), "w");
    if (!f) return;
    // the symbolizer needs the load base of the executable (PIE)
    FILE* maps = fopen("/proc/self/maps", "r");
    if (maps) {
      char line[512];
      while (fgets(line, sizeof(line), maps)) fprintf(f, "#map %s", line);
      fclose(maps);
    }
    const int n = s.next.load() < kMaxSamples ? s.next.load() : kMaxSamples;
    for (int i = 0; i < n; ++i) {
      fprintf(f, "t%d ", s.tid[i]);
      for (int d = 2; d < s.depth[i]; ++d) fprintf(f, "%p ", s.frames[i][d]);  // skip handler frames
      fputc('\n', f);
    }
    fclose(f);
  }
};

}  // namespace ps
#endif  // PS_CORE_SAMPLER_H_
