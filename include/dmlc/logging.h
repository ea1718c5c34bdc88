/**
 * \file logging.h
 * \brief Stream-style logging and invariant checks for the B200 parameter server.
 *
 * Source-compatible with the macro surface user code expects from the reference
 * (LOG(sev), CHECK*, LOG_IF, CHECK_NOTNULL, DLOG; parity: reference
 * include/dmlc/logging.h:74-143,168-190,249-286) but implemented from scratch:
 * one LogLine object formats into a thread-local buffer and emits with a single
 * write(2) so lines from the van / customer / app threads never interleave.
 * A failed CHECK or LOG(FATAL) throws dmlc::Error carrying a demangled
 * backtrace (override depth with PS_BACKTRACE_DEPTH, default 12).
 */
#ifndef DMLC_LOGGING_H_
#define DMLC_LOGGING_H_

#include <unistd.h>
#include <cxxabi.h>
#include <execinfo.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>

#include "dmlc/base.h"

namespace dmlc {

/*! \brief exception thrown by LOG(FATAL) and failed CHECKs */
struct Error : public std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

/*! \brief kept for API parity: logging needs no global initialisation here */
inline void InitLogging(const char* /*argv0*/) {}

namespace log_detail {

enum Severity { kInfo = 0, kWarning = 1, kError = 2, kFatal = 3 };

inline int MinLevel() {
  static int lvl = [] {
    const char* v = getenv("PS_MIN_LOG_LEVEL");
    return v ? atoi(v) : 0;
  }();
  return lvl;
}

inline std::string Backtrace(int skip) {
  int depth = 12;
  if (const char* v = getenv("PS_BACKTRACE_DEPTH")) depth = atoi(v);
  if (depth <= 0) return "";
  if (depth > 64) depth = 64;
  void* frames[64 + 4];
  int n = backtrace(frames, depth + skip);
  char** syms = backtrace_symbols(frames, n);
  std::ostringstream os;
  os << "\nStack trace:\n";
  for (int i = skip; syms && i < n; ++i) {
    std::string line(syms[i]);
    // "module(mangled+0x12) [0xaddr]" -> demangle the bit between '(' and '+'
    auto lp = line.find('('), plus = line.find('+', lp == std::string::npos ? 0 : lp);
    if (lp != std::string::npos && plus != std::string::npos && plus > lp + 1) {
      std::string sym = line.substr(lp + 1, plus - lp - 1);
      int status = 0;
      char* dem = abi::__cxa_demangle(sym.c_str(), nullptr, nullptr, &status);
      if (status == 0 && dem) line = line.substr(0, lp + 1) + dem + line.substr(plus);
      free(dem);
    }
    os << "  [" << (i - skip) << "] " << line << "\n";
  }
  free(syms);
  return os.str();
}

/*! \brief one log record; emitted (or thrown) when it goes out of scope */
class LogLine {
 public:
  LogLine(const char* file, int line, Severity sev) : sev_(sev) {
    time_t now = time(nullptr);
    struct tm tmv;
    localtime_r(&now, &tmv);
    char stamp[16];
    snprintf(stamp, sizeof(stamp), "%02d:%02d:%02d", tmv.tm_hour, tmv.tm_min, tmv.tm_sec);
    const char* base = strrchr(file, '/');
    static const char kTag[] = {'I', 'W', 'E', 'F'};
    os_ << "[" << stamp << "] " << kTag[sev] << " " << (base ? base + 1 : file) << ":" << line
        << ": ";
  }
  std::ostringstream& stream() { return os_; }
  ~LogLine() noexcept(false) {
    if (sev_ == kFatal) {
      std::string what = os_.str() + Backtrace(2);
      std::string out = what + "\n";
      ssize_t r = ::write(STDERR_FILENO, out.data(), out.size());
      (void)r;
      throw Error(what);
    }
    if (sev_ < MinLevel()) return;
    os_ << "\n";
    const std::string s = os_.str();
    ssize_t r = ::write(STDERR_FILENO, s.data(), s.size());
    (void)r;
  }

 private:
  std::ostringstream os_;
  Severity sev_;
};

/*! \brief swallows a stream expression so `cond ? (void)0 : Voidify() & stream` type-checks */
struct Voidify {
  void operator&(std::ostream&) {}
};

/*! \brief builds the "(a vs. b)" suffix of a failed binary CHECK, only on failure */
template <typename A, typename B>
inline std::unique_ptr<std::string> FormatCmp(const A& a, const B& b) {
  std::ostringstream os;
  os << " (" << a << " vs. " << b << ") ";
  return std::unique_ptr<std::string>(new std::string(os.str()));
}

#define PS_DEFINE_CMP_(name, op)                                               \
  template <typename A, typename B>                                            \
  inline std::unique_ptr<std::string> Cmp##name(const A& a, const B& b) {      \
    if (a op b) return nullptr;                                                \
    return FormatCmp(a, b);                                                    \
  }                                                                            \
  inline std::unique_ptr<std::string> Cmp##name(int a, int b) {                \
    if (a op b) return nullptr;                                                \
    return FormatCmp(a, b);                                                    \
  }
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wsign-compare"
PS_DEFINE_CMP_(EQ, ==)
PS_DEFINE_CMP_(NE, !=)
PS_DEFINE_CMP_(LT, <)
PS_DEFINE_CMP_(LE, <=)
PS_DEFINE_CMP_(GT, >)
PS_DEFINE_CMP_(GE, >=)
#pragma GCC diagnostic pop
#undef PS_DEFINE_CMP_

template <typename T>
inline T* NotNull(const char* file, int line, const char* expr, T* p) {
  if (p == nullptr) {
    LogLine(file, line, kFatal).stream() << "Check notnull failed: " << expr;
  }
  return p;
}

}  // namespace log_detail
}  // namespace dmlc

#define PS_LOG_SEV_INFO ::dmlc::log_detail::kInfo
#define PS_LOG_SEV_WARNING ::dmlc::log_detail::kWarning
#define PS_LOG_SEV_ERROR ::dmlc::log_detail::kError
#define PS_LOG_SEV_FATAL ::dmlc::log_detail::kFatal

#define LOG(sev) ::dmlc::log_detail::LogLine(__FILE__, __LINE__, PS_LOG_SEV_##sev).stream()
#define LOG_IF(sev, cond) \
  !(cond) ? (void)0 : ::dmlc::log_detail::Voidify() & LOG(sev)

#define CHECK(x)                                                              \
  if (!(x))                                                                   \
  ::dmlc::log_detail::LogLine(__FILE__, __LINE__, ::dmlc::log_detail::kFatal) \
          .stream()                                                           \
      << "Check failed: " #x << ' '

#define PS_CHECK_BINARY_(name, x, y, opstr)                                    \
  if (auto ps_cmp_msg_ = ::dmlc::log_detail::Cmp##name((x), (y)))              \
  ::dmlc::log_detail::LogLine(__FILE__, __LINE__, ::dmlc::log_detail::kFatal)  \
          .stream()                                                            \
      << "Check failed: " #x " " opstr " " #y << *ps_cmp_msg_

#define CHECK_EQ(x, y) PS_CHECK_BINARY_(EQ, x, y, "==")
#define CHECK_NE(x, y) PS_CHECK_BINARY_(NE, x, y, "!=")
#define CHECK_LT(x, y) PS_CHECK_BINARY_(LT, x, y, "<")
#define CHECK_LE(x, y) PS_CHECK_BINARY_(LE, x, y, "<=")
#define CHECK_GT(x, y) PS_CHECK_BINARY_(GT, x, y, ">")
#define CHECK_GE(x, y) PS_CHECK_BINARY_(GE, x, y, ">=")
#define CHECK_NOTNULL(x) \
  ::dmlc::log_detail::NotNull(__FILE__, __LINE__, #x, (x))

#ifdef NDEBUG
#define DLOG(sev) true ? (void)0 : ::dmlc::log_detail::Voidify() & LOG(sev)
#define DCHECK(x) while (false) CHECK(x)
#define DCHECK_EQ(x, y) while (false) CHECK_EQ(x, y)
#define DCHECK_NE(x, y) while (false) CHECK_NE(x, y)
#define DCHECK_LT(x, y) while (false) CHECK_LT(x, y)
#define DCHECK_LE(x, y) while (false) CHECK_LE(x, y)
#define DCHECK_GT(x, y) while (false) CHECK_GT(x, y)
#define DCHECK_GE(x, y) while (false) CHECK_GE(x, y)
#else
#define DLOG(sev) LOG(sev)
#define DCHECK(x) CHECK(x)
#define DCHECK_EQ(x, y) CHECK_EQ(x, y)
#define DCHECK_NE(x, y) CHECK_NE(x, y)
#define DCHECK_LT(x, y) CHECK_LT(x, y)
#define DCHECK_LE(x, y) CHECK_LE(x, y)
#define DCHECK_GT(x, y) CHECK_GT(x, y)
#define DCHECK_GE(x, y) CHECK_GE(x, y)
#endif

#endif  // DMLC_LOGGING_H_
