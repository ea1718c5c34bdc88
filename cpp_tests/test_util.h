// Minimal unit-test scaffolding shared by the C++ tests.
#ifndef PS_CPP_TESTS_TEST_UTIL_H_
#define PS_CPP_TESTS_TEST_UTIL_H_
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>
#include "ps/ps.h"

struct TestCase {
  const char* name;
  std::function<void()> fn;
};
inline std::vector<TestCase>& Registry() {
  static std::vector<TestCase> r;
  return r;
}
struct Registrar {
  Registrar(const char* n, std::function<void()> f) { Registry().push_back({n, std::move(f)}); }
};
#define TEST(name)                                   \
  static void test_##name();                         \
  static Registrar reg_##name(#name, test_##name);   \
  static void test_##name()

#define EXPECT_THROW(stmt)                                             \
  do {                                                                 \
    bool threw_ = false;                                               \
    setenv("PS_BACKTRACE_DEPTH", "0", 1);                              \
    try { stmt; } catch (const dmlc::Error&) { threw_ = true; }        \
    CHECK(threw_) << "expected a failure from: " #stmt;                \
  } while (0)

inline int RunAllTests() {
  int failed = 0;
  for (auto& t : Registry()) {
    try {
      t.fn();
      fprintf(stderr, "[  OK  ] %s\n", t.name);
    } catch (const std::exception& e) {
      ++failed;
      fprintf(stderr, "[ FAIL ] %s: %s\n", t.name, e.what());
    }
  }
  fprintf(stderr, "%zu tests, %d failed\n", Registry().size(), failed);
  return failed ? 1 : 0;
}
#endif
