/**
 * \file utils.h
 * \brief Small helpers shared by every layer: typed env lookup, LL.
 * Parity: reference include/ps/internal/utils.h:29-54.
 */
#ifndef PS_INTERNAL_UTILS_H_
#define PS_INTERNAL_UTILS_H_
#include <cstdlib>
#include <string>
#include "dmlc/logging.h"
#include "ps/internal/env.h"

namespace ps {

/*! \brief integer env lookup through ps::Environment (user map first, then getenv) */
template <typename V>
inline V GetEnv(const char* name, V fallback) {
  const char* v = Environment::Get()->find(name);
  return v == nullptr ? fallback : static_cast<V>(atoll(v));
}
inline const char* GetEnv(const char* name, const char* fallback) {
  const char* v = Environment::Get()->find(name);
  return v == nullptr ? fallback : v;
}
inline std::string GetEnvStr(const char* name, const std::string& fallback = "") {
  const char* v = Environment::Get()->find(name);
  return v == nullptr ? fallback : std::string(v);
}

#ifndef LL
/*! \brief always-visible log line used by the benchmarks */
#define LL LOG(ERROR)
#endif

}  // namespace ps
#endif  // PS_INTERNAL_UTILS_H_
