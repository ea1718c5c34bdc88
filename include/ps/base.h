/**
 * \file base.h
 * \brief Key type and the bit-OR-able node-group ids.
 * Parity: reference include/ps/base.h:11-25 (same public names and values).
 */
#ifndef PS_BASE_H_
#define PS_BASE_H_
#include <cstdint>
#include <limits>
#include "ps/internal/utils.h"

namespace ps {

/*! \brief keys are 64-bit unless the build asks for 32 (make USE_KEY32=1) */
#if USE_KEY32
typedef uint32_t Key;
#else
typedef uint64_t Key;
#endif

/*! \brief one past the largest key a server range can end at: ranges tile [0, kMaxKey) */
constexpr Key kMaxKey = std::numeric_limits<Key>::max();

/*!
 * \brief node-group ids. Each is a single bit, so any union of groups is the sum (or OR)
 *        of its members: kWorkerGroup + kServerGroup addresses every worker and server.
 */
enum NodeGroup : int {
  kScheduler = 1 << 0,
  kServerGroup = 1 << 1,
  kWorkerGroup = 1 << 2,
};

}  // namespace ps
#endif  // PS_BASE_H_
