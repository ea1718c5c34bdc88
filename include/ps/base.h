/**
 * \file base.h
 * \brief Key type and the bit-OR-able node-group ids.
 * Parity: reference include/ps/base.h:11-25.
 */
#ifndef PS_BASE_H_
#define PS_BASE_H_
#include <cstdint>
#include <limits>
#include "ps/internal/utils.h"

namespace ps {

#if USE_KEY32
using Key = uint32_t;
#else
using Key = uint64_t;
#endif
/*! \brief largest representable key; server key ranges partition [0, kMaxKey) */
static const Key kMaxKey = std::numeric_limits<Key>::max();

/*! \brief group ids are single bits so that groups compose with + or | */
static const int kScheduler = 1;
static const int kServerGroup = 2;
static const int kWorkerGroup = 4;

}  // namespace ps
#endif  // PS_BASE_H_
