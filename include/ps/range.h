/**
 * \file range.h
 * \brief Half-open interval [begin, end) over uint64 keys.
 * Parity: reference include/ps/range.h:12-23.
 */
#ifndef PS_RANGE_H_
#define PS_RANGE_H_
#include <cstdint>
#include "ps/internal/utils.h"

namespace ps {

class Range {
 public:
  Range() : lo_(0), hi_(0) {}
  Range(uint64_t begin, uint64_t end) : lo_(begin), hi_(end) {}
  uint64_t begin() const { return lo_; }
  uint64_t end() const { return hi_; }
  uint64_t size() const { return hi_ - lo_; }
  bool contains(uint64_t k) const { return k >= lo_ && k < hi_; }
  bool operator==(const Range& o) const { return lo_ == o.lo_ && hi_ == o.hi_; }

 private:
  uint64_t lo_, hi_;
};

}  // namespace ps
#endif  // PS_RANGE_H_
