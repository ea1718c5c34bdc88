/**
 * \file range.h
 * \brief Half-open interval [begin, end) over uint64 keys: what one server owns of the key space.
 * Parity: reference include/ps/range.h:12-23 (begin / end / size), plus the helpers the
 * slicer and the tests use (contains, empty, intersection, comparison, printing).
 */
#ifndef PS_RANGE_H_
#define PS_RANGE_H_
#include <algorithm>
#include <cstdint>
#include <ostream>

#include "ps/internal/utils.h"

namespace ps {

class Range {
 public:
  constexpr Range() = default;
  constexpr Range(uint64_t first, uint64_t past_last) : lo_(first), hi_(past_last) {}

  constexpr uint64_t begin() const { return lo_; }
  constexpr uint64_t end() const { return hi_; }
  constexpr uint64_t size() const { return hi_ > lo_ ? hi_ - lo_ : 0; }
  constexpr bool empty() const { return hi_ <= lo_; }
  constexpr bool contains(uint64_t key) const { return lo_ <= key && key < hi_; }

  /*! \brief the part both ranges share (empty if they are disjoint) */
  Range intersect(const Range& other) const {
    const uint64_t lo = std::max(lo_, other.lo_), hi = std::min(hi_, other.hi_);
    return hi > lo ? Range(lo, hi) : Range();
  }

  friend constexpr bool operator==(const Range& a, const Range& b) {
    return a.lo_ == b.lo_ && a.hi_ == b.hi_;
  }
  friend constexpr bool operator!=(const Range& a, const Range& b) { return !(a == b); }
  friend std::ostream& operator<<(std::ostream& os, const Range& r) {
    return os << "[" << r.lo_ << ", " << r.hi_ << ")";
  }

 private:
  uint64_t lo_ = 0;
  uint64_t hi_ = 0;
};

}  // namespace ps
#endif  // PS_RANGE_H_
