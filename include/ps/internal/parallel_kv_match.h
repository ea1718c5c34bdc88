/**
 * \file parallel_kv_match.h
 * \brief Merge a sorted (key, value) list into another by key, in parallel.
 *
 * For every key present in both lists: dst_val[k] <op>= src_val[k] (k values per key).
 * Parity: reference include/ps/internal/parallel_kv_match.h:29-120 (unused there and no
 * longer compiling); this version works and is unit-tested.
 */
#ifndef PS_INTERNAL_PARALLEL_KV_MATCH_H_
#define PS_INTERNAL_PARALLEL_KV_MATCH_H_
#include <algorithm>
#include <thread>
#include "ps/internal/assign_op.h"
#include "ps/sarray.h"

namespace ps {
namespace match_detail {

template <typename K, typename V>
void MatchRange(const K* src_key, const K* src_key_end, const V* src_val, const K* dst_key,
                const K* dst_key_end, V* dst_val, int k, AssignOp op, size_t grainsize,
                size_t* matched) {
  const size_t src_len = static_cast<size_t>(src_key_end - src_key);
  const size_t dst_len = static_cast<size_t>(dst_key_end - dst_key);
  if (src_len == 0 || dst_len == 0) return;
  // trim to the overlapping key interval
  src_key = std::lower_bound(src_key, src_key_end, *dst_key);
  src_val += (src_key - (src_key_end - src_len)) * k;
  if (src_key == src_key_end) return;
  const K* dst_begin = dst_key;
  dst_key = std::lower_bound(dst_key, dst_key_end, *src_key);
  dst_val += (dst_key - dst_begin) * k;
  if (dst_key == dst_key_end) return;

  if (static_cast<size_t>(src_key_end - src_key) + static_cast<size_t>(dst_key_end - dst_key) <=
      grainsize) {
    while (src_key != src_key_end && dst_key != dst_key_end) {
      if (*src_key < *dst_key) {
        ++src_key;
        src_val += k;
      } else if (*dst_key < *src_key) {
        ++dst_key;
        dst_val += k;
      } else {
        for (int i = 0; i < k; ++i) AssignFunc(src_val[i], op, &dst_val[i]);
        ++src_key; ++dst_key;
        src_val += k; dst_val += k;
        ++*matched;
      }
    }
    return;
  }
  // split the source in half; each half only touches a disjoint part of dst
  const K* mid = src_key + (src_key_end - src_key) / 2;
  size_t left_n = 0, right_n = 0;
  std::thread left([&] {
    MatchRange(src_key, mid, src_val, dst_key, dst_key_end, dst_val, k, op, grainsize, &left_n);
  });
  MatchRange(mid, src_key_end, src_val + (mid - src_key) * k, dst_key, dst_key_end, dst_val, k, op,
             grainsize, &right_n);
  left.join();
  *matched += left_n + right_n;
}

}  // namespace match_detail

/*!
 * \brief dst_val[key] <op>= src_val[key] for every key in both (sorted, unique) lists.
 * \param k values per key
 * \return number of matched keys
 */
template <typename K, typename V>
size_t ParallelOrderedMatch(const SArray<K>& src_key, const SArray<V>& src_val,
                            const SArray<K>& dst_key, SArray<V>* dst_val, int k = 1,
                            AssignOp op = ASSIGN, int num_threads = 2) {
  CHECK_GT(num_threads, 0);
  CHECK_EQ(src_key.size() * static_cast<size_t>(k), src_val.size());
  CHECK_NOTNULL(dst_val)->resize(dst_key.size() * static_cast<size_t>(k));
  if (dst_key.empty()) return 0;
  const size_t grain = std::max<size_t>((src_key.size() + dst_key.size()) / num_threads + 5,
                                        static_cast<size_t>(1024));
  size_t matched = 0;
  match_detail::MatchRange(src_key.begin(), src_key.end(), src_val.begin(), dst_key.begin(),
                           dst_key.end(), dst_val->begin(), k, op, grain, &matched);
  return matched;
}

}  // namespace ps
#endif  // PS_INTERNAL_PARALLEL_KV_MATCH_H_
