/**
 * \file parallel_sort.h
 * \brief Multi-threaded sort of an SArray: sort 2^k runs concurrently, then merge.
 * Parity: reference include/ps/internal/parallel_sort.h:23-55 (recursive
 * thread-per-split merge sort).
 */
#ifndef PS_INTERNAL_PARALLEL_SORT_H_
#define PS_INTERNAL_PARALLEL_SORT_H_
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>
#include "ps/sarray.h"

namespace ps {

namespace sort_detail {
template <typename T, class Fn>
void SortRange(T* data, size_t len, size_t grainsize, const Fn& cmp) {
  if (len <= grainsize) {
    std::sort(data, data + len, cmp);
    return;
  }
  const size_t half = len / 2;
  std::thread left(SortRange<T, Fn>, data, half, grainsize, std::cref(cmp));
  SortRange(data + half, len - half, grainsize, cmp);
  left.join();
  std::inplace_merge(data, data + half, data + len, cmp);
}
}  // namespace sort_detail

/*! \brief sort *arr with about `num_threads` threads */
template <typename T, class Fn>
void ParallelSort(SArray<T>* arr, int num_threads = 2, const Fn& cmp = std::less<T>()) {
  CHECK_GT(num_threads, 0);
  CHECK(arr);
  const size_t grain = std::max<size_t>(arr->size() / static_cast<size_t>(num_threads) + 5,
                                        static_cast<size_t>(1024 * 16));
  sort_detail::SortRange(arr->data(), arr->size(), grain, cmp);
}

}  // namespace ps
#endif  // PS_INTERNAL_PARALLEL_SORT_H_
