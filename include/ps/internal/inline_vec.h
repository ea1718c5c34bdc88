/**
 * \file inline_vec.h
 * \brief InlineVec<T, N>: the few-element sequences of a Message (its 2-3 payload segments, their
 *        data types) without a heap allocation per message.
 *
 * A push + pull used to cost 23 (server) / 30 (worker) mallocs, a third of them for the growth of
 * the std::vectors inside Meta and Message. The first N elements live inside the object; anything
 * beyond spills into a std::vector (a message with more than N segments still works).
 * Only what the code base uses of the vector interface is provided.
 */
#ifndef PS_INTERNAL_INLINE_VEC_H_
#define PS_INTERNAL_INLINE_VEC_H_
#include <array>
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace ps {

template <typename T, size_t N>
class InlineVec {
 public:
  template <typename V, typename Ref>
  class Iter {
   public:
    Iter(V* v, size_t i) : v_(v), i_(i) {}
    Ref operator*() const { return (*v_)[i_]; }
    Iter& operator++() {
      ++i_;
      return *this;
    }
    bool operator!=(const Iter& o) const { return i_ != o.i_; }
    bool operator==(const Iter& o) const { return i_ == o.i_; }

   private:
    V* v_;
    size_t i_;
  };
  using iterator = Iter<InlineVec, T&>;
  using const_iterator = Iter<const InlineVec, const T&>;
  using value_type = T;

  InlineVec() {}
  InlineVec(std::initializer_list<T> init) {
    for (const T& v : init) push_back(v);
  }
  InlineVec& operator=(std::initializer_list<T> init) {
    clear();
    for (const T& v : init) push_back(v);
    return *this;
  }
  bool operator==(const InlineVec& o) const {
    if (n_ != o.n_) return false;
    for (size_t i = 0; i < n_; ++i) {
      if (!((*this)[i] == o[i])) return false;
    }
    return true;
  }
  bool operator!=(const InlineVec& o) const { return !(*this == o); }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  void reserve(size_t) {}
  void clear() {
    for (size_t i = 0; i < n_ && i < N; ++i) inl_[i] = T();  // drop references held by the elements
    over_.clear();
    n_ = 0;
  }
  void push_back(const T& v) {
    if (n_ < N) inl_[n_] = v; else over_.push_back(v);
    ++n_;
  }
  void push_back(T&& v) {
    if (n_ < N) inl_[n_] = std::move(v); else over_.push_back(std::move(v));
    ++n_;
  }
  void resize(size_t n) {
    for (size_t i = n; i < n_ && i < N; ++i) inl_[i] = T();
    over_.resize(n > N ? n - N : 0);
    n_ = n;
  }
  T& operator[](size_t i) { return i < N ? inl_[i] : over_[i - N]; }
  const T& operator[](size_t i) const { return i < N ? inl_[i] : over_[i - N]; }
  T& back() { return (*this)[n_ - 1]; }
  const T& back() const { return (*this)[n_ - 1]; }
  iterator begin() { return iterator(this, 0); }
  iterator end() { return iterator(this, n_); }
  const_iterator begin() const { return const_iterator(this, 0); }
  const_iterator end() const { return const_iterator(this, n_); }

 private:
  size_t n_ = 0;
  std::array<T, N> inl_{};
  std::vector<T> over_;
};

}  // namespace ps
#endif  // PS_INTERNAL_INLINE_VEC_H_
