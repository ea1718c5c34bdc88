/**
 * \file sarray.h
 * \brief SArray<V>: a reference-counted, zero-copy view over a contiguous buffer
 *        that may live in host memory or in B200 HBM.
 *
 * Design (B200-first): the array is {owner handle, element pointer, length,
 * capacity, placement}. The owner handle is an aliasing std::shared_ptr, so views
 * of different element types share one control block, a segment is a pointer
 * bump, and the deleter can be anything (delete[], cudaFree, arena release, a
 * torch::Tensor keep-alive). `placement` says where the bytes are *now*
 * (src_device_*) and where the consumer wants them to land (dst_device_*); the
 * NVLink van keys its routing on it. Element access (operator[], DebugString of
 * the contents, resize-fill) is only defined for host-resident arrays.
 *
 * API parity: reference include/ps/sarray.h:46-324 (ctor set, CopyFrom, reset,
 * resize/reserve, segment, public device fields :320-323, FindRange :344-350).
 */
#ifndef PS_SARRAY_H_
#define PS_SARRAY_H_
#include <algorithm>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "ps/internal/utils.h"
#include "ps/range.h"

namespace ps {

/*! \brief where a buffer lives */
enum DeviceType { UNK, CPU, GPU };
static const char* const DeviceTypeName[] = {"UNK", "CPU", "GPU"};

template <typename V>
class SArray {
 public:
  using Deleter = std::function<void(V*)>;

  /*! \brief empty array */
  SArray() {}
  ~SArray() {}

  /*! \brief host array of `size` elements, each set to `val` */
  explicit SArray(size_t size, V val = 0) { resize(size, val); }

  /*! \brief zero-copy reinterpretation of another element type */
  template <typename W>
  explicit SArray(const SArray<W>& other) {
    *this = other;
  }
  template <typename W>
  void operator=(const SArray<W>& other) {
    const size_t bytes = other.size() * sizeof(W);
    CHECK_EQ(bytes % sizeof(V), (size_t)0) << "cannot view " << bytes << " bytes as "
                                           << sizeof(V) << "-byte elements";
    size_ = bytes / sizeof(V);
    capacity_ = other.capacity() * sizeof(W) / sizeof(V);
    // aliasing ctor: share the control block, point at the same bytes
    ptr_ = std::shared_ptr<V>(other.ptr(), reinterpret_cast<V*>(other.data()));
    CopyPlacement(other);
  }

  /*! \brief wrap a raw host pointer; freed with delete[] iff deletable */
  SArray(V* data, size_t size, bool deletable = false) {
    Adopt(data, size, deletable);
  }
  /*! \brief wrap a raw pointer with explicit placement (host or device memory) */
  SArray(V* data, size_t size, DeviceType src_device_type, int src_device_id,
         DeviceType dst_device_type, int dst_device_id, bool deletable = false) {
    Adopt(data, size, deletable);
    src_device_type_ = src_device_type;
    src_device_id_ = src_device_id;
    dst_device_type_ = dst_device_type;
    dst_device_id_ = dst_device_id;
  }

  /*! \brief deep copies (host only) */
  void CopyFrom(const V* data, size_t size) {
    resize(size);
    if (size) memcpy(this->data(), data, size * sizeof(V));
  }
  void CopyFrom(const SArray<V>& other) {
    if (this == &other) return;
    CopyFrom(other.data(), other.size());
  }
  template <typename ForwardIt>
  void CopyFrom(const ForwardIt& first, const ForwardIt& last) {
    size_t n = static_cast<size_t>(std::distance(first, last));
    V* buf = new V[n + 1];
    reset(buf, n, [](V* p) { delete[] p; });
    V* out = buf;
    for (auto it = first; it != last; ++it) *out++ = *it;
  }

  explicit SArray(const std::vector<V>& vec) { CopyFrom(vec.data(), vec.size()); }
  /*! \brief zero-copy view of a shared vector */
  explicit SArray(const std::shared_ptr<std::vector<V>>& vec) {
    ptr_ = std::shared_ptr<V>(vec, vec->data());
    size_ = capacity_ = vec->size();
  }
  template <typename W>
  SArray(const std::initializer_list<W>& list) {
    CopyFrom(list.begin(), list.end());
  }
  template <typename W>
  void operator=(const std::initializer_list<W>& list) {
    CopyFrom(list.begin(), list.end());
  }

  /*! \brief take over `data` with a custom deleter and placement */
  template <typename Del>
  void reset(V* data, size_t size, Del del, DeviceType src_device_type = CPU,
             int src_device_id = 0, DeviceType dst_device_type = CPU, int dst_device_id = 0) {
    size_ = capacity_ = size;
    ptr_.reset(data, del);
    src_device_type_ = src_device_type;
    src_device_id_ = src_device_id;
    dst_device_type_ = dst_device_type;
    dst_device_id_ = dst_device_id;
  }

  /*!
   * \brief host array whose control block and storage are ONE allocation (std::allocate_shared
   *        on a buffer struct + aliasing pointer). For the short-lived small segments of the
   *        receive path (a key, a length): half the malloc / free traffic of reset(new V[n]).
   */
  static SArray<V> Compact(size_t size) {
    struct Block {
      explicit Block(size_t n) : data(new V[n]) {}
      ~Block() { delete[] data; }
      V* data;
    };
    SArray<V> a;
    if (size == 0) return a;
    if (size * sizeof(V) <= kInline) {
      struct Small {
        alignas(16) unsigned char bytes[kInline];
      };
      auto blk = std::make_shared<Small>();
      a.ptr_ = std::shared_ptr<V>(blk, reinterpret_cast<V*>(blk->bytes));
    } else {
      auto blk = std::make_shared<Block>(size);
      a.ptr_ = std::shared_ptr<V>(blk, blk->data);
    }
    a.size_ = a.capacity_ = size;
    return a;
  }

  /*!
   * \brief host resize. Within capacity this only moves the end; growing
   *        reallocates, copies, and fills the new tail with `val`.
   */
  void resize(size_t size, V val = 0) {
    const size_t old = size_;
    if (size > capacity_) {
      CHECK(src_device_type_ != GPU) << "resize of a device-resident SArray";
      V* buf = new V[size + kSlack];
      if (old) memcpy(buf, data(), old * sizeof(V));
      reset(buf, size, [](V* p) { delete[] p; }, src_device_type_, src_device_id_,
            dst_device_type_, dst_device_id_);
      capacity_ = size + kSlack;
    }
    size_ = size;
    if (size > old) std::fill(data() + old, data() + size, val);
  }
  /*! \brief grow capacity without changing size */
  void reserve(size_t cap) {
    if (cap <= capacity_) return;
    size_t n = size_;
    resize(cap);
    size_ = n;
  }
  void clear() {
    ptr_.reset();
    size_ = capacity_ = 0;
  }

  bool empty() const { return size_ == 0; }
  size_t size() const { return size_; }
  size_t capacity() const { return capacity_; }
  /*! \brief payload size in bytes */
  size_t bytes() const { return size_ * sizeof(V); }

  V* begin() { return data(); }
  const V* begin() const { return data(); }
  V* end() { return data() + size_; }
  const V* end() const { return data() + size_; }
  V* data() const { return ptr_.get(); }
  std::shared_ptr<V>& ptr() { return ptr_; }
  const std::shared_ptr<V>& ptr() const { return ptr_; }

  V back() const {
    CHECK(!empty());
    return data()[size_ - 1];
  }
  V front() const {
    CHECK(!empty());
    return data()[0];
  }
  V& operator[](size_t i) { return data()[i]; }
  const V& operator[](size_t i) const { return data()[i]; }

  void push_back(const V& val) {
    if (size_ == capacity_) reserve(size_ * 2 + kSlack);
    data()[size_++] = val;
  }
  void pop_back() {
    if (size_) --size_;
  }
  void append(const SArray<V>& tail) {
    if (tail.empty()) return;
    const size_t old = size_, add = tail.size();
    // tail may alias *this, so capture its pointer after the resize
    SArray<V> keep = tail;
    reserve(old + add);
    size_ = old + add;
    memcpy(data() + old, keep.data(), add * sizeof(V));
  }

  /*! \brief zero-copy view of [begin, end); shares ownership with *this */
  SArray<V> segment(size_t begin, size_t end) const {
    CHECK_GE(end, begin);
    CHECK_LE(end, size_);
    SArray<V> out;
    out.ptr_ = std::shared_ptr<V>(ptr_, data() + begin);
    out.size_ = end - begin;
    out.capacity_ = end - begin;
    out.CopyPlacement(*this);
    return out;
  }

  /*! \brief true if the bytes currently live in GPU memory */
  bool on_gpu() const { return src_device_type_ == GPU; }

  std::string DebugString() const {
    std::ostringstream os;
    os << "SArray{ptr=" << static_cast<const void*>(data()) << ", n=" << size_
       << ", " << DeviceTypeName[src_device_type_] << "[" << src_device_id_ << "]->"
       << DeviceTypeName[dst_device_type_] << "[" << dst_device_id_ << "]}";
    return os.str();
  }

  /*! \brief placement (public for parity with the reference's fields) */
  DeviceType src_device_type_ = CPU;
  int src_device_id_ = 0;
  DeviceType dst_device_type_ = CPU;
  int dst_device_id_ = 0;

  template <typename W>
  void CopyPlacement(const SArray<W>& o) {
    src_device_type_ = o.src_device_type_;
    src_device_id_ = o.src_device_id_;
    dst_device_type_ = o.dst_device_type_;
    dst_device_id_ = o.dst_device_id_;
  }

 private:
  template <typename W>
  friend class SArray;
  // growth slack so that push_back on a just-sized array does not reallocate
  static constexpr size_t kSlack = 4;

  void Adopt(V* data, size_t size, bool deletable) {
    if (deletable) {
      reset(data, size, [](V* p) { delete[] p; });
    } else {
      reset(data, size, [](V*) {});
    }
  }

  size_t size_ = 0;
  size_t capacity_ = 0;
  static constexpr size_t kInline = 64;  // bytes stored inside the control block by Compact()
  std::shared_ptr<V> ptr_;
};

/*!
 * \brief index range of the entries of a sorted array that fall in [lower, upper)
 *  e.g. FindRange({1,3,5,7,9}, 2, 7) == Range(1,3)
 */
template <typename V>
Range FindRange(const SArray<V>& arr, V lower, V upper) {
  if (upper <= lower) return Range(0, 0);
  const V* b = arr.begin();
  const V* e = arr.end();
  return Range(std::lower_bound(b, e, lower) - b, std::lower_bound(b, e, upper) - b);
}

/*! \brief "[n]: a b c ... x y z" for host arrays */
template <typename V>
inline std::string DebugStr(const V* data, int n, int m = 5) {
  std::ostringstream os;
  os << "[" << n << "]: ";
  for (int i = 0; i < n; ++i) {
    if (n >= 2 * m && i == m) {
      os << "... ";
      i = n - m;
    }
    os << data[i] << " ";
  }
  return os.str();
}

template <typename V>
std::ostream& operator<<(std::ostream& os, const SArray<V>& a) {
  if (a.on_gpu()) return os << a.DebugString();
  return os << DebugStr(a.data(), static_cast<int>(a.size()));
}

}  // namespace ps
#endif  // PS_SARRAY_H_
