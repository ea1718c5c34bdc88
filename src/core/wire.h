/**
 * \file wire.h
 * \brief Binary codec for Meta (control + data descriptors).
 *
 * Self-describing little-endian field stream, NOT a struct dump: the reference
 * memcpy's a POD RawMeta/RawNode (src/meta.h:12-96, src/van.cc:689-831), which
 * bakes compiler padding and 32-bit byte counts into the protocol and ships 552
 * bytes per node. Here a Node costs ~40 bytes + hostname, byte counts are 64-bit,
 * and every frame starts with a magic/version word so a mismatched peer fails
 * loudly instead of mis-parsing.
 */
#ifndef PS_CORE_WIRE_H_
#define PS_CORE_WIRE_H_
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "ps/internal/message.h"

namespace ps {
namespace wire {

static const uint32_t kMetaMagic = 0x32425350u;  // "PSB2"
static const uint16_t kMetaVersion = 1;

/*!
 * \brief append-only little-endian writer over a byte vector. The vector is grown in large steps
 *        and written through a cursor (a resize per field costs a capacity check plus a zero fill:
 *        ~30 fields made PackMeta half a microsecond per message); Finish() trims it to size.
 */
class Writer {
 public:
  explicit Writer(std::vector<char>* out) : out_(out), pos_(out->size()) {}
  /*! \brief write from offset `at` of a vector the caller has already sized generously */
  Writer(std::vector<char>* out, size_t at) : out_(out), pos_(at) {}
  ~Writer() { Finish(); }
  template <typename T>
  void Put(T v) {
    static_assert(std::is_trivially_copyable<T>::value, "POD only");
    Ensure(sizeof(T));
    memcpy(out_->data() + pos_, &v, sizeof(T));
    pos_ += sizeof(T);
  }
  void PutBytes(const void* p, size_t n) {
    Ensure(n);
    if (n) memcpy(out_->data() + pos_, p, n);
    pos_ += n;
  }
  void PutString(const std::string& s) {
    Put<uint32_t>(static_cast<uint32_t>(s.size()));
    PutBytes(s.data(), s.size());
  }
  /*! \brief make the vector exactly as long as what was written */
  void Finish() { out_->resize(pos_); }

 private:
  void Ensure(size_t n) {
    if (pos_ + n > out_->size()) out_->resize(std::max(out_->size() * 2, pos_ + n + 256));
  }
  std::vector<char>* out_;
  size_t pos_;
};

/*! \brief bounds-checked reader; ok() turns false on the first short read */
class Reader {
 public:
  Reader(const char* p, size_t n) : p_(p), end_(p + n) {}
  template <typename T>
  T Get() {
    T v{};
    if (static_cast<size_t>(end_ - p_) < sizeof(T)) {
      ok_ = false;
      p_ = end_;
      return v;
    }
    memcpy(&v, p_, sizeof(T));
    p_ += sizeof(T);
    return v;
  }
  bool GetBytes(void* dst, size_t n) {
    if (static_cast<size_t>(end_ - p_) < n) {
      ok_ = false;
      p_ = end_;
      return false;
    }
    if (n) memcpy(dst, p_, n);
    p_ += n;
    return true;
  }
  std::string GetString() {
    uint32_t n = Get<uint32_t>();
    if (!ok_ || static_cast<size_t>(end_ - p_) < n) {
      ok_ = false;
      return std::string();
    }
    std::string s(p_, n);
    p_ += n;
    return s;
  }
  bool ok() const { return ok_; }
  size_t remaining() const { return static_cast<size_t>(end_ - p_); }

 private:
  const char* p_;
  const char* end_;
  bool ok_ = true;
};

/*! \brief serialise `meta` (everything except sender/recver, which the transport frames) */
void PackMeta(const Meta& meta, std::vector<char>* out);
/*! \brief inverse of PackMeta; returns false on a malformed buffer */
bool UnpackMeta(const char* buf, size_t len, Meta* meta);
/*! \brief exact size PackMeta will produce */
size_t PackedMetaSize(const Meta& meta);

void PackNode(const Node& n, Writer* w);
bool UnpackNode(Reader* r, Node* n);

}  // namespace wire
}  // namespace ps
#endif  // PS_CORE_WIRE_H_
