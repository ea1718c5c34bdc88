/**
 * \file assign_op.h
 * \brief Element-wise assignment operators used when merging key-value lists.
 * Parity: reference include/ps/internal/assign_op.h:12-68 (which no longer compiles:
 * it references undeclared `right`/`left`). This is a working version.
 */
#ifndef PS_INTERNAL_ASSIGN_OP_H_
#define PS_INTERNAL_ASSIGN_OP_H_
#include "ps/internal/utils.h"

namespace ps {

enum AssignOp { ASSIGN, PLUS, MINUS, TIMES, DIVIDE, AND, OR, XOR };

namespace assign_detail {
template <typename T, bool IsInt>
struct Bitwise {
  static void Apply(AssignOp, const T&, T*) { LOG(FATAL) << "bitwise op on a non-integral type"; }
};
template <typename T>
struct Bitwise<T, true> {
  static void Apply(AssignOp op, const T& rhs, T* lhs) {
    if (op == AND) *lhs &= rhs;
    else if (op == OR) *lhs |= rhs;
    else *lhs ^= rhs;
  }
};
}  // namespace assign_detail

/*! \brief *lhs = *lhs <op> rhs */
template <typename T>
inline void AssignFunc(const T& rhs, AssignOp op, T* lhs) {
  switch (op) {
    case ASSIGN: *lhs = rhs; break;
    case PLUS: *lhs += rhs; break;
    case MINUS: *lhs -= rhs; break;
    case TIMES: *lhs *= rhs; break;
    case DIVIDE: *lhs /= rhs; break;
    default: assign_detail::Bitwise<T, std::is_integral<T>::value>::Apply(op, rhs, lhs);
  }
}

}  // namespace ps
#endif  // PS_INTERNAL_ASSIGN_OP_H_
