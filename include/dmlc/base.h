/**
 * \file base.h
 * \brief Tiny portability header kept for include-compatibility with user code
 *        written against the reference (`#include "dmlc/base.h"`).
 *        Parity: reference include/dmlc/base.h (feature macros only).
 */
#ifndef DMLC_BASE_H_
#define DMLC_BASE_H_

#ifndef DISALLOW_COPY_AND_ASSIGN
#define DISALLOW_COPY_AND_ASSIGN(T) \
  T(T const&) = delete;             \
  T(T&&) = delete;                  \
  T& operator=(T const&) = delete;  \
  T& operator=(T&&) = delete
#endif

#define DMLC_USE_CXX11 1
#define DMLC_THROW_EXCEPTION noexcept(false)

#endif  // DMLC_BASE_H_
