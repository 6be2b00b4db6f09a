// qureg_impl.hpp -- helpers shared by the QubitRegister translation units (not installed).
#pragma once
#include <stdexcept>
#include <string>

#include "../include/qureg.hpp"
#include "iqsb.h"

namespace iqs {
namespace detail {

// C ABI status -> C++ exception (CUDA/NCCL status -> int codes -> std::runtime_error)
inline void Check(int rc, const char *what) {
  if (rc != IQSB_OK) throw std::runtime_error(std::string("iqs (B200 engine): ") + what + ": " + iqsb_last_error());
}

template <class Type>
inline void M8(TM2x2<Type> const &m, double out[8]) {
  out[0] = m(0, 0).real(); out[1] = m(0, 0).imag();
  out[2] = m(0, 1).real(); out[3] = m(0, 1).imag();
  out[4] = m(1, 0).real(); out[5] = m(1, 0).imag();
  out[6] = m(1, 1).real(); out[7] = m(1, 1).imag();
}

template <class Type> struct DType;
template <> struct DType<ComplexDP> { static const int value = IQSB_F64; };
template <> struct DType<ComplexSP> { static const int value = IQSB_F32; };

inline bool IsOne(double re, double im) { return re == 1.0 && im == 0.0; }

}  // namespace detail
}  // namespace iqs
