// utils.hpp -- basic aliases of the Intel-QS API (interface of reference include/utils.hpp:17-52),
// re-authored for the B200 engine.
#ifndef IQS_UTILS_HPP
#define IQS_UTILS_HPP

#include <complex>
#include <cstddef>

#define UL(x) ((std::size_t)(x))
#define sec() time_in_seconds()
#ifndef TODO
#define TODO(x)
#endif
#ifndef INFO
#define INFO(x)
#endif

using ComplexSP = std::complex<float>;
using ComplexDP = std::complex<double>;

namespace iqs {

// extract_value_type<std::complex<double>>::value_type == double
template <typename T>
struct extract_value_type {
  typedef T value_type;
};
template <template <typename> class X, typename T>
struct extract_value_type<X<T>> {
  typedef T value_type;
};

double time_in_seconds(void);
void WhatCompileDefinitions();

}  // namespace iqs
#endif
