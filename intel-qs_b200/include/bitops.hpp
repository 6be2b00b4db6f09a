// bitops.hpp -- integer helpers (interface of reference include/bitops.hpp:28-137).
#ifndef IQS_BITOPS_HPP
#define IQS_BITOPS_HPP
#include <cassert>
#include <cstdint>
namespace iqs {

template <class Integral>
unsigned floor_power_of_two(Integral x) {
  unsigned p = 1;
  while (x >>= 1) p <<= 1;
  return p;
}

template <class Integral>
constexpr unsigned highestBit(Integral i) {
  return i <= 1 ? 0u : 1u + highestBit(i >> 1);
}

// exact log2; asserts when n is not a power of two
template <class Integral>
unsigned int ilog2(Integral n) {
  for (unsigned w = 0; w < 8 * sizeof(Integral); ++w)
    if ((static_cast<Integral>(1) << w) == n) return w;
  assert(false && "ilog2: not a power of two");
  return 0;
}

template <class Integral>
inline constexpr bool isPowerOf2(Integral i) {
  return i > 0 && (i & (i - 1)) == 0;
}

inline long popcnt(uint32_t x) { return __builtin_popcount(x); }
inline long popcnt(uint64_t x) { return __builtin_popcountll(x); }

}  // namespace iqs
#endif
