#pragma once
// Small integer helpers of the stencil API (API-compatible with the reference's
// include/stencil/numeric.hpp; implementation in src/numeric.cpp).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <vector>

// smallest power of two >= x (x >= 1)
inline int64_t nextPowerOfTwo(int64_t x) {
  int64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// prime factors of n, largest first (empty for n == 0; {} for n == 1)
template <typename T> std::vector<T> prime_factors(T n);

inline int64_t div_ceil(int64_t n, int64_t d) { return (n + d - 1) / d; }

template <typename T> T get_max_abs_error(const T *a, const T *b, const size_t n) {
  T worst = std::numeric_limits<T>::lowest();
  for (size_t i = 0; i < n; ++i) {
    const T e = std::abs(a[i] - b[i]);
    if (e > worst) worst = e;
  }
  return worst;
}
