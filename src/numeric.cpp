#include "stencil/numeric.hpp"

#include <algorithm>
#include <functional>

// trial division; factors returned largest first (the partitioners split by the big factors first)
template <typename T> std::vector<T> prime_factors(T n) {
  std::vector<T> f;
  if (n == 0) return f;
  for (T p = 2; p * p <= n; p += (p == 2 ? 1 : 2)) {
    while (n % p == 0) {
      f.push_back(p);
      n /= p;
    }
  }
  if (n > 1) f.push_back(n);
  std::sort(f.begin(), f.end(), std::greater<T>());
  return f;
}

template std::vector<int> prime_factors(int n);
template std::vector<int64_t> prime_factors(int64_t n);
