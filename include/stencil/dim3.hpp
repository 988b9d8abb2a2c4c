#pragma once
// Dim3: a signed 64-bit 3-vector used for sizes, positions, directions and subdomain indices.
// Layout {x, y, z} of int64_t is ABI: it is passed by value into user kernels inside Accessor and
// Rect3 (reference include/stencil/dim3.hpp:17-22).

#include "stencil/numeric.hpp"

#include <algorithm>
#include <cassert>
#include <iostream>
#include <ostream>

#ifdef __CUDACC__
#define STENCIL_HD __host__ __device__
#else
#define STENCIL_HD
#endif

class Dim3 {
public:
  int64_t x;
  int64_t y;
  int64_t z;

  Dim3() = default;
  Dim3(const Dim3 &) = default;
  Dim3(Dim3 &&) = default;
  Dim3 &operator=(const Dim3 &) = default;
  Dim3 &operator=(Dim3 &&) = default;
  STENCIL_HD Dim3(int64_t x_, int64_t y_, int64_t z_) : x(x_), y(y_), z(z_) {}

#ifdef __CUDACC__
  STENCIL_HD Dim3(const dim3 d) : x(d.x), y(d.y), z(d.z) {}
  STENCIL_HD operator dim3() const {
    assert(x > 0 && y > 0 && z > 0);
    return dim3((unsigned int)x, (unsigned int)y, (unsigned int)z);
  }
#endif

  void swap(Dim3 &o) {
    std::swap(x, o.x);
    std::swap(y, o.y);
    std::swap(z, o.z);
  }

  STENCIL_HD int64_t &operator[](const size_t i) {
    assert(i < 3 && "only 3 dimensions!");
    return i == 0 ? x : (i == 1 ? y : z);
  }
  STENCIL_HD const int64_t &operator[](const size_t i) const {
    assert(i < 3 && "only 3 dimensions!");
    return i == 0 ? x : (i == 1 ? y : z);
  }

  // element-wise max
  Dim3 max(const Dim3 &o) const { return Dim3(std::max(x, o.x), std::max(y, o.y), std::max(z, o.z)); }

  STENCIL_HD bool any() const { return x != 0 || y != 0 || z != 0; }
  STENCIL_HD bool all() const { return x != 0 && y != 0 && z != 0; }
  STENCIL_HD size_t flatten() const { return x * y * z; }

  // lexicographic (x, then y, then z): the ordering of messages and std::map keys
  STENCIL_HD bool operator<(const Dim3 &r) const {
    if (x != r.x) return x < r.x;
    if (y != r.y) return y < r.y;
    return z < r.z;
  }
  STENCIL_HD bool operator==(const Dim3 &r) const { return x == r.x && y == r.y && z == r.z; }
  STENCIL_HD bool operator!=(const Dim3 &r) const { return !(*this == r); }

  STENCIL_HD bool all_lt(const int64_t r) const { return x < r && y < r && z < r; }
  STENCIL_HD bool all_lt(const Dim3 r) const { return x < r.x && y < r.y && z < r.z; }
  STENCIL_HD bool all_gt(const int64_t r) const { return x > r && y > r && z > r; }
  STENCIL_HD bool all_ge(const int64_t r) const { return x >= r && y >= r && z >= r; }
  STENCIL_HD bool any_lt(const int64_t r) const { return x < r || y < r || z < r; }
  STENCIL_HD bool any_gt(const int64_t r) const { return x > r || y > r || z > r; }

#define STENCIL_DIM3_OP(OP)                                                                                            \
  STENCIL_HD Dim3 &operator OP##=(const Dim3 &r) {                                                                     \
    x OP## = r.x;                                                                                                      \
    y OP## = r.y;                                                                                                      \
    z OP## = r.z;                                                                                                      \
    return *this;                                                                                                      \
  }                                                                                                                    \
  STENCIL_HD Dim3 operator OP(const Dim3 &r) const {                                                                   \
    Dim3 t(*this);                                                                                                     \
    t OP## = r;                                                                                                        \
    return t;                                                                                                          \
  }
  STENCIL_DIM3_OP(+)
  STENCIL_DIM3_OP(-)
  STENCIL_DIM3_OP(*)
  STENCIL_DIM3_OP(/)
  STENCIL_DIM3_OP(%)
#undef STENCIL_DIM3_OP

  STENCIL_HD Dim3 operator-(int64_t r) const { return Dim3(x - r, y - r, z - r); }
  STENCIL_HD Dim3 operator*(int64_t r) const { return Dim3(x * r, y * r, z * r); }
  STENCIL_HD Dim3 &operator*=(const double &r) {
    x = int64_t(x * r);
    y = int64_t(y * r);
    z = int64_t(z * r);
    return *this;
  }
  STENCIL_HD Dim3 &operator/=(const double &r) {
    x = int64_t(x / r);
    y = int64_t(y / r);
    z = int64_t(z / r);
    return *this;
  }

  // periodic wrap of each component into [0, lims) -- modifies and returns *this
  STENCIL_HD Dim3 wrap(const Dim3 &lims) {
    for (int i = 0; i < 3; ++i) {
      int64_t &c = (*this)[i];
      const int64_t l = lims[i];
      c %= l;
      if (c < 0) c += l;
    }
    return *this;
  }

  // shape `threads` threads like `extent`: x first, then y, then z; power-of-two sides,
  // capped by the CUDA block limits (1024, 1024, 64)
  static Dim3 make_block_dim(const Dim3 extent, int64_t threads) {
    assert(extent.x >= 0 && extent.y >= 0 && extent.z >= 0);
    threads = std::min<int64_t>(threads, 1024);
    Dim3 b;
    b.x = std::min(threads, nextPowerOfTwo(extent.x));
    threads /= b.x;
    b.y = std::min(threads, nextPowerOfTwo(extent.y));
    threads /= b.y;
    b.z = std::min(threads, nextPowerOfTwo(extent.z));
    b.x = std::min<int64_t>(b.x, 1024);
    b.y = std::min<int64_t>(b.y, 1024);
    b.z = std::min<int64_t>(b.z, 64);
    assert(b.x * b.y * b.z <= 1024);
    return b;
  }
};

inline std::ostream &operator<<(std::ostream &os, const Dim3 &d) { return os << '[' << d.x << ',' << d.y << ',' << d.z << ']'; }

#undef STENCIL_HD
