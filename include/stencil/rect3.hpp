#pragma once
// Rect3: half-open box [lo, hi) of grid coordinates.  Layout {Dim3 lo, Dim3 hi} is ABI (passed by
// value into user kernels, reference bin/jacobi3d.cu:40-43).

#include <ostream>

#include "stencil/dim3.hpp"

class Rect3 {
public:
  Dim3 lo;
  Dim3 hi;

  Rect3() {}
  Rect3(const Dim3 &lo_, const Dim3 &hi_) : lo(lo_), hi(hi_) {}

  Dim3 extent() const noexcept { return hi - lo; }
};

inline std::ostream &operator<<(std::ostream &os, const Rect3 &r) { return os << r.lo << "..<" << r.hi; }
