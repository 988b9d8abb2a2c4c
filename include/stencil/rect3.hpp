#pragma once
// Rect3 -- an axis-aligned box of grid points, closed at `lo`, open at `hi`.
// The two public Dim3 members in this order are part of the ABI: user kernels take Rect3 by value
// (reference bin/jacobi3d.cu:40-43), and drivers add offsets to lo / hi directly (astaroth/astaroth.cu:563-566).

#include <ostream>

#include "stencil/dim3.hpp"

class Rect3 {
public:
  Dim3 lo; // first point inside
  Dim3 hi; // first point outside

  Rect3() = default;
  Rect3(const Dim3 &first, const Dim3 &past_last) : lo(first), hi(past_last) {}

  // points per axis
  Dim3 extent() const noexcept { return hi - lo; }
};

inline std::ostream &operator<<(std::ostream &out, const Rect3 &box) {
  out << box.lo << "..<" << box.hi;
  return out;
}
