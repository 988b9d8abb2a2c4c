#pragma once
// Radius: the stencil reach in each of the 27 directions, independently settable
// (reference include/stencil/radius.hpp).  x(d)/y(d)/z(d) read the face entries.

#include <cassert>
#include <cstdint>
#include <cstdlib>

#include "stencil/dim3.hpp"
#include "stencil/direction_map.hpp"

class Radius {
  DirectionMap<size_t> r_;

  void set_where(int nonzeros, size_t v) {
    for (int z = -1; z <= 1; ++z)
      for (int y = -1; y <= 1; ++y)
        for (int x = -1; x <= 1; ++x)
          if ((x != 0) + (y != 0) + (z != 0) == nonzeros) r_.at_dir(x, y, z) = v;
  }

public:
  size_t &dir(int x, int y, int z) noexcept { return r_.at_dir(x, y, z); }
  const size_t &dir(int x, int y, int z) const noexcept { return r_.at_dir(x, y, z); }
  size_t &dir(const Dim3 &d) noexcept { return r_.at_dir(int(d.x), int(d.y), int(d.z)); }
  const size_t &dir(const Dim3 &d) const noexcept { return r_.at_dir(int(d.x), int(d.y), int(d.z)); }

  const size_t &x(int d) const noexcept { return dir(d, 0, 0); }
  const size_t &y(int d) const noexcept { return dir(0, d, 0); }
  const size_t &z(int d) const noexcept { return dir(0, 0, d); }

  bool operator==(const Radius &o) const noexcept { return r_ == o.r_; }

  void set_face(const size_t r) { set_where(1, r); }
  void set_edge(const size_t r) { set_where(2, r); }
  void set_corner(const size_t r) { set_where(3, r); }

  // same radius in all 27 slots (centre included)
  static Radius constant(const size_t r) {
    Radius out;
    for (int n = 0; n <= 3; ++n) out.set_where(n, r);
    return out;
  }

  // faces / edges / corners; centre 0
  static Radius face_edge_corner(const size_t face, const size_t edge, const size_t corner) {
    Radius out;
    out.set_where(0, 0);
    out.set_face(face);
    out.set_edge(edge);
    out.set_corner(corner);
    return out;
  }

  // the 27 values in storage order [z+1][y+1][x+1]
  const size_t *data() const noexcept { return r_.data(); }
};
