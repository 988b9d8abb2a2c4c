#pragma once
// Halo geometry shared by LocalDomain, DistributedDomain, the packers and the C ABI
// (include/stencil_b200.h).  Host-only integer math; positions are allocation-relative elements.

#include <vector>

#include "stencil/dim3.hpp"
#include "stencil/radius.hpp"
#include "stencil/rect3.hpp"

namespace stencil {
namespace geom {

// Where the region on the `dir` side starts.  halo=true: the ghost cells outside the compute
// region; halo=false: the outermost compute cells (what a send in `dir` reads).  A zero component
// means "the whole compute extent on that axis" and starts after the negative-side ghost cells.
// (semantics of the reference's LocalDomain::halo_pos, src/local_domain.cu:86-125)
inline Dim3 halo_pos(const Dim3 &dir, const Dim3 &sz, const Radius &radius, const bool halo) noexcept {
  assert(dir.all_gt(-2) && dir.all_lt(2));
  const int64_t lowGhost[3] = {int64_t(radius.x(-1)), int64_t(radius.y(-1)), int64_t(radius.z(-1))};
  Dim3 p;
  for (int a = 0; a < 3; ++a) {
    const int64_t d = dir[a];
    if (d > 0)
      p[a] = sz[a] + (halo ? lowGhost[a] : 0);
    else if (d < 0)
      p[a] = halo ? 0 : lowGhost[a];
    else
      p[a] = lowGhost[a];
  }
  return p;
}

// Size of the region on the `dir` side: the face radius on axes where dir != 0, the compute
// extent elsewhere (reference include/stencil/local_domain.cuh:212-222).
inline Dim3 halo_extent(const Dim3 &dir, const Dim3 &sz, const Radius &radius) noexcept {
  assert(dir.all_gt(-2) && dir.all_lt(2));
  return Dim3(dir.x == 0 ? sz.x : int64_t(radius.x(int(dir.x))), dir.y == 0 ? sz.y : int64_t(radius.y(int(dir.y))),
              dir.z == 0 ? sz.z : int64_t(radius.z(int(dir.z))));
}

inline Dim3 raw_size(const Dim3 &sz, const Radius &radius) noexcept {
  return Dim3(sz.x + radius.x(-1) + radius.x(1), sz.y + radius.y(-1) + radius.y(1), sz.z + radius.z(-1) + radius.z(1));
}

// Region of `compute` that no halo exchange can touch the inputs of: every face/edge/corner
// direction with a non-zero radius pulls the matching sides in (reference src/stencil.cu:878-921).
inline Rect3 interior(const Rect3 &compute, const Radius &radius) noexcept {
  Rect3 in = compute;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        if (0 == dx && 0 == dy && 0 == dz) continue;
        const int64_t r = int64_t(radius.dir(dx, dy, dz));
        const int d[3] = {dx, dy, dz};
        for (int a = 0; a < 3; ++a) {
          if (d[a] < 0)
            in.lo[a] = std::max(compute.lo[a] + r, in.lo[a]);
          else if (d[a] > 0)
            in.hi[a] = std::min(compute.hi[a] - r, in.hi[a]);
        }
      }
  return in;
}

// compute \ interior as disjoint slabs, peeled +x,+y,+z,-x,-y,-z, each against the box left by the
// previous peel (reference src/stencil.cu:927-977).
inline std::vector<Rect3> exterior(const Rect3 &compute, const Radius &radius) {
  const Rect3 in = interior(compute, radius);
  Rect3 box = compute;
  std::vector<Rect3> slabs;
  for (int a = 0; a < 3; ++a) {
    if (in.hi[a] != box.hi[a]) {
      Rect3 s = box;
      s.lo[a] = in.hi[a];
      slabs.push_back(s);
      box.hi[a] = in.hi[a];
    }
  }
  for (int a = 0; a < 3; ++a) {
    if (in.lo[a] != box.lo[a]) {
      Rect3 s = box;
      s.hi[a] = in.lo[a];
      slabs.push_back(s);
      box.lo[a] = in.lo[a];
    }
  }
  return slabs;
}

} // namespace geom
} // namespace stencil
