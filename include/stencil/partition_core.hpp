#pragma once
// RankPartition / NodePartition: cut a 3-D box into a grid of subdomains by the prime factors of
// the subdomain count.  Pure host integer math (no MPI, no CUDA); stencil/partition.hpp adds the
// Placement classes on top.  Behaviour matches the reference's include/stencil/partition.hpp:20-256
// (pinned by test/test_cpu_partition.cpp and tests/golden/ref_geometry.json).

#include <cassert>
#include <vector>

#include "stencil/dim3.hpp"
#include "stencil/numeric.hpp"
#include "stencil/radius.hpp"

namespace stencil {
namespace detail {

// A grid of dim subdomains of ceil-size `cell`; the first rem[a] indices along axis a keep the
// full cell, the rest are one shorter (rem = total % dim, 0 means "divides evenly").
struct CutGrid {
  Dim3 cell;
  Dim3 rem;

  Dim3 size_of(const Dim3 &idx) const {
    Dim3 s = cell;
    for (int a = 0; a < 3; ++a)
      if (rem[a] != 0 && idx[a] >= rem[a]) s[a] -= 1;
    return s;
  }
  Dim3 origin_of(const Dim3 &idx) const {
    Dim3 o = cell * idx;
    for (int a = 0; a < 3; ++a)
      if (rem[a] != 0 && idx[a] >= rem[a]) o[a] -= idx[a] - rem[a];
    return o;
  }
};

inline int64_t linearize(const Dim3 &idx, const Dim3 &dim) {
  assert(idx.all_ge(0) && idx.all_lt(dim));
  return idx.x + dim.x * (idx.y + dim.y * idx.z);
}

inline Dim3 dimensionize(int64_t i, const Dim3 &dim) {
  assert(i >= 0 && i < int64_t(dim.flatten()));
  const int64_t x = i % dim.x;
  i /= dim.x;
  return Dim3(x, i % dim.y, i / dim.y);
}

} // namespace detail
} // namespace stencil

// Split the LONGEST remaining axis by each prime factor of n (largest factor first).
class RankPartition {
  Dim3 dim_;
  stencil::detail::CutGrid grid_;

public:
  RankPartition(const Dim3 &size, const int64_t n) : dim_(1, 1, 1) {
    Dim3 cell = size;
    for (int64_t f : prime_factors(n)) {
      if (f < 2) continue;
      const int a = (cell.x >= cell.y && cell.x >= cell.z) ? 0 : (cell.y >= cell.z ? 1 : 2);
      cell[a] = div_ceil(cell[a], f);
      dim_[a] *= f;
    }
    grid_.cell = cell;
    grid_.rem = size % dim_;
  }
  RankPartition() : RankPartition(Dim3(0, 0, 0), 0) {}
  virtual ~RankPartition() {}

  virtual Dim3 dim() const { return dim_; }
  virtual Dim3 subdomain_size(const Dim3 &idx) const { return grid_.size_of(idx); }
  Dim3 subdomain_origin(const Dim3 &idx) const noexcept { return grid_.origin_of(idx); }

  size_t linearize(Dim3 idx) const { return size_t(stencil::detail::linearize(idx, dim())); }
  Dim3 dimensionize(int64_t i) { return stencil::detail::dimensionize(i, dim()); }
};

// Two-level split: first among nodes, then among the GPUs of a node; each prime factor cuts the
// axis whose cut plane carries the least halo traffic (area x (r+ + r-)), ties to x then y.
class NodePartition {
  Dim3 sysDim_;
  Dim3 nodeDim_;
  stencil::detail::CutGrid grid_;

  static void split(Dim3 &cell, Dim3 &dim, const Radius &radius, int64_t count) {
    for (int64_t f : prime_factors(count)) {
      if (f < 2) continue;
      const int64_t xi = cell.y * cell.z * int64_t(radius.dir(1, 0, 0) + radius.dir(-1, 0, 0));
      const int64_t yi = cell.x * cell.z * int64_t(radius.dir(0, 1, 0) + radius.dir(0, -1, 0));
      const int64_t zi = cell.x * cell.y * int64_t(radius.dir(0, 0, 1) + radius.dir(0, 0, -1));
      const int a = (xi <= yi && xi <= zi) ? 0 : (yi <= zi ? 1 : 2);
      cell[a] = div_ceil(cell[a], f);
      dim[a] *= f;
    }
  }

public:
  NodePartition(const Dim3 &size, const Radius radius, const int64_t nodes, const int64_t gpus)
      : sysDim_(1, 1, 1), nodeDim_(1, 1, 1) {
    Dim3 cell = size;
    split(cell, sysDim_, radius, nodes);
    split(cell, nodeDim_, radius, gpus);
    grid_.cell = cell;
    grid_.rem = size % (sysDim_ * nodeDim_);
  }
  NodePartition() : NodePartition(Dim3(0, 0, 0), Radius::constant(0), 0, 0) {}

  Dim3 sys_dim() const noexcept { return sysDim_; }
  Dim3 node_dim() const noexcept { return nodeDim_; }
  Dim3 dim() const noexcept { return sysDim_ * nodeDim_; }

  Dim3 subdomain_size(const Dim3 &idx) const { return grid_.size_of(idx); }
  Dim3 subdomain_origin(const Dim3 &idx) const noexcept { return grid_.origin_of(idx); }

  Dim3 sys_idx(int64_t i) const noexcept { return stencil::detail::dimensionize(i, sysDim_); }
  Dim3 node_idx(int64_t i) const noexcept { return stencil::detail::dimensionize(i, nodeDim_); }
};
