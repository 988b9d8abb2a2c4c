#pragma once
// The grid of subdomains as a graph: Topology::get_neighbor(index, dir) answers "which subdomain lies one step in
// direction dir (each component -1, 0 or +1) from subdomain `index`?".  Only periodic grids exist (as in the
// reference, topology.hpp:12-15), so the answer always exists; OptionalNeighbor keeps the reference's two fields.

#include <cassert>

#include "stencil/dim3.hpp"
#include "stencil/logging.hpp"

class Topology {
public:
  enum class Boundary { NONE, PERIODIC };

  struct OptionalNeighbor {
    Dim3 index;
    bool exists = false;
    explicit operator bool() const noexcept { return exists; }
  };

  Topology() : grid_(0, 0, 0), kind_(Boundary::NONE) {}
  Topology(const Dim3 &extent, const Boundary &boundary) : grid_(extent), kind_(boundary) {}

  OptionalNeighbor get_neighbor(const Dim3 &index, const Dim3 &dir) const noexcept;

  const Dim3 &extent() const noexcept { return grid_; }

private:
  Dim3 grid_;     // subdomains per axis
  Boundary kind_; // how the grid closes at its faces
};
