#include "stencil/topology.hpp"

Topology::Topology() : Topology(Dim3(0, 0, 0), Boundary::NONE) {}

Topology::OptionalNeighbor Topology::get_neighbor(const Dim3 &index, const Dim3 &dir) const noexcept {
  assert(dir.all_gt(-2) && dir.all_lt(2));
  assert(index.all_ge(0));
  if (Boundary::PERIODIC != boundary_) {
    LOG_FATAL("unexpected Boundary type");
  }
  OptionalNeighbor nbr;
  nbr.exists = true; // periodic: everybody has a neighbour everywhere
  nbr.index = (index + dir).wrap(extent_);
  return nbr;
}
