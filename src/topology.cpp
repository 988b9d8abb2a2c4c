#include "stencil/topology.hpp"

Topology::OptionalNeighbor Topology::get_neighbor(const Dim3 &index, const Dim3 &dir) const noexcept {
  assert(index.all_ge(0) && "subdomain indices are non-negative");
  assert(dir.all_gt(-2) && dir.all_lt(2) && "a direction has components in {-1, 0, +1}");
  OptionalNeighbor answer;
  switch (kind_) {
  case Boundary::PERIODIC:
    // a periodic grid wraps: stepping off one face re-enters through the opposite one
    answer.index = (index + dir).wrap(grid_);
    answer.exists = true;
    break;
  case Boundary::FIXED: {
    // the grid ends at its faces: a step that leaves it has no neighbour
    const Dim3 to = index + dir;
    answer.exists = to.all_ge(0) && to.all_lt(grid_);
    answer.index = answer.exists ? to : index;
    break;
  }
  default:
    LOG_FATAL("Topology::get_neighbor: the topology has no boundary kind (default-constructed?)");
  }
  return answer;
}
