#pragma once
// Topology: neighbour lookup in the grid of subdomains (periodic boundaries only).

#include "stencil/dim3.hpp"
#include "stencil/logging.hpp"

class Topology {
public:
  enum class Boundary { NONE, PERIODIC };

  struct OptionalNeighbor {
    Dim3 index;
    bool exists;
  };

  Topology();
  Topology(const Dim3 &extent, const Boundary &boundary) : extent_(extent), boundary_(boundary) {}

  // index of the subdomain one step in `dir` from `index`
  OptionalNeighbor get_neighbor(const Dim3 &index, const Dim3 &dir) const noexcept;

private:
  Dim3 extent_;
  Boundary boundary_;
};
