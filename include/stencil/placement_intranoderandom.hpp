#pragma once
// IntraNodeRandom: NodePartition, but the subdomains of each node land on its GPUs in a random order
// (the control experiment for NodeAware).

#include <random>

#include "stencil/partition.hpp"

class IntraNodeRandom : public Placement {
  NodePartition partition_;
  std::mt19937 generator_;
  stencil::detail::OwnerTable owners_;

public:
  IntraNodeRandom(const Dim3 &size, MpiTopology &mpiTopo, Radius radius, const std::vector<int> &rankCudaIds);

  Dim3 get_idx(int rank, int domId) override { return owners_.idx(rank, domId); }
  int get_rank(const Dim3 &idx) override { return owners_.rank(idx); }
  int get_subdomain_id(const Dim3 &idx) override { return owners_.id(idx); }
  int get_cuda(const Dim3 &idx) override { return owners_.cuda(idx); }
  Dim3 subdomain_size(const Dim3 &idx) override { return partition_.subdomain_size(idx); }
  Dim3 subdomain_origin(const Dim3 &idx) override { return partition_.subdomain_origin(idx); }
  Dim3 dim() override { return partition_.dim(); }
};
