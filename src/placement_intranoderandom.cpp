#include "stencil/placement_intranoderandom.hpp"

#include <algorithm>

// Same node-level structure as NodeAware, but inside a node the subdomains are dealt to the node's
// (rank, gpu) components in a shuffled order.  The shuffle is seeded identically on rank 0 only and
// broadcast, so all ranks agree.
IntraNodeRandom::IntraNodeRandom(const Dim3 &size, MpiTopology &mpiTopo, Radius radius, const std::vector<int> &rankCudaIds)
    : generator_(0) {
  MPI_Barrier(MPI_COMM_WORLD);
  const int gpusPerRank = int(rankCudaIds.size());
  const int ranksPerNode = mpiTopo.colocated_size();
  const int gpusPerNode = gpusPerRank * ranksPerNode;
  const int numNodes = mpiTopo.size() / ranksPerNode;
  const int numSubdomains = numNodes * gpusPerNode;
  partition_ = NodePartition(size, radius, numNodes, gpusPerNode);

  char name[MPI_MAX_PROCESSOR_NAME] = {0};
  int nameLen = 0;
  MPI_Get_processor_name(name, &nameLen);
  std::vector<char> names;
  if (0 == mpiTopo.rank()) names.resize(size_t(MPI_MAX_PROCESSOR_NAME) * mpiTopo.size());
  MPI_Gather(name, MPI_MAX_PROCESSOR_NAME, MPI_CHAR, names.data(), MPI_MAX_PROCESSOR_NAME, MPI_CHAR, 0, MPI_COMM_WORLD);

  std::vector<int> globalCudaIds(numSubdomains);
  MPI_Allgather(rankCudaIds.data(), gpusPerRank, MPI_INT, globalCudaIds.data(), gpusPerRank, MPI_INT, mpiTopo.comm());

  std::vector<int> rankOf(numSubdomains), idOf(numSubdomains), cudaOf(numSubdomains);
  if (0 == mpiTopo.rank()) {
    std::map<std::string, int> nodeOfName;
    std::vector<std::vector<int>> nodeRanks;
    for (int r = 0; r < mpiTopo.size(); ++r) {
      const std::string nm(names.data() + size_t(r) * MPI_MAX_PROCESSOR_NAME);
      auto it = nodeOfName.find(nm);
      if (it == nodeOfName.end()) {
        it = nodeOfName.emplace(nm, int(nodeRanks.size())).first;
        nodeRanks.emplace_back();
      }
      nodeRanks[it->second].push_back(r);
    }
    for (int node = 0; node < numNodes; ++node) {
      std::vector<size_t> component(gpusPerNode);
      for (int i = 0; i < gpusPerNode; ++i) component[i] = size_t(i);
      std::shuffle(component.begin(), component.end(), generator_);
      for (int id = 0; id < gpusPerNode; ++id) {
        const int c = int(component[id]);
        const int rank = nodeRanks[node][c / gpusPerRank];
        const size_t gi = size_t(node) * gpusPerNode + id;
        rankOf[gi] = rank;
        idOf[gi] = c % gpusPerRank;
        cudaOf[gi] = globalCudaIds[rank * gpusPerRank + c % gpusPerRank];
      }
    }
  }
  MPI_Bcast(rankOf.data(), int(rankOf.size()), MPI_INT, 0, MPI_COMM_WORLD);
  MPI_Bcast(idOf.data(), int(idOf.size()), MPI_INT, 0, MPI_COMM_WORLD);
  MPI_Bcast(cudaOf.data(), int(cudaOf.size()), MPI_INT, 0, MPI_COMM_WORLD);
  for (size_t gi = 0; gi < rankOf.size(); ++gi) {
    const Dim3 idx = partition_.sys_idx(int64_t(gi / gpusPerNode)) * partition_.node_dim() + partition_.node_idx(int64_t(gi % gpusPerNode));
    owners_.assign(idx, rankOf[gi], idOf[gi], cudaOf[gi]);
  }
}
