#pragma once
// MpiTopology: which ranks of a communicator share this node (shared-memory split).

#include <mpi.h>

#include <cassert>
#include <set>
#include <utility>
#include <vector>

class MpiTopology {
  MPI_Comm comm_;
  MPI_Comm nodeComm_;
  std::set<int> onNode_; // global ranks colocated with the caller

public:
  // collective over comm
  MpiTopology(MPI_Comm comm) : comm_(comm), nodeComm_{} {
    if (!comm_) {
      onNode_.insert(0);
      return;
    }
    MPI_Comm_split_type(comm_, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, &nodeComm_);
    std::vector<int> ranks(colocated_size());
    int me = 0;
    MPI_Comm_rank(comm_, &me);
    MPI_Allgather(&me, 1, MPI_INT, ranks.data(), 1, MPI_INT, nodeComm_);
    onNode_.insert(ranks.begin(), ranks.end());
  }
  MpiTopology() : MpiTopology(MPI_Comm{}) {}
  MpiTopology(const MpiTopology &) = delete;
  MpiTopology(MpiTopology &&) = delete;
  ~MpiTopology() {
    if (nodeComm_) MPI_Comm_free(&nodeComm_);
  }

  MpiTopology &operator=(MpiTopology &&o) {
    comm_ = o.comm_;
    nodeComm_ = o.nodeComm_;
    o.nodeComm_ = 0;
    onNode_ = std::move(o.onNode_);
    return *this;
  }

  int rank() const noexcept {
    assert(comm_);
    int r = 0;
    MPI_Comm_rank(comm_, &r);
    return r;
  }
  int size() const noexcept {
    assert(comm_);
    int s = 1;
    MPI_Comm_size(comm_, &s);
    return s;
  }
  MPI_Comm comm() const noexcept { return comm_; }
  MPI_Comm colocated_comm() const noexcept {
    assert(nodeComm_);
    return nodeComm_;
  }
  int colocated_rank() const noexcept {
    int r = 0;
    if (nodeComm_) MPI_Comm_rank(nodeComm_, &r);
    return r;
  }
  int colocated_size() const noexcept {
    int s = 1;
    if (nodeComm_) MPI_Comm_size(nodeComm_, &s);
    return s;
  }
  // is global rank `rank` on the caller's node?
  bool colocated(int rank) const noexcept { return onNode_.count(rank) != 0; }
};
