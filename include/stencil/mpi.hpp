#pragma once
// mpi:: convenience wrappers.  With STENCIL_USE_MPI != 1 they describe a single-rank world.

#if STENCIL_USE_MPI == 1
#include <mpi.h>
#endif

#include <cassert>
#include <limits>
#include <string>
#include <vector>

namespace mpi {

inline int comm_rank(MPI_Comm comm) {
#if STENCIL_USE_MPI == 1
  int r = 0;
  MPI_Comm_rank(comm, &r);
  return r;
#else
  (void)comm;
  return 0;
#endif
}

inline int comm_size(MPI_Comm comm) {
#if STENCIL_USE_MPI == 1
  int s = 1;
  MPI_Comm_size(comm, &s);
  return s;
#else
  (void)comm;
  return 1;
#endif
}

// largest usable tag
inline int tag_ub(MPI_Comm comm) {
#if STENCIL_USE_MPI == 1
  int *ub = nullptr;
  int found = 0;
  MPI_Comm_get_attr(comm, MPI_TAG_UB, &ub, &found);
  assert(found);
  return found ? *ub : -1;
#else
  (void)comm;
  return std::numeric_limits<int>::max();
#endif
}

inline int world_rank() { return comm_rank(MPI_COMM_WORLD); }
inline int world_size() { return comm_size(MPI_COMM_WORLD); }

inline std::string processor_name() {
  char name[MPI_MAX_PROCESSOR_NAME] = {0};
  int len = 0;
  MPI_Get_processor_name(name, &len);
  return std::string(name);
}

struct ColocatedInfo {
  MPI_Comm comm;          // shared-memory communicator
  std::vector<int> ranks; // ranks on this node
};

} // namespace mpi
