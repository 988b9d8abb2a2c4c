#pragma once
// Small mpi:: queries used all over the library and by the drivers (mpi::world_rank() in every log line).
// Built without MPI (STENCIL_USE_MPI != 1) they describe a world of exactly one rank.  In this repository MPI is the
// single-process shim include/mpi_shim/mpi.h unless a real implementation is put first on the include path.

#if STENCIL_USE_MPI == 1
#include <mpi.h>
#endif

#include <cassert>
#include <limits>
#include <string>
#include <vector>

namespace mpi {

#if STENCIL_USE_MPI == 1

namespace detail {
template <typename Query> inline int ask(MPI_Comm comm, Query query, int fallback) {
  int answer = fallback;
  query(comm, &answer);
  return answer;
}
} // namespace detail

inline int comm_rank(MPI_Comm comm) { return detail::ask(comm, MPI_Comm_rank, 0); }
inline int comm_size(MPI_Comm comm) { return detail::ask(comm, MPI_Comm_size, 1); }

// the largest tag the implementation accepts on `comm` (MPI_TAG_UB attribute), -1 if it does not say
inline int tag_ub(MPI_Comm comm) {
  int *value = nullptr;
  int present = 0;
  MPI_Comm_get_attr(comm, MPI_TAG_UB, &value, &present);
  assert(present && "MPI_TAG_UB is a predefined attribute");
  return (present && value) ? *value : -1;
}

inline std::string processor_name() {
  std::string name(MPI_MAX_PROCESSOR_NAME, '\0');
  int used = 0;
  MPI_Get_processor_name(&name[0], &used);
  name.resize(used > 0 ? size_t(used) : 0);
  return name;
}

#else // a world of one

inline int comm_rank(MPI_Comm) { return 0; }
inline int comm_size(MPI_Comm) { return 1; }
inline int tag_ub(MPI_Comm) { return std::numeric_limits<int>::max(); }
inline std::string processor_name() { return "localhost"; }

#endif

inline int world_rank() { return comm_rank(MPI_COMM_WORLD); }
inline int world_size() { return comm_size(MPI_COMM_WORLD); }

// the ranks sharing this node, and the communicator that connects them
struct ColocatedInfo {
  MPI_Comm comm;
  std::vector<int> ranks;
};

} // namespace mpi
