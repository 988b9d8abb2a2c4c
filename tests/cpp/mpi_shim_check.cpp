// Exercises the node-local MPI shim (src/mpi_shim.cpp) under bin/sb_mpirun: the collectives and point-to-point calls the
// reference library and drivers use.  Exit code 0 = every check passed on this rank.
#include <mpi.h>

#include <cstdio>
#include <vector>

#define CHECK(c)                                                                                                       \
  do {                                                                                                                 \
    if (!(c)) {                                                                                                        \
      std::fprintf(stderr, "rank %d: check failed: %s (line %d)\n", r, #c, __LINE__);                                  \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int r = -1, n = 0;
  MPI_Comm_rank(MPI_COMM_WORLD, &r);
  MPI_Comm_size(MPI_COMM_WORLD, &n);
  MPI_Comm shm;
  MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, &shm);
  int sr = -1, sn = 0;
  MPI_Comm_rank(shm, &sr);
  MPI_Comm_size(shm, &sn);
  CHECK(sr == r && sn == n);

  int v = r + 1, sum = 0;
  MPI_Allreduce(&v, &sum, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  CHECK(sum == n * (n + 1) / 2);
  double mx = r;
  MPI_Allreduce(MPI_IN_PLACE, &mx, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
  CHECK(mx == n - 1);
  uint64_t big = uint64_t(1) << 40, bsum = 0;
  MPI_Reduce(&big, &bsum, 1, MPI_UINT64_T, MPI_SUM, 0, MPI_COMM_WORLD);
  CHECK(r != 0 || bsum == big * uint64_t(n));

  std::vector<int> all(size_t(n), -1);
  MPI_Allgather(&v, 1, MPI_INT, all.data(), 1, MPI_INT, shm);
  for (int i = 0; i < n; ++i) CHECK(all[size_t(i)] == i + 1);
  char name[MPI_MAX_PROCESSOR_NAME];
  int len = 0;
  MPI_Get_processor_name(name, &len);
  std::vector<char> names(size_t(n) * MPI_MAX_PROCESSOR_NAME);
  MPI_Allgather(name, MPI_MAX_PROCESSOR_NAME, MPI_CHAR, names.data(), MPI_MAX_PROCESSOR_NAME, MPI_CHAR, MPI_COMM_WORLD);
  CHECK(names[size_t(n - 1) * MPI_MAX_PROCESSOR_NAME] == name[0]);

  // ring exchange of a message larger than the whole channel (8 x 32 KiB), non-blocking both ways, plus out-of-order tags
  const int N = 100000;
  std::vector<double> out, in;
  out.assign(size_t(N), double(r));
  in.assign(size_t(N), -1.0);
  int small_out = 100 + r, small_in = -1;
  MPI_Request q[4];
  const int next = (r + 1) % n, prev = (r + n - 1) % n;
  MPI_Irecv(&small_in, 1, MPI_INT, prev, 9, MPI_COMM_WORLD, &q[0]); // posted first, sent last
  MPI_Irecv(in.data(), N, MPI_DOUBLE, prev, 7, MPI_COMM_WORLD, &q[1]);
  MPI_Isend(out.data(), N, MPI_DOUBLE, next, 7, MPI_COMM_WORLD, &q[2]);
  MPI_Isend(&small_out, 1, MPI_INT, next, 9, MPI_COMM_WORLD, &q[3]);
  MPI_Status st[4];
  MPI_Waitall(4, q, st);
  CHECK(in[0] == prev && in[size_t(N - 1)] == prev && small_in == 100 + prev);
  int cnt = 0;
  MPI_Get_count(&st[1], MPI_DOUBLE, &cnt);
  CHECK(cnt == N && st[1].MPI_SOURCE == prev && st[1].MPI_TAG == 7);

  // blocking self-send (eager), then the matching receive
  int selfv = 77, selfgot = 0;
  MPI_Send(&selfv, 1, MPI_INT, r, 3, MPI_COMM_WORLD);
  selfv = 0;
  MPI_Recv(&selfgot, 1, MPI_INT, r, 3, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
  CHECK(selfgot == 77);

  int b = r == 0 ? 42 : 0;
  MPI_Bcast(&b, 1, MPI_INT, 0, MPI_COMM_WORLD);
  CHECK(b == 42);
  std::vector<int> g(size_t(n), -1);
  MPI_Gather(&v, 1, MPI_INT, g.data(), 1, MPI_INT, 0, MPI_COMM_WORLD);
  CHECK(r != 0 || g[size_t(n - 1)] == n);
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Comm_free(&shm);
  MPI_Finalize();
  if (0 == r) std::printf("mpi_shim_check OK: %d rank(s)\n", n);
  return 0;
}
