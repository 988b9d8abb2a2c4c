// C++ API check on ALL visible GPUs -- 1 process x N GPUs, or one rank per GPU under bin/sb_mpirun -n N (every rank
// checks its own subdomains; the verdict is all-reduced): fill every subdomain's compute region with a
// function of the global coordinate, exchange (+ swap + exchange again), and verify that the WHOLE
// allocation of every quantity -- ghost cells included -- holds the periodically wrapped function.
// This is the check of the reference's test/test_exchange.cu:153-187 extended to several GPUs, several
// quantities of different element sizes, asymmetric radii and exchange-after-swap.
// Build: make bin/test_exchange_multigpu      Run: bin/test_exchange_multigpu
#include <cstdio>
#include <cstring>
#include <vector>

#include "stencil/stencil.hpp"

template <typename T> __host__ __device__ T field(int q, int64_t x, int64_t y, int64_t z, int rep) {
  return T((x * 7 + y * 131 + z * 1009 + q * 3 + rep * 17) % 8191);
}

template <typename T> __global__ void fill(Accessor<T> acc, Rect3 reg, int q, int rep) {
  for (int64_t z = reg.lo.z + blockIdx.z; z < reg.hi.z; z += gridDim.z)
    for (int64_t y = reg.lo.y + blockIdx.y; y < reg.hi.y; y += gridDim.y)
      for (int64_t x = reg.lo.x + threadIdx.x; x < reg.hi.x; x += blockDim.x) acc[Dim3(x, y, z)] = field<T>(q, x, y, z, rep);
}

template <typename T>
static long check_quantity(DistributedDomain &dd, LocalDomain &d, size_t q, const Radius &radius, int rep, bool fixed) {
  const std::vector<unsigned char> raw = d.quantity_to_host(q);
  const Dim3 rs = d.raw_size();
  const Dim3 org = d.origin() - Dim3(radius.x(-1), radius.y(-1), radius.z(-1));
  const T *v = reinterpret_cast<const T *>(raw.data());
  long bad = 0;
  for (int64_t z = 0; z < rs.z; ++z)
    for (int64_t y = 0; y < rs.y; ++y)
      for (int64_t x = 0; x < rs.x; ++x) {
        // which direction's halo (if any) does this cell belong to?
        const Dim3 sz = d.size();
        const int dx = x < int64_t(radius.x(-1)) ? -1 : (x >= int64_t(radius.x(-1)) + sz.x ? 1 : 0);
        const int dy = y < int64_t(radius.y(-1)) ? -1 : (y >= int64_t(radius.y(-1)) + sz.y ? 1 : 0);
        const int dz = z < int64_t(radius.z(-1)) ? -1 : (z >= int64_t(radius.z(-1)) + sz.z ? 1 : 0);
        if ((dx || dy || dz) && 0 == radius.dir(dx, dy, dz)) continue; // no message fills this ghost region
        Dim3 p = (org + Dim3(x, y, z));
        // a FIXED grid ends at its faces: no message crosses them, those ghost cells belong to the application
        if (fixed && !(p.all_ge(0) && p.all_lt(dd.size()))) continue;
        p.wrap(dd.size());
        const T want = field<T>(int(q), p.x, p.y, p.z, rep);
        if (v[(z * rs.y + y) * rs.x + x] != want) {
          if (bad < 5)
            std::fprintf(stderr, "MISMATCH gpu%d q%zu at alloc(%ld,%ld,%ld): got %g want %g\n", d.gpu(), q, long(x), long(y), long(z),
                         double(v[(z * rs.y + y) * rs.x + x]), double(want));
          ++bad;
        }
      }
  return bad;
}

static long run_case(size_t X, size_t Y, size_t Z, const Radius &radius, const char *name, bool fixed = false) {
  DistributedDomain dd(X, Y, Z);
  dd.set_radius(radius);
  if (fixed) dd.set_boundary(Topology::Boundary::FIXED);
  auto h0 = dd.add_data<float>("f");
  auto h1 = dd.add_data<double>("d");
  auto h2 = dd.add_data<char>("c");
  dd.realize();
  long bad = 0;
  for (int rep = 0; rep < 3; ++rep) {
    for (auto &d : dd.domains()) {
      d.set_device();
      const Rect3 reg = d.get_compute_region();
      fill<float><<<dim3(1, 8, 8), 64>>>(d.get_curr_accessor(h0), reg, 0, rep);
      fill<double><<<dim3(1, 8, 8), 64>>>(d.get_curr_accessor(h1), reg, 1, rep);
      fill<char><<<dim3(1, 8, 8), 64>>>(d.get_curr_accessor(h2), reg, 2, rep);
      CUDA_RUNTIME(cudaDeviceSynchronize());
    }
    dd.exchange();
    for (auto &d : dd.domains()) {
      bad += check_quantity<float>(dd, d, 0, radius, rep, fixed);
      bad += check_quantity<double>(dd, d, 1, radius, rep, fixed);
      bad += check_quantity<char>(dd, d, 2, radius, rep, fixed);
    }
    dd.swap(); // the next round fills and exchanges the other buffer
  }
  long long all = bad;
  MPI_Allreduce(MPI_IN_PLACE, &all, 1, MPI_LONG_LONG, MPI_SUM, MPI_COMM_WORLD);
  if (0 == mpi::world_rank())
    std::printf("%-28s %zux%zux%zu on %d rank(s) x %zu subdomain(s): %s (%lld mismatches)\n", name, X, Y, Z, mpi::world_size(), dd.domains().size(),
                all ? "FAIL" : "ok", all);
  return long(all);
}

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  long bad = 0;
  bad += run_case(40, 36, 44, Radius::constant(1), "uniform r=1");
  bad += run_case(64, 48, 40, Radius::constant(3), "uniform r=3");
  bad += run_case(37, 29, 41, Radius::face_edge_corner(2, 1, 0), "faces 2, edges 1");
  {
    Radius r = Radius::constant(1);
    r.dir(1, 0, 0) = 2;
    r.dir(0, -1, 0) = 3;
    bad += run_case(48, 40, 32, r, "asymmetric +x2 -y3");
  }
  bad += run_case(128, 128, 128, Radius::face_edge_corner(1, 0, 0), "jacobi faces r=1");
  bad += run_case(48, 44, 40, Radius::constant(2), "non-periodic (FIXED) r=2", true);
  if (0 == mpi::world_rank()) std::printf(bad ? "FAILED\n" : "ALL OK\n");
  MPI_Finalize();
  return bad ? 1 : 0;
}
