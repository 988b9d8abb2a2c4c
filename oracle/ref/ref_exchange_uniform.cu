// Baseline driver: times dd.exchange(); dd.swap() of the UNMODIFIED reference library exactly like
// the reference's bin/bench_exchange.cu:12-55 does, but only for radius shapes the reference can run in
// 1 rank x N GPUs mode.  (bin/bench_exchange.cu starts with the asymmetric "+x only" shape, for which
// PeerAccessSender::send uses halo_extent(dir) instead of halo_extent(-dir) (tx_cuda.cuh:84) and
// make_block_dim divides by zero -- the stock binary dumps core on one GPU.)
// usage: ref_exchange_uniform <x> <y> <z> <nQuants> <radius> <iters> [default|cudampi|peer]
// TEST / BASELINE INFRASTRUCTURE ONLY.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>

#include "statistics.hpp"
#include "stencil/stencil.hpp"

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  const size_t x = argc > 1 ? atoi(argv[1]) : 512, y = argc > 2 ? atoi(argv[2]) : 512, z = argc > 3 ? atoi(argv[3]) : 512;
  const int nQuants = argc > 4 ? atoi(argv[4]) : 3;
  const int r = argc > 5 ? atoi(argv[5]) : 2;
  const int iters = argc > 6 ? atoi(argv[6]) : 30;
  const std::string how = argc > 7 ? argv[7] : "default";
  {
    DistributedDomain dd(x, y, z);
    dd.set_radius(Radius::constant(r));
    for (int i = 0; i < nQuants; ++i) dd.add_data<float>("d");
    Method m = Method::Default;
    if (how == "cudampi") m = Method::CudaMpi;
    if (how == "peer") m = Method::CudaMemcpyPeer | Method::CudaMpi;
    dd.set_methods(m);
    dd.realize();
    Statistics stats;
    for (int i = 0; i < iters + 3; ++i) {
      MPI_Barrier(MPI_COMM_WORLD);
      const double start = MPI_Wtime();
      dd.exchange();
      dd.swap();
      const double elapsed = MPI_Wtime() - start;
      if (i >= 3) stats.insert(elapsed);
    }
    const uint64_t bytes = dd.exchange_bytes_for_method(Method::Default);
    std::cout << "ref_exchange," << how << "," << x << "," << y << "," << z << "," << nQuants << "," << r << "," << dd.domains().size()
              << "," << bytes << "," << stats.trimean() << "," << stats.min() << "\n";
  }
  MPI_Finalize();
  return 0;
}
