#pragma once
// rt::time / rt::launch / mpirt::time: call-through wrappers with (optional) API timing hooks.

#include <cuda_runtime.h>

#include "stencil/timer.hpp"

namespace rt {

template <typename Fn, typename... Args> cudaError_t time(Fn fn, Args... args) {
  CR_TIC();
  const cudaError_t err = fn(args...);
  CR_TOC();
  return err;
}

#if __CUDACC__
template <typename Fn, typename... Args>
void launch(Fn fn, const dim3 &grid, const dim3 &block, const int shmem, cudaStream_t stream, Args... args) {
  CR_TIC();
  fn<<<grid, block, shmem, stream>>>(args...);
  CR_TOC();
}
#endif

} // namespace rt

namespace mpirt {

template <typename Fn, typename... Args> int time(Fn fn, Args... args) {
  MPI_TIC();
  const int err = fn(args...);
  MPI_TOC();
  return err;
}

} // namespace mpirt
