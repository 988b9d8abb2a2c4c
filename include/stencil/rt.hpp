#pragma once
// Call-through helpers the reference API exposes for "timed" CUDA / MPI calls:
//   rt::time(fn, args...)                      -> cudaError_t
//   rt::launch(kernel, grid, block, shmem, stream, args...)
//   mpirt::time(fn, args...)                   -> int
// Here the accounting is a scope guard around the call (stencil::detail::ApiScope); with per-call accounting
// disabled (the default, like the reference's Release build) the guard is empty and the helpers cost nothing.

#include <cuda_runtime.h>

#include <utility>

#include "stencil/timer.hpp"

namespace stencil {
namespace detail {

#ifdef STENCIL_TIME_API_CALLS
struct ApiScope {
  Timer &t_;
  explicit ApiScope(Timer &t) : t_(t) { t_.resume(); }
  ~ApiScope() { t_.pause(); }
};
#else
struct ApiScope {
  explicit ApiScope(Timer &) {}
};
#endif

} // namespace detail
} // namespace stencil

namespace rt {

template <typename Call, typename... Ts> inline cudaError_t time(Call call, Ts... ts) {
  stencil::detail::ApiScope scope(timers::cudaRuntime);
  (void)scope;
  return call(ts...);
}

#if defined(__CUDACC__)
template <typename Kernel, typename... Ts>
inline void launch(Kernel kernel, const dim3 &grid, const dim3 &block, const int shmem, cudaStream_t stream, Ts... ts) {
  stencil::detail::ApiScope scope(timers::cudaRuntime);
  (void)scope;
  kernel<<<grid, block, shmem, stream>>>(ts...);
}
#endif

} // namespace rt

namespace mpirt {

template <typename Call, typename... Ts> inline int time(Call call, Ts... ts) {
  stencil::detail::ApiScope scope(timers::mpi);
  (void)scope;
  return call(ts...);
}

} // namespace mpirt
