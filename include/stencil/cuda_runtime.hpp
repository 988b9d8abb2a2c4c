#pragma once
// Error convention of the library for CUDA runtime calls: report file:line and terminate.
//   CUDA_RUNTIME(call);          wraps any expression returning cudaError_t
//   CudaErrorsFatal::{YES, NO}   lets a few entry points (LocalDomain::set_device) report instead of exiting

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "stencil/timer.hpp"

enum class CudaErrorsFatal { NO, YES };

namespace stencil {
namespace detail {
[[noreturn]] inline void cuda_failure(cudaError_t err, const char *where, int line) {
  std::fprintf(stderr, "%s:%d: CUDA Runtime Error %d: %s\n", where, line, static_cast<int>(err), cudaGetErrorString(err));
  std::exit(-1);
}
} // namespace detail
} // namespace stencil

// kept as a function as well: the reference's helper of the same name is part of what drivers may call
inline void checkCuda(cudaError_t status, const char *where, const int line) {
  if (status != cudaSuccess) stencil::detail::cuda_failure(status, where, line);
}

#define CUDA_RUNTIME(call) checkCuda((call), __FILE__, __LINE__);
