#pragma once
// CUDA_RUNTIME(stmt): every CUDA runtime error is fatal (the library's error convention: print, exit).

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "stencil/timer.hpp"

inline void checkCuda(cudaError_t result, const char *file, const int line) {
  if (cudaSuccess == result) return;
  std::fprintf(stderr, "%s:%d: CUDA Runtime Error %d: %s\n", file, line, int(result), cudaGetErrorString(result));
  std::exit(-1);
}

#define CUDA_RUNTIME(stmt) checkCuda(stmt, __FILE__, __LINE__);

enum class CudaErrorsFatal { NO, YES };
