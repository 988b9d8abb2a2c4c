#pragma once
// Stand-alone strided->strided copy kernels with the reference's signatures (see pack_kernel.cuh).

#include <cuda_runtime.h>

#include "stencil/dim3.hpp"

// dst[dstPos + p] = src[srcPos + p] for p in [0, extent)
__global__ void translate(cudaPitchedPtr dst, const Dim3 dstPos, cudaPitchedPtr src, const Dim3 srcPos, const Dim3 extent,
                          const size_t elemSize);

// the same region of n quantities (dsts / srcs / elemSizes are DEVICE arrays of length n)
__global__ void multi_translate(cudaPitchedPtr *dsts, const Dim3 dstPos, const cudaPitchedPtr *srcs, const Dim3 srcPos,
                                const Dim3 extent, const size_t *__restrict__ elemSizes, const size_t n);
