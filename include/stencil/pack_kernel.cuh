#pragma once
// Compatibility kernels: pack_kernel / unpack_kernel (and their grid_* device helpers) keep the signatures the
// reference exports (src/pack_kernel.cu:3-108) because its tests and user code launch them with their own
// <<<grid, block>>>.  A box of `extent` elements at `pos` of a pitched allocation <-> a dense x-fastest buffer.
// Implemented in src/compat_kernels.cu; the library's own data movement never calls them (box-copy engine,
// stencil_b200/csrc/box_copy.cu).

#include <cuda_runtime.h>

#include <cstddef>

#include "stencil/dim3.hpp"

// strided -> dense.  Works for any launch shape: threads stride over the box in all three dimensions.
__global__ void pack_kernel(void *__restrict__ dst, const cudaPitchedPtr src, const Dim3 srcPos, const Dim3 srcExtent,
                            const size_t elemSize);
__device__ void grid_pack(void *__restrict__ dst, const cudaPitchedPtr src, const Dim3 srcPos, const Dim3 srcExtent,
                          const size_t elemSize);

// dense -> strided, the mirror image
__global__ void unpack_kernel(cudaPitchedPtr dst, const void *src, const Dim3 dstPos, const Dim3 dstExtent,
                              const size_t elemSize);
__device__ void grid_unpack(cudaPitchedPtr dst, const void *__restrict__ src, const Dim3 dstPos, const Dim3 dstExtent,
                            const size_t elemSize);
