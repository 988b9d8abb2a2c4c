#pragma once
// Stand-alone pack / unpack kernels with the reference's signatures (tests and user code launch
// them directly with their own <<<grid, block>>>).  The library's own data movement does NOT use
// these: it goes through the box-copy engine (stencil_b200/csrc/box_copy.cu).

#include <cuda_runtime.h>

#include "stencil/dim3.hpp"

// dst[zo*ey*ex + yo*ex + xo] = src(pos + (xo,yo,zo)), any launch shape (grid-stride in 3-D)
__device__ void grid_pack(void *__restrict__ dst, const cudaPitchedPtr src, const Dim3 srcPos, const Dim3 srcExtent,
                          const size_t elemSize);

__global__ void pack_kernel(void *__restrict__ dst, const cudaPitchedPtr src, const Dim3 srcPos, const Dim3 srcExtent,
                            const size_t elemSize);

__device__ void grid_unpack(cudaPitchedPtr dst, const void *__restrict__ src, const Dim3 dstPos, const Dim3 dstExtent,
                            const size_t elemSize);

__global__ void unpack_kernel(cudaPitchedPtr dst, const void *src, const Dim3 dstPos, const Dim3 dstExtent,
                              const size_t elemSize);
