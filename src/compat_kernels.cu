// pack_kernel / unpack_kernel / translate / multi_translate with the reference's signatures, for code
// that launches them directly (the reference's tests do).  Straightforward 3-D grid-stride copies; the
// library's hot path never calls these (see stencil_b200/csrc/box_copy.cu).
#include "stencil/copy.cuh"
#include "stencil/pack_kernel.cuh"

namespace {

// copy `n` bytes with the widest access the size allows (elements are naturally aligned)
__device__ __forceinline__ void copy_elem(char *d, const char *s, size_t n) {
  switch (n) {
  case 1:
    *d = *s;
    break;
  case 2:
    *reinterpret_cast<uint16_t *>(d) = *reinterpret_cast<const uint16_t *>(s);
    break;
  case 4:
    *reinterpret_cast<uint32_t *>(d) = *reinterpret_cast<const uint32_t *>(s);
    break;
  case 8:
    *reinterpret_cast<uint64_t *>(d) = *reinterpret_cast<const uint64_t *>(s);
    break;
  case 16:
    *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
    break;
  default:
    for (size_t i = 0; i < n; ++i) d[i] = s[i];
  }
}

// visit every (x,y,z) of `extent` once with whatever 3-D launch shape the caller chose
template <typename F> __device__ __forceinline__ void for_each_cell(const Dim3 &extent, F f) {
  const int64_t sx = int64_t(blockDim.x) * gridDim.x, sy = int64_t(blockDim.y) * gridDim.y, sz = int64_t(blockDim.z) * gridDim.z;
  for (int64_t z = int64_t(blockDim.z) * blockIdx.z + threadIdx.z; z < extent.z; z += sz)
    for (int64_t y = int64_t(blockDim.y) * blockIdx.y + threadIdx.y; y < extent.y; y += sy)
      for (int64_t x = int64_t(blockDim.x) * blockIdx.x + threadIdx.x; x < extent.x; x += sx) f(x, y, z);
}

__device__ __forceinline__ size_t strided_offset(const cudaPitchedPtr &p, int64_t x, int64_t y, int64_t z, size_t es) {
  return (size_t(z) * p.ysize + size_t(y)) * p.pitch + size_t(x) * es;
}

} // namespace

__device__ void grid_pack(void *__restrict__ dst, const cudaPitchedPtr src, const Dim3 srcPos, const Dim3 srcExtent,
                          const size_t elemSize) {
  char *out = static_cast<char *>(dst);
  const char *in = static_cast<const char *>(src.ptr);
  for_each_cell(srcExtent, [&](int64_t x, int64_t y, int64_t z) {
    const size_t dense = (size_t(z) * srcExtent.y + size_t(y)) * srcExtent.x + size_t(x);
    copy_elem(out + dense * elemSize, in + strided_offset(src, x + srcPos.x, y + srcPos.y, z + srcPos.z, elemSize), elemSize);
  });
}

__global__ void pack_kernel(void *__restrict__ dst, const cudaPitchedPtr src, const Dim3 srcPos, const Dim3 srcExtent,
                            const size_t elemSize) {
  grid_pack(dst, src, srcPos, srcExtent, elemSize);
}

__device__ void grid_unpack(cudaPitchedPtr dst, const void *__restrict__ src, const Dim3 dstPos, const Dim3 dstExtent,
                            const size_t elemSize) {
  char *out = static_cast<char *>(dst.ptr);
  const char *in = static_cast<const char *>(src);
  for_each_cell(dstExtent, [&](int64_t x, int64_t y, int64_t z) {
    const size_t dense = (size_t(z) * dstExtent.y + size_t(y)) * dstExtent.x + size_t(x);
    copy_elem(out + strided_offset(dst, x + dstPos.x, y + dstPos.y, z + dstPos.z, elemSize), in + dense * elemSize, elemSize);
  });
}

__global__ void unpack_kernel(cudaPitchedPtr dst, const void *src, const Dim3 dstPos, const Dim3 dstExtent, const size_t elemSize) {
  grid_unpack(dst, src, dstPos, dstExtent, elemSize);
}

namespace {
__device__ __forceinline__ void translate_cells(cudaPitchedPtr dst, const Dim3 &dstPos, const cudaPitchedPtr &src,
                                                const Dim3 &srcPos, const Dim3 &extent, size_t elemSize) {
  char *out = static_cast<char *>(dst.ptr);
  const char *in = static_cast<const char *>(src.ptr);
  for_each_cell(extent, [&](int64_t x, int64_t y, int64_t z) {
    copy_elem(out + strided_offset(dst, x + dstPos.x, y + dstPos.y, z + dstPos.z, elemSize),
              in + strided_offset(src, x + srcPos.x, y + srcPos.y, z + srcPos.z, elemSize), elemSize);
  });
}
} // namespace

__global__ void translate(cudaPitchedPtr dst, const Dim3 dstPos, cudaPitchedPtr src, const Dim3 srcPos, const Dim3 extent,
                          const size_t elemSize) {
  translate_cells(dst, dstPos, src, srcPos, extent, elemSize);
}

__global__ void multi_translate(cudaPitchedPtr *dsts, const Dim3 dstPos, const cudaPitchedPtr *srcs, const Dim3 srcPos,
                                const Dim3 extent, const size_t *__restrict__ elemSizes, const size_t n) {
  for (size_t q = 0; q < n; ++q) translate_cells(dsts[q], dstPos, srcs[q], srcPos, extent, elemSizes[q]);
}
