// TMA probe: which combination of descriptor / instruction form works on this box?  argv[1] = variant
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap smap, const __grid_constant__ CUtensorMap dmap, int x, int y, int z, int dx, int dy, int dz,
                      unsigned bytes) {
  __shared__ alignas(128) unsigned char buf[8192];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(buf)),
                   "l"(&smap), "r"(s32(&bar)), "r"(x), "r"(y)
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(buf)),
                   "l"(&smap), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z)
                   : "memory");
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(s32(&bar)) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&dmap), "r"(s32(buf)), "r"(dx), "r"(dy) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&dmap), "r"(s32(buf)), "r"(dx), "r"(dy),
                   "r"(dz)
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  // an allocation like a 66^3 double subdomain
  const int nx = 66, ny = 66, nz = 66;
  double *src, *dst;
  CK(cudaMalloc(&src, sizeof(double) * nx * ny * nz));
  CK(cudaMalloc(&dst, sizeof(double) * nx * ny * nz));
  std::vector<double> h(nx * ny * nz);
  for (size_t i = 0; i < h.size(); ++i) h[i] = double(i);
  CK(cudaMemcpy(src, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dst, 0, h.size() * 8));
  CUtensorMap ms, md;
  CUresult r1, r2;
  unsigned bytes = 0;
  int rank = 3;
  if (variant == 0) { // guide-like 2D: dims (66, 66*66), box 64 x 1, UINT64
    rank = 2;
    cuuint64_t dims[2] = {66, 66 * 66};
    cuuint64_t strides[1] = {528};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    r1 = enc(&ms, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    r2 = enc(&md, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = 512;
  } else {
    cuuint64_t zdim = (variant == 1) ? 66 : (variant == 2 ? (1ull << 20) : 66);
    cuuint64_t dims[3] = {66, 66, zdim};
    cuuint64_t strides[2] = {528, 528 * 66};
    cuuint32_t bx = (variant == 3) ? 32 : 64;
    cuuint32_t box[3] = {bx, 1, 1}, es[3] = {1, 1, 1};
    CUtensorMapDataType dt = (variant == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
    CUtensorMapL2promotion l2 = (variant == 5) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    r1 = enc(&ms, dt, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    r2 = enc(&md, dt, 3, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    bytes = bx * 8;
  }
  printf("variant %d: encode %d %d\n", variant, (int)r1, (int)r2);
  if (r1 || r2) return 1;
  if (rank == 2)
    probe<2><<<1, 32>>>(ms, md, 1, 1 + 66 * 1, 0, 1, 0 + 66 * 0, 0, bytes);
  else
    probe<3><<<1, 32>>>(ms, md, 1, 1, 1, 1, 0, 0, bytes);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d: kernel -> %s\n", variant, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    CK(cudaMemcpy(h.data(), dst, h.size() * 8, cudaMemcpyDeviceToHost));
    // expect dst row (y=0,z=0) x=1.. == src (x=1.., y=1, z=1)
    double want = double((1 * 66 + 1) * 66 + 1);
    printf("variant %d: dst[1] = %.0f (want %.0f), dst[%u] = %.0f\n", variant, h[1], want, bytes / 8, h[bytes / 8]);
  }
  return 0;
}
