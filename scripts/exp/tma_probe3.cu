// probe 2: my PTX sequence with the guide's descriptor shape, then vary towards the library's shape.
// argv: etype(4|8) bx by expect_first(0|1) fence_init(0|1)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap smap, const __grid_constant__ CUtensorMap dmap, unsigned bytes, int expect_first, int fence_init, int cx, int cy) {
  __shared__ alignas(128) unsigned char buf[16384];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    if (fence_init) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (expect_first) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(buf)),
                 "l"(&smap), "r"(s32(&bar)), "r"(cx), "r"(cy)
                 : "memory");
    if (!expect_first) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(s32(&bar)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&dmap), "r"(s32(buf)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char **argv) {
  const int et = atoi(argv[1]), bx = atoi(argv[2]), by = atoi(argv[3]), ef = atoi(argv[4]), fi = atoi(argv[5]);
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  const size_t W = atoi(argv[6]), H = 1024; const int cx = atoi(argv[7]), cy = atoi(argv[8]);
  char *src, *dst;
  cudaMalloc(&src, W * H * et);
  cudaMalloc(&dst, W * H * et);
  std::vector<unsigned char> h(W * H * et);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + 3);
  cudaMemcpy(src, h.data(), h.size(), cudaMemcpyHostToDevice);
  cudaMemset(dst, 0, h.size());
  CUtensorMap ms, md;
  cuuint64_t dims[2] = {W, H};
  cuuint64_t strides[1] = {W * et};
  cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by}, es[2] = {1, 1};
  CUtensorMapDataType dt = et == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_INT32;
  CUresult r1 = enc(&ms, dt, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&md, dt, 2, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  probe<<<1, 32>>>(ms, md, bx * by * et, ef, fi, cx, cy);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned char> g(64);
  if (e == cudaSuccess) cudaMemcpy(g.data(), dst, 64, cudaMemcpyDeviceToHost);
  printf("W=%d c=(%d,%d) et=%d box=%dx%d expect_first=%d fence_init=%d: encode %d %d kernel -> %s ; first bytes %s\n", (int)W, cx, cy, et, bx, by, ef, fi, (int)r1, (int)r2,
         cudaGetErrorString(e), (e == cudaSuccess && g[0] == h[0] && g[5] == h[5]) ? "match" : "n/a");
  return 0;
}
