// The CUDA programming guide's TMA example, verbatim in spirit (cuda::barrier + cuda::device::experimental), to
// find out whether TMA works at all on this box.
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

constexpr int SMEM_WIDTH = 32, SMEM_HEIGHT = 8, GMEM_WIDTH = 1024, GMEM_HEIGHT = 1024;

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y) {
  __shared__ alignas(128) int smem_buffer[SMEM_HEIGHT][SMEM_WIDTH];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  smem_buffer[0][threadIdx.x % SMEM_WIDTH] += threadIdx.x;
  cde::fence_proxy_async_shared_cta();
  __syncthreads();
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_shared_to_global(&tensor_map, x, y, &smem_buffer);
    cde::cp_async_bulk_commit_group();
    cde::cp_async_bulk_wait_group_read<0>();
  }
  if (threadIdx.x == 0) {
    (&bar)->~barrier();
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  int *t;
  cudaMalloc(&t, sizeof(int) * GMEM_WIDTH * GMEM_HEIGHT);
  cudaMemset(t, 0, sizeof(int) * GMEM_WIDTH * GMEM_HEIGHT);
  CUtensorMap tm{};
  cuuint64_t size[2] = {GMEM_WIDTH, GMEM_HEIGHT};
  cuuint64_t stride[1] = {GMEM_WIDTH * sizeof(int)};
  cuuint32_t box[2] = {SMEM_WIDTH, SMEM_HEIGHT}, es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, t, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d (query %d)\n", (int)r, (int)q);
  kernel<<<1, 128>>>(tm, 0, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("guide kernel -> %s\n", cudaGetErrorString(e));
  int h[4];
  if (e == cudaSuccess) {
    cudaMemcpy(h, t, sizeof(h), cudaMemcpyDeviceToHost);
    printf("t[0..3] = %d %d %d %d (expect 0+96? sums of thread ids mod 32)\n", h[0], h[1], h[2], h[3]);
  }
  int drv = 0, rt = 0;
  cudaDriverGetVersion(&drv);
  cudaRuntimeGetVersion(&rt);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("driver %d runtime %d device %s cc %d.%d\n", drv, rt, p.name, p.major, p.minor);
  return 0;
}
