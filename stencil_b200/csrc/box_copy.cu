// Box-copy kernels (see box_copy.cuh for the design).  sm_100a only.
#include "box_copy.cuh"

namespace sb {

namespace {

// Halo source cells are read exactly once and never written by the same launch (sources are compute
// cells, destinations ghost cells or dense buffers).  Measured on B200 (scripts/time_pack.py, 512^3):
// wide rows are fastest through the read-only path (ld.global.nc: 4.2 us vs 6.2 us per 3 MB face),
// thin strided rows with plain loads (8.4 us vs 12.5 us with evict-first .cs loads).
template <typename V> __device__ __forceinline__ V ld_wide(const V *p) { return __ldg(p); }
template <> __device__ __forceinline__ unsigned char ld_wide(const unsigned char *p) { return *p; }
template <> __device__ __forceinline__ unsigned short ld_wide(const unsigned short *p) { return *p; }
template <typename V> __device__ __forceinline__ V ld_thin(const V *p) { return *p; }

// Row index -> (plane, row-in-plane) without an integer divide.
__device__ __forceinline__ void split_row(const Seg &s, unsigned R, unsigned &z, unsigned &y) {
  z = (unsigned)(((unsigned long long)R * s.ny_magic) >> s.ny_shift);
  y = R - z * s.ny;
}

template <typename V> __device__ __forceinline__ void copy_tile(const Seg &s, unsigned row0, unsigned nrows) {
  const unsigned lg = s.lg_group;
  const unsigned g = 1u << lg;
  const unsigned lane = threadIdx.x & (g - 1);
  const unsigned rsub = threadIdx.x >> lg;
  const unsigned rows_per_pass = kCopyThreads >> lg;
  const unsigned nvec = s.row_bytes / (unsigned)sizeof(V);

  if (nvec <= g) {
    // Thin rows (x-faces, edges, corners): one access per row; keep 4 rows in flight per thread.
    if (lane >= nvec) return;
    for (unsigned r = rsub; r < nrows; r += 4 * rows_per_pass) {
      V v[4];
      char *d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned rr = r + k * rows_per_pass;
        d[k] = nullptr;
        if (rr < nrows) {
          unsigned z, y;
          split_row(s, row0 + rr, z, y);
          const V *sp = reinterpret_cast<const V *>(s.src + z * s.src_slice + y * s.src_pitch) + lane;
          d[k] = s.dst + z * s.dst_slice + y * s.dst_pitch;
          v[k] = ld_thin(sp);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (d[k]) reinterpret_cast<V *>(d[k])[lane] = v[k];
      }
    }
  } else {
    // Wide rows (y/z faces, yz edges, dense buffers): a lane group streams each row.
    for (unsigned r = rsub; r < nrows; r += rows_per_pass) {
      unsigned z, y;
      split_row(s, row0 + r, z, y);
      const V *sp = reinterpret_cast<const V *>(s.src + z * s.src_slice + y * s.src_pitch);
      V *dp = reinterpret_cast<V *>(s.dst + z * s.dst_slice + y * s.dst_pitch);
      for (unsigned c = lane; c < nvec; c += 4 * g) {
        V v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned cc = c + k * g;
          if (cc < nvec) v[k] = ld_wide(sp + cc);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned cc = c + k * g;
          if (cc < nvec) dp[cc] = v[k];
        }
      }
    }
  }
}

__device__ __forceinline__ void copy_tile_dispatch(const Seg &s, unsigned row0, unsigned nrows) {
  switch (s.vec) {
  case 16:
    copy_tile<uint4>(s, row0, nrows);
    break;
  case 8:
    copy_tile<uint2>(s, row0, nrows);
    break;
  case 4:
    copy_tile<unsigned>(s, row0, nrows);
    break;
  case 2:
    copy_tile<unsigned short>(s, row0, nrows);
    break;
  default:
    copy_tile<unsigned char>(s, row0, nrows);
    break;
  }
}

// Persistent walk over the tile table: grid is a multiple of the SM count, tiles are taken
// round-robin so consecutive CTAs stream consecutive rows of the same segment.
__global__ void __launch_bounds__(kCopyThreads, 3) box_copy_kernel(const Seg *__restrict__ segs, const Tile *__restrict__ tiles,
                                                               unsigned ntiles) {
  __shared__ Seg seg;
  __shared__ Tile tile;
  unsigned cached = 0xFFFFFFFFu;
  for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
    __syncthreads(); // previous tile done before the descriptors are replaced
    if (threadIdx.x == 0) tile = tiles[t];
    __syncthreads();
    if (tile.seg != cached) {
      // cooperative 80-byte descriptor fetch
      if (threadIdx.x < sizeof(Seg) / 4)
        reinterpret_cast<unsigned *>(&seg)[threadIdx.x] = reinterpret_cast<const unsigned *>(&segs[tile.seg])[threadIdx.x];
      cached = tile.seg;
      __syncthreads();
    }
    copy_tile_dispatch(seg, tile.row0, tile.nrows);
  }
}

// One segment passed by value: the pack_kernel / unpack_kernel / translate one-shots.
__global__ void __launch_bounds__(kCopyThreads, 3) box_copy_single_kernel(const __grid_constant__ Seg seg, unsigned rows_per_tile) {
  const unsigned total = seg.ny * seg.nz;
  for (unsigned long long row0 = (unsigned long long)blockIdx.x * rows_per_tile; row0 < total;
       row0 += (unsigned long long)gridDim.x * rows_per_tile) {
    // NOTE: written as end = min(row0 + rpt, total); nrows = end - row0.  The natural
    // min(total - row0, rpt) is miscompiled by ptxas 12.9 for sm_100a: the subtract and the minimum
    // are fused into VIADDMNMX.U32 with the negation dropped (it computes min(row0 + total, rpt)),
    // which turned every partial last tile into a full one (caught by tests/test_gpu_copy.py cases 7, 11).
    const unsigned long long stop = row0 + rows_per_tile;
    const unsigned end = stop < total ? (unsigned)stop : total;
    copy_tile_dispatch(seg, (unsigned)row0, end - (unsigned)row0);
  }
}

} // namespace

void launch_box_copy(const Seg *segs_dev, const Tile *tiles_dev, unsigned ntiles, int grid, cudaStream_t stream) {
  box_copy_kernel<<<grid, kCopyThreads, 0, stream>>>(segs_dev, tiles_dev, ntiles);
}

unsigned rows_per_tile_for(unsigned row_bytes) {
  // a thin row costs at least one 32-byte sector on each side however few bytes it carries
  const unsigned cost = row_bytes < 32u ? 32u : row_bytes;
  const unsigned rows = kTileBytes / cost;
  return rows < 1 ? 1 : rows;
}

void preload_box_copy_kernels() {
  // CUDA loads kernels lazily: resolve them at plan-creation time so the first exchange does not pay
  // the module load (bench_pack of the reference has no warm-up iteration)
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, box_copy_kernel);
  cudaFuncGetAttributes(&a, box_copy_single_kernel);
}

void launch_box_copy_single(const Seg &seg, cudaStream_t stream) {
  const unsigned total = seg.ny * seg.nz;
  const unsigned rows_per_tile = rows_per_tile_for(seg.row_bytes);
  unsigned ntiles = (total + rows_per_tile - 1) / rows_per_tile;
  if (ntiles == 0) return;
  const unsigned cap = 148u * 8u;
  box_copy_single_kernel<<<ntiles < cap ? ntiles : cap, kCopyThreads, 0, stream>>>(seg, rows_per_tile);
}

} // namespace sb
