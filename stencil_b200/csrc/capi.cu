// C ABI of stencil_b200 (include/stencil_b200.h): geometry wrappers over the C++ host mirror
// (include/stencil/*.hpp), copy plans over box_copy.cu, jacobi launches, device/peer/IPC plumbing.
#include "stencil_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "box_copy.cuh"
#include "jacobi.cuh"

#include "stencil/geometry.hpp"
#include "stencil/partition_core.hpp"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(sb_status code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return int(code);
}

#define SB_CUDA(expr)                                                                                                  \
  do {                                                                                                                 \
    cudaError_t e_ = (expr);                                                                                           \
    if (e_ != cudaSuccess) return fail(SB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      ok = false;
      return;
    }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

Dim3 d3(const int64_t v[3]) { return Dim3(v[0], v[1], v[2]); }
void put(const Dim3 &d, int64_t out[3]) {
  out[0] = d.x;
  out[1] = d.y;
  out[2] = d.z;
}
Radius radius_of(const int64_t r27[27]) {
  Radius r = Radius::constant(0);
  int i = 0;
  for (int z = -1; z <= 1; ++z)
    for (int y = -1; y <= 1; ++y)
      for (int x = -1; x <= 1; ++x) r.dir(x, y, z) = size_t(r27[i++]);
  return r;
}
bool dir_ok(const int64_t d[3]) {
  for (int a = 0; a < 3; ++a)
    if (d[a] < -1 || d[a] > 1) return false;
  return true;
}

unsigned ceil_log2(unsigned v) {
  unsigned b = 0;
  while ((1ull << b) < v) ++b;
  return b;
}

// Turn one sb_box_copy into >= 1 kernel segments (see box_copy.cuh).
int build_segments(const sb_box_copy &c, std::vector<sb::Seg> &out) {
  const int64_t es = c.elem_size;
  if (es <= 0 || (es & (es - 1)) != 0 || es > 16) return fail(SB_ERR_INVALID, "elem_size %lld must be 1,2,4,8 or 16", (long long)es);
  for (int a = 0; a < 3; ++a) {
    if (c.extent[a] < 0 || c.src_pos[a] < 0 || c.dst_pos[a] < 0) return fail(SB_ERR_INVALID, "negative extent/position");
  }
  if (c.extent[0] == 0 || c.extent[1] == 0 || c.extent[2] == 0) return SB_OK; // empty box: nothing to do
  if (!c.src.ptr || !c.dst.ptr) return fail(SB_ERR_INVALID, "null pointer in box copy");
  if (c.src.pitch <= 0 || c.dst.pitch <= 0 || c.src.ysize <= 0 || c.dst.ysize <= 0)
    return fail(SB_ERR_INVALID, "pitch/ysize must be positive");

  long long row_bytes = c.extent[0] * es;
  long long ny = c.extent[1], nz = c.extent[2];
  long long sp = c.src.pitch, ss = c.src.pitch * c.src.ysize;
  long long dp = c.dst.pitch, ds = c.dst.pitch * c.dst.ysize;
  const char *src = static_cast<const char *>(c.src.ptr) + (c.src_pos[2] * c.src.ysize + c.src_pos[1]) * c.src.pitch + c.src_pos[0] * es;
  char *dst = static_cast<char *>(c.dst.ptr) + (c.dst_pos[2] * c.dst.ysize + c.dst_pos[1]) * c.dst.pitch + c.dst_pos[0] * es;

  // merge dimensions that are contiguous on BOTH sides (full-width rows / full planes)
  const long long kMaxRow = 1ll << 30;
  if (ny > 1 && sp == row_bytes && dp == row_bytes && row_bytes * ny < kMaxRow) {
    row_bytes *= ny;
    ny = 1;
  }
  if (ny == 1 && nz > 1 && ss == row_bytes && ds == row_bytes && row_bytes * nz < kMaxRow) {
    row_bytes *= nz;
    nz = 1;
  }
  if (ny == 1 && nz > 1) { // a single row per plane: planes become the rows
    ny = nz;
    nz = 1;
    sp = ss;
    dp = ds;
  }
  if (row_bytes >= (1ll << 31)) return fail(SB_ERR_INVALID, "row of %lld bytes too long", row_bytes);

  auto emit = [&](const char *s, char *d, long long rb, long long rows_y, long long rows_z, long long spitch, long long sslice,
                  long long dpitch, long long dslice) -> int {
    // keep rows per segment below 2^24 (range of the divide-free row split)
    const long long kMaxRows = (1ll << 24) - 1;
    if (rows_y > kMaxRows) return fail(SB_ERR_INVALID, "more than 2^24 rows per plane");
    long long zmax = kMaxRows / rows_y;
    for (long long z0 = 0; z0 < rows_z; z0 += zmax) {
      const long long zn = (rows_z - z0 < zmax) ? rows_z - z0 : zmax;
      sb::Seg g{};
      g.src = s + z0 * sslice;
      g.dst = d + z0 * dslice;
      g.src_pitch = spitch;
      g.src_slice = sslice;
      g.dst_pitch = dpitch;
      g.dst_slice = dslice;
      g.row_bytes = unsigned(rb);
      g.ny = unsigned(rows_y);
      g.nz = unsigned(zn);
      unsigned long long align = (unsigned long long)(uintptr_t)g.src | (unsigned long long)(uintptr_t)g.dst | (unsigned long long)rb;
      if (rows_y > 1) align |= (unsigned long long)spitch | (unsigned long long)dpitch;
      if (zn > 1) align |= (unsigned long long)sslice | (unsigned long long)dslice;
      unsigned vec = 16;
      while (vec > 1 && (align & (vec - 1))) vec >>= 1;
      g.vec = vec;
      const unsigned nvec = unsigned(rb) / vec;
      unsigned lg = ceil_log2(nvec);
      if (lg > 5) lg = 5;
      g.lg_group = lg;
      const unsigned bits = ceil_log2(unsigned(rows_y));
      g.ny_shift = 24 + bits;
      g.ny_magic = ((1ull << g.ny_shift) + unsigned(rows_y) - 1) / unsigned(rows_y);
      out.push_back(g);
    }
    return SB_OK;
  };

  if (ny == 1 && nz == 1 && row_bytes > 2 * (long long)sb::kTileBytes) {
    // one long contiguous run: cut it into tile-sized "rows"
    const long long chunk = sb::kTileBytes;
    const long long q = row_bytes / chunk, rem = row_bytes % chunk;
    int rc = emit(src, dst, chunk, q, 1, chunk, chunk * q, chunk, chunk * q);
    if (rc != SB_OK) return rc;
    if (rem) rc = emit(src + q * chunk, dst + q * chunk, rem, 1, 1, rem, rem, rem, rem);
    return rc;
  }
  return emit(src, dst, row_bytes, ny, nz, sp, ss, dp, ds);
}

void build_tiles(const std::vector<sb::Seg> &segs, std::vector<sb::Tile> &tiles) {
  for (size_t si = 0; si < segs.size(); ++si) {
    const sb::Seg &g = segs[si];
    const unsigned rows_per_tile = sb::rows_per_tile_for(g.row_bytes);
    const unsigned total = g.ny * g.nz;
    for (unsigned r = 0; r < total; r += rows_per_tile) {
      sb::Tile t{};
      t.seg = unsigned(si);
      t.row0 = r;
      t.nrows = (total - r < rows_per_tile) ? total - r : rows_per_tile;
      tiles.push_back(t);
    }
  }
}

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

struct SlotTable {
  uint32_t *p[64];
};

__global__ void signal_kernel(const __grid_constant__ SlotTable slots, int n, uint32_t value) {
  // all earlier work of this stream is complete (stream order); publish it system-wide
  __threadfence_system();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(slots.p[i]), "r"(value) : "memory");
  }
}

__global__ void wait_kernel(const uint32_t *slots, int n, uint32_t value) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slots + i) : "memory");
      if ((int32_t)(v - value) >= 0) break; // wrap-safe "v >= value"
      __nanosleep(64);
    } while (true);
  }
  __threadfence_system();
}

} // namespace

struct sb_copy_plan {
  int device = 0;
  sb::Seg *segs_dev = nullptr;
  sb::Tile *tiles_dev = nullptr;
  unsigned ntiles = 0;
  unsigned nsegs = 0;
  int grid = 0;
  int64_t bytes = 0;
};

extern "C" {

const char *sb_last_error(void) { return g_err.c_str(); }
int sb_version(void) { return SB_VERSION; }
uint64_t sb_launch_count(void) { return g_launches.load(); }

// ----------------------------------------------------------------------------------------- geometry
int sb_halo_pos(const int64_t dir[3], const int64_t size[3], const int64_t radius27[27], int halo, int64_t out[3]) {
  if (!dir_ok(dir)) return fail(SB_ERR_INVALID, "direction components must be in -1..1");
  put(stencil::geom::halo_pos(d3(dir), d3(size), radius_of(radius27), halo != 0), out);
  return SB_OK;
}
int sb_halo_extent(const int64_t dir[3], const int64_t size[3], const int64_t radius27[27], int64_t out[3]) {
  if (!dir_ok(dir)) return fail(SB_ERR_INVALID, "direction components must be in -1..1");
  put(stencil::geom::halo_extent(d3(dir), d3(size), radius_of(radius27)), out);
  return SB_OK;
}
int sb_raw_size(const int64_t size[3], const int64_t radius27[27], int64_t out[3]) {
  put(stencil::geom::raw_size(d3(size), radius_of(radius27)), out);
  return SB_OK;
}
int sb_prime_factors(int64_t n, int64_t *out, int cap) {
  if (n < 0) return fail(SB_ERR_INVALID, "n must be >= 0");
  const std::vector<int64_t> f = prime_factors<int64_t>(n);
  if (int(f.size()) > cap) return fail(SB_ERR_INVALID, "output capacity %d too small for %zu factors", cap, f.size());
  for (size_t i = 0; i < f.size(); ++i) out[i] = f[i];
  return int(f.size());
}
int sb_rank_partition(const int64_t size[3], int64_t n, int64_t dim[3], int64_t base[3], int64_t rem[3]) {
  if (n < 1) return fail(SB_ERR_INVALID, "n must be >= 1");
  RankPartition p(d3(size), n);
  const Dim3 d = p.dim();
  put(d, dim);
  put(p.subdomain_size(Dim3(0, 0, 0)), base);
  put(d3(size) % d, rem);
  return SB_OK;
}
int sb_node_partition(const int64_t size[3], const int64_t radius27[27], int64_t nodes, int64_t gpus, int64_t sys_dim[3],
                      int64_t node_dim[3], int64_t base[3], int64_t rem[3]) {
  if (nodes < 1 || gpus < 1) return fail(SB_ERR_INVALID, "nodes and gpus must be >= 1");
  NodePartition p(d3(size), radius_of(radius27), nodes, gpus);
  put(p.sys_dim(), sys_dim);
  put(p.node_dim(), node_dim);
  put(p.subdomain_size(Dim3(0, 0, 0)), base);
  put(d3(size) % p.dim(), rem);
  return SB_OK;
}
int sb_subdomain_size(const int64_t base[3], const int64_t rem[3], const int64_t idx[3], int64_t out[3]) {
  stencil::detail::CutGrid g{d3(base), d3(rem)};
  put(g.size_of(d3(idx)), out);
  return SB_OK;
}
int sb_subdomain_origin(const int64_t base[3], const int64_t rem[3], const int64_t idx[3], int64_t out[3]) {
  stencil::detail::CutGrid g{d3(base), d3(rem)};
  put(g.origin_of(d3(idx)), out);
  return SB_OK;
}
int sb_interior(const int64_t lo[3], const int64_t hi[3], const int64_t radius27[27], int64_t int_lo[3], int64_t int_hi[3]) {
  const Rect3 in = stencil::geom::interior(Rect3(d3(lo), d3(hi)), radius_of(radius27));
  put(in.lo, int_lo);
  put(in.hi, int_hi);
  return SB_OK;
}
int sb_exterior(const int64_t lo[3], const int64_t hi[3], const int64_t radius27[27], int64_t *ext_lo, int64_t *ext_hi) {
  const std::vector<Rect3> slabs = stencil::geom::exterior(Rect3(d3(lo), d3(hi)), radius_of(radius27));
  for (size_t i = 0; i < slabs.size(); ++i) {
    put(slabs[i].lo, ext_lo + 3 * i);
    put(slabs[i].hi, ext_hi + 3 * i);
  }
  return int(slabs.size());
}

// ----------------------------------------------------------------------------------------- box copies
static int one_shot(const sb_box_copy &c, void *stream) {
  std::vector<sb::Seg> segs;
  int rc = build_segments(c, segs);
  if (rc != SB_OK) return rc;
  for (const sb::Seg &g : segs) {
    sb::launch_box_copy_single(g, static_cast<cudaStream_t>(stream));
    ++g_launches;
  }
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int sb_pack(void *dst, sb_pitched src, const int64_t pos[3], const int64_t extent[3], int64_t elem_size, void *stream) {
  sb_box_copy c{};
  c.dst = sb_pitched{dst, extent[0] * elem_size, extent[1]};
  c.src = src;
  for (int a = 0; a < 3; ++a) {
    c.src_pos[a] = pos[a];
    c.extent[a] = extent[a];
  }
  c.elem_size = elem_size;
  return one_shot(c, stream);
}

int sb_unpack(sb_pitched dst, const void *src, const int64_t pos[3], const int64_t extent[3], int64_t elem_size, void *stream) {
  sb_box_copy c{};
  c.dst = dst;
  c.src = sb_pitched{const_cast<void *>(src), extent[0] * elem_size, extent[1]};
  for (int a = 0; a < 3; ++a) {
    c.dst_pos[a] = pos[a];
    c.extent[a] = extent[a];
  }
  c.elem_size = elem_size;
  return one_shot(c, stream);
}

int sb_translate(sb_pitched dst, const int64_t dst_pos[3], sb_pitched src, const int64_t src_pos[3], const int64_t extent[3],
                 int64_t elem_size, void *stream) {
  sb_box_copy c{};
  c.dst = dst;
  c.src = src;
  for (int a = 0; a < 3; ++a) {
    c.dst_pos[a] = dst_pos[a];
    c.src_pos[a] = src_pos[a];
    c.extent[a] = extent[a];
  }
  c.elem_size = elem_size;
  return one_shot(c, stream);
}

int sb_copy_plan_create(sb_copy_plan **out, int device, const sb_box_copy *copies, int64_t n) {
  if (!out || (n > 0 && !copies) || n < 0) return fail(SB_ERR_INVALID, "bad arguments to sb_copy_plan_create");
  std::vector<sb::Seg> segs;
  int64_t bytes = 0;
  for (int64_t i = 0; i < n; ++i) {
    int rc = build_segments(copies[i], segs);
    if (rc != SB_OK) return rc;
    bytes += copies[i].extent[0] * copies[i].extent[1] * copies[i].extent[2] * copies[i].elem_size;
  }
  std::vector<sb::Tile> tiles;
  build_tiles(segs, tiles);

  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  sb::preload_box_copy_kernels();
  sb_copy_plan *p = new sb_copy_plan();
  p->device = device;
  p->bytes = bytes;
  p->nsegs = unsigned(segs.size());
  p->ntiles = unsigned(tiles.size());
  if (p->ntiles) {
    SB_CUDA(cudaMalloc(&p->segs_dev, segs.size() * sizeof(sb::Seg)));
    SB_CUDA(cudaMalloc(&p->tiles_dev, tiles.size() * sizeof(sb::Tile)));
    SB_CUDA(cudaMemcpy(p->segs_dev, segs.data(), segs.size() * sizeof(sb::Seg), cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(p->tiles_dev, tiles.data(), tiles.size() * sizeof(sb::Tile), cudaMemcpyHostToDevice));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int per_sm = env_int("SB_COPY_CTAS_PER_SM", 8);
    const long long cap = (long long)sms * per_sm;
    p->grid = int(p->ntiles < cap ? p->ntiles : cap);
  }
  *out = p;
  return SB_OK;
}

int sb_copy_plan_launch(sb_copy_plan *p, void *stream) {
  if (!p) return fail(SB_ERR_INVALID, "null plan");
  if (0 == p->ntiles) return SB_OK;
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", p->device);
  sb::launch_box_copy(p->segs_dev, p->tiles_dev, p->ntiles, p->grid, static_cast<cudaStream_t>(stream));
  ++g_launches;
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int64_t sb_copy_plan_bytes(const sb_copy_plan *p) { return p ? p->bytes : 0; }
int64_t sb_copy_plan_num_tiles(const sb_copy_plan *p) { return p ? p->ntiles : 0; }

int sb_copy_plan_destroy(sb_copy_plan *p) {
  if (!p) return SB_OK;
  DeviceGuard guard(p->device);
  if (p->segs_dev) cudaFree(p->segs_dev);
  if (p->tiles_dev) cudaFree(p->tiles_dev);
  delete p;
  return SB_OK;
}

int sb_signal(uint32_t *const *remote_slots, int n, uint32_t value, int device, void *stream) {
  if (n <= 0) return SB_OK;
  if (n > 64) return fail(SB_ERR_INVALID, "at most 64 slots per signal");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SlotTable tab{};
  for (int i = 0; i < n; ++i) tab.p[i] = remote_slots[i];
  signal_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(tab, n, value);
  ++g_launches;
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int sb_wait(const uint32_t *local_slots, int n, uint32_t value, int device, void *stream) {
  if (n <= 0) return SB_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  wait_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(local_slots, n, value);
  ++g_launches;
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

// ----------------------------------------------------------------------------------------- jacobi
static int to_alloc_box(const sb_pitched &a, int dtype_size, const int64_t acc_origin[3], const int64_t lo[3],
                        const int64_t hi[3], int alo[3], int ahi[3]) {
  if (dtype_size != 4 && dtype_size != 8) return fail(SB_ERR_INVALID, "dtype_size must be 4 (float) or 8 (double)");
  if (!a.ptr || a.pitch <= 0 || a.ysize <= 0) return fail(SB_ERR_INVALID, "bad pitched pointer");
  for (int k = 0; k < 3; ++k) {
    const int64_t l = lo[k] - acc_origin[k], h = hi[k] - acc_origin[k];
    if (l < 0 || h > (1ll << 30)) return fail(SB_ERR_INVALID, "region outside the allocation on axis %d", k);
    alo[k] = int(l);
    ahi[k] = int(h);
  }
  return SB_OK;
}

static int jacobi_common(sb::JacobiParams &p, const sb_pitched &dst, const sb_pitched &src, int dtype_size,
                         const int64_t acc_origin[3], const int64_t clo[3], const int64_t chi[3]) {
  if (!dst.ptr || dst.pitch != src.pitch || dst.ysize != src.ysize)
    return fail(SB_ERR_INVALID, "dst and src must have the same pitch and ysize");
  p.dst = static_cast<char *>(dst.ptr);
  p.src = static_cast<const char *>(src.ptr);
  p.pitch = src.pitch;
  p.slice = src.pitch * src.ysize;
  p.raw[0] = int(src.pitch / dtype_size);
  p.raw[1] = int(src.ysize);
  for (int k = 0; k < 3; ++k) p.org[k] = int(acc_origin[k]);
  // sphere placement, bin/jacobi3d.cu:46-51
  const int64_t ex = chi[0] - clo[0];
  p.hot_x = int(clo[0] + ex / 3);
  p.cold_x = int(clo[0] + ex * 2 / 3);
  p.cy = int((clo[1] + chi[1]) / 2);
  p.cz = int((clo[2] + chi[2]) / 2);
  p.rad = int(ex / 10);
  p.zchunk = 0;
  return SB_OK;
}

int sb_jacobi3d_regions(sb_pitched dst, sb_pitched src, int dtype_size, const int64_t acc_origin[3], int n, const int64_t *lo,
                        const int64_t *hi, const int64_t clo[3], const int64_t chi[3], void *stream) {
  if (n < 0 || n > 8) return fail(SB_ERR_INVALID, "between 0 and 8 regions per launch");
  sb::JacobiParams p{};
  sb::JacobiRegions r{};
  int zmax = 0;
  r.first[0] = 0;
  int m = 0;
  for (int i = 0; i < n; ++i) {
    int alo[3], ahi[3];
    int rc = to_alloc_box(src, dtype_size, acc_origin, lo + 3 * i, hi + 3 * i, alo, ahi);
    if (rc != SB_OK) return rc;
    long long cells = 1;
    for (int k = 0; k < 3; ++k) {
      if (ahi[k] > alo[k] && alo[k] < 1) return fail(SB_ERR_INVALID, "region needs one ghost cell below it on axis %d", k);
      cells *= (ahi[k] > alo[k]) ? (ahi[k] - alo[k]) : 0;
    }
    if (cells == 0) continue;
    if (cells >= (1ll << 32)) return fail(SB_ERR_INVALID, "region too large for sb_jacobi3d_regions");
    for (int k = 0; k < 3; ++k) {
      r.lo[m][k] = alo[k];
      r.ext[m][k] = ahi[k] - alo[k];
    }
    if (ahi[2] > zmax) zmax = ahi[2];
    r.first[m + 1] = r.first[m] + cells;
    ++m;
  }
  r.n = m;
  if (m == 0) return SB_OK;
  int rc = jacobi_common(p, dst, src, dtype_size, acc_origin, clo, chi);
  if (rc != SB_OK) return rc;
  p.raw[2] = zmax + 1;
  g_launches += uint64_t(sb::launch_jacobi_regions(p, r, dtype_size, static_cast<cudaStream_t>(stream)));
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int sb_jacobi3d(sb_pitched dst, sb_pitched src, int dtype_size, const int64_t acc_origin[3], const int64_t lo[3],
                const int64_t hi[3], const int64_t clo[3], const int64_t chi[3], void *stream) {
  sb::JacobiParams p{};
  int rc = to_alloc_box(src, dtype_size, acc_origin, lo, hi, p.lo, p.hi);
  if (rc != SB_OK) return rc;
  for (int k = 0; k < 3; ++k) {
    if (p.lo[k] < 1 && p.hi[k] > p.lo[k]) return fail(SB_ERR_INVALID, "region needs one ghost cell below it on axis %d", k);
  }
  rc = jacobi_common(p, dst, src, dtype_size, acc_origin, clo, chi);
  if (rc != SB_OK) return rc;
  p.dst = static_cast<char *>(dst.ptr);
  p.src = static_cast<const char *>(src.ptr);
  p.pitch = src.pitch;
  p.slice = src.pitch * src.ysize;
  p.raw[0] = int(src.pitch / dtype_size);
  p.raw[1] = int(src.ysize);
  p.raw[2] = p.hi[2] + 1; // planes: the caller guarantees plane hi.z (one ghost plane above) exists
  const int n = sb::launch_jacobi(p, dtype_size, static_cast<cudaStream_t>(stream));
  g_launches += uint64_t(n);
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int sb_fill(sb_pitched dst, int dtype_size, const int64_t acc_origin[3], const int64_t lo[3], const int64_t hi[3], double value,
            void *stream) {
  int alo[3], ahi[3];
  int rc = to_alloc_box(dst, dtype_size, acc_origin, lo, hi, alo, ahi);
  if (rc != SB_OK) return rc;
  g_launches += uint64_t(sb::launch_fill(static_cast<char *>(dst.ptr), dst.pitch, dst.pitch * dst.ysize, alo, ahi, dtype_size, value,
                                         static_cast<cudaStream_t>(stream)));
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

int sb_sqdiff(sb_pitched a, sb_pitched b, int dtype_size, const int64_t acc_origin[3], const int64_t lo[3], const int64_t hi[3],
              double *out_dev, void *stream) {
  int alo[3], ahi[3];
  int rc = to_alloc_box(a, dtype_size, acc_origin, lo, hi, alo, ahi);
  if (rc != SB_OK) return rc;
  if (b.pitch != a.pitch || b.ysize != a.ysize || !b.ptr || !out_dev) return fail(SB_ERR_INVALID, "mismatched operands");
  g_launches += uint64_t(sb::launch_sqdiff(static_cast<const char *>(a.ptr), static_cast<const char *>(b.ptr), a.pitch,
                                           a.pitch * a.ysize, alo, ahi, dtype_size, out_dev, static_cast<cudaStream_t>(stream)));
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

// ----------------------------------------------------------------------------------------- device plumbing
int sb_device_count(int *count) {
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return fail(SB_ERR_NOGPU, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return SB_OK;
}

int sb_malloc(void **ptr, size_t bytes, int device) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaMalloc(ptr, bytes));
  return SB_OK;
}

int sb_free(void *ptr, int device) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaFree(ptr));
  return SB_OK;
}

int sb_memset(void *ptr, int value, size_t bytes, int device, void *stream) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaMemsetAsync(ptr, value, bytes, static_cast<cudaStream_t>(stream)));
  return SB_OK;
}

int sb_memcpy(void *dst, const void *src, size_t bytes, int device, void *stream) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return SB_OK;
}

int sb_stream_sync(int device, void *stream) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return SB_OK;
}

int sb_device_sync(int device) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaDeviceSynchronize());
  return SB_OK;
}

int sb_enable_peer(int src_device, int dst_device, int *ok) {
  *ok = 0;
  if (src_device == dst_device) {
    *ok = 1;
    return SB_OK;
  }
  int can = 0;
  SB_CUDA(cudaDeviceCanAccessPeer(&can, src_device, dst_device));
  if (!can) return SB_OK;
  DeviceGuard guard(src_device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", src_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(dst_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  SB_CUDA(e);
  *ok = 1;
  return SB_OK;
}

int sb_ipc_export(void *ptr, void *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  SB_CUDA(cudaIpcGetMemHandle(&h, ptr));
  std::memcpy(handle64, &h, sizeof(h));
  return SB_OK;
}

int sb_ipc_import(const void *handle64, int device, void **ptr) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  SB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SB_OK;
}

int sb_ipc_close(void *ptr, int device) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(SB_ERR_NOGPU, "cannot select CUDA device %d", device);
  SB_CUDA(cudaIpcCloseMemHandle(ptr));
  return SB_OK;
}

} // extern "C"
