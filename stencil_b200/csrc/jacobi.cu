// 7-point jacobi kernels for sm_100a (see jacobi.cuh).
//
// jacobi_march_kernel: register-blocked z-march.  A warp owns a strip of 32*VX cells in x and RY
// rows in y; each lane keeps the z-1 / z / z+1 values of its VX*RY columns in registers, so every
// cell of the subdomain is fetched from HBM exactly once (as the z+1 "centre" of its own column)
// with 16-byte vector loads.  x-neighbours come from the adjacent lane by warp shuffle (lanes 0
// and 31 fetch one scalar), y-neighbours inside the RY rows from registers and across warps from
// L1 (the neighbouring warp loaded that row one step earlier).  Eight warps are stacked in y, the
// z axis is cut into chunks so that the grid has several waves on 148 SMs, and an optional
// prefetch.global.L2 runs a few planes ahead of the march to deepen the memory pipeline without
// spending registers.  Algorithmic traffic: one read + one write per cell (2*sizeof(T) B/cell).
//
// Numerics: sum order ((((((0+px)+mx)+py)+my)+pz)+mz) as in bin/jacobi3d.cu:66-77, then an exact
// IEEE division by 6 done as q = v*c; r = fma(-6,q,v); q' = fma(r,c,q) with c = RN(1/6)
// (Markstein's correction: q' is the correctly rounded quotient), so results are bit-identical
// to the CPU oracle's `v / 6` in both FP32 and FP64.
#include "jacobi.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace sb {
namespace {

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float div6(float v) {
    const float c = 1.0f / 6.0f;
    const float q = __fmul_rn(v, c);
    const float r = __fmaf_rn(-6.0f, q, v);
    return __fmaf_rn(r, c, q);
  }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double div6(double v) {
    const double c = 1.0 / 6.0;
    const double q = __dmul_rn(v, c);
    const double r = __fma_rn(-6.0, q, v);
    return __fma_rn(r, c, q);
  }
};

template <typename T, int N> struct alignas(sizeof(T) * N) Vec { T v[N]; };

// dist() of bin/jacobi3d.cu:31-33: int64(__fsqrt_rn(float(d2)))
__device__ __forceinline__ bool in_sphere(int d2, int rad) { return (int)__fsqrt_rn((float)d2) <= rad; }

template <typename T> __device__ __forceinline__ T stencil_value(T px, T mx, T py, T my, T pz, T mz) {
  T v = Num<T>::add(T(0), px);
  v = Num<T>::add(v, mx);
  v = Num<T>::add(v, py);
  v = Num<T>::add(v, my);
  v = Num<T>::add(v, pz);
  v = Num<T>::add(v, mz);
  return Num<T>::div6(v);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Register-blocked z-march (see the file header).  Everything that does not change along z is hoisted
// out of the loop -- clamped row offsets, store masks, the (y - cy)^2 term of the sphere test -- so a
// z-step is: RY centre loads for plane z+1, 2 halo-row loads and RY edge scalars for plane z, the
// shuffles, 9 FP64 ops per cell and RY stores.  The loop is unrolled by three with the roles of the
// three register planes rotating, so there are no register-to-register plane copies.
//
// SHIFT variant (RY = 1): when the row pitch is only half-vector aligned (FP32 rows of 514 floats =
// 2056 B are 8 mod 16) consecutive rows alternate between two 16-byte phases.  Each warp then
// starts its strip VX/2 cells earlier on the odd-phase rows, so the HBM-facing accesses (centre
// loads, stores) stay full 128-bit vectors on every row; only the two L1-resident neighbour rows
// (y-1, y+1), which have the opposite phase, are fetched as two half vectors.
//
// EDGE variants (RY = 1; the boundary CTAs of the fused kernel): identical to the plain loop except for where the cells
// just outside the subdomain come from -- a periodic self-neighbour is read in place from the opposite face of src
// (pointer set-up before the loop), and the x neighbour of the first / last column may come from a dense received array
// instead of the ghost column (EDGE = 3).  Nothing is pushed inside the loop: the faces are shipped afterwards by the
// last CTA of each face group (jacobi_fused_kernel).
template <typename T, int VX, int RY, bool SHIFT, int PUSH> // PUSH (= EDGE): 0 plain, 2 boundary CTA, 3 boundary CTA with dense x ghosts
__device__ __forceinline__ void march_body(const JacobiParams &p, const int bx, const int by, const int bz) {
  static_assert(!SHIFT || (RY == 1 && VX >= 2), "the phase-shifted variant handles one row per warp");
  static_assert(!PUSH || RY == 1, "the push variant handles one row per warp");
  using V = Vec<T, VX>;
  constexpr int WY = 8; // warps stacked in y
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;

  const long long S = p.slice, P = p.pitch;
  const int y = p.lo[1] + (by * WY + warp) * RY;         // first row of this warp
  // phase of this row inside a 16-byte vector, in elements (0 unless SHIFT)
  const int shift = SHIFT ? (int)((((unsigned long long)y * (unsigned long long)P) & (sizeof(T) * VX - 1)) / sizeof(T)) : 0;
  const int x0w = p.x0a + bx * 32 * VX - shift; // first cell of this warp's strip (allocation index)
  const int x = x0w + lane * VX;                              // first cell of this lane
  const int z0 = p.lo[2] + bz * p.zchunk;
  const int z1 = min(z0 + p.zchunk, p.hi[2]);
  if (y >= p.hi[1] || x0w >= p.hi[0]) return; // warp-uniform

  // does this lane's vector lie inside the allocation row?  (SHIFT: a vector may hang over either end
  // of its row into the neighbouring row of the same allocation -- valid memory, masked cells)
  const bool xin = SHIFT ? true : (x + VX <= p.raw[0]);
  const int xs = xin ? x : (p.x0a >= 0 ? p.x0a : p.x0a + VX); // out-of-row lanes read (and discard) an aligned in-row vector
  const long long xoff = (long long)xs * (long long)sizeof(T);

  // row byte offsets inside a plane; rows beyond the allocation are clamped (their results are masked)
  auto yo = [&](int yy) { return (long long)clampi(yy, 0, p.raw[1] - 1) * P; };
  const char *__restrict__ src = p.src;
  const char *pc[RY]; // centre rows, plane z+1
  const char *ph[RY]; // edge scalar of each row, plane z
  char *pw[RY];       // output rows, plane z
  const bool edge_lane = (lane == 0) || (lane == 31);
  long long phstep = S; // per-plane advance of ph (mode 3: a dense x-ghost array is z fastest)
  int hxv = lane == 0 ? x - 1 : x + VX;
  if (PUSH && p.xwrap) { // periodic self-neighbour: the cell beyond a face is the first / last cell of the same row
    if (hxv == p.lo[0] - 1) hxv = p.hi[0] - 1;
    else if (hxv == p.hi[0]) hxv = p.lo[0];
  }
  const int hx = clampi(hxv, 0, p.raw[0] - 1);
#pragma unroll
  for (int j = 0; j < RY; ++j) {
    pc[j] = src + (long long)(z0 + 1) * S + yo(y + j) + xoff;
    ph[j] = src + (long long)z0 * S + yo(y + j) + (long long)hx * (long long)sizeof(T);
    if (PUSH == 3) { // the x neighbour of the first / last compute cell may come from a dense received array [y][z]
      const int side = (lane == 0 && x == p.lo[0]) ? 0 : ((lane == 31 && x + VX == p.hi[0]) ? 1 : -1);
      if (side >= 0 && p.xghost_ptr[side]) {
        ph[j] = p.xghost_ptr[side] + (long long)(y + j) * p.xghost_pitch[side] + (long long)z0 * (long long)sizeof(T);
        phstep = (long long)sizeof(T);
      }
    }
    pw[j] = p.dst + (long long)z0 * S + (long long)(y + j) * P + (long long)x * (long long)sizeof(T);
  }
  int yu = y - 1, yd = y + RY;
  if (PUSH && p.ywrap) { // periodic self-neighbour: the row beyond a face is the opposite face row
    if (yu == p.lo[1] - 1) yu = p.hi[1] - 1;
    if (yd == p.hi[1]) yd = p.lo[1];
  }
  const char *pu = src + (long long)z0 * S + yo(yu) + xoff; // row above the strip, plane z
  const char *pd = src + (long long)z0 * S + yo(yd) + xoff; // row below the strip, plane z

  // store masks
  bool row_ok[RY];
#pragma unroll
  for (int j = 0; j < RY; ++j) row_ok[j] = (y + j < p.hi[1]);
  const bool full = (x >= p.lo[0]) && (x + VX <= p.hi[0]);
  unsigned cell_ok = 0;
#pragma unroll
  for (int i = 0; i < VX; ++i)
    if (x + i >= p.lo[0] && x + i < p.hi[0]) cell_ok |= 1u << i;

  // sphere test pieces that are constant along z
  const int rr = (p.rad + 1) * (p.rad + 1);
  int dy2[RY];
#pragma unroll
  for (int j = 0; j < RY; ++j) {
    const int dyv = y + j + p.org[1] - p.cy;
    dy2[j] = dyv * dyv;
  }
  const int gx0 = x + p.org[0];

  V A[RY], B[RY], C[RY];
#pragma unroll
  for (int j = 0; j < RY; ++j) {
    A[j] = *reinterpret_cast<const V *>(pc[j] - 2 * S); // plane z0-1
    B[j] = *reinterpret_cast<const V *>(pc[j] - S);     // plane z0
  }
  const int zlast = p.raw[2] - 1; // last plane that may be touched (prefetch guard)
  int z = z0;

  auto step = [&](const V(&prev)[RY], const V(&cur)[RY], V(&nxt)[RY]) {
#pragma unroll
    for (int j = 0; j < RY; ++j) nxt[j] = *reinterpret_cast<const V *>(pc[j]);
    if (p.prefetch > 0 && z + 1 + p.prefetch <= zlast) {
#pragma unroll
      for (int j = 0; j < RY; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc[j] + (long long)p.prefetch * S));
    }
    V up, dn;
    if (SHIFT) { // the rows above / below have the other phase: two aligned half vectors each
      using H = Vec<T, (VX >= 2 ? VX / 2 : 1)>;
      const H u0 = reinterpret_cast<const H *>(pu)[0], u1 = reinterpret_cast<const H *>(pu)[1];
      const H d0 = reinterpret_cast<const H *>(pd)[0], d1 = reinterpret_cast<const H *>(pd)[1];
#pragma unroll
      for (int i = 0; i < VX / 2; ++i) {
        up.v[i] = u0.v[i];
        up.v[i + VX / 2] = u1.v[i];
        dn.v[i] = d0.v[i];
        dn.v[i + VX / 2] = d1.v[i];
      }
    } else {
      up = *reinterpret_cast<const V *>(pu);
      dn = *reinterpret_cast<const V *>(pd);
    }
    const int dzv = z + p.org[2] - p.cz;
    const int dz2 = dzv * dzv;
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      T h = T(0);
      if (edge_lane) h = *reinterpret_cast<const T *>(ph[j]);
      T left = __shfl_up_sync(0xffffffffu, cur[j].v[VX - 1], 1);
      T right = __shfl_down_sync(0xffffffffu, cur[j].v[0], 1);
      if (lane == 0) left = h;
      if (lane == 31) right = h;
      V out;
#pragma unroll
      for (int i = 0; i < VX; ++i) {
        const T px = (i < VX - 1) ? cur[j].v[i + 1 < VX ? i + 1 : i] : right;
        const T mx = (i > 0) ? cur[j].v[i > 0 ? i - 1 : 0] : left;
        const T py = (j < RY - 1) ? cur[j + 1 < RY ? j + 1 : j].v[i] : dn.v[i];
        const T my = (j > 0) ? cur[j > 0 ? j - 1 : 0].v[i] : up.v[i];
        out.v[i] = stencil_value<T>(px, mx, py, my, nxt[j].v[i], prev[j].v[i]);
      }
      const int dyz2 = dy2[j] + dz2;
      if (dyz2 < rr) { // rare: this row crosses a sphere (exact pre-filter, see oracle/stencil_oracle.c)
#pragma unroll
        for (int i = 0; i < VX; ++i) {
          const int dh = (gx0 + i - p.hot_x) * (gx0 + i - p.hot_x) + dyz2;
          const int dc = (gx0 + i - p.cold_x) * (gx0 + i - p.cold_x) + dyz2;
          if (in_sphere(dh, p.rad)) {
            out.v[i] = T(1);
          } else if (in_sphere(dc, p.rad)) {
            out.v[i] = T(0);
          }
        }
      }
      if (row_ok[j]) {
        if (full) {
          *reinterpret_cast<V *>(pw[j]) = out;
        } else {
#pragma unroll
          for (int i = 0; i < VX; ++i)
            if (cell_ok & (1u << i)) reinterpret_cast<T *>(pw[j])[i] = out.v[i];
        }
      }
      pc[j] += S;
      ph[j] += (PUSH == 3) ? phstep : S;
      pw[j] += S;
    }
    pu += S;
    pd += S;
    ++z;
  };

  while (z < z1) {
    step(A, B, C);
    if (z >= z1) break;
    step(B, C, A);
    if (z >= z1) break;
    step(C, A, B);
  }

}

template <typename T, int VX, int RY, int MB, bool SHIFT>
__global__ void __launch_bounds__(256, MB) jacobi_march_kernel(const __grid_constant__ JacobiParams p, int tiles_x, int tiles_y) {
  int b = blockIdx.x;
  const int bx = b % tiles_x;
  b /= tiles_x;
  march_body<T, VX, RY, SHIFT, 0>(p, bx, b % tiles_y, b / tiles_y);
}

// ---------------------------------------------------------------------------------------------------- fused iteration
// launch_jacobi_fused: the march over the WHOLE compute region of a subdomain, with the halo exchange of the next
// iteration and the ordering between ranks inside the kernel.
//
//  * Blocks run in the natural order (x fastest): measured, gathering the boundary CTAs at the start or the end of the
//    grid costs 10 % (their 512-byte row segments lose the DRAM locality of whole rows).  CTAs that touch no face of the
//    subdomain (64 % at 512^3) run the plain loop; boundary CTAs run the EDGE variant, which differs only in where the
//    cells just outside the subdomain come from.  No push code inside any loop (it cost 36 % more issued instructions).
//  * Face groups.  The cells of face f computed by one group of boundary CTAs form a slab of ~2048 cells that is shipped
//    in one piece (struct Groups below): x faces -- 64 rows x zchunk planes of the column, y faces -- one strip of the row x
//    zchunk planes, z faces -- 8 rows of the plane.  Every boundary CTA counts itself into the
//    groups it belongs to (release at GPU scope); the LAST one to arrive re-reads the slab from dst (L2) and stores it
//    into the neighbour -- ghost rows / planes, the ghost column, or the neighbour's dense x array -- then, if the
//    neighbour is another rank, fences at system scope and publishes the iteration number in the neighbour's mailbox word
//    for that group (st.release.sys).  ~220 of 8192 CTAs ship; only they pay an NVLink round trip.
//  * Waiting.  Before marching, a boundary CTA polls the mailbox words of ITS groups (ld.acquire.sys, own memory) until
//    the neighbour has shipped the previous iteration's slab: that one flag says both "the ghost cells I read are
//    filled" and "the neighbour's group is done reading the ghost cells my group's shipper will overwrite".  Neighbours
//    walk their grids in the same order, so the word a CTA needs was written a whole iteration earlier: x / y groups
//    by construction, z groups because the z order is rotated by half the chunks (s.zrot) -- without the rotation the
//    first chunk of iteration e+1 would need what the last chunk of iteration e ships.
constexpr int kMaxGroups = SB_FUSED_MAX_GROUPS;

// Group geometry (the same on both sides of a face: neighbours across a face have equal extents on the other two axes).
// Slabs are kept to ~2048 cells so that the CTA that ships one reads it in one or two batches of independent loads:
//   x face: group (z chunk, 8 tile rows)  = 64 rows x zchunk planes of the column,  members: the 8 (x nxhi) CTAs of those rows
//   y face: group (z chunk, x tile)       = zchunk planes x one strip of the row,    members: 1
//           (phase-shifted rows, SHIFT: strips of neighbouring rows do not line up -> one group per z chunk, members nx)
//   z face: group (tile row)              = 8 rows x all x of the plane,             members: the nx CTAs of that tile row
struct Groups {
  int ny8; // x-face groups per z chunk
  int nx;  // y-face groups per z chunk
};
template <bool SHIFT> __device__ __forceinline__ int group_of(int f, int bx, int by, int bz, const Groups &g) {
  return f < 2 ? bz * g.ny8 + (by >> 3) : (f < 4 ? (SHIFT ? bz : bz * g.nx + bx) : by);
}

// all 256 threads: copy `total` cells; cell i is read from src_of(i) and stored to dst_of(i).  Eight independent loads per
// thread are in flight before the first store (a shipper is the tail of its group: latency, not bandwidth, is what counts)
template <typename T, typename SrcOf, typename DstOf> __device__ __forceinline__ void ship_cells(int total, SrcOf src_of, DstOf dst_of) {
  for (int base = threadIdx.x; base < total; base += 256 * 8) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * 256;
      if (i < total) v[k] = __ldcg(reinterpret_cast<const T *>(src_of(i)));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * 256;
      if (i < total) *reinterpret_cast<T *>(dst_of(i)) = v[k];
    }
  }
}

// copy the slab of face f that belongs to the group of tile (bx, by, bz) from dst into the neighbour
template <typename T, int VX, bool SHIFT> __device__ __forceinline__ void ship_face(const JacobiParams &p, int f, int bx, int by, int bz) {
  const long long S = p.slice, P = p.pitch;
  const long long es = (long long)sizeof(T);
  const int z0 = p.lo[2] + bz * p.zchunk, z1 = min(z0 + p.zchunk, p.hi[2]);
  if (f < 2) { // x face: column xc, the 64 rows of 8 tile rows, planes [z0, z1); consecutive threads along z (dense target: z fastest)
    const int xc = f == 0 ? p.lo[0] : p.hi[0] - 1;
    const int np = z1 - z0;
    const int y0 = p.lo[1] + (by >> 3) * 64, rows = min(64, p.hi[1] - y0);
    const char *src = p.dst + (long long)xc * es;
    char *tgt = p.push_ptr[f];
    const long long tp = p.push_pitch[f], ts = p.xdense[f] ? es : p.push_slice[f];
    ship_cells<T>(
        rows * np, [&](int i) { return src + (long long)(z0 + i % np) * S + (long long)(y0 + i / np) * P; },
        [&](int i) { return tgt + (long long)(y0 + i / np) * tp + (long long)(z0 + i % np) * ts; });
  } else if (f < 4) { // y face: row yf, planes [z0, z1), the cells of this x strip (SHIFT: the whole row)
    const int yf = f == 2 ? p.lo[1] : p.hi[1] - 1;
    const int xa = SHIFT ? p.lo[0] : max(p.lo[0], p.x0a + bx * 32 * VX), xb = SHIFT ? p.hi[0] : min(p.hi[0], p.x0a + (bx + 1) * 32 * VX);
    const int ex = xb - xa;
    if (ex <= 0) return;
    const char *src = p.dst + (long long)yf * P + (long long)xa * es;
    char *tgt = p.push_ptr[f] + (long long)xa * es;
    const long long ts = p.push_slice[f];
    ship_cells<T>(
        ex * (z1 - z0), [&](int i) { return src + (long long)(z0 + i / ex) * S + (long long)(i % ex) * es; },
        [&](int i) { return tgt + (long long)(z0 + i / ex) * ts + (long long)(i % ex) * es; });
  } else { // z face: plane zf, the 8 rows of tile row `by`, all x
    const int zf = f == 4 ? p.lo[2] : p.hi[2] - 1;
    const int ex = p.hi[0] - p.lo[0];
    const int y0 = p.lo[1] + by * 8, rows = min(8, p.hi[1] - y0);
    const char *src = p.dst + (long long)zf * S + (long long)p.lo[0] * es;
    char *tgt = p.push_ptr[f] + (long long)p.lo[0] * es;
    const long long tp = p.push_pitch[f];
    ship_cells<T>(
        ex * rows, [&](int i) { return src + (long long)(y0 + i / ex) * P + (long long)(i % ex) * es; },
        [&](int i) { return tgt + (long long)(y0 + i / ex) * tp + (long long)(i % ex) * es; });
  }
}

template <typename T, int VX, bool SHIFT, int EDGE>
__global__ void __launch_bounds__(256, 4)
    jacobi_fused_kernel(const __grid_constant__ JacobiParams p, const __grid_constant__ FusedSync s, int nx, int ny, int nz) {
  int b = blockIdx.x;
  const int bx = b % nx;
  b /= nx;
  const int by = b % ny;
  int bz = b / ny + s.zrot;
  if (bz >= nz) bz -= nz;
  const Groups g{(ny + 7) >> 3, nx};
  // the faces this CTA's cells lie on (phase-shifted rows end one strip later than the others, so with SHIFT the last TWO
  // strips along x may hold cells of the +x face); only faces with something to push form groups
  const int nxhi = SHIFT ? min(nx, 2) : 1;
  unsigned touch = 0;
  if (bx == 0) touch |= 1u;
  if (bx >= nx - nxhi) touch |= 2u;
  if (by == 0) touch |= 4u;
  if (by == ny - 1) touch |= 8u;
  if (bz == 0) touch |= 16u;
  if (bz == nz - 1) touch |= 32u;
  if (!touch) {
    march_body<T, VX, 1, SHIFT, 0>(p, bx, by, bz);
    return;
  }
  unsigned faces = 0; // faces with a push
#pragma unroll
  for (int f = 0; f < 6; ++f)
    if ((touch >> f & 1u) && p.push_ptr[f]) faces |= 1u << f;

  if (s.any_wait && faces) {
    const int f = threadIdx.x;
    if (f < 6 && (faces >> f & 1u) && s.wait_row[f]) {
      const uint32_t want = s.wait_value;
      const uint32_t *slot = s.wait_row[f] + group_of<SHIFT>(f, bx, by, bz, g);
      uint32_t v;
      while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
        if ((int32_t)(v - want) >= 0) break; // wrap-safe "v >= want"
        __nanosleep(100);
      }
    }
    __syncthreads();
  }

  march_body<T, VX, 1, SHIFT, EDGE>(p, bx, by, bz);
  if (!faces) return;

  __shared__ unsigned last_mask;
  __syncthreads(); // every thread's stores to dst are issued ...
  if (threadIdx.x == 0) {
    __threadfence(); // ... and visible GPU-wide before this CTA counts itself in
    unsigned m = 0;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      if (!(faces >> f & 1u)) continue;
      // x: the tile rows of this group (the last group may be short) x the strips holding the face; y: one strip; z: all strips
      const unsigned members = f < 2 ? unsigned(min(8, ny - (by & ~7))) * (f == 1 ? unsigned(nxhi) : 1u) : (f < 4 && !SHIFT ? 1u : unsigned(nx));
      uint32_t *cnt = s.counters + f * kMaxGroups + group_of<SHIFT>(f, bx, by, bz, g);
      if (atomicAdd(cnt, 1u) == members - 1u) {
        *cnt = 0; // the next launch on this stream starts from zero
        m |= 1u << f;
      }
    }
    if (m) __threadfence(); // acquire: the other members' stores
    last_mask = m;
  }
  __syncthreads();
  const unsigned mine = last_mask;
  if (!mine) return;
#pragma unroll
  for (int f = 0; f < 6; ++f)
    if (mine >> f & 1u) ship_face<T, VX, SHIFT>(p, f, bx, by, bz);
  bool signal = false;
#pragma unroll
  for (int f = 0; f < 6; ++f)
    if ((mine >> f & 1u) && s.signal_row[f]) signal = true;
  if (!signal) return; // neighbours inside this process are ordered by stream events
  __syncthreads();     // all of the slab is on its way ...
  if (threadIdx.x == 0) {
    __threadfence_system(); // ... and has landed in the neighbour before the flag does
    const uint32_t value = s.signal_value;
#pragma unroll
    for (int f = 0; f < 6; ++f)
      if ((mine >> f & 1u) && s.signal_row[f])
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(s.signal_row[f] + group_of<SHIFT>(f, bx, by, bz, g)), "r"(value) : "memory");
  }
}

// One thread per cell: thin regions (the +-x exterior slabs are 1..r cells wide in x).
// Threads run fastest along y there so a warp's accesses land in 32 different rows but the
// same few sectors per row; z-neighbours of consecutive planes hit L2.
template <typename T> __global__ void __launch_bounds__(256) jacobi_cell_kernel(const __grid_constant__ JacobiParams p) {
  const int ex = p.hi[0] - p.lo[0], ey = p.hi[1] - p.lo[1], ez = p.hi[2] - p.lo[2];
  const long long total = (long long)ex * ey * ez;
  const int rr = (p.rad + 1) * (p.rad + 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // y fastest, then x, then z
    const int yy = (int)(i % ey);
    const long long t = i / ey;
    const int xx = (int)(t % ex);
    const int zz = (int)(t / ex);
    const int x = p.lo[0] + xx, y = p.lo[1] + yy, z = p.lo[2] + zz;
    const char *c = p.src + (long long)z * p.slice + (long long)y * p.pitch + (long long)x * (long long)sizeof(T);
    const T px = *reinterpret_cast<const T *>(c + sizeof(T));
    const T mx = *reinterpret_cast<const T *>(c - sizeof(T));
    const T py = *reinterpret_cast<const T *>(c + p.pitch);
    const T my = *reinterpret_cast<const T *>(c - p.pitch);
    const T pz = *reinterpret_cast<const T *>(c + p.slice);
    const T mz = *reinterpret_cast<const T *>(c - p.slice);
    T val = stencil_value<T>(px, mx, py, my, pz, mz);
    const int gx = x + p.org[0], gy = y + p.org[1], gz = z + p.org[2];
    const int dyz2 = (gy - p.cy) * (gy - p.cy) + (gz - p.cz) * (gz - p.cz);
    if (dyz2 < rr) {
      const int dh = (gx - p.hot_x) * (gx - p.hot_x) + dyz2;
      const int dc = (gx - p.cold_x) * (gx - p.cold_x) + dyz2;
      if (in_sphere(dh, p.rad)) {
        val = T(1);
      } else if (in_sphere(dc, p.rad)) {
        val = T(0);
      }
    }
    *reinterpret_cast<T *>(p.dst + (long long)z * p.slice + (long long)y * p.pitch + (long long)x * (long long)sizeof(T)) = val;
  }
}

// All exterior slabs of a subdomain in one launch (bin/jacobi3d.cu:324-342 issues up to six).  One thread
// per cell; inside a region threads run fastest along x when the slab is wide in x (coalesced rows)
// and along y when it is an x-face (1..r cells wide), so a warp always touches few sectors per row.
template <typename T>
__global__ void __launch_bounds__(256) jacobi_regions_kernel(const __grid_constant__ JacobiParams p, const __grid_constant__ JacobiRegions rg) {
  const long long total = rg.first[rg.n];
  const int rr = (p.rad + 1) * (p.rad + 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (k + 1 < rg.n && i >= rg.first[k + 1]) ++k;
    const unsigned li = (unsigned)(i - rg.first[k]);
    const unsigned ex = rg.ext[k][0], ey = rg.ext[k][1];
    unsigned xx, yy, zz;
    if (ex >= 32) { // x fastest
      xx = li % ex;
      const unsigned t = li / ex;
      yy = t % ey;
      zz = t / ey;
    } else { // y fastest
      yy = li % ey;
      const unsigned t = li / ey;
      xx = t % ex;
      zz = t / ex;
    }
    const int x = rg.lo[k][0] + (int)xx, y = rg.lo[k][1] + (int)yy, z = rg.lo[k][2] + (int)zz;
    const long long off = (long long)z * p.slice + (long long)y * p.pitch + (long long)x * (long long)sizeof(T);
    const char *c = p.src + off;
    const T px = *reinterpret_cast<const T *>(c + sizeof(T));
    const T mx = *reinterpret_cast<const T *>(c - sizeof(T));
    const T py = *reinterpret_cast<const T *>(c + p.pitch);
    const T my = *reinterpret_cast<const T *>(c - p.pitch);
    const T pz = *reinterpret_cast<const T *>(c + p.slice);
    const T mz = *reinterpret_cast<const T *>(c - p.slice);
    T val = stencil_value<T>(px, mx, py, my, pz, mz);
    const int gx = x + p.org[0], gy = y + p.org[1], gz = z + p.org[2];
    const int dyz2 = (gy - p.cy) * (gy - p.cy) + (gz - p.cz) * (gz - p.cz);
    if (dyz2 < rr) {
      const int dh = (gx - p.hot_x) * (gx - p.hot_x) + dyz2;
      const int dc = (gx - p.cold_x) * (gx - p.cold_x) + dyz2;
      if (in_sphere(dh, p.rad)) {
        val = T(1);
      } else if (in_sphere(dc, p.rad)) {
        val = T(0);
      }
    }
    *reinterpret_cast<T *>(p.dst + off) = val;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(char *dst, long long pitch, long long slice, int lx, int ly, int lz, int ex,
                                                   int ey, int ez, T value) {
  const long long total = (long long)ex * ey * ez;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % ex);
    const long long t = i / ex;
    const int yy = (int)(t % ey);
    const int zz = (int)(t / ey);
    reinterpret_cast<T *>(dst + (long long)(lz + zz) * slice + (long long)(ly + yy) * pitch)[lx + xx] = value;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) sqdiff_kernel(const char *a, const char *b, long long pitch, long long slice, int lx,
                                                     int ly, int lz, int ex, int ey, int ez, double *out) {
  const long long total = (long long)ex * ey * ez;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % ex);
    const long long t = i / ex;
    const int yy = (int)(t % ey);
    const int zz = (int)(t / ey);
    const long long off = (long long)(lz + zz) * slice + (long long)(ly + yy) * pitch;
    const double d = (double)reinterpret_cast<const T *>(a + off)[lx + xx] - (double)reinterpret_cast<const T *>(b + off)[lx + xx];
    acc += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    atomicAdd(out, s);
  }
}

int env_int(const char *name, int dflt);

template <typename T, int VX, bool SHIFT> int launch_fused(const JacobiParams &p, FusedSync s, cudaStream_t stream) {
  const int x0a = p.x0a;
  const int tiles_x = (p.hi[0] - x0a + (SHIFT ? VX / 2 : 0) + 32 * VX - 1) / (32 * VX);
  const int ny = p.hi[1] - p.lo[1], nz = p.hi[2] - p.lo[2];
  const int tiles_z = (nz + p.zchunk - 1) / p.zchunk;
  const int tiles_y = (ny + 7) / 8;
  if ((long long)tiles_z * ((tiles_y + 7) / 8) > kMaxGroups || (long long)tiles_z * tiles_x > kMaxGroups || tiles_y > kMaxGroups) return -2;
  const long long blocks = (long long)tiles_x * tiles_y * tiles_z;
  // start in the middle of z when a z face waits for another rank (see the kernel's header)
  s.zrot = (s.wait_row[4] || s.wait_row[5]) ? tiles_z / 2 : 0;
  s.any_wait = 0;
  for (int f = 0; f < 6; ++f)
    if (s.wait_row[f]) s.any_wait = 1;
  if (p.xghost_ptr[0] || p.xghost_ptr[1])
    jacobi_fused_kernel<T, VX, SHIFT, 3><<<(unsigned)blocks, 256, 0, stream>>>(p, s, tiles_x, tiles_y, tiles_z);
  else
    jacobi_fused_kernel<T, VX, SHIFT, 2><<<(unsigned)blocks, 256, 0, stream>>>(p, s, tiles_x, tiles_y, tiles_z);
  return 1;
}

template <typename T, int VX, bool SHIFT> int launch_march(const JacobiParams &p, int ry, int mb, cudaStream_t stream) {
  const int x0a = p.x0a;
  // a shifted row starts VX/2 cells early, so the last tile must reach VX/2 cells further
  const int tiles_x = (p.hi[0] - x0a + (SHIFT ? VX / 2 : 0) + 32 * VX - 1) / (32 * VX);
  const int ny = p.hi[1] - p.lo[1], nz = p.hi[2] - p.lo[2];
  const int tiles_z = (nz + p.zchunk - 1) / p.zchunk;
  auto go = [&](auto kern, int RY) {
    const int tiles_y = (ny + 8 * RY - 1) / (8 * RY);
    const long long blocks = (long long)tiles_x * tiles_y * tiles_z;
    kern<<<(unsigned)blocks, 256, 0, stream>>>(p, tiles_x, tiles_y);
  };
  // (rows per warp, min CTAs/SM): fewer rows -> fewer registers -> more warps in flight
  if (SHIFT) {
    if (mb == 5)
      go(jacobi_march_kernel<T, VX, 1, 5, SHIFT>, 1);
    else
      go(jacobi_march_kernel<T, VX, 1, 4, SHIFT>, 1);
    return 1;
  }
  const int key = ry * 10 + mb;
  switch (key) {
  case 14:
    go(jacobi_march_kernel<T, VX, 1, 4, false>, 1);
    break;
  case 15:
    go(jacobi_march_kernel<T, VX, 1, 5, false>, 1);
    break;
  case 16:
    go(jacobi_march_kernel<T, VX, 1, 6, false>, 1);
    break;
  case 24:
    go(jacobi_march_kernel<T, VX, 2, 4, false>, 2);
    break;
  case 42:
    go(jacobi_march_kernel<T, VX, 4, 2, false>, 4);
    break;
  default:
    go(jacobi_march_kernel<T, VX, 2, 3, false>, 2);
    break;
  }
  return 1;
}

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

} // namespace

int launch_jacobi_regions(const JacobiParams &p, const JacobiRegions &r, int dtype_size, cudaStream_t stream) {
  const long long total = r.first[r.n];
  if (r.n <= 0 || total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  static const int per_sm = env_int("SB_JACOBI_EXT_CTAS_PER_SM", 32);
  if (blocks > 148ll * per_sm) blocks = 148ll * per_sm;
  if (dtype_size == 4)
    jacobi_regions_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(p, r);
  else
    jacobi_regions_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(p, r);
  return 1;
}

static int pick_vectors(JacobiParams &p, int dtype_size, bool allow_shift, bool *shift);

int launch_jacobi(const JacobiParams &p_in, int dtype_size, cudaStream_t stream) {
  JacobiParams p = p_in;
  const int ex = p.hi[0] - p.lo[0], ey = p.hi[1] - p.lo[1], ez = p.hi[2] - p.lo[2];
  if (ex <= 0 || ey <= 0 || ez <= 0) return 0;

  static const int ry = env_int("SB_JACOBI_RY", 1);
  static const int mb = env_int("SB_JACOBI_MB", 4);
  static const int zchunk_env = env_int("SB_JACOBI_ZCHUNK", 0);
  static const int pf = env_int("SB_JACOBI_PREFETCH", 2);
  static const int thin = env_int("SB_JACOBI_THIN_X", 16);

  if (ex < thin) {
    const long long total = (long long)ex * ey * ez;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dtype_size == 4)
      jacobi_cell_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else
      jacobi_cell_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(p);
    return 1;
  }

  // z chunk: enough CTAs for >= ~6 waves of 148 SMs x 4 resident CTAs, but chunks no shorter than 16 planes
  if (zchunk_env > 0) {
    p.zchunk = zchunk_env;
  } else if (p.zchunk <= 0) {
    p.zchunk = 32;
  }
  if (p.zchunk > ez) p.zchunk = ez;
  p.prefetch = pf;

  static const int allow_shift = env_int("SB_JACOBI_SHIFT", 1);
  bool shift = false;
  const int vx = pick_vectors(p, dtype_size, allow_shift != 0, &shift);
  if (dtype_size == 8) {
    if (vx == 2) return shift ? launch_march<double, 2, true>(p, ry, mb, stream) : launch_march<double, 2, false>(p, ry, mb, stream);
    return launch_march<double, 1, false>(p, ry, mb, stream);
  }
  if (vx == 4) return shift ? launch_march<float, 4, true>(p, ry, mb, stream) : launch_march<float, 4, false>(p, ry, mb, stream);
  if (vx == 2) return launch_march<float, 2, false>(p, ry, mb, stream);
  return launch_march<float, 1, false>(p, ry, mb, stream);
}

// Vector width and strip origin for a launch.  Rows whose pitch and slice are multiples of 16 bytes all have the phase
// of the allocation base: LocalDomain places the base so that the first COMPUTE cell is 16-byte aligned (the whole
// compute region of 512 cells is then exactly 8 strips of 32 lanes x 2 doubles); any other common phase of src and dst
// works as well.  Returns VX (1 = scalar fallback) and sets p.x0a; `shift` reports the alternating-phase case.
static int pick_vectors(JacobiParams &p, int dtype_size, bool allow_shift, bool *shift) {
  const unsigned long long s = (unsigned long long)(uintptr_t)p.src, d = (unsigned long long)(uintptr_t)p.dst;
  const int full = 16 / dtype_size;
  *shift = false;
  auto origin = [&](int vx, unsigned phase) {
    const int xal = int(((16u - phase) & 15u) / unsigned(dtype_size)) % vx; // x (mod vx) of the aligned cells
    int r = (p.lo[0] - xal) % vx;
    if (r < 0) r += vx;
    p.x0a = p.lo[0] - r;
    return vx;
  };
  if (p.pitch % 16 == 0 && p.slice % 16 == 0 && (s & 15) == (d & 15) && (s & 15) % dtype_size == 0) {
    // with a non-zero phase the first vector may start before x = 0: still inside the (256-byte aligned) allocation,
    // whose first `phase` bytes are the lead pad
    return origin(full, unsigned(s & 15));
  }
  if (allow_shift && (s & 15) == 0 && (d & 15) == 0 && p.slice % 16 == 0 && p.pitch % 16 == 8) {
    *shift = true;
    return origin(full, 0);
  }
  if (dtype_size == 4 && ((s | d | (unsigned long long)p.pitch | (unsigned long long)p.slice) % 8 == 0)) return origin(2, 0);
  return origin(1, 0);
}

// Arrival counters of the face groups: one zeroed block per (device, stream).  Launches on one stream are serialised and
// every group's last arrival resets its counter, so the block is all zero again whenever the next launch starts.
static uint32_t *group_counters(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, uint32_t *> pool;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  uint32_t *&c = pool[std::make_pair(dev, stream)];
  if (!c) {
    if (cudaMalloc(&c, 6 * kMaxGroups * sizeof(uint32_t)) != cudaSuccess) return nullptr;
    cudaMemsetAsync(c, 0, 6 * kMaxGroups * sizeof(uint32_t), stream);
  }
  return c;
}

int launch_jacobi_fused(const JacobiParams &p_in, const FusedSync &sync_in, int dtype_size, cudaStream_t stream) {
  JacobiParams p = p_in;
  const int ex = p.hi[0] - p.lo[0], ey = p.hi[1] - p.lo[1], ez = p.hi[2] - p.lo[2];
  if (ex <= 0 || ey <= 0 || ez <= 0) return 0;
  static const int zchunk_env = env_int("SB_JACOBI_ZCHUNK", 0);
  static const int pf = env_int("SB_JACOBI_PREFETCH", 2);
  static const int allow_shift = env_int("SB_JACOBI_SHIFT", 1);
  p.zchunk = zchunk_env > 0 ? zchunk_env : 32;
  if (p.zchunk > ez) p.zchunk = ez;
  p.prefetch = pf;
  FusedSync sync = sync_in;
  if (!sync.counters) sync.counters = group_counters(stream);
  if (!sync.counters) return -3;
  bool shift = false;
  const int vx = pick_vectors(p, dtype_size, allow_shift != 0, &shift);
  // x wrap (periodic self-neighbour read in place) works through the edge lanes' scalar load, so the first compute cell
  // must open lane 0 of the first strip and the last one must close lane 31 of the last strip; otherwise the ghost
  // column is read by a vector load or a shuffle and the x faces are pushed into the ghost cells like any other face
  if (p.xwrap && !(!shift && p.x0a == p.lo[0] && (p.hi[0] - p.lo[0]) % (32 * vx) == 0)) p.xwrap = 0;
  if (p.xwrap) p.push_ptr[0] = p.push_ptr[1] = nullptr;
  if (p.xghost_ptr[0] || p.xghost_ptr[1]) {
    // a dense received x array is read by the edge lanes' scalar load: same layout conditions as the wrap
    if (shift || p.x0a != p.lo[0] || (p.hi[0] - p.lo[0]) % (32 * vx) != 0) return -1;
  }
  if (dtype_size == 8) {
    if (vx == 2) return shift ? launch_fused<double, 2, true>(p, sync, stream) : launch_fused<double, 2, false>(p, sync, stream);
    return launch_fused<double, 1, false>(p, sync, stream);
  }
  if (vx == 4) return shift ? launch_fused<float, 4, true>(p, sync, stream) : launch_fused<float, 4, false>(p, sync, stream);
  if (vx == 2) return launch_fused<float, 2, false>(p, sync, stream);
  return launch_fused<float, 1, false>(p, sync, stream);
}

int launch_fill(char *dst, long long pitch, long long slice, const int lo[3], const int hi[3], int dtype_size, double value,
                cudaStream_t stream) {
  const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  if (ex <= 0 || ey <= 0 || ez <= 0) return 0;
  long long blocks = ((long long)ex * ey * ez + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype_size == 4)
    fill_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(dst, pitch, slice, lo[0], lo[1], lo[2], ex, ey, ez, (float)value);
  else
    fill_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(dst, pitch, slice, lo[0], lo[1], lo[2], ex, ey, ez, value);
  return 1;
}

int launch_sqdiff(const char *a, const char *b, long long pitch, long long slice, const int lo[3], const int hi[3],
                  int dtype_size, double *out_dev, cudaStream_t stream) {
  cudaMemsetAsync(out_dev, 0, sizeof(double), stream);
  const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  if (ex <= 0 || ey <= 0 || ez <= 0) return 0;
  long long blocks = ((long long)ex * ey * ez + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype_size == 4)
    sqdiff_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(a, b, pitch, slice, lo[0], lo[1], lo[2], ex, ey, ez, out_dev);
  else
    sqdiff_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(a, b, pitch, slice, lo[0], lo[1], lo[2], ex, ey, ez, out_dev);
  return 1;
}

} // namespace sb
