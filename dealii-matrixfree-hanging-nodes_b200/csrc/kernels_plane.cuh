// Register-tiled fused cell kernel for Cartesian cells (the fast path).
//
// Work decomposition: n = k+1 threads per cell, 32/n cells per warp, one warp
// per batch of cells; a warp never synchronises with another warp.  Each
// thread owns one n x n plane of the cell's n^3 values in registers:
//
//   P1 (thread = z, plane (x,y)):  gather -> [hanging-node interpolation: in-place
//        directional passes on a shared-memory copy, only in warps with a
//        constrained cell] -> a = M_y M_x u,  b = (M_y K_x + K_y M_x) u  -> shared memory
//   P2 (thread = x, plane (y,z)):  r = h (M_z b + K_z a)      -> shared memory
//   P3 (thread = z, plane (x,y)):  interpolation^T -> atomic scatter-add
//
// On a Cartesian cell the Laplace cell matrix produced by the reference's
// evaluate / submit_gradient / integrate sequence with QGauss(k+1)
// (benchmark_03.h:305-312) is exactly h (K x M x M + M x K x M + M x M x K) with
// the 1D Gauss-integrated mass M and stiffness K; applying it in this form
// needs 7 one-dimensional sweeps instead of 12 and no quadrature-point data.
// M and K are persymmetric, so every sweep uses the even-odd decomposition.
//
// The DoF indices are stored warp-interleaved ([batch][plane slot][lane]) so
// that every index load is one coalesced 128-byte request.  The kernel's thread
// axis Z is the physical x direction (kernel axes (X,Y,Z) = physical (y,z,x)):
// the threads of a cell then read consecutive x, the most contiguous direction
// of deal.II's object-wise DoF numbering (fewer 128-byte lines per request).
#pragma once
#include "kernels_generic.cuh"
#include "shape_tables.cuh"

#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <vector>

namespace mfhn
{
constexpr int round_up_mod(int v, int r, int mod)
{
  while (v % mod != r % mod) ++v;
  return v;
}

template <int n, typename Number>
struct PlaneCfg
{
  static constexpr int cpw   = 32 / n;     // cells per warp
  static constexpr int lanes = cpw * n;    // active lanes
  static constexpr int mod   = sizeof(Number) == 8 ? 16 : 32;
  // plane / cell strides chosen so that lane (cell c, thread t) hits bank
  // (c n + t) mod `mod` in every phase: conflict-free shared memory traffic
  static constexpr int ps    = round_up_mod(n * n, 1, mod);
  static constexpr int cs    = round_up_mod(n * ps, n, mod);
  static constexpr int warps = 4;
  static constexpr int smem_per_warp = cpw * cs * (int)sizeof(Number); // one array: a and b cross it one after the other
  static constexpr int smem  = warps * smem_per_warp;
};

constexpr bool plane_supported(int n) { return n >= 2 && n <= 6; }

struct PlaneParams
{
  const uint32_t *pidx; // [n_batches][n*n][32]
  const uint8_t *masks; // [n_cells]
  const void *h;        // Number[n_cells]
  const void *src;
  void *dst;
  long long cell_begin, cell_end, batch_begin, batch_end;
  int apply_constraints;
  cudaTextureObject_t src_tex; // src bound as a linear texture: gathers go through the TEX pipe
  // peer mode (boundary cells of a partitioned operator): ghost entries (index >= n_owned) are read from /
  // added to the OWNER's vectors through peer-mapped pointers over NVLink instead of a local ghost section
  long long n_owned;
  const void *const *ghost_src; // [n_ghost] address of the entry in the owner's src
  void *const *ghost_dst;       // [n_ghost] address of the entry in the owner's dst
};

template <typename Number>
__device__ __forceinline__ Number tex_fetch(cudaTextureObject_t tex, uint32_t i);
template <>
__device__ __forceinline__ double tex_fetch<double>(cudaTextureObject_t tex, uint32_t i)
{
  const int2 v = tex1Dfetch<int2>(tex, (int)i);
  return __hiloint2double(v.y, v.x);
}
template <>
__device__ __forceinline__ float tex_fetch<float>(cudaTextureObject_t tex, uint32_t i)
{
  return tex1Dfetch<float>(tex, (int)i);
}

// ---- even-odd application of a persymmetric matrix --------------------------
template <int n, typename Number>
__device__ __forceinline__ void eo_split(const Number (&x)[n], Number (&xs)[(n + 1) / 2], Number (&xd)[(n + 1) / 2])
{
  constexpr int h = n / 2;
#pragma unroll
  for (int j = 0; j < h; ++j)
    {
      xs[j] = x[j] + x[n - 1 - j];
      xd[j] = x[j] - x[n - 1 - j];
    }
  if (n % 2)
    {
      xs[h] = x[h];
      xd[h] = Number(0);
    }
}
template <int n, int TE, int TO, bool INIT, typename Number>
__device__ __forceinline__ void eo_mac(const Number (&xs)[(n + 1) / 2], const Number (&xd)[(n + 1) / 2],
                                       Number (&se)[(n + 1) / 2], Number (&so)[(n + 1) / 2])
{
  constexpr int h = n / 2, he = (n + 1) / 2;
#pragma unroll
  for (int i = 0; i < he; ++i)
#pragma unroll
    for (int j = 0; j < he; ++j)
      {
        const Number c = Shape<Number>::template eo<n, TE>(i * he + j);
        se[i]          = (INIT && j == 0) ? c * xs[j] : se[i] + c * xs[j];
      }
#pragma unroll
  for (int i = 0; i < h; ++i)
#pragma unroll
    for (int j = 0; j < h; ++j)
      {
        const Number c = Shape<Number>::template eo<n, TO>(i * he + j);
        so[i]          = (INIT && j == 0) ? c * xd[j] : so[i] + c * xd[j];
      }
}
template <int n, typename Number>
__device__ __forceinline__ void eo_merge(const Number (&se)[(n + 1) / 2], const Number (&so)[(n + 1) / 2], Number (&y)[n])
{
  constexpr int h = n / 2;
#pragma unroll
  for (int i = 0; i < h; ++i)
    {
      y[i]         = se[i] + so[i];
      y[n - 1 - i] = se[i] - so[i];
    }
  if (n % 2) y[h] = se[h];
}

// p = M x, q = K x
template <int n, typename Number>
__device__ __forceinline__ void apply_MK(const Number (&x)[n], Number (&p)[n], Number (&q)[n])
{
  if (n >= 4)
    {
      Number xs[(n + 1) / 2], xd[(n + 1) / 2], se[(n + 1) / 2], so[(n + 1) / 2];
      eo_split<n>(x, xs, xd);
      eo_mac<n, T_ME, T_MO, true>(xs, xd, se, so);
      eo_merge<n>(se, so, p);
      eo_mac<n, T_KE, T_KO, true>(xs, xd, se, so);
      eo_merge<n>(se, so, q);
    }
  else
    {
      mat_vec<n, T_M, false>(x, p);
      mat_vec<n, T_K, false>(x, q);
    }
}
// a = M p, b = M q + K p
template <int n, typename Number>
__device__ __forceinline__ void apply_M_MK(const Number (&p)[n], const Number (&q)[n], Number (&a)[n], Number (&b)[n])
{
  if (n >= 4)
    {
      constexpr int he = (n + 1) / 2;
      Number ps[he], pd[he], qs[he], qd[he], se[he], so[he];
      eo_split<n>(p, ps, pd);
      eo_split<n>(q, qs, qd);
      eo_mac<n, T_ME, T_MO, true>(ps, pd, se, so);
      eo_merge<n>(se, so, a);
      eo_mac<n, T_ME, T_MO, true>(qs, qd, se, so);
      eo_mac<n, T_KE, T_KO, false>(ps, pd, se, so);
      eo_merge<n>(se, so, b);
    }
  else
    {
      Number t[n];
      mat_vec<n, T_M, false>(p, a);
      mat_vec<n, T_M, false>(q, b);
      mat_vec<n, T_K, false>(p, t);
#pragma unroll
      for (int i = 0; i < n; ++i) b[i] += t[i];
    }
}
// r = M b + K a
template <int n, typename Number>
__device__ __forceinline__ void apply_Mb_Ka(const Number (&a)[n], const Number (&b)[n], Number (&r)[n])
{
  if (n >= 4)
    {
      constexpr int he = (n + 1) / 2;
      Number as[he], ad[he], bs[he], bd[he], se[he], so[he];
      eo_split<n>(a, as, ad);
      eo_split<n>(b, bs, bd);
      eo_mac<n, T_ME, T_MO, true>(bs, bd, se, so);
      eo_mac<n, T_KE, T_KO, false>(as, ad, se, so);
      eo_merge<n>(se, so, r);
    }
  else
    {
      Number t[n];
      mat_vec<n, T_M, false>(b, r);
      mat_vec<n, T_K, false>(a, t);
#pragma unroll
      for (int i = 0; i < n; ++i) r[i] += t[i];
    }
}

// In-place hanging-node interpolation (or its transpose) on the cell arrays of
// one warp: three directional passes over kernel axes (X,Y,Z) = shared-memory
// strides (1, n, ps).  face / edge / cb are the constraint bits already permuted
// to kernel axes.  In a pass the n^2 lines are shared by the n threads of the
// cell; the assignment is chosen per cell so that the lines of a constrained
// face land on n different threads of the same iteration.
template <int n, bool transpose, typename Number>
__device__ __forceinline__ void hn_smem(Number *cellA, unsigned face, unsigned edge, unsigned cb, int t)
{
  constexpr int k = n - 1;
  using Cfg = PlaneCfg<n, Number>;
#pragma unroll 1
  for (int d = 0; d < 3; ++d)
    {
      const int t0 = (d == 0) ? 1 : 0, t1 = (d == 2) ? 1 : 2;
      const int c0 = (int)((cb >> t0) & 1u) * k, c1 = (int)((cb >> t1) & 1u) * k;
      const bool f0 = (face >> t0) & 1u, f1 = (face >> t1) & 1u, ed = (edge >> d) & 1u;
      const bool work = f0 || f1 || ed;
      const bool upper = (cb >> d) & 1u;
      const int stride = d == 0 ? 1 : d == 1 ? n : Cfg::ps;
      // f1 (lines b == c1, all a): thread = a, iteration = b; otherwise thread = b, iteration = a
      const int last = !work ? 0 : (f0 && f1) ? n : 1; // a single face / edge needs one iteration
      const int first = f1 ? c1 : c0;
#pragma unroll 1
      for (int it = 0; it < last; ++it)
        {
          const int i = (first + it) % n; // start with the iteration that holds the whole face
          const int a = f1 ? t : i, b = f1 ? i : t;
          const bool on0 = a == c0, on1 = b == c1;
          const bool sel = (f0 && on0) || (f1 && on1) || (ed && on0 && on1);
          if (sel)
            {
              const int base = d == 0 ? b * Cfg::ps + a * n : d == 1 ? b * Cfg::ps + a : b * n + a;
              Number *line   = cellA + base;
              Number v[n], w[n];
#pragma unroll
              for (int i2 = 0; i2 < n; ++i2) v[i2] = line[(upper ? k - i2 : i2) * stride];
              mat_vec<n, T_W0, transpose>(v, w);
#pragma unroll
              for (int i2 = 0; i2 < n; ++i2) line[(upper ? k - i2 : i2) * stride] = w[i2];
            }
        }
      __syncwarp(); // the next pass reads lines written by other threads
    }
}

// Constraint bits of a compressed mask in KERNEL axes.  The plane / patch kernels
// use x as the thread axis (threads of a cell read consecutive x: the most
// contiguous direction of deal.II's object-wise numbering), i.e. kernel axes
// (X,Y,Z) = physical (y,z,x).
__device__ __forceinline__ unsigned rot3(unsigned b) { return ((b >> 1) | (b << 2)) & 7u; }
__device__ __forceinline__ void decode_mask_kernel_axes(unsigned m, unsigned &face, unsigned &edge, unsigned &cb)
{
  decode_mask(m, face, edge, cb);
  face = rot3(face);
  edge = rot3(edge);
  cb   = rot3(cb);
}

// The seven sweeps of one cell: u = this thread's plane [Y][X] (thread = Z); on return the warp's shared-memory
// array holds h (K x M x M + M x K x M + M x M x K) u in the layout cellA[Z * ps + Y * n + X] (warp-synchronised).
template <int n, typename Number>
__device__ __forceinline__ void plane_sweeps(const Number (&u)[n][n], Number *cellA, const int t, const Number h)
{
  using Cfg = PlaneCfg<n, Number>;
  constexpr int ps = Cfg::ps;
  // ---- P1: x and y sweeps (thread = Z) ----------------------------------------------
  // a = M_Y M_X u and b = (M_Y K_X + K_Y M_X) u cross the transpose through the same
  // shared-memory array one after the other (half the shared memory, more warps per SM)
  Number az[n][n], bz[n][n]; // P2 operands of this thread: [Y][Z]
  {
    Number bb[n][n];
    {
      Number pp[n][n], qq[n][n];
#pragma unroll
      for (int y = 0; y < n; ++y) apply_MK<n>(u[y], pp[y], qq[y]);
#pragma unroll
      for (int x = 0; x < n; ++x)
        {
          Number pc[n], qc[n], a[n], b[n];
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              pc[i] = pp[i][x];
              qc[i] = qq[i][x];
            }
          apply_M_MK<n>(pc, qc, a, b);
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              cellA[t * ps + i * n + x] = a[i];
              bb[i][x]                  = b[i];
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int y = 0; y < n; ++y)
#pragma unroll
      for (int z = 0; z < n; ++z) az[y][z] = cellA[z * ps + y * n + t];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < n * n; ++j) cellA[t * ps + j] = bb[j / n][j % n];
    __syncwarp();
#pragma unroll
    for (int y = 0; y < n; ++y)
#pragma unroll
      for (int z = 0; z < n; ++z) bz[y][z] = cellA[z * ps + y * n + t];
  }
  // ---- P2: Z sweep (thread = X) ------------------------------------------------------
#pragma unroll
  for (int y = 0; y < n; ++y)
    {
      Number r[n];
      apply_Mb_Ka<n>(az[y], bz[y], r);
#pragma unroll
      for (int z = 0; z < n; ++z) cellA[z * ps + y * n + t] = h * r[z];
    }
  __syncwarp();
}

template <int n, typename Number, bool TEX, bool PEER = false>
__global__ void __launch_bounds__(PlaneCfg<n, Number>::warps * 32, (n <= 5 ? 4 : 3)) plane_cell_kernel(const PlaneParams p)
{
  using Cfg = PlaneCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = p.batch_begin + (long long)blockIdx.x * Cfg::warps + warp;
  if (batch >= p.batch_end) return; // warps are independent: no block-level barrier below
  Number *A = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * Cfg::cpw * cs;

  // the 32 - cpw n idle lanes mirror lane - 16 (same loads, same values stored to the same
  // shared-memory addresses; that lane sits in the other half-warp, so no bank conflict arises),
  // hence the arithmetic needs no per-lane predicate; they do not scatter
  const bool active = lane < Cfg::lanes;
  const int ml = active ? lane : lane - 16;
  const int c = ml / n, t = ml - c * n;
  const long long cell = batch * Cfg::cpw + c;
  const bool valid = cell >= p.cell_begin && cell < p.cell_end;
  const Number *__restrict__ src = static_cast<const Number *>(p.src);
  Number *__restrict__ dst = static_cast<Number *>(p.dst);
  const uint32_t *ip = p.pidx + batch * (long long)(n * n * 32) + (c * n + t);
  Number *cellA = A + c * cs;

  // ---- P1: gather (thread = z, plane (x,y)) ---------------------------------------
  Number u[n][n];
  {
    uint32_t idx[n * n];
#pragma unroll
    for (int j = 0; j < n * n; ++j) idx[j] = valid ? __ldg(ip + j * 32) : 0u; // every entry of a valid cell is a valid index
#pragma unroll
    for (int j = 0; j < n * n; ++j)
      {
        if (PEER && valid && idx[j] >= (uint32_t)p.n_owned) // remote entry: plain load through the peer mapping
          u[j / n][j % n] = *static_cast<const Number *>(p.ghost_src[idx[j] - (uint32_t)p.n_owned]);
        else
          u[j / n][j % n] = valid ? (TEX ? tex_fetch<Number>(p.src_tex, idx[j]) : __ldg(src + idx[j])) : Number(0);
      }
  }
  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
  const bool any_hn   = __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);
  if (any_hn)
    {
      // hanging-node interpolation as in-place directional passes on the shared-memory copy
#pragma unroll
      for (int j = 0; j < n * n; ++j) cellA[t * ps + j] = u[j / n][j % n];
      __syncwarp();
      hn_smem<n, false>(cellA, hn_face, hn_edge, hn_cb, t);
#pragma unroll
      for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
      __syncwarp();
    }

  plane_sweeps<n>(u, cellA, t, h);
  if (any_hn) hn_smem<n, true>(cellA, hn_face, hn_edge, hn_cb, t);
  // ---- P3: scatter (thread = z) --------------------------------------------------------
  if (active && valid)
    {
#pragma unroll
      for (int j = 0; j < n * n; ++j)
        {
          const uint32_t g = __ldg(ip + j * 32);
          if (PEER && g >= (uint32_t)p.n_owned) // red over NVLink into the owner's dst
            atomicAdd(static_cast<Number *>(p.ghost_dst[g - (uint32_t)p.n_owned]), cellA[t * ps + j]);
          else
            atomicAdd(dst + g, cellA[t * ps + j]);
        }
    }
}

// ---- host side ------------------------------------------------------------------
struct PlaneLayout
{
  int n = 0;
  long long n_cells = 0, n_batches = 0;
  uint32_t *d_pidx = nullptr;

  void free()
  {
    cudaFree(d_pidx);
    d_pidx = nullptr;
  }
  void build(int n_, long long n_cells_, const uint32_t *idx);
};

struct PeerTables
{
  long long n_owned           = 0;
  const void *const *ghost_src = nullptr;
  void *const *ghost_dst       = nullptr;
};

template <int n, typename Number>
void launch_plane_impl(const PlaneLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream, cudaTextureObject_t tex,
                       const PeerTables *peer = nullptr)
{
  using Cfg = PlaneCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(plane_cell_kernel<n, Number, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(plane_cell_kernel<n, Number, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
      if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      attr[device] = true;
    }
  PlaneParams p;
  p.pidx              = L.d_pidx;
  p.masks             = cp.masks;
  p.h                 = cp.geom;
  p.src               = cp.src;
  p.dst               = cp.dst;
  p.cell_begin        = cp.cell_begin;
  p.cell_end          = cp.cell_end;
  p.batch_begin       = cp.cell_begin / Cfg::cpw;
  p.batch_end         = (cp.cell_end + Cfg::cpw - 1) / Cfg::cpw;
  p.apply_constraints = cp.apply_constraints;
  p.src_tex           = tex;
  p.n_owned           = peer ? peer->n_owned : 0;
  p.ghost_src         = peer ? peer->ghost_src : nullptr;
  p.ghost_dst         = peer ? peer->ghost_dst : nullptr;
  const long long nb  = p.batch_end - p.batch_begin;
  if (nb <= 0) return;
  const unsigned grid = (unsigned)((nb + Cfg::warps - 1) / Cfg::warps);
  if (peer)
    {
      static bool pattr[64] = {};
      if (!pattr[device])
        {
          cudaError_t e = cudaFuncSetAttribute(plane_cell_kernel<n, Number, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
          if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
          pattr[device] = true;
        }
      plane_cell_kernel<n, Number, false, true><<<grid, Cfg::warps * 32, Cfg::smem, stream>>>(p);
    }
  else if (tex)
    plane_cell_kernel<n, Number, true><<<grid, Cfg::warps * 32, Cfg::smem, stream>>>(p);
  else
    plane_cell_kernel<n, Number, false><<<grid, Cfg::warps * 32, Cfg::smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("plane kernel launch: ") + cudaGetErrorString(e));
}

template <int n, typename Number>
void launch_plane(const PlaneLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream, cudaTextureObject_t tex,
                  const PeerTables *peer = nullptr)
{
  if constexpr (plane_supported(n))
    launch_plane_impl<n, Number>(L, cp, device, stream, tex, peer);
  else
    throw std::runtime_error("plane kernel not available for this degree");
}
} // namespace mfhn
