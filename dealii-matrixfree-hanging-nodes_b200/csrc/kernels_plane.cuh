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
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef MFHN_PLANE_DEFAULT_OCC
#define MFHN_PLANE_DEFAULT_OCC 4
#endif
#ifndef MFHN_HN_INLINE
#define MFHN_HN_INLINE __noinline__
#endif
// (-DMFHN_HN_INLINE=__forceinline__ inlines the hanging-node passes into the cell kernels)

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


struct PlaneParams
{
  const uint32_t *pidx; // [n_batches][n*n][32]
  const uint8_t *masks; // [n_cells]
  const void *h;        // Number[n_cells]
  const void *src;
  void *dst;
  long long cell_begin, cell_end, batch_begin, batch_end;
  int apply_constraints;
  int hn_mask_strategy; // every warp takes the interpolation passes
  // peer mode (boundary cells of a partitioned operator): ghost entries (index >= n_owned) are read from /
  // added to the OWNER's vectors through peer-mapped pointers over NVLink instead of a local ghost section
  long long n_owned;
  const void *const *ghost_src; // [n_ghost] address of the entry in the owner's src
  void *const *ghost_dst;       // [n_ghost] address of the entry in the owner's dst
};

// ---- fast diagonalisation of the Cartesian cell matrix -------------------------
// With M = T^T T and K = T^T diag(lambda) T (fe1d.hpp)
//   h (K x M x M + M x K x M + M x M x K) = h (T x T x T)^T diag(lambda_i + lambda_j + lambda_k) (T x T x T):
// three forward sweeps with T, one scaling, three backward sweeps with T^T -- six one-matrix sweeps, and only
// ONE array crosses each change of the thread axis.  The rows of T are symmetric / antisymmetric, so a sweep is
// an even-odd product: (n+1)/2 x (n+1)/2 + n/2 x n/2 multiply-adds per line.
template <int n, typename Number>
__device__ __forceinline__ void fdm_fwd(const Number (&x)[n], Number (&y)[n]) // y = T x
{
  constexpr int h = n / 2, he = (n + 1) / 2;
  Number xs[he], xd[he];
#pragma unroll
  for (int j = 0; j < h; ++j)
    {
      xs[j] = x[j] + x[n - 1 - j];
      xd[j] = x[j] - x[n - 1 - j];
    }
  if (n % 2) xs[h] = x[h];
  Number r[n];
#pragma unroll
  for (int i = 0; i < he; ++i)
    {
      Number s = Shape<Number>::template eo<n, T_TE>(i * he) * xs[0];
#pragma unroll
      for (int j = 1; j < he; ++j) s += Shape<Number>::template eo<n, T_TE>(i * he + j) * xs[j];
      r[i] = s;
    }
#pragma unroll
  for (int i = 0; i < h; ++i)
    {
      Number s = Shape<Number>::template eo<n, T_TO>(i * he) * xd[0];
#pragma unroll
      for (int j = 1; j < h; ++j) s += Shape<Number>::template eo<n, T_TO>(i * he + j) * xd[j];
      r[he + i] = s;
    }
#pragma unroll
  for (int i = 0; i < n; ++i) y[i] = r[i];
}
template <int n, typename Number>
__device__ __forceinline__ void fdm_bwd(const Number (&w)[n], Number (&y)[n]) // y = T^T w
{
  constexpr int h = n / 2, he = (n + 1) / 2;
  Number e[he], o[he];
#pragma unroll
  for (int j = 0; j < he; ++j)
    {
      Number s = Shape<Number>::template eo<n, T_TE>(j) * w[0];
#pragma unroll
      for (int i = 1; i < he; ++i) s += Shape<Number>::template eo<n, T_TE>(i * he + j) * w[i];
      e[j] = s;
    }
#pragma unroll
  for (int j = 0; j < h; ++j)
    {
      Number s = Shape<Number>::template eo<n, T_TO>(j) * w[he];
#pragma unroll
      for (int i = 1; i < h; ++i) s += Shape<Number>::template eo<n, T_TO>(i * he + j) * w[he + i];
      o[j] = s;
    }
#pragma unroll
  for (int j = 0; j < h; ++j)
    {
      y[j]         = e[j] + o[j];
      y[n - 1 - j] = e[j] - o[j];
    }
  if (n % 2) y[h] = e[h];
}
// One line along the cross-thread axis: forward, eigenvalue scaling, backward.  hl = h (lambda_a + lambda_b) of
// the two other indices; the eigenvalue of the line's own index is a compile-time constant.
template <int n, typename Number>
__device__ __forceinline__ void fdm_mid(Number (&v)[n], const Number h, const Number hl)
{
  Number w[n];
  fdm_fwd<n>(v, w);
#pragma unroll
  for (int z = 0; z < n; ++z) w[z] *= h * Shape<Number>::template lam<n>(z) + hl;
  fdm_bwd<n>(w, v);
}

// In-place hanging-node interpolation (or its transpose) on the cell arrays of
// one warp: three directional passes over kernel axes (X,Y,Z) = shared-memory
// strides (1, n, ps).  face / edge / cb are the constraint bits already permuted
// to kernel axes.  In a pass the n^2 lines are shared by the n threads of the
// cell; the assignment is chosen per cell so that the lines of a constrained
// face land on n different threads of the same iteration.
template <int n, bool transpose, typename Number>
__device__ MFHN_HN_INLINE void hn_smem(Number *cellA, unsigned face, unsigned edge, unsigned cb, int t, const bool active = true)
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
          // idle lanes mirror an active lane's coordinates: they must not repeat its in-place update
          const bool sel = active && ((f0 && on0) || (f1 && on1) || (ed && on0 && on1));
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

// The six sweeps of one cell (fast diagonalisation).  u = this thread's plane [Y][X] (thread = Z); on return u
// holds h (K x M x M + M x K x M + M x M x K) u in the same layout.  The warp's shared-memory array carries the
// plane across the two changes of the thread axis (Z -> X -> Z); idle lanes compute along but do not store.
template <int n, typename Number>
__device__ __forceinline__ void plane_sweeps(Number (&u)[n][n], Number *cellA, const int t, const Number h, const bool active)
{
  using Cfg = PlaneCfg<n, Number>;
  constexpr int ps = Cfg::ps;
  // ---- P1 (thread = Z): forward sweeps along X (rows) and Y (columns) ----------------
#pragma unroll
  for (int y = 0; y < n; ++y) fdm_fwd<n>(u[y], u[y]);
#pragma unroll
  for (int x = 0; x < n; ++x)
    {
      Number c[n];
#pragma unroll
      for (int y = 0; y < n; ++y) c[y] = u[y][x];
      fdm_fwd<n>(c, c);
      if (active)
        {
#pragma unroll
          for (int y = 0; y < n; ++y) cellA[t * ps + y * n + x] = c[y];
        }
    }
  __syncwarp();
  // ---- P2 (thread = X): forward along Z, eigenvalue scaling, backward along Z ---------
  const Number hlt = h * Shape<Number>::template lam<n>(t); // runtime index: one constant-memory load
#pragma unroll
  for (int y = 0; y < n; ++y)
    {
      Number v[n];
#pragma unroll
      for (int z = 0; z < n; ++z) v[z] = cellA[z * ps + y * n + t];
      fdm_mid<n>(v, h, h * Shape<Number>::template lam<n>(y) + hlt);
      if (active)
        {
#pragma unroll
          for (int z = 0; z < n; ++z) cellA[z * ps + y * n + t] = v[z]; // only this thread read these entries
        }
    }
  __syncwarp();
  // ---- P3 (thread = Z): backward sweeps along Y and X ---------------------------------
#pragma unroll
  for (int x = 0; x < n; ++x)
    {
      Number c[n];
#pragma unroll
      for (int y = 0; y < n; ++y) c[y] = cellA[t * ps + y * n + x];
      fdm_bwd<n>(c, c);
#pragma unroll
      for (int y = 0; y < n; ++y) u[y][x] = c[y];
    }
#pragma unroll
  for (int y = 0; y < n; ++y) fdm_bwd<n>(u[y], u[y]);
}

// OCC = CTAs per SM the register allocation is limited for (4: 128, 5: 96, 6: 80 registers per thread)
template <int n, typename Number, bool PEER = false, int OCC = 4>
__global__ void __launch_bounds__(PlaneCfg<n, Number>::warps * 32, OCC) plane_cell_kernel(const PlaneParams p)
{
  using Cfg = PlaneCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = p.batch_begin + (long long)blockIdx.x * Cfg::warps + warp;
  if (batch >= p.batch_end) return; // warps are independent: no block-level barrier below
  Number *A = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * Cfg::cpw * cs;

  // the 32 - cpw n idle lanes mirror lane - 16 (same loads and arithmetic, so the warp needs no per-lane
  // predicate around the sweeps); they neither store to shared memory nor scatter
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
          u[j / n][j % n] = valid ? __ldg(src + idx[j]) : Number(0);
      }
  }
  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
  const bool any_hn   = p.hn_mask_strategy || __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);
  if (any_hn)
    {
      // hanging-node interpolation as in-place directional passes on the shared-memory copy
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n * n; ++j) cellA[t * ps + j] = u[j / n][j % n];
        }
      __syncwarp();
      hn_smem<n, false>(cellA, hn_face, hn_edge, hn_cb, t, active);
#pragma unroll
      for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
      __syncwarp();
    }

  plane_sweeps<n>(u, cellA, t, h, active);
  if (any_hn)
    {
      __syncwarp(); // every lane has read its plane back
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n * n; ++j) cellA[t * ps + j] = u[j / n][j % n];
        }
      __syncwarp();
      hn_smem<n, true>(cellA, hn_face, hn_edge, hn_cb, t, active);
#pragma unroll
      for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
    }
  // ---- P3: scatter (thread = z), straight from registers --------------------------------
  if (active && valid)
    {
#pragma unroll
      for (int j = 0; j < n * n; ++j)
        {
          const uint32_t g = __ldg(ip + j * 32);
          if (PEER && g >= (uint32_t)p.n_owned) // red over NVLink into the owner's dst
            atomicAdd(static_cast<Number *>(p.ghost_dst[g - (uint32_t)p.n_owned]), u[j / n][j % n]);
          else
            atomicAdd(dst + g, u[j / n][j % n]);
        }
    }
}

// ---- host side ------------------------------------------------------------------
// CTAs per SM to compile for: the double-precision kernels of degree >= 4 are register-bound, the others fit 4+ CTAs anyway
inline int occupancy_choice(const int n, const bool f64, const int fallback)
{
  const char *e  = std::getenv("MFHN_OCC"); // experiments only (read per launch so that one process can compare)
  const int env = e ? std::atoi(e) : 0;
  if (!(f64 && n >= 5)) return 4;
  return env >= 4 && env <= 6 ? env : fallback;
}

template <int n, typename Number, bool PEER, int OCC>
void launch_plane_occ(const PlaneParams &p, const unsigned grid, int device, cudaStream_t stream)
{
  using Cfg = PlaneCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(plane_cell_kernel<n, Number, PEER, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
      if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      attr[device] = true;
    }
  plane_cell_kernel<n, Number, PEER, OCC><<<grid, Cfg::warps * 32, Cfg::smem, stream>>>(p);
}

template <int n, typename Number>
void launch_plane_impl(const PlaneLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream, const PeerTables *peer = nullptr)
{
  using Cfg = PlaneCfg<n, Number>;
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
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
  p.hn_mask_strategy  = cp.hn_mask_strategy && cp.apply_constraints;
  p.n_owned           = peer ? peer->n_owned : 0;
  p.ghost_src         = peer ? peer->ghost_src : nullptr;
  p.ghost_dst         = peer ? peer->ghost_dst : nullptr;
  const long long nb  = p.batch_end - p.batch_begin;
  if (nb <= 0) return;
  const unsigned grid = (unsigned)((nb + Cfg::warps - 1) / Cfg::warps);
  constexpr bool reg_bound = sizeof(Number) == 8 && n >= 5;
  const int occ = occupancy_choice(n, sizeof(Number) == 8, MFHN_PLANE_DEFAULT_OCC);
  if (peer)
    launch_plane_occ<n, Number, true, 4>(p, grid, device, stream);
  else if (reg_bound && occ == 5)
    launch_plane_occ<n, Number, false, reg_bound ? 5 : 4>(p, grid, device, stream);
  else if (reg_bound && occ == 6)
    launch_plane_occ<n, Number, false, reg_bound ? 6 : 4>(p, grid, device, stream);
  else
    launch_plane_occ<n, Number, false, 4>(p, grid, device, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("plane kernel launch: ") + cudaGetErrorString(e));
}

template <int n, typename Number>
void launch_plane(const PlaneLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream, const PeerTables *peer = nullptr)
{
  if constexpr (plane_supported(n))
    launch_plane_impl<n, Number>(L, cp, device, stream, peer);
  else
    throw std::runtime_error("plane kernel not available for this degree");
}
} // namespace mfhn
