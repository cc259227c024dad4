// Plane kernel for the high degrees (k = 6, 7, 8): same decomposition as
// kernels_plane.cuh (n threads per cell, thread owns an n x n plane, warps are
// independent, Cartesian cell matrix by fast diagonalisation), but the plane
// lives in the warp's shared memory instead of registers -- an n x n plane of
// doubles does not fit the register file for n >= 7.  The X and Y sweeps work
// row by row / column by column on the thread's PRIVATE plane (no
// synchronisation needed), only the Z sweep crosses threads.  One array per
// warp: every sweep is in place.
//
//   P0 (thread = Z): gather rows, forward X sweep -> A        [hanging-node passes on A first]
//   P1 (thread = Z): columns: forward Y sweep                   (in place)
//   P2 (thread = X): Z lines: forward, eigenvalue scaling, backward  (in place)
//   P3 (thread = Z): columns: backward Y; rows: backward X -> RED scatter from registers
//                    [transposed hanging-node passes on A, then scatter from A]
#pragma once
#include "kernels_plane.cuh"

#ifndef MFHN_PLANE_SMEM_OCC7
#define MFHN_PLANE_SMEM_OCC7 1
#endif

namespace mfhn
{
template <int n, typename Number>
struct PlaneSmemCfg
{
  using Plane = PlaneCfg<n, Number>;
  static constexpr int cpw = Plane::cpw, ps = Plane::ps, cs = Plane::cs;
  static constexpr int warps = (cpw * cs * (int)sizeof(Number) > 20 * 1024) ? 2 : 4; // 2-warp CTAs pack the SM better when a warp needs > 20 KB
  static constexpr int smem  = warps * cpw * cs * (int)sizeof(Number);
  static constexpr int rows_in_flight = n <= 7 ? 4 : 3; // gather rows issued before the first use
  // CTAs per SM the register allocation is limited for: at k = 6 shared memory leaves room for 5 CTAs (20 warps)
  static constexpr int min_ctas = (n == 7 && sizeof(Number) == 8) ? MFHN_PLANE_SMEM_OCC7 : 1;
};


template <int n, typename Number, bool PEER = false>
__global__ void __launch_bounds__(PlaneSmemCfg<n, Number>::warps * 32, PlaneSmemCfg<n, Number>::min_ctas) plane_smem_kernel(const PlaneParams p)
{
  using Cfg = PlaneSmemCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs, cpw = Cfg::cpw, RF = Cfg::rows_in_flight;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = p.batch_begin + (long long)blockIdx.x * Cfg::warps + warp;
  if (batch >= p.batch_end) return; // warps are independent: no block-level barrier below
  Number *A = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * cpw * cs;

  const bool active = lane < cpw * n;
  const int ml = active ? lane : lane - 16; // idle lanes mirror a lane of the other half-warp; they never store
  const int c = ml / n, t = ml - c * n;
  const long long cell = batch * cpw + c;
  const bool valid = cell >= p.cell_begin && cell < p.cell_end;
  const Number *__restrict__ src = static_cast<const Number *>(p.src);
  Number *__restrict__ dst = static_cast<Number *>(p.dst);
  const uint32_t *ip = p.pidx + batch * (long long)(n * n * 32) + (c * n + t);
  Number *cellA = A + c * cs, *planeA = cellA + t * ps; // planeA: this thread's private plane (P0, P1, P3)

  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
  const bool any_hn   = p.hn_mask_strategy || __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);

  // ---- P0: gather, RF rows in flight.  Warps without a constrained cell fuse the X sweep into the
  // gather (rows go straight from registers through T); the others stage u in A for the interpolation.
#pragma unroll 1
  for (int y0 = 0; y0 < n; y0 += RF)
    {
      uint32_t idx[RF][n];
      Number v[RF][n];
#pragma unroll
      for (int r = 0; r < RF; ++r)
#pragma unroll
        for (int x = 0; x < n; ++x) idx[r][x] = (valid && y0 + r < n) ? __ldg(ip + ((y0 + r) * n + x) * 32) : 0u;
#pragma unroll
      for (int r = 0; r < RF; ++r)
#pragma unroll
        for (int x = 0; x < n; ++x)
          {
            if (PEER && valid && y0 + r < n && idx[r][x] >= (uint32_t)p.n_owned) // remote entry: plain load through the peer mapping
              v[r][x] = *static_cast<const Number *>(p.ghost_src[idx[r][x] - (uint32_t)p.n_owned]);
            else
              v[r][x] = (valid && y0 + r < n) ? __ldg(src + idx[r][x]) : Number(0);
          }
#pragma unroll
      for (int r = 0; r < RF; ++r)
        if (y0 + r < n)
          {
            if (!any_hn) fdm_fwd<n>(v[r], v[r]);
            if (active)
              {
#pragma unroll
                for (int x = 0; x < n; ++x) planeA[(y0 + r) * n + x] = v[r][x];
              }
          }
    }
  if (any_hn)
    {
      __syncwarp();
      hn_smem<n, false>(cellA, hn_face, hn_edge, hn_cb, t, active);
      // ---- P1: X sweep on the rows of the private plane ---------------------------------
#pragma unroll 1
      for (int y = 0; y < n; ++y)
        {
          Number u[n];
#pragma unroll
          for (int x = 0; x < n; ++x) u[x] = planeA[y * n + x];
          fdm_fwd<n>(u, u);
          if (active)
            {
#pragma unroll
              for (int x = 0; x < n; ++x) planeA[y * n + x] = u[x];
            }
        }
    }
  // ---- P1: Y sweep on the columns of the private plane ---------------------------------
#pragma unroll 1
  for (int x = 0; x < n; ++x)
    {
      Number u[n];
#pragma unroll
      for (int y = 0; y < n; ++y) u[y] = planeA[y * n + x];
      fdm_fwd<n>(u, u);
      if (active)
        {
#pragma unroll
          for (int y = 0; y < n; ++y) planeA[y * n + x] = u[y];
        }
    }
  __syncwarp();
  // ---- P2: Z lines (thread = X): forward, eigenvalue scaling, backward -------------------
  {
    const Number hlt = h * Shape<Number>::template lam<n>(t);
#pragma unroll 1
    for (int y = 0; y < n; ++y)
      {
        Number v[n];
#pragma unroll
        for (int z = 0; z < n; ++z) v[z] = cellA[z * ps + y * n + t];
        fdm_mid<n>(v, h, h * Shape<Number>::template lam<n>(y) + hlt);
        if (active)
          {
#pragma unroll
            for (int z = 0; z < n; ++z) cellA[z * ps + y * n + t] = v[z];
          }
      }
  }
  __syncwarp();
  // ---- P3 (thread = Z): backward Y on the columns, backward X on the rows ----------------
#pragma unroll 1
  for (int x = 0; x < n; ++x)
    {
      Number u[n];
#pragma unroll
      for (int y = 0; y < n; ++y) u[y] = planeA[y * n + x];
      fdm_bwd<n>(u, u);
      if (active)
        {
#pragma unroll
          for (int y = 0; y < n; ++y) planeA[y * n + x] = u[y];
        }
    }
  if (any_hn)
    {
#pragma unroll 1
      for (int y = 0; y < n; ++y)
        {
          Number u[n];
#pragma unroll
          for (int x = 0; x < n; ++x) u[x] = planeA[y * n + x];
          fdm_bwd<n>(u, u);
          if (active)
            {
#pragma unroll
              for (int x = 0; x < n; ++x) planeA[y * n + x] = u[x];
            }
        }
      __syncwarp();
      hn_smem<n, true>(cellA, hn_face, hn_edge, hn_cb, t, active);
    }
  // ---- scatter (thread = Z) ---------------------------------------------------------------
#pragma unroll 1
  for (int y = 0; y < n; ++y)
    {
      Number u[n];
      uint32_t g[n];
#pragma unroll
      for (int x = 0; x < n; ++x) g[x] = (active && valid) ? __ldg(ip + (y * n + x) * 32) : 0u;
#pragma unroll
      for (int x = 0; x < n; ++x) u[x] = planeA[y * n + x];
      if (!any_hn) fdm_bwd<n>(u, u); // rows go from the last sweep straight to the vector
      if (active && valid)
        {
#pragma unroll
          for (int x = 0; x < n; ++x)
            {
              if (PEER && g[x] >= (uint32_t)p.n_owned) // red over NVLink into the owner's dst
                atomicAdd(static_cast<Number *>(p.ghost_dst[g[x] - (uint32_t)p.n_owned]), u[x]);
              else
                atomicAdd(dst + g[x], u[x]);
            }
        }
    }
}

template <int n, typename Number, bool PEER>
void launch_plane_smem_variant(const PlaneParams &p, const unsigned grid, int device, cudaStream_t stream)
{
  using Cfg = PlaneSmemCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(plane_smem_kernel<n, Number, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
      if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      attr[device] = true;
    }
  plane_smem_kernel<n, Number, PEER><<<grid, Cfg::warps * 32, Cfg::smem, stream>>>(p);
}

template <int n, typename Number>
void launch_plane_smem(const PlaneLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream, const PeerTables *peer = nullptr)
{
  using Cfg = PlaneSmemCfg<n, Number>;
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
  if (peer)
    launch_plane_smem_variant<n, Number, true>(p, grid, device, stream);
  else
    launch_plane_smem_variant<n, Number, false>(p, grid, device, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("plane (smem) kernel launch: ") + cudaGetErrorString(e));
}
} // namespace mfhn
