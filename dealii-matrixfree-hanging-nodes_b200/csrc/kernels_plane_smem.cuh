// Plane kernel for the high degrees (k = 6, 7, 8): same decomposition as
// kernels_plane.cuh (n threads per cell, thread owns an n x n plane, warps are
// independent, separable Cartesian operator), but the plane lives in the
// warp's shared memory instead of registers -- an n x n plane of doubles with
// its two intermediates does not fit the register file for n >= 7.  The X and Y
// sweeps work row by row / column by column on the thread's PRIVATE plane (no
// synchronisation needed), only the Z sweep crosses threads.
//
//   P0 (thread = Z): gather rows -> A                      [hanging-node passes on A]
//   P1 (thread = Z): rows:    p = M_X u, q = K_X u          -> A, B  (in place)
//                    columns: a = M_Y p, b = M_Y q + K_Y p   -> A, B  (in place)
//   P2 (thread = X): r = h (M_Z b + K_Z a)                   -> A
//   P3 (thread = Z): [transposed hanging-node passes on A]  RED scatter from A
#pragma once
#include "kernels_plane.cuh"

namespace mfhn
{
template <int n, typename Number>
struct PlaneSmemCfg
{
  using Plane = PlaneCfg<n, Number>;
  static constexpr int cpw = Plane::cpw, ps = Plane::ps, cs = Plane::cs;
  static constexpr int warps = (2 * cpw * cs * (int)sizeof(Number) > 20 * 1024) ? 2 : 4; // 2-warp CTAs pack the SM better when a warp needs > 20 KB
  static constexpr int smem  = warps * 2 * cpw * cs * (int)sizeof(Number);
  static constexpr int rows_in_flight = n <= 7 ? 4 : 3; // gather rows issued before the first use
};

constexpr bool plane_smem_supported(int n) { return n >= 2 && n <= 9; }

template <int n, typename Number>
__global__ void __launch_bounds__(PlaneSmemCfg<n, Number>::warps * 32) plane_smem_kernel(const PlaneParams p)
{
  using Cfg = PlaneSmemCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs, cpw = Cfg::cpw, RF = Cfg::rows_in_flight;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = p.batch_begin + (long long)blockIdx.x * Cfg::warps + warp;
  if (batch >= p.batch_end) return; // warps are independent: no block-level barrier below
  Number *A = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * 2 * cpw * cs;
  Number *B = A + cpw * cs;

  const bool active = lane < cpw * n;
  const int ml = active ? lane : lane - 16; // idle lanes mirror a lane of the other half-warp
  const int c = ml / n, t = ml - c * n;
  const long long cell = batch * cpw + c;
  const bool valid = cell >= p.cell_begin && cell < p.cell_end;
  const Number *__restrict__ src = static_cast<const Number *>(p.src);
  Number *__restrict__ dst = static_cast<Number *>(p.dst);
  const uint32_t *ip = p.pidx + batch * (long long)(n * n * 32) + (c * n + t);
  Number *planeA = A + c * cs + t * ps, *planeB = B + c * cs + t * ps; // this thread's private planes (P0, P1, P3)
  Number *cellA = A + c * cs, *cellB = B + c * cs;

  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
  const bool any_hn   = __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);

  // ---- P0: gather, RF rows in flight.  Warps without a constrained cell fuse the X sweep into the
  // gather (rows go straight from registers to p, q); the others stage u in A for the interpolation.
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
        for (int x = 0; x < n; ++x) v[r][x] = (valid && y0 + r < n) ? __ldg(src + idx[r][x]) : Number(0);
#pragma unroll
      for (int r = 0; r < RF; ++r)
        if (y0 + r < n)
          {
            if (any_hn)
              {
#pragma unroll
                for (int x = 0; x < n; ++x) planeA[(y0 + r) * n + x] = v[r][x];
              }
            else
              {
                Number pr[n], qr[n];
                apply_MK<n>(v[r], pr, qr);
#pragma unroll
                for (int x = 0; x < n; ++x)
                  {
                    planeA[(y0 + r) * n + x] = pr[x];
                    planeB[(y0 + r) * n + x] = qr[x];
                  }
              }
          }
    }
  if (any_hn)
    {
      __syncwarp();
      hn_smem<n, false>(cellA, hn_face, hn_edge, hn_cb, t);
      // ---- P1: X sweep on the rows of the private plane ---------------------------------
#pragma unroll 1
      for (int y = 0; y < n; ++y)
        {
          Number u[n], pr[n], qr[n];
#pragma unroll
          for (int x = 0; x < n; ++x) u[x] = planeA[y * n + x];
          apply_MK<n>(u, pr, qr);
#pragma unroll
          for (int x = 0; x < n; ++x)
            {
              planeA[y * n + x] = pr[x];
              planeB[y * n + x] = qr[x];
            }
        }
    }
  // ---- P1: Y sweep on the columns of the private plane ---------------------------------
#pragma unroll 1
  for (int x = 0; x < n; ++x)
    {
      Number pc[n], qc[n], a[n], b[n];
#pragma unroll
      for (int y = 0; y < n; ++y)
        {
          pc[y] = planeA[y * n + x];
          qc[y] = planeB[y * n + x];
        }
      apply_M_MK<n>(pc, qc, a, b);
#pragma unroll
      for (int y = 0; y < n; ++y)
        {
          planeA[y * n + x] = a[y];
          planeB[y * n + x] = b[y];
        }
    }
  __syncwarp();
  // ---- P2: Z sweep (thread = X) ------------------------------------------------------
#pragma unroll 1
  for (int y = 0; y < n; ++y)
    {
      Number a[n], b[n], r[n];
#pragma unroll
      for (int z = 0; z < n; ++z)
        {
          a[z] = cellA[z * ps + y * n + t];
          b[z] = cellB[z * ps + y * n + t];
        }
      apply_Mb_Ka<n>(a, b, r);
#pragma unroll
      for (int z = 0; z < n; ++z) cellA[z * ps + y * n + t] = h * r[z];
    }
  __syncwarp();
  if (any_hn) hn_smem<n, true>(cellA, hn_face, hn_edge, hn_cb, t);
  // ---- P3: scatter (thread = Z) ------------------------------------------------------
  if (active && valid)
    {
#pragma unroll 1
      for (int y = 0; y < n; ++y)
        {
          uint32_t g[n];
#pragma unroll
          for (int x = 0; x < n; ++x) g[x] = __ldg(ip + (y * n + x) * 32);
#pragma unroll
          for (int x = 0; x < n; ++x) atomicAdd(dst + g[x], planeA[y * n + x]);
        }
    }
}

template <int n, typename Number>
void launch_plane_smem(const PlaneLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  using Cfg = PlaneSmemCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(plane_smem_kernel<n, Number>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
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
  p.src_tex           = 0;
  const long long nb  = p.batch_end - p.batch_begin;
  if (nb <= 0) return;
  const unsigned grid = (unsigned)((nb + Cfg::warps - 1) / Cfg::warps);
  plane_smem_kernel<n, Number><<<grid, Cfg::warps * 32, Cfg::smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("plane (smem) kernel launch: ") + cudaGetErrorString(e));
}
} // namespace mfhn
