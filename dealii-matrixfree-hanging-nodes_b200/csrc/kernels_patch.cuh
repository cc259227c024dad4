// Patch kernel: the fast Cartesian path (supersedes the per-element gather of
// kernels_plane.cuh, whose arithmetic core it reuses).
//
// A CTA processes a patch of 4 x (32/n) Morton-consecutive cells.  The setup
// turns the patch's n^3-per-cell DoF indices into
//   uidx[U]      the patch's UNIQUE global DoF indices, sorted by (multiplicity, address)
//   off[U+1]     CSR offsets into
//   ent[...]     shared-memory slots (cell, local dof) that each unique DoF feeds
// so that
//   * the gather reads every DoF of the patch once, with consecutive lanes on
//     ascending addresses (few 128-byte lines per request instead of ~17), and
//     pushes it into the cells' shared-memory arrays;
//   * the scatter pulls the contributions of all cells of the patch to a DoF
//     out of shared memory, sums them and issues ONE red.global.add per unique
//     DoF (about 2/3 of the per-cell count at degree 4, again address-sorted).
// Between the two, each warp runs the register-tiled separable operator on its
// own cells (see kernels_plane.cuh) and the hanging-node interpolation /
// its transpose as in-place directional passes on the shared-memory arrays.
#pragma once
#include "kernels_plane.cuh"

#include <cstdint>

namespace mfhn
{
struct PatchInfo
{
  long long cell_begin; // first cell of the patch
  long long uidx_start; // offset of the patch's unique list in uidx; its CSR offsets start at uidx_start + patch id
  int n_cells;
  int n_unique;
};

struct PatchParams
{
  const PatchInfo *patches;
  const uint32_t *uidx;
  const uint16_t *off;
  const uint16_t *ent;
  const uint8_t *masks;
  const void *h;
  const void *src;
  void *dst;
  long long patch_begin;
  int apply_constraints;
};

template <int n, typename Number>
struct PatchCfg
{
  using Plane = PlaneCfg<n, Number>;
  static constexpr int warps = 4;
  static constexpr int cpw   = Plane::cpw;
  static constexpr int cells = warps * cpw;             // cells per patch
  static constexpr int ps = Plane::ps, cs = Plane::cs;
  static constexpr int warp_stride = 2 * cpw * cs;      // two arrays (A, B) per warp
  static constexpr int smem = warps * warp_stride * (int)sizeof(Number);
  static constexpr int ent_stride = cells * n * n * n;  // entries reserved per patch
  // shared-memory slot of (cell slot s in the patch, local dof (x,y,z)), in units of Number
  static constexpr int slot(int s, int x, int y, int z) { return (s / cpw) * warp_stride + (s % cpw) * cs + z * ps + y * n + x; }
};

// In-place hanging-node interpolation (or its transpose) on the cell arrays of
// one warp: three directional passes, every thread of a cell takes n of the
// n^2 lines of a pass.
template <int n, bool transpose, typename Number>
__device__ __forceinline__ void hn_smem(Number *cellA, unsigned mask, int t, bool active)
{
  constexpr int k = n - 1;
  using Cfg = PlaneCfg<n, Number>;
  unsigned face, edge, cb;
  decode_mask(mask, face, edge, cb);
#pragma unroll 1
  for (int d = 0; d < 3; ++d)
    {
      const int t0 = (d == 0) ? 1 : 0, t1 = (d == 2) ? 1 : 2;
      const int c0 = (int)((cb >> t0) & 1u) * k, c1 = (int)((cb >> t1) & 1u) * k;
      const bool f0 = (face >> t0) & 1u, f1 = (face >> t1) & 1u, ed = (edge >> d) & 1u;
      const bool upper = (cb >> d) & 1u;
      const int stride = d == 0 ? 1 : d == 1 ? n : Cfg::ps;
      const int b      = t;
#pragma unroll 1
      for (int a = 0; a < n; ++a)
        {
          const bool on0 = a == c0, on1 = b == c1;
          const bool sel = active && mask != 0u && ((f0 && on0) || (f1 && on1) || (ed && on0 && on1));
          if (sel)
            {
              const int base = d == 0 ? b * Cfg::ps + a * n : d == 1 ? b * Cfg::ps + a : b * n + a;
              Number *line   = cellA + base;
              Number v[n], w[n];
#pragma unroll
              for (int i = 0; i < n; ++i) v[i] = line[(upper ? k - i : i) * stride];
              mat_vec<n, T_W0, transpose>(v, w);
#pragma unroll
              for (int i = 0; i < n; ++i) line[(upper ? k - i : i) * stride] = w[i];
            }
        }
      __syncwarp();
    }
}

template <int n, typename Number>
__global__ void __launch_bounds__(PatchCfg<n, Number>::warps * 32, 3) patch_cell_kernel(const PatchParams p)
{
  using Cfg = PatchCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs, cpw = Cfg::cpw, nthreads = Cfg::warps * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *sm = reinterpret_cast<Number *>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long patch = p.patch_begin + blockIdx.x;
  const PatchInfo info  = p.patches[patch];
  const uint32_t *__restrict__ uidx = p.uidx + info.uidx_start;
  const uint16_t *__restrict__ off  = p.off + info.uidx_start + patch;
  const uint16_t *__restrict__ ent  = p.ent + patch * (long long)Cfg::ent_stride;
  const Number *__restrict__ src    = static_cast<const Number *>(p.src);
  Number *__restrict__ dst          = static_cast<Number *>(p.dst);
  const int U                       = info.n_unique;

  // ---- gather: every unique DoF of the patch once, pushed to the cells that use it
  for (int i0 = tid; i0 < U; i0 += 4 * nthreads)
    {
      Number v[4];
      int e0[4], e1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        {
          const int i = i0 + q * nthreads;
          if (i < U)
            {
              v[q]  = __ldg(src + __ldg(uidx + i));
              e0[q] = __ldg(off + i);
              e1[q] = __ldg(off + i + 1);
            }
          else
            e0[q] = e1[q] = 0;
        }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        for (int e = e0[q]; e < e1[q]; ++e) sm[__ldg(ent + e)] = v[q];
    }
  __syncthreads();

  // ---- per-warp cell operator ----------------------------------------------------
  {
    Number *A = sm + warp * Cfg::warp_stride;
    Number *B = A + cpw * cs;
    const int c = lane / n, t = lane - c * n;
    const int slot    = warp * cpw + c;
    const bool active = lane < cpw * n;
    const bool valid  = active && slot < info.n_cells;
    const long long cell = info.cell_begin + slot;
    const unsigned mask  = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
    const Number h       = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
    const bool any_hn    = __any_sync(0xffffffffu, mask != 0u);
    Number *cellA        = A + c * cs;
    if (any_hn) hn_smem<n, false>(cellA, mask, t, active);

    // P1 (thread = z): plane (x,y) -> a = M_y M_x u, b = (M_y K_x + K_y M_x) u
    Number u[n][n];
    if (active)
      {
#pragma unroll
        for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
      }
    else
      {
#pragma unroll
        for (int j = 0; j < n * n; ++j) u[j / n][j % n] = Number(0);
      }
    __syncwarp();
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
          if (active)
            {
#pragma unroll
              for (int i = 0; i < n; ++i)
                {
                  A[c * cs + t * ps + i * n + x] = a[i];
                  B[c * cs + t * ps + i * n + x] = b[i];
                }
            }
        }
    }
    __syncwarp();
    // P2 (thread = x): r = h (M_z b + K_z a), back into A
    if (active)
      {
#pragma unroll
        for (int y = 0; y < n; ++y)
          {
            Number a[n], b[n], r[n];
#pragma unroll
            for (int z = 0; z < n; ++z)
              {
                a[z] = A[c * cs + z * ps + y * n + t];
                b[z] = B[c * cs + z * ps + y * n + t];
              }
            apply_Mb_Ka<n>(a, b, r);
#pragma unroll
            for (int z = 0; z < n; ++z) A[c * cs + z * ps + y * n + t] = h * r[z];
          }
      }
    __syncwarp();
    if (any_hn) hn_smem<n, true>(cellA, mask, t, active);
  }
  __syncthreads();

  // ---- scatter: sum the patch's contributions per unique DoF, one RED each -------
  for (int i0 = tid; i0 < U; i0 += 4 * nthreads)
    {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        {
          const int i = i0 + q * nthreads;
          if (i < U)
            {
              const int e0 = __ldg(off + i), e1 = __ldg(off + i + 1);
              Number s = sm[__ldg(ent + e0)];
              for (int e = e0 + 1; e < e1; ++e) s += sm[__ldg(ent + e)];
              atomicAdd(dst + __ldg(uidx + i), s);
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------
struct PatchLayout
{
  int n = 0, number = 0;
  long long n_patches = 0;
  std::vector<long long> patch_cell_begin; // host copy, n_patches + 1 (last = n_cells), for range launches
  PatchInfo *d_patches = nullptr;
  uint32_t *d_uidx     = nullptr;
  uint16_t *d_off      = nullptr;
  uint16_t *d_ent      = nullptr;
  double unique_per_cell = 0;
  long long index_bytes  = 0;

  void free()
  {
    cudaFree(d_patches);
    cudaFree(d_uidx);
    cudaFree(d_off);
    cudaFree(d_ent);
    d_patches = nullptr;
    d_uidx    = nullptr;
    d_off = d_ent = nullptr;
  }
  void build(int n_, int number_, long long n_cells, const uint32_t *idx, const std::vector<long long> &segments);
};

template <int n, typename Number>
void launch_patch_impl(const PatchLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  using Cfg = PatchCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(patch_cell_kernel<n, Number>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem);
      if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      attr[device] = true;
    }
  // the cell range must coincide with patch boundaries (segments given at creation)
  const auto &pb = L.patch_cell_begin;
  const auto lo  = std::lower_bound(pb.begin(), pb.end(), cp.cell_begin);
  const auto hi  = std::lower_bound(pb.begin(), pb.end(), cp.cell_end);
  if (lo == pb.end() || *lo != cp.cell_begin || hi == pb.end() || *hi != cp.cell_end)
    throw std::invalid_argument("cell range does not coincide with the segments given at operator creation");
  const long long pbeg = lo - pb.begin(), pend = hi - pb.begin();
  if (pend <= pbeg) return;
  PatchParams p;
  p.patches           = L.d_patches;
  p.uidx              = L.d_uidx;
  p.off               = L.d_off;
  p.ent               = L.d_ent;
  p.masks             = cp.masks;
  p.h                 = cp.geom;
  p.src               = cp.src;
  p.dst               = cp.dst;
  p.patch_begin       = pbeg;
  p.apply_constraints = cp.apply_constraints;
  patch_cell_kernel<n, Number><<<(unsigned)(pend - pbeg), Cfg::warps * 32, Cfg::smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("patch kernel launch: ") + cudaGetErrorString(e));
}

template <int n, typename Number>
void launch_patch(const PatchLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  if constexpr (plane_supported(n))
    launch_patch_impl<n, Number>(L, cp, device, stream);
  else
    throw std::runtime_error("patch kernel not available for this degree");
}
} // namespace mfhn
