// Patch kernel (experimental alternative to the per-element gather of
// kernels_plane.cuh, whose arithmetic core and thread layout it shares).
//
// A warp processes a patch of 32/n Morton-consecutive cells and never
// synchronises with another warp.  The setup turns the patch's n^3-per-cell DoF
// indices into
//   uidx[U]       the patch's UNIQUE global DoF indices, grouped into classes by
//                 multiplicity m in {1,2,4,8} (number of cell-local DoFs they feed;
//                 other multiplicities are split, e.g. 3 = 2 + 1) and sorted by
//                 address inside a class (padded with the last entry);
//   lidx[j][lane] for every thread-plane slot the position of its DoF in uidx (uint16);
//   ent[m][n_m]   per class a rectangular array of shared-memory slots (cell,
//                 local dof): column i lists the m slots of the class's i-th DoF.
// Gather: consecutive lanes read consecutive entries of the sorted list (few
// 128-byte lines per request instead of ~15) into shared memory; every thread
// then picks its plane out of shared memory through lidx.  Scatter: the
// contributions of all cells of the patch to a DoF are summed out of shared
// memory and ONE red.global.add per unique DoF is issued, again address-sorted.
#pragma once
#include "kernels_plane.cuh"

#include <cstdint>

namespace mfhn
{
constexpr int N_CLASSES = 4; // multiplicities 1, 2, 4, 8
struct PatchInfo
{
  unsigned short n_unique;
  unsigned short count[N_CLASSES]; // number of unique DoFs per multiplicity class
  unsigned short pad[3];
};
static_assert(sizeof(PatchInfo) == 16, "PatchInfo layout");

struct PatchParams
{
  const PatchInfo *patches;
  const uint32_t *uidx; // [n_patches][u_stride]
  const uint16_t *lidx; // [n_patches][n*n][32]
  const uint16_t *ent;  // [n_patches][ent_stride]
  const uint8_t *masks;
  const void *h;
  const void *src;
  void *dst;
  long long cell_begin, cell_end, patch_begin, patch_end;
  int apply_constraints;
};

template <int n, typename Number>
struct PatchCfg
{
  using Plane = PlaneCfg<n, Number>;
  static constexpr int warps = 4;                       // independent warps per CTA
  static constexpr int cpw   = Plane::cpw;              // cells per patch
  static constexpr int ps = Plane::ps, cs = Plane::cs;
  static constexpr int ent_stride = cpw * n * n * n;    // slots of a patch
  static constexpr int rounds     = (ent_stride + 31) / 32; // gather rounds covering the worst case U = ent_stride
  static constexpr int u_stride   = rounds * 32;
  static constexpr int warp_stride = cpw * cs > u_stride ? cpw * cs : u_stride; // one array per warp (values, then a / b / r)
  static constexpr int smem = warps * warp_stride * (int)sizeof(Number);
  // shared-memory slot of (cell s of the patch, local dof (x,y,z)): kernel axes (X,Y,Z) = physical (y,z,x)
  static constexpr int slot(int s, int x, int y, int z) { return s * cs + x * ps + z * n + y; }
};

// Rounds (DoFs per lane) reserved per multiplicity class of the scatter so that one
// pass covers the class sizes observed on the reference's meshes; larger classes
// fall into a (rare) remainder loop.
template <int n>
struct PatchRounds;
template <> struct PatchRounds<2> { static constexpr int r1 = 1, r2 = 1, r4 = 1, r8 = 1; };
template <> struct PatchRounds<3> { static constexpr int r1 = 5, r2 = 3, r4 = 1, r8 = 1; };
template <> struct PatchRounds<4> { static constexpr int r1 = 11, r2 = 4, r4 = 2, r8 = 1; };
template <> struct PatchRounds<5> { static constexpr int r1 = 18, r2 = 5, r4 = 2, r8 = 1; };
template <> struct PatchRounds<6> { static constexpr int r1 = 24, r2 = 7, r4 = 2, r8 = 1; };

// One multiplicity class of the scatter: sum the M contributions of each DoF out of the
// shared-memory array, one RED per DoF.  All index loads of the class are issued first.
template <int M, int R, typename Number>
__device__ __forceinline__ void class_pull(Number *__restrict__ dst, const uint32_t *__restrict__ uidx,
                                           const uint16_t *__restrict__ ent, int &ubase, int &ebase, const int count,
                                           const Number *A, const int lane)
{
  if (count > 0)
    {
      uint32_t g[R];
      unsigned short e[R][M];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r * 32 < count) // warp-uniform
          {
            const int i = min(lane + r * 32, count - 1); // lanes past the end re-read the last DoF (benign)
            g[r]        = __ldg(uidx + ubase + i);
#pragma unroll
            for (int q = 0; q < M; ++q) e[r][q] = __ldg(ent + ebase + q * count + i);
          }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r * 32 < count)
          {
            Number s = A[e[r][0]];
#pragma unroll
            for (int q = 1; q < M; ++q) s += A[e[r][q]];
            if (lane + r * 32 < count) atomicAdd(dst + g[r], s);
          }
      for (int i = lane + R * 32; i < count; i += 32)
        {
          Number s = A[__ldg(ent + ebase + i)];
#pragma unroll
          for (int q = 1; q < M; ++q) s += A[__ldg(ent + ebase + q * count + i)];
          atomicAdd(dst + __ldg(uidx + ubase + i), s);
        }
    }
  ubase += count;
  ebase += M * count;
}

template <int n, typename Number>
__global__ void __launch_bounds__(PatchCfg<n, Number>::warps * 32, (n <= 5 ? 4 : 3)) patch_cell_kernel(const PatchParams p)
{
  using Cfg = PatchCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs, cpw = Cfg::cpw;
  constexpr int R1 = PatchRounds<n>::r1, R2 = PatchRounds<n>::r2, R4 = PatchRounds<n>::r4, R8 = PatchRounds<n>::r8;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long patch = p.patch_begin + (long long)blockIdx.x * Cfg::warps + warp;
  if (patch >= p.patch_end) return; // warps are independent: no block-level barrier below
  Number *A = reinterpret_cast<Number *>(smem_raw) + warp * Cfg::warp_stride;

  const uint32_t *__restrict__ uidx = p.uidx + patch * (long long)Cfg::u_stride;
  const uint16_t *__restrict__ ent  = p.ent + patch * (long long)Cfg::ent_stride;
  const Number *__restrict__ src    = static_cast<const Number *>(p.src);
  Number *__restrict__ dst          = static_cast<Number *>(p.dst);
  const PatchInfo info              = p.patches[patch]; // needed only for the scatter

  const bool active = lane < cpw * n;
  const int ml = active ? lane : lane - 16; // idle lanes mirror a lane of the other half-warp
  const int c = ml / n, t = ml - c * n;
  const long long cell = patch * cpw + c;
  const bool valid = cell >= p.cell_begin && cell < p.cell_end;
  Number *cellA = A + c * cs;

  // ---- gather ------------------------------------------------------------------------
  // local indices of this thread's plane (coalesced 16-bit loads), issued first
  unsigned short li[n * n];
  {
    const uint16_t *lp = p.lidx + patch * (long long)(n * n * 32) + (c * n + t);
#pragma unroll
    for (int j = 0; j < n * n; ++j) li[j] = __ldg(lp + j * 32);
  }
  // every unique DoF of the patch once: lanes on consecutive entries of the address-sorted list
  // (the list is padded with its last entry, so no bound is needed)
#pragma unroll 1
  for (int r0 = 0; r0 < Cfg::rounds; r0 += 8)
    {
      uint32_t g[8];
      Number v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < Cfg::rounds) g[r] = __ldg(uidx + (r0 + r) * 32 + lane);
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < Cfg::rounds) v[r] = __ldg(src + g[r]);
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < Cfg::rounds) A[(r0 + r) * 32 + lane] = v[r];
    }
  __syncwarp();
  Number u[n][n];
#pragma unroll
  for (int j = 0; j < n * n; ++j) u[j / n][j % n] = A[li[j]];
  __syncwarp();

  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0); // h = 0 silences cells outside the range
  const bool any_hn   = __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);
  if (any_hn)
    {
#pragma unroll
      for (int j = 0; j < n * n; ++j) cellA[t * ps + j] = u[j / n][j % n];
      __syncwarp();
      hn_smem<n, false>(cellA, hn_face, hn_edge, hn_cb, t);
#pragma unroll
      for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
      __syncwarp();
    }

  // ---- cell operator (as in kernels_plane.cuh) -----------------------------------------
  Number az[n][n], bz[n][n];
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
#pragma unroll
  for (int y = 0; y < n; ++y)
    {
      Number r[n];
      apply_Mb_Ka<n>(az[y], bz[y], r);
#pragma unroll
      for (int z = 0; z < n; ++z) cellA[z * ps + y * n + t] = h * r[z];
    }
  __syncwarp();
  if (any_hn) hn_smem<n, true>(cellA, hn_face, hn_edge, hn_cb, t);

  // ---- scatter: sum the patch's contributions per unique DoF, one RED each -------------
  {
    int ubase = 0, ebase = 0;
    class_pull<1, R1>(dst, uidx, ent, ubase, ebase, info.count[0], A, lane);
    class_pull<2, R2>(dst, uidx, ent, ubase, ebase, info.count[1], A, lane);
    class_pull<4, R4>(dst, uidx, ent, ubase, ebase, info.count[2], A, lane);
    class_pull<8, R8>(dst, uidx, ent, ubase, ebase, info.count[3], A, lane);
  }
}

// ---- host side ------------------------------------------------------------------
struct PatchLayout
{
  int n = 0, number = 0;
  long long n_patches = 0, n_cells = 0;
  PatchInfo *d_patches = nullptr;
  uint32_t *d_uidx     = nullptr;
  uint16_t *d_lidx     = nullptr;
  uint16_t *d_ent      = nullptr;
  double unique_per_cell = 0;
  long long index_bytes  = 0;

  void free()
  {
    cudaFree(d_patches);
    cudaFree(d_uidx);
    cudaFree(d_lidx);
    cudaFree(d_ent);
    d_patches = nullptr;
    d_uidx    = nullptr;
    d_lidx = d_ent = nullptr;
  }
  void build(int n_, int number_, long long n_cells, const uint32_t *idx);
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
  PatchParams p;
  p.patches           = L.d_patches;
  p.uidx              = L.d_uidx;
  p.lidx              = L.d_lidx;
  p.ent               = L.d_ent;
  p.masks             = cp.masks;
  p.h                 = cp.geom;
  p.src               = cp.src;
  p.dst               = cp.dst;
  p.cell_begin        = cp.cell_begin;
  p.cell_end          = cp.cell_end;
  p.patch_begin       = cp.cell_begin / Cfg::cpw;
  p.patch_end         = (cp.cell_end + Cfg::cpw - 1) / Cfg::cpw;
  p.apply_constraints = cp.apply_constraints;
  const long long np  = p.patch_end - p.patch_begin;
  if (np <= 0) return;
  patch_cell_kernel<n, Number><<<(unsigned)((np + Cfg::warps - 1) / Cfg::warps), Cfg::warps * 32, Cfg::smem, stream>>>(p);
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
