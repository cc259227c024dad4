// Patch kernel: the fast Cartesian path (supersedes the per-element gather of
// kernels_plane.cuh, whose arithmetic core it reuses).
//
// A warp processes a patch of 32/n Morton-consecutive cells and never
// synchronises with another warp.  The setup turns the patch's n^3-per-cell DoF
// indices into
//   uidx[U]      the patch's UNIQUE global DoF indices, grouped into classes by
//                multiplicity m in {1,2,4,8} (number of cell-local DoFs they feed;
//                other multiplicities are split, e.g. 3 = 2 + 1) and sorted by
//                address inside a class
//   ent[m][n_m]  per class a rectangular array of shared-memory slots (cell,
//                local dof): column i lists the m slots of the class's i-th DoF
// (fixed stride per patch, so every address follows from the patch number alone
// and all index loads are independent of each other), so that
//   * the gather reads every DoF of the patch once, with consecutive lanes on
//     ascending addresses (few 128-byte lines per request instead of ~17), and
//     pushes it into the cells' shared-memory arrays;
//   * the scatter pulls the contributions of all cells of the patch to a DoF
//     out of shared memory, sums them and issues ONE red.global.add per unique
//     DoF, again address-sorted.
// Between the two, the warp runs the register-tiled separable operator on its
// cells (see kernels_plane.cuh) and the hanging-node interpolation / its
// transpose as in-place directional passes on the shared-memory arrays.
#pragma once
#include "kernels_plane.cuh"

#include <cstdint>

namespace mfhn
{
constexpr int N_CLASSES = 4; // multiplicities 1, 2, 4, 8
struct PatchInfo
{
  long long cell_begin;             // first cell of the patch
  int n_cells;
  unsigned short count[N_CLASSES];  // number of unique DoFs per multiplicity class
  int pad[3];
};
static_assert(sizeof(PatchInfo) == 32, "PatchInfo layout");

struct PatchParams
{
  const PatchInfo *patches;
  const uint32_t *uidx; // [n_patches][ent_stride]
  const uint16_t *ent;  // [n_patches][ent_stride]
  const uint8_t *masks;
  const void *h;
  const void *src;
  void *dst;
  long long patch_begin, patch_end, n_patches_total;
  int apply_constraints;
};

template <int n, typename Number>
struct PatchCfg
{
  using Plane = PlaneCfg<n, Number>;
  static constexpr int warps = 4;                       // independent warps per CTA
  static constexpr int cpw   = Plane::cpw;              // cells per patch
  static constexpr int ps = Plane::ps, cs = Plane::cs;
  static constexpr int warp_stride = 2 * cpw * cs;      // two arrays (A, B) per warp
  static constexpr int smem = warps * warp_stride * (int)sizeof(Number);
  static constexpr int ent_stride = cpw * n * n * n;    // entries reserved per patch
  // shared-memory slot of (cell s of the patch, local dof (x,y,z)) inside the warp's array A
  static constexpr int slot(int s, int x, int y, int z) { return s * cs + z * ps + y * n + x; }
};

// Rounds (DoFs per lane) reserved per multiplicity class so that one pass covers
// the class sizes observed on the reference's meshes; larger classes fall into a
// (rare) remainder loop.
template <int n>
struct PatchRounds;
template <> struct PatchRounds<2> { static constexpr int r1 = 1, r2 = 1, r4 = 1, r8 = 1; };
template <> struct PatchRounds<3> { static constexpr int r1 = 5, r2 = 3, r4 = 1, r8 = 1; };
template <> struct PatchRounds<4> { static constexpr int r1 = 11, r2 = 4, r4 = 2, r8 = 1; };
template <> struct PatchRounds<5> { static constexpr int r1 = 18, r2 = 5, r4 = 2, r8 = 1; };
template <> struct PatchRounds<6> { static constexpr int r1 = 24, r2 = 7, r4 = 2, r8 = 1; };

// Index registers of one multiplicity class: R DoFs per lane, M slots each.
template <int M, int R>
struct ClassIdx
{
  uint32_t g[R];
  unsigned short e[R][M];
  int count, ubase, ebase;
};

// stage 1: all index loads (independent of each other and of the data).  Lanes
// past the end of a class re-read its last DoF (benign duplicate), so the loads
// and the shared-memory stores need no per-lane predicate.
template <int M, int R>
__device__ __forceinline__ void class_load_idx(ClassIdx<M, R> &c, const uint32_t *__restrict__ uidx,
                                               const uint16_t *__restrict__ ent, int &ubase, int &ebase, const int count,
                                               const int lane)
{
  c.count = count;
  c.ubase = ubase;
  c.ebase = ebase;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (r * 32 < count) // warp-uniform
      {
        const int i = min(lane + r * 32, count - 1);
        c.g[r]      = __ldg(uidx + ubase + i);
#pragma unroll
        for (int q = 0; q < M; ++q) c.e[r][q] = __ldg(ent + ebase + q * count + i);
      }
  ubase += count;
  ebase += M * count;
}
// stage 2: value loads
template <int M, int R, typename Number>
__device__ __forceinline__ void class_load_val(const ClassIdx<M, R> &c, Number (&v)[R], const Number *__restrict__ src)
{
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (r * 32 < c.count) v[r] = __ldg(src + c.g[r]);
}
// stage 3: push to the cells' shared-memory arrays (+ remainder of an oversized class)
template <int M, int R, typename Number>
__device__ __forceinline__ void class_push(const ClassIdx<M, R> &c, const Number (&v)[R], const Number *__restrict__ src,
                                           const uint32_t *__restrict__ uidx, const uint16_t *__restrict__ ent, Number *A,
                                           const int lane)
{
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (r * 32 < c.count)
      {
#pragma unroll
        for (int q = 0; q < M; ++q) A[c.e[r][q]] = v[r];
      }
  for (int i = lane + R * 32; i < c.count; i += 32)
    {
      const Number val = __ldg(src + __ldg(uidx + c.ubase + i));
#pragma unroll
      for (int q = 0; q < M; ++q) A[__ldg(ent + c.ebase + q * c.count + i)] = val;
    }
}
// scatter: sum the M contributions, one RED per DoF
template <int M, int R, typename Number>
__device__ __forceinline__ void class_pull(const ClassIdx<M, R> &c, Number *__restrict__ dst, const uint32_t *__restrict__ uidx,
                                           const uint16_t *__restrict__ ent, const Number *A, const int lane)
{
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (r * 32 < c.count)
      {
        Number s = A[c.e[r][0]];
#pragma unroll
        for (int q = 1; q < M; ++q) s += A[c.e[r][q]];
        if (lane + r * 32 < c.count) atomicAdd(dst + c.g[r], s);
      }
  for (int i = lane + R * 32; i < c.count; i += 32)
    {
      Number s = A[__ldg(ent + c.ebase + i)];
#pragma unroll
      for (int q = 1; q < M; ++q) s += A[__ldg(ent + c.ebase + q * c.count + i)];
      atomicAdd(dst + __ldg(uidx + c.ubase + i), s);
    }
}

#define MFHN_FOR_CLASSES(X) X(1, R1, c1, v1, 0) X(2, R2, c2, v2, 1) X(4, R4, c4, v4, 2) X(8, R8, c8, v8, 3)

template <int n, typename Number>
__global__ void __launch_bounds__(PatchCfg<n, Number>::warps * 32, 3) patch_cell_kernel(const PatchParams p)
{
  using Cfg = PatchCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs, cpw = Cfg::cpw;
  constexpr int R1 = PatchRounds<n>::r1, R2 = PatchRounds<n>::r2, R4 = PatchRounds<n>::r4, R8 = PatchRounds<n>::r8;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long patch = p.patch_begin + (long long)blockIdx.x * Cfg::warps + warp;
  if (patch >= p.patch_end) return; // warps are independent: no block-level barrier below
  Number *A = reinterpret_cast<Number *>(smem_raw) + warp * Cfg::warp_stride;
  Number *B = A + cpw * cs;

  const PatchInfo info = p.patches[patch];
  const uint32_t *__restrict__ uidx = p.uidx + patch * (long long)Cfg::ent_stride;
  const uint16_t *__restrict__ ent  = p.ent + patch * (long long)Cfg::ent_stride;
  const Number *__restrict__ src    = static_cast<const Number *>(p.src);
  Number *__restrict__ dst          = static_cast<Number *>(p.dst);

  // ---- gather: every unique DoF of the patch once, pushed to the cells that use it.
  // All index loads first, then all value loads, then the shared-memory stores: one
  // dependent chain (index -> value) per patch.
  {
    // warm the L2 with the index blocks of a patch that a later warp will process
    const long long ahead = patch + 3LL * 148 * Cfg::warps;
    if (ahead < p.n_patches_total)
      {
        const char *pu = reinterpret_cast<const char *>(p.uidx + ahead * (long long)Cfg::ent_stride);
        const char *pe = reinterpret_cast<const char *>(p.ent + ahead * (long long)Cfg::ent_stride);
        constexpr int lines_u = (Cfg::ent_stride * 7 / 2) / 128 + 1; // ~7/8 of the reserved uint32 block is used
        constexpr int lines_e = (Cfg::ent_stride * 2) / 128 + 1;
        for (int l = lane; l < lines_u + lines_e; l += 32)
          {
            const char *a = l < lines_u ? pu + 128 * l : pe + 128 * (l - lines_u);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.patches + ahead));
      }
    ClassIdx<1, R1> c1;
    ClassIdx<2, R2> c2;
    ClassIdx<4, R4> c4;
    ClassIdx<8, R8> c8;
    int ubase = 0, ebase = 0;
#define X(M, R, c, v, ci) class_load_idx<M, R>(c, uidx, ent, ubase, ebase, info.count[ci], lane);
    MFHN_FOR_CLASSES(X)
#undef X
#define X(M, R, c, v, ci) \
  Number v[R];            \
  class_load_val<M, R>(c, v, src);
    MFHN_FOR_CLASSES(X)
#undef X
#define X(M, R, c, v, ci) class_push<M, R>(c, v, src, uidx, ent, A, lane);
    MFHN_FOR_CLASSES(X)
#undef X
  }
  __syncwarp();

  // ---- cell operator ----------------------------------------------------------------
  {
    // the 32 - cpw n idle lanes mirror lane 0 (same loads, same values stored to the
    // same addresses), so the arithmetic below needs no per-lane predicate
    const int ml = lane < cpw * n ? lane : lane - 16; // idle lanes mirror a lane of the other half-warp
    const int c = ml / n, t = ml - c * n;
    const bool valid     = c < info.n_cells;
    const long long cell = info.cell_begin + c;
    const unsigned mask  = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
    const Number h       = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
    const bool any_hn    = __any_sync(0xffffffffu, mask != 0u);
    unsigned hn_face, hn_edge, hn_cb;
    decode_mask(mask, hn_face, hn_edge, hn_cb); // the patch layout keeps physical axes
    Number *cellA        = A + c * cs;
    Number *cellB        = B + c * cs;
    if (any_hn) hn_smem<n, false>(cellA, hn_face, hn_edge, hn_cb, t);

    // P1 (thread = z): plane (x,y) -> a = M_y M_x u, b = (M_y K_x + K_y M_x) u
    Number u[n][n];
#pragma unroll
    for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
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
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              cellA[t * ps + i * n + x] = a[i];
              cellB[t * ps + i * n + x] = b[i];
            }
        }
    }
    __syncwarp();
    // P2 (thread = x): r = h (M_z b + K_z a), back into A
#pragma unroll
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
  }
  __syncwarp();

  // ---- scatter: sum the patch's contributions per unique DoF, one RED each -------
  {
    ClassIdx<1, R1> c1;
    ClassIdx<2, R2> c2;
    ClassIdx<4, R4> c4;
    ClassIdx<8, R8> c8;
    int ubase = 0, ebase = 0;
#define X(M, R, c, v, ci) class_load_idx<M, R>(c, uidx, ent, ubase, ebase, info.count[ci], lane);
    MFHN_FOR_CLASSES(X)
#undef X
#define X(M, R, c, v, ci) class_pull<M, R>(c, dst, uidx, ent, A, lane);
    MFHN_FOR_CLASSES(X)
#undef X
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
  uint16_t *d_ent      = nullptr;
  double unique_per_cell = 0;
  long long index_bytes  = 0;

  void free()
  {
    cudaFree(d_patches);
    cudaFree(d_uidx);
    cudaFree(d_ent);
    d_patches = nullptr;
    d_uidx    = nullptr;
    d_ent     = nullptr;
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
  p.ent               = L.d_ent;
  p.masks             = cp.masks;
  p.h                 = cp.geom;
  p.src               = cp.src;
  p.dst               = cp.dst;
  p.patch_begin       = pbeg;
  p.patch_end         = pend;
  p.n_patches_total   = L.n_patches;
  p.apply_constraints = cp.apply_constraints;
  patch_cell_kernel<n, Number><<<(unsigned)((pend - pbeg + Cfg::warps - 1) / Cfg::warps), Cfg::warps * 32, Cfg::smem, stream>>>(p);
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
