// Plane kernel with block-wise vector access through the bulk-copy engine (TMA).
//
// The plane kernel (kernels_plane.cuh) is bound by the L1 data stage: every DoF is
// an 8-byte gather and an 8-byte RED, about 15 distinct 128-byte lines per warp
// request.  In deal.II's object-wise numbering the (k-1)^3 DoFs of the cell interior
// and the (k-1)^2 DoFs of every face are CONTIGUOUS in the vector, in a fixed order
// (also for a hanging face: the slots then hold the coarse neighbour's face DoFs,
// again a whole quad).  At k = 4 these seven blocks carry 81 of the 125 DoFs of a
// cell.  This kernel moves them with 1D bulk copies instead of per-thread accesses:
//
//   gather : cp.async.bulk  global -> shared   (7 blocks per cell, mbarrier completion)
//   scatter: cp.reduce.async.bulk.add  shared -> global  (element-wise atomic add at L2)
//
// Bulk copies need 16-byte aligned addresses and sizes: a block that starts at an odd
// entry is widened to the enclosing aligned range; the extra entries are ignored on
// the way in and ZERO on the way out (adding 0 leaves the neighbour untouched).  The
// remaining 12 (k-1) + 8 line / vertex entries of a cell go through ordinary
// coalesced-index gathers and REDs, balanced over the lanes of the warp.
//
// Everything between gather and scatter (hanging-node passes, seven sweeps) is the
// plane kernel's.  Cells whose index array does not show the expected blocks (other
// numberings, non-standard orientation, a block ending at the last vector entry) are
// flagged at setup and run through the plane kernel instead.
#pragma once
#include "kernels_plane.cuh"

#ifndef MFHN_BULK_DEFAULT_OCC
#define MFHN_BULK_DEFAULT_OCC 5
#endif

namespace mfhn
{
__host__ __device__ constexpr int round_up_to(int v, int m) { return (v + m - 1) / m * m; }
constexpr uint32_t bulk_invalid = 0xffffffffu;

template <int n, typename Number>
struct BulkCfg
{
  using P                    = PlaneCfg<n, Number>;
  static constexpr int k     = n - 1, m = n - 2;
  static constexpr int E     = 16 / (int)sizeof(Number); // entries per 16 bytes
  static constexpr int OB    = E == 2 ? 1 : 2;           // bits of a block's alignment offset
  static constexpr int hex   = m * m * m, quad = m * m;
  static constexpr int Hs    = round_up_to(hex + E - 1, E);  // staging entries of the hex block
  static constexpr int Qs    = round_up_to(quad + E - 1, E); // ... of one quad block
  static constexpr int lv    = 12 * m + 8;                   // line and vertex entries of a cell
  static constexpr int LVoff = Hs + 6 * Qs;
  static constexpr int Ss    = round_up_to(LVoff + lv, E); // staging entries per cell
  static constexpr int nblk  = 7 * P::cpw;                 // bulk blocks per warp
  static constexpr int brow  = 64;                         // block descriptors per warp batch (padded)
  static constexpr int NQ    = (6 * P::cpw + 31) / 32;      // quad descriptors per lane
  static constexpr int R     = (P::cpw * lv + 31) / 32; // line / vertex requests per warp
  static constexpr int elems = P::cpw * (P::cs > Ss ? P::cs : Ss);
  static constexpr int warps = 4;
  static constexpr int smem_per_warp = round_up_to(elems * (int)sizeof(Number), 16) + 16; // + mbarrier
  static constexpr int smem  = warps * smem_per_warp;
  static_assert(nblk <= brow, "descriptor row too short");

  // staging offset of block o (0: hex, 1..6: quads x-,x+,y-,y+,z-,z+) and its entry count
  __host__ __device__ static constexpr int block_offset(int o) { return o == 0 ? 0 : Hs + (o - 1) * Qs; }
  __host__ __device__ static constexpr int block_count(int o) { return o == 0 ? hex : quad; }
};

// Position of the lexicographic cell entry (x,y,z) in the staging array of its cell;
// offs = the alignment offsets of the seven blocks, OB bits each.
template <int n, typename Number>
__host__ __device__ __forceinline__ int bulk_pos(const int x, const int y, const int z, const unsigned offs)
{
  using C         = BulkCfg<n, Number>;
  constexpr int k = C::k, m = C::m;
  const bool bx = x == 0 || x == k, by = y == 0 || y == k, bz = z == 0 || z == k;
  const int nb  = (int)bx + (int)by + (int)bz;
  const int hx = x > 0, hy = y > 0, hz = z > 0;
  constexpr unsigned om = (1u << C::OB) - 1u;
  if (nb == 0) return (int)(offs & om) + (x - 1) + m * (y - 1) + m * m * (z - 1);
  if (nb == 1)
    {
      // deal.II face-local axes: x-faces (y,z), y-faces (z,x), z-faces (x,y)
      const int f = bx ? hx : by ? 2 + hy : 4 + hz;
      const int l = bx ? (y - 1) + m * (z - 1) : by ? (z - 1) + m * (x - 1) : (x - 1) + m * (y - 1);
      return C::Hs + f * C::Qs + (int)((offs >> (C::OB * (1 + f))) & om) + l;
    }
  if (nb == 2)
    {
      if (!bx) return C::LVoff + (hy + 2 * hz) * m + (x - 1);
      if (!by) return C::LVoff + 4 * m + (hx + 2 * hz) * m + (y - 1);
      return C::LVoff + 8 * m + (hx + 2 * hy) * m + (z - 1);
    }
  return C::LVoff + 12 * m + hx + 2 * hy + 4 * hz;
}

// Per-thread form of bulk_pos for x = t fixed: the position of slot (y,z) is a base plus compile-time
// multiples of a stride, both depending only on whether t is an interior or a boundary coordinate.
template <int n, typename Number>
struct BulkAddr
{
  int pII, sYII, sZII; // y, z interior: hex (t interior) or an x-face quad
  int pYB[2];          // y = 0 / k, z interior: y-face quad or a z-line
  int pZB[2], sYZB;    // z = 0 / k, y interior: z-face quad or a y-line
  int pC, sC;          // y, z on the boundary: x-line or a vertex

  __host__ __device__ __forceinline__ BulkAddr(const int t, const unsigned offs)
  {
    using C         = BulkCfg<n, Number>;
    constexpr int k = C::k, m = C::m;
    constexpr unsigned om = (1u << C::OB) - 1u;
    const bool ti = t > 0 && t < k;
    const int hx  = t > 0;
    auto oq = [&](const int f) { return (int)((offs >> (C::OB * (1 + f))) & om); };
    pII  = ti ? (int)(offs & om) + (t - 1) : C::Hs + hx * C::Qs + oq(hx);
    sYII = ti ? m : 1;
    sZII = ti ? m * m : m;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      {
        pYB[h] = ti ? C::Hs + (2 + h) * C::Qs + oq(2 + h) + m * (t - 1) : C::LVoff + 8 * m + (hx + 2 * h) * m;
        pZB[h] = ti ? C::Hs + (4 + h) * C::Qs + oq(4 + h) + (t - 1) : C::LVoff + 4 * m + (hx + 2 * h) * m;
      }
    sYZB = ti ? m : 1;
    pC   = ti ? C::LVoff + (t - 1) : C::LVoff + 12 * m + hx;
    sC   = ti ? m : 2;
  }
  __host__ __device__ __forceinline__ int operator()(const int y, const int z) const
  {
    constexpr int k = n - 1;
    const bool by = y == 0 || y == k, bz = z == 0 || z == k;
    if (!by && !bz) return pII + (y - 1) * sYII + (z - 1) * sZII;
    if (by && !bz) return pYB[y > 0] + (z - 1);
    if (!by && bz) return pZB[z > 0] + (y - 1) * sYZB;
    return pC + ((y > 0) + 2 * (z > 0)) * sC;
  }
};

struct BulkParams
{
  const uint32_t *bidx;  // [n_batches][64]: first vector entry of each block (cpw hex blocks, then 6 cpw quads), bulk_invalid = none
  const uint32_t *lvidx; // [n_batches][R][32]: vector entries of the line / vertex slots
  const uint32_t *cinfo; // [n_cells]: alignment offsets of the seven blocks | irregular << 31
  const uint8_t *masks;
  const void *h;
  const void *src;
  void *dst;
  long long batch_begin, batch_end; // whole warp batches (cpw cells each)
  int apply_constraints;
  int hn_mask_strategy; // every warp takes the interpolation passes
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool LOAD, typename Number>
__device__ __forceinline__ void bulk_copy(const unsigned smem, const Number *gmem, const unsigned bar, const unsigned bytes)
{
  if (LOAD)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem), "l"(gmem), "r"(bytes), "r"(bar)
                 : "memory");
  else if (sizeof(Number) == 8)
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gmem), "r"(smem), "r"(bytes) : "memory");
  else
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem), "r"(smem), "r"(bytes) : "memory");
}

// The block descriptors of this lane (one hex block for lanes < cpw, NQ quads), bulk_invalid = none
template <int n, typename Number>
__device__ __forceinline__ void bulk_descriptors(const BulkParams &p, const long long batch, const int lane, uint32_t &hexfirst,
                                                 uint32_t (&qfirst)[BulkCfg<n, Number>::NQ])
{
  using B  = BulkCfg<n, Number>;
  constexpr int cpw = B::P::cpw;
  const uint32_t *bp = p.bidx + batch * (long long)B::brow;
  hexfirst = lane < cpw ? __ldg(bp + lane) : bulk_invalid;
#pragma unroll
  for (int r = 0; r < B::NQ; ++r)
    {
      const int q = r * 32 + lane;
      qfirst[r]   = q < 6 * cpw ? __ldg(bp + cpw + q) : bulk_invalid;
    }
}

// One bulk copy per block, issued by the lane that holds its descriptor.  Sizes are fixed per block type
// (the widened block [first - off, first - off + Hs) resp. Qs): only addresses differ between lanes.
template <int n, typename Number, bool LOAD>
__device__ __forceinline__ void bulk_issue(Number *A, const Number *vec, const unsigned bar, const int lane, const uint32_t hexfirst,
                                           const uint32_t (&qfirst)[BulkCfg<n, Number>::NQ])
{
  using B = BulkCfg<n, Number>;
  constexpr uint32_t am = ~(uint32_t)(B::E - 1);
  if (hexfirst != bulk_invalid) bulk_copy<LOAD>(smem_u32(A + lane * B::Ss), vec + (hexfirst & am), bar, (unsigned)(B::Hs * sizeof(Number)));
#pragma unroll
  for (int r = 0; r < B::NQ; ++r)
    if (qfirst[r] != bulk_invalid)
      {
        const int q = r * 32 + lane, c = q / 6, f = q - 6 * c;
        bulk_copy<LOAD>(smem_u32(A + c * B::Ss + B::Hs + f * B::Qs), vec + (qfirst[r] & am), bar, (unsigned)(B::Qs * sizeof(Number)));
      }
}

// OCC = CTAs per SM the register allocation is limited for (4: 128, 5: 96, 6: 80 registers per thread)
template <int n, typename Number, int OCC = 4>
__global__ void __launch_bounds__(BulkCfg<n, Number>::warps * 32, OCC) bulk_cell_kernel(const BulkParams p)
{
  using Cfg = PlaneCfg<n, Number>;
  using B   = BulkCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs, E = B::E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = p.batch_begin + (long long)blockIdx.x * B::warps + warp;
  if (batch >= p.batch_end) return; // warps are independent: no block-level barrier below
  // one array per warp, used first as the object-ordered staging area S (stride Ss per cell), then as the
  // lexicographic cell arrays A of the plane kernel (stride cs), finally as S again
  Number *A          = reinterpret_cast<Number *>(smem_raw + (size_t)warp * B::smem_per_warp);
  const unsigned bar = smem_u32(smem_raw + (size_t)warp * B::smem_per_warp + (B::smem_per_warp - 16));

  const bool active = lane < Cfg::lanes;
  const int ml = active ? lane : lane - 16; // idle lanes mirror lane - 16 (see the plane kernel)
  const int c = ml / n, t = ml - c * n;
  const long long cell = batch * Cfg::cpw + c; // the launcher passes whole batches only: every cell exists
  const Number *__restrict__ src = static_cast<const Number *>(p.src);
  Number *__restrict__ dst = static_cast<Number *>(p.dst);
  const uint32_t *lvp = p.lvidx + batch * (long long)(B::R * 32) + lane;

  // ---- gather ---------------------------------------------------------------------------
  // all index loads first: their latency overlaps the barrier set-up
  uint32_t hexfirst, qfirst[B::NQ], g[B::R];
  bulk_descriptors<n, Number>(p, batch, lane, hexfirst, qfirst);
#pragma unroll
  for (int r = 0; r < B::R; ++r) g[r] = __ldg(lvp + r * 32);
  const unsigned info = __ldg(p.cinfo + cell);
  const bool valid    = !(info >> 31);
  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
  if (lane == 0)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncwarp();
  {
    unsigned total = hexfirst != bulk_invalid ? (unsigned)(B::Hs * sizeof(Number)) : 0u;
#pragma unroll
    for (int r = 0; r < B::NQ; ++r) total += qfirst[r] != bulk_invalid ? (unsigned)(B::Qs * sizeof(Number)) : 0u;
    total = __reduce_add_sync(0xffffffffu, total);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
    __syncwarp();
    bulk_issue<n, Number, true>(A, src, bar, lane, hexfirst, qfirst);
  }
  {
    // line and vertex entries: R balanced requests, values to the staging area
    Number v[B::R];
#pragma unroll
    for (int r = 0; r < B::R; ++r)
      {
        v[r] = g[r] != bulk_invalid ? __ldg(src + g[r]) : Number(0);
      }
#pragma unroll
    for (int r = 0; r < B::R; ++r)
      {
        const int e = r * 32 + lane, ec = e / B::lv;
        if (g[r] != bulk_invalid) A[ec * B::Ss + B::LVoff + (e - ec * B::lv)] = v[r];
      }
  }
  const bool any_hn = p.hn_mask_strategy || __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);
  __syncwarp();
  asm volatile("{\n.reg .pred pw;\nBULK_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 pw, [%0], 0;\n@!pw bra BULK_WAIT;\n}" ::"r"(bar) : "memory");

  // thread = x, plane slot j = y + n z.  (Cells left to the plane kernel compute on whatever the staging
  // area holds and never store.)
  Number u[n][n];
  {
    const BulkAddr<n, Number> pos(t, info);
    const Number *S = A + c * B::Ss;
#pragma unroll
    for (int j = 0; j < n * n; ++j) u[j / n][j % n] = S[pos(j % n, j / n)];
  }
  __syncwarp(); // the staging area is dead: the array now holds the lexicographic cell arrays
  Number *cellA = A + c * cs;
  if (any_hn)
    {
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
  __syncwarp(); // every lane has read its plane back
  if (any_hn)
    {
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n * n; ++j) cellA[t * ps + j] = u[j / n][j % n];
        }
      __syncwarp();
      hn_smem<n, true>(cellA, hn_face, hn_edge, hn_cb, t, active);
#pragma unroll
      for (int j = 0; j < n * n; ++j) u[j / n][j % n] = cellA[t * ps + j];
      __syncwarp();
    }

  // ---- scatter ---------------------------------------------------------------------------
  // the results go from registers straight to the staging layout
  bulk_descriptors<n, Number>(p, batch, lane, hexfirst, qfirst);
#pragma unroll
  for (int r = 0; r < B::R; ++r) g[r] = __ldg(lvp + r * 32);
  {
    const unsigned info2 = __ldg(p.cinfo + cell); // loaded again instead of living through the sweeps
    if (active && !(info2 >> 31))
      {
        const BulkAddr<n, Number> pos(t, info2);
        Number *S = A + c * B::Ss;
#pragma unroll
        for (int j = 0; j < n * n; ++j) S[pos(j % n, j / n)] = u[j / n][j % n];
      }
  }
  // the widened blocks add into neighbouring vector entries: those staging slots must hold zero
  if (hexfirst != bulk_invalid)
    {
      Number *blk   = A + lane * B::Ss;
      const int off = (int)(hexfirst & (uint32_t)(E - 1));
#pragma unroll
      for (int e = 0; e < E - 1; ++e)
        if (e < off) blk[e] = Number(0);
#pragma unroll
      for (int e = B::hex; e < B::Hs; ++e)
        if (e >= off + B::hex) blk[e] = Number(0);
    }
#pragma unroll
  for (int r = 0; r < B::NQ; ++r)
    if (qfirst[r] != bulk_invalid)
      {
        const int q = r * 32 + lane, qc = q / 6, f = q - 6 * qc;
        Number *blk   = A + qc * B::Ss + B::Hs + f * B::Qs;
        const int off = (int)(qfirst[r] & (uint32_t)(E - 1));
#pragma unroll
        for (int e = 0; e < E - 1; ++e)
          if (e < off) blk[e] = Number(0);
#pragma unroll
        for (int e = B::quad; e < B::Qs; ++e)
          if (e >= off + B::quad) blk[e] = Number(0);
      }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // staging writes -> visible to the bulk engine
  __syncwarp();
  bulk_issue<n, Number, false>(A, dst, 0u, lane, hexfirst, qfirst);
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#pragma unroll
  for (int r = 0; r < B::R; ++r)
    {
      const int e = r * 32 + lane, ec = e / B::lv;
      if (g[r] != bulk_invalid) atomicAdd(dst + g[r], A[ec * B::Ss + B::LVoff + (e - ec * B::lv)]);
    }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the staging area must outlive the bulk reads
}

// ---- host side ------------------------------------------------------------------
template <int n, typename Number, int OCC>
void launch_bulk_occ(const BulkParams &p, const unsigned grid, int device, cudaStream_t stream)
{
  using B = BulkCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(bulk_cell_kernel<n, Number, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, B::smem);
      if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      attr[device] = true;
    }
  bulk_cell_kernel<n, Number, OCC><<<grid, B::warps * 32, B::smem, stream>>>(p);
}

template <int n, typename Number>
void launch_bulk_impl(const BulkLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  using Cfg = PlaneCfg<n, Number>;
  using B   = BulkCfg<n, Number>;
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
  BulkParams p;
  p.bidx              = L.d_bidx;
  p.lvidx             = L.d_lvidx;
  p.cinfo             = L.d_cinfo;
  p.masks             = cp.masks;
  p.h                 = cp.geom;
  p.src               = cp.src;
  p.dst               = cp.dst;
  if (cp.cell_begin % Cfg::cpw || cp.cell_end % Cfg::cpw || cp.cell_end > L.n_batches * Cfg::cpw)
    throw std::runtime_error("bulk kernel: cell range must consist of whole warp batches");
  p.batch_begin       = cp.cell_begin / Cfg::cpw;
  p.batch_end         = cp.cell_end / Cfg::cpw;
  p.apply_constraints = cp.apply_constraints;
  p.hn_mask_strategy  = cp.hn_mask_strategy && cp.apply_constraints;
  const long long nb  = p.batch_end - p.batch_begin;
  if (nb <= 0) return;
  const unsigned grid = (unsigned)((nb + B::warps - 1) / B::warps);
  // the double-precision kernels of degree >= 4 are register-bound: 5 CTAs (96 registers) measured best on B200
  constexpr bool reg_bound = sizeof(Number) == 8 && n >= 5;
  const int occ = occupancy_choice(n, sizeof(Number) == 8, n == 5 ? MFHN_BULK_DEFAULT_OCC : 4); // k = 5: 128 registers measured best
  if (reg_bound && occ == 5)
    launch_bulk_occ<n, Number, reg_bound ? 5 : 4>(p, grid, device, stream);
  else if (reg_bound && occ == 6)
    launch_bulk_occ<n, Number, reg_bound ? 6 : 4>(p, grid, device, stream);
  else
    launch_bulk_occ<n, Number, 4>(p, grid, device, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("bulk kernel launch: ") + cudaGetErrorString(e));
}

template <int n, typename Number>
void launch_bulk(const BulkLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  if constexpr (bulk_supported(n))
    launch_bulk_impl<n, Number>(L, cp, device, stream);
  else
    throw std::runtime_error("bulk kernel not available for this degree");
}

// Host analysis of the reference index array: block descriptors, line / vertex index lists, irregular cells.
template <int n, typename Number>
void bulk_analyze_impl(BulkHostLayout &L, long long n_cells, long long n_vec, const uint32_t *idx)
{
  using Cfg = PlaneCfg<n, Number>;
  using B   = BulkCfg<n, Number>;
  constexpr int n3 = n * n * n, k = B::k, m = B::m, E = B::E;
  L.n         = n;
  L.E         = E;
  L.n_cells   = n_cells;
  L.n_batches = (n_cells + Cfg::cpw - 1) / Cfg::cpw;
  L.bidx.assign((size_t)std::max<long long>(L.n_batches, 1) * B::brow, bulk_invalid);
  L.lvidx.assign((size_t)std::max<long long>(L.n_batches, 1) * B::R * 32, bulk_invalid);
  L.cinfo.assign((size_t)std::max<long long>(L.n_batches, 1) * Cfg::cpw, 0x80000000u); // padded to whole batches
  std::vector<unsigned char> irr((size_t)std::max<long long>(n_cells, 1), 0);
#pragma omp parallel for schedule(static)
  for (long long c = 0; c < n_cells; ++c)
    {
      const uint32_t *ci = idx + c * n3;
      const long long batch = c / Cfg::cpw;
      const int slot        = (int)(c % Cfg::cpw);
      auto lex = [](int x, int y, int z) { return x + n * (y + n * z); };
      // first entries of the seven blocks: hex, quads x-,x+,y-,y+,z-,z+ (first face-local entry = (1,1))
      uint32_t first[7];
      first[0] = ci[lex(1, 1, 1)];
      first[1] = ci[lex(0, 1, 1)];
      first[2] = ci[lex(k, 1, 1)];
      first[3] = ci[lex(1, 0, 1)];
      first[4] = ci[lex(1, k, 1)];
      first[5] = ci[lex(1, 1, 0)];
      first[6] = ci[lex(1, 1, k)];
      unsigned offs = 0;
      bool regular  = true;
      for (int o = 0; o < 7; ++o)
        {
          const unsigned off = first[o] & (unsigned)(E - 1);
          offs |= off << (B::OB * o);
          // the widened block (fixed size per block type) must stay inside the vector
          if ((long long)(first[o] - off) + (o == 0 ? B::Hs : B::Qs) > n_vec) regular = false;
        }
      // every entry of the cell must sit where the block pattern says
      for (int z = 0; z < n && regular; ++z)
        for (int y = 0; y < n && regular; ++y)
          for (int x = 0; x < n; ++x)
            {
              const int nb = (x == 0 || x == k) + (y == 0 || y == k) + (z == 0 || z == k);
              if (nb >= 2) continue;
              const int pos = bulk_pos<n, Number>(x, y, z, offs);
              int o = 0;
              while (o < 6 && pos >= B::block_offset(o + 1)) ++o;
              const long long expect = (long long)first[o] - (first[o] & (unsigned)(E - 1)) + (pos - B::block_offset(o));
              if ((long long)ci[lex(x, y, z)] != expect)
                {
                  regular = false;
                  break;
                }
            }
      if (!regular)
        {
          irr[c] = 1;
          continue;
        }
      L.cinfo[c] = offs;
      L.bidx[(size_t)batch * B::brow + slot] = first[0];
      for (int f = 0; f < 6; ++f) L.bidx[(size_t)batch * B::brow + Cfg::cpw + slot * 6 + f] = first[1 + f];
      for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
          for (int x = 0; x < n; ++x)
            {
              const int nb = (x == 0 || x == k) + (y == 0 || y == k) + (z == 0 || z == k);
              if (nb < 2) continue;
              const int e = slot * B::lv + (bulk_pos<n, Number>(x, y, z, offs) - B::LVoff);
              L.lvidx[((size_t)batch * B::R + e / 32) * 32 + e % 32] = ci[lex(x, y, z)];
            }
    }
  for (long long c = 0; c < n_cells; ++c)
    if (irr[c]) L.irregular.push_back(c);
}

// Emulation of the kernel's gather through the layout with src[i] = i: returns the number of cell entries
// that do not come out as the reference index array says (0 = the layout reproduces it).
template <int n, typename Number>
long long bulk_verify_impl(const BulkHostLayout &L, const uint32_t *idx)
{
  using Cfg = PlaneCfg<n, Number>;
  using B   = BulkCfg<n, Number>;
  constexpr int n3 = n * n * n, E = B::E;
  long long bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (long long batch = 0; batch < L.n_batches; ++batch)
    {
      std::vector<long long> S((size_t)Cfg::cpw * B::Ss, -1);
      for (int i = 0; i < B::nblk; ++i)
        {
          const uint32_t f = L.bidx[(size_t)batch * B::brow + i];
          if (f == bulk_invalid) continue;
          const int cw = i < Cfg::cpw ? i : (i - Cfg::cpw) / 6, o = i < Cfg::cpw ? 0 : 1 + (i - Cfg::cpw) % 6;
          const int off = (int)(f & (unsigned)(E - 1)), size = o == 0 ? B::Hs : B::Qs;
          for (int e = 0; e < size; ++e) S[(size_t)cw * B::Ss + B::block_offset(o) + e] = (long long)(f - off) + e;
        }
      for (int e = 0; e < Cfg::cpw * B::lv; ++e)
        {
          const uint32_t g = L.lvidx[((size_t)batch * B::R + e / 32) * 32 + e % 32];
          if (g != bulk_invalid) S[(size_t)(e / B::lv) * B::Ss + B::LVoff + e % B::lv] = g;
        }
      for (int s = 0; s < Cfg::cpw; ++s)
        {
          const long long c = batch * Cfg::cpw + s;
          if (c >= L.n_cells || (L.cinfo[c] >> 31)) continue;
          for (int j = 0; j < n3; ++j)
            {
              const int x = j % n, y = (j / n) % n, z = j / (n * n);
              const BulkAddr<n, Number> pos(x, L.cinfo[c]); // the kernel's per-thread form must agree with bulk_pos
              if (pos(y, z) != bulk_pos<n, Number>(x, y, z, L.cinfo[c])) ++bad;
              if (S[(size_t)s * B::Ss + pos(y, z)] != (long long)idx[c * n3 + j]) ++bad;
            }
        }
    }
  return bad;
}
} // namespace mfhn
