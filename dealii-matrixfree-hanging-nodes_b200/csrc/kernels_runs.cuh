// Plane kernel with run-wise vector access through the bulk-copy engine (TMA).
//
// deal.II numbers the DoFs cell by cell: the vertices, lines, quads and the interior a cell sees first form ONE
// contiguous chunk of the vector, and what it shares with earlier cells sits in a few contiguous pieces of their
// chunks (a face quad followed by lines of that face, ...).  The bulk-copy kernel (kernels_bulk.cuh) knows seven
// fixed blocks per cell and moves the remaining 12 (k-1) + 8 line / vertex entries one by one.  This kernel makes
// no assumption about the numbering: at setup the sorted index list of every cell is cut into RUNS of (nearly)
// consecutive vector entries.  Long runs become bulk copies
//
//   gather : cp.async.bulk  global -> shared   (mbarrier completion)
//   scatter: cp.reduce.async.bulk.add  shared -> global  (element-wise atomic add at L2)
//
// of the enclosing 16-byte aligned range; entries of that range the cell does not use are ignored on the way in and
// ZERO on the way out.  Short runs are single entries (coalesced index list, one gather / RED per entry).  On the
// annulus mesh at k = 4 a cell without hanging nodes needs about 6 bulk copies and 18 single entries instead of 7
// and 44.  A per-thread table of byte positions says where each plane slot sits in the cell's staging area; the
// positions are chosen at setup so that the lanes of a half-warp hit different shared-memory banks where possible.
// All index data of a warp batch sit in one fixed-size record (no dependent index loads).
//
// Everything between gather and scatter (hanging-node passes, six sweeps) is the plane kernel's.
#pragma once
#include "kernels_bulk.cuh"

#include <algorithm>
#include <cstring>


#ifndef MFHN_RUNS_CAP_PCT
#define MFHN_RUNS_CAP_PCT 50 // staging room beyond the (k+1)^3 cell entries: copied foreign entries, alignment, placement slack
#endif

namespace mfhn
{
template <int n, typename Number>
struct RunsCfg
{
  using P                    = PlaneCfg<n, Number>;
  static constexpr int E     = 16 / (int)sizeof(Number); // entries per 16 bytes
  static constexpr int n3    = n * n * n;
  static constexpr int cap0  = round_up_to(n3 + n3 * MFHN_RUNS_CAP_PCT / 100, 8);
  static constexpr int cap   = cap0 > 256 ? 256 : cap0; // staging entries per cell (positions are bytes)
  static constexpr int NT    = (n * n + 3) / 4;         // position words per thread
  static constexpr int BR    = 2;                       // descriptor rounds: at most 64 bulk copies per warp batch
  static constexpr int maxb  = (32 * BR) / P::cpw;      // bulk copies per cell
  static constexpr int SU    = 4;                       // single-entry requests in flight per lane
  static constexpr int RW    = 4 + 64 * BR + 32 * NT;   // 32-bit words of a batch record
  static constexpr int idle  = 32 - P::lanes;           // idle lanes mirror lane - idle (same half-warp: broadcast loads)
  static constexpr int elems = P::cpw * (P::cs > cap ? P::cs : cap);
#ifndef MFHN_RUNS_SCATTER_HOIST
#define MFHN_RUNS_SCATTER_HOIST 1
#endif
#ifndef MFHN_RUNS_WARPS
#define MFHN_RUNS_WARPS 4
#endif
  static constexpr int warps = MFHN_RUNS_WARPS;
  static constexpr int smem_per_warp = round_up_to(elems * (int)sizeof(Number), 16) + 16; // + mbarrier
  static constexpr int smem  = warps * smem_per_warp;
  static_assert(n3 <= cap, "staging area too small");
  static_assert(idle < 16, "idle lanes must have their twin in the same half-warp");
};

struct RunsParams
{
  // [n_batches][RW] records: {first overflow single, first zero entry, n_zero | n_overflow << 16, 0},
  //   BR x 32 bulk copies {first vector entry (16-byte aligned), staging entry of the warp | 16-byte units << 16},
  //   NT x 32 position words (staging positions of the lane's plane slots, 4 bytes per word)
  const uint32_t *rec;
  const uint32_t *srow;   // [n_batches][sr][32] vector entries of the single entries, bulk_invalid = none
  const uint16_t *sprow;  // their staging entries
  const uint32_t *ov_idx; // single entries beyond sr x 32 of a batch
  const uint16_t *ov_pos;
  const uint16_t *zpos;   // staging entries that must be zero before the scatter
  const uint8_t *masks;
  const void *h;
  const void *src;
  void *dst;
  long long batch_begin, batch_end, n_cells;
  int sr;
  int apply_constraints;
  int hn_mask_strategy; // every warp takes the interpolation passes
};

// OCC = CTAs per SM the register allocation is limited for (4: 128, 5: 96, 6: 80 registers per thread)
template <int n, typename Number, int OCC = 4>
__global__ void __launch_bounds__(RunsCfg<n, Number>::warps * 32, OCC) runs_cell_kernel(const RunsParams p)
{
  using Cfg = PlaneCfg<n, Number>;
  using R   = RunsCfg<n, Number>;
  constexpr int ps = Cfg::ps, cs = Cfg::cs;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = p.batch_begin + (long long)blockIdx.x * R::warps + warp;
  if (batch >= p.batch_end) return; // warps are independent: no block-level barrier below
  // one array per warp, used first as the staging area S (stride cap per cell), then as the lexicographic cell
  // arrays A of the plane kernel (stride cs), finally as S again
  Number *A          = reinterpret_cast<Number *>(smem_raw + (size_t)warp * R::smem_per_warp);
  const unsigned bar = smem_u32(smem_raw + (size_t)warp * R::smem_per_warp + (R::smem_per_warp - 16));

  const bool active = lane < Cfg::lanes;
  const int ml = active ? lane : lane - R::idle;
  const int c = ml / n, t = ml - c * n;
  const long long cell = batch * Cfg::cpw + c;
  const bool valid     = cell < p.n_cells;
  const Number *__restrict__ src = static_cast<const Number *>(p.src);
  Number *__restrict__ dst = static_cast<Number *>(p.dst);
  const uint32_t *rec  = p.rec + batch * (long long)R::RW;
  const uint2 *bp      = reinterpret_cast<const uint2 *>(rec + 4) + lane;
  const uint32_t *tp   = rec + 4 + 64 * R::BR + lane;
  const uint32_t *sp_i = p.srow + batch * (long long)p.sr * 32 + lane;
  const uint16_t *sp_p = p.sprow + batch * (long long)p.sr * 32 + lane;

  // ---- gather ---------------------------------------------------------------------------
  // all index loads first (they are independent of each other): one exposed memory latency, not three
  uint2 bd[R::BR];
  unsigned total = 0;
#pragma unroll
  for (int r = 0; r < R::BR; ++r) bd[r] = __ldg(bp + r * 32);
  uint32_t g0[R::SU];
  unsigned sp0[R::SU];
#pragma unroll
  for (int u = 0; u < R::SU; ++u)
    {
      g0[u]  = u < p.sr ? __ldg(sp_i + u * 32) : bulk_invalid;
      sp0[u] = u < p.sr ? __ldg(sp_p + u * 32) : 0u;
    }
  const uint4 hdr = __ldg(reinterpret_cast<const uint4 *>(rec));
  uint32_t tab[R::NT];
#pragma unroll
  for (int q = 0; q < R::NT; ++q) tab[q] = __ldg(tp + q * 32);
  const unsigned mask = (valid && p.apply_constraints) ? p.masks[cell] : 0u;
  const Number h      = valid ? static_cast<const Number *>(p.h)[cell] : Number(0);
  if (lane == 0)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < R::BR; ++r) total += (bd[r].y >> 16) * 16u;
  total = __reduce_add_sync(0xffffffffu, total);
  if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
  __syncwarp();
#pragma unroll
  for (int r = 0; r < R::BR; ++r)
    if (bd[r].y >> 16) bulk_copy<true>(smem_u32(A + (bd[r].y & 0xffffu)), src + bd[r].x, bar, (bd[r].y >> 16) * 16u);
  // single entries: SU independent requests per lane and round
  {
    Number v[R::SU];
#pragma unroll
    for (int u = 0; u < R::SU; ++u) v[u] = g0[u] != bulk_invalid ? __ldg(src + g0[u]) : Number(0);
#pragma unroll
    for (int u = 0; u < R::SU; ++u)
      if (g0[u] != bulk_invalid) A[sp0[u]] = v[u];
  }
  for (int r0 = R::SU; r0 < p.sr; r0 += R::SU)
    {
      uint32_t g[R::SU];
      unsigned sp[R::SU];
      Number v[R::SU];
#pragma unroll
      for (int u = 0; u < R::SU; ++u)
        {
          const bool in = r0 + u < p.sr;
          g[u]          = in ? __ldg(sp_i + (r0 + u) * 32) : bulk_invalid;
          sp[u]         = in ? __ldg(sp_p + (r0 + u) * 32) : 0u;
        }
#pragma unroll
      for (int u = 0; u < R::SU; ++u) v[u] = g[u] != bulk_invalid ? __ldg(src + g[u]) : Number(0);
#pragma unroll
      for (int u = 0; u < R::SU; ++u)
        if (g[u] != bulk_invalid) A[sp[u]] = v[u];
    }
  const int nov = (int)(hdr.z >> 16), nz = (int)(hdr.z & 0xffffu);
  for (int i = lane; i < nov; i += 32) A[__ldg(p.ov_pos + hdr.x + i)] = __ldg(src + __ldg(p.ov_idx + hdr.x + i));
  const bool any_hn   = p.hn_mask_strategy || __any_sync(0xffffffffu, mask != 0u);
  unsigned hn_face, hn_edge, hn_cb;
  decode_mask_kernel_axes(mask, hn_face, hn_edge, hn_cb);
  __syncwarp();
  asm volatile("{\n.reg .pred pw;\nRUNS_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 pw, [%0], 0;\n@!pw bra RUNS_WAIT;\n}" ::"r"(bar) : "memory");

  // thread = x, plane slot j = y + n z
  Number u[n][n];
  {
    const Number *S = A + c * R::cap;
#pragma unroll
    for (int j = 0; j < n * n; ++j) u[j / n][j % n] = S[(tab[j / 4] >> (8 * (j % 4))) & 0xffu];
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

#ifndef MFHN_RUNS_NO_PREFETCH
  // the scatter reads the record again: keep its lines close (one 128-byte line per lane)
  if (lane * 32 < R::RW) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + lane * 32));
  if (lane < p.sr) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp_i - lane + lane * 32));
  if (lane * 2 < p.sr) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp_p - lane + lane * 64));
  if (lane * 64 < nz) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.zpos + hdr.y + lane * 64));
#endif
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
  // the results go from registers straight to the staging layout (descriptors and positions are loaded again
  // instead of living through the sweeps); all index loads of the scatter are issued up front
#pragma unroll
  for (int q = 0; q < R::NT; ++q) tab[q] = __ldg(tp + q * 32);
#pragma unroll
  for (int r = 0; r < R::BR; ++r) bd[r] = __ldg(bp + r * 32);
#if MFHN_RUNS_SCATTER_HOIST
#pragma unroll
  for (int u2 = 0; u2 < R::SU; ++u2)
    {
      g0[u2]  = u2 < p.sr ? __ldg(sp_i + u2 * 32) : bulk_invalid;
      sp0[u2] = u2 < p.sr ? __ldg(sp_p + u2 * 32) : 0u;
    }
#endif
  const unsigned z0 = lane < nz ? __ldg(p.zpos + hdr.y + lane) : 0xffffffffu;
  if (active && valid)
    {
      Number *S = A + c * R::cap;
#pragma unroll
      for (int j = 0; j < n * n; ++j) S[(tab[j / 4] >> (8 * (j % 4))) & 0xffu] = u[j / n][j % n];
    }
  // entries of the copied ranges that do not belong to the cell add into foreign vector entries: they must hold zero
  if (z0 != 0xffffffffu) A[z0] = Number(0);
  for (int i = lane + 32; i < nz; i += 32) A[__ldg(p.zpos + hdr.y + i)] = Number(0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // staging writes -> visible to the bulk engine
  __syncwarp();
#pragma unroll
  for (int r = 0; r < R::BR; ++r)
    if (bd[r].y >> 16) bulk_copy<false>(smem_u32(A + (bd[r].y & 0xffffu)), dst + bd[r].x, 0u, (bd[r].y >> 16) * 16u);
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#if !MFHN_RUNS_SCATTER_HOIST
#pragma unroll
  for (int u2 = 0; u2 < R::SU; ++u2)
    {
      g0[u2]  = u2 < p.sr ? __ldg(sp_i + u2 * 32) : bulk_invalid;
      sp0[u2] = u2 < p.sr ? __ldg(sp_p + u2 * 32) : 0u;
    }
#endif
#pragma unroll
  for (int u2 = 0; u2 < R::SU; ++u2)
    if (g0[u2] != bulk_invalid) atomicAdd(dst + g0[u2], A[sp0[u2]]);
  for (int r0 = R::SU; r0 < p.sr; r0 += R::SU)
    {
      uint32_t g[R::SU];
      unsigned sp[R::SU];
#pragma unroll
      for (int u2 = 0; u2 < R::SU; ++u2)
        {
          const bool in = r0 + u2 < p.sr;
          g[u2]         = in ? __ldg(sp_i + (r0 + u2) * 32) : bulk_invalid;
          sp[u2]        = in ? __ldg(sp_p + (r0 + u2) * 32) : 0u;
        }
#pragma unroll
      for (int u2 = 0; u2 < R::SU; ++u2)
        if (g[u2] != bulk_invalid) atomicAdd(dst + g[u2], A[sp[u2]]);
    }
  for (int i = lane; i < nov; i += 32) atomicAdd(dst + __ldg(p.ov_idx + hdr.x + i), A[__ldg(p.ov_pos + hdr.x + i)]);
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the staging area must outlive the bulk reads
}

// ---- host side ------------------------------------------------------------------
inline int runs_env(const char *name, int fallback) // development switches of the placement
{
  const char *e = std::getenv(name);
  return e && *e ? std::atoi(e) : fallback;
}

template <int n, typename Number, int OCC>
void launch_runs_occ(const RunsParams &p, const unsigned grid, int device, cudaStream_t stream)
{
  using R = RunsCfg<n, Number>;
  static bool attr[64] = {};
  if (!attr[device])
    {
      cudaError_t e = cudaFuncSetAttribute(runs_cell_kernel<n, Number, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, R::smem);
      if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      attr[device] = true;
    }
  runs_cell_kernel<n, Number, OCC><<<grid, R::warps * 32, R::smem, stream>>>(p);
}

template <int n, typename Number>
void launch_runs_impl(const RunsLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  using Cfg = PlaneCfg<n, Number>;
  using R   = RunsCfg<n, Number>;
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
  RunsParams p;
  p.rec    = L.d_rec;
  p.srow   = L.d_srow;
  p.sprow  = L.d_sprow;
  p.ov_idx = L.d_ov_idx;
  p.ov_pos = L.d_ov_pos;
  p.zpos   = L.d_zpos;
  p.sr     = L.sr;
  p.masks  = cp.masks;
  p.h      = cp.geom;
  p.src    = cp.src;
  p.dst    = cp.dst;
  if (cp.cell_begin % Cfg::cpw || (cp.cell_end % Cfg::cpw && cp.cell_end != L.n_cells) || cp.cell_end > L.n_cells)
    throw std::runtime_error("runs kernel: cell range must consist of whole warp batches");
  p.batch_begin       = cp.cell_begin / Cfg::cpw;
  p.batch_end         = (cp.cell_end + Cfg::cpw - 1) / Cfg::cpw;
  p.n_cells           = L.n_cells;
  p.apply_constraints = cp.apply_constraints;
  p.hn_mask_strategy  = cp.hn_mask_strategy && cp.apply_constraints;
  const long long nb  = p.batch_end - p.batch_begin;
  if (nb <= 0) return;
  const unsigned grid = (unsigned)((nb + R::warps - 1) / R::warps);
  constexpr bool reg_bound = sizeof(Number) == 8 && n >= 5;
  const int occ = occupancy_choice(n, sizeof(Number) == 8, n == 5 ? 5 : 4); // measured on B200 (k = 4: 5 CTAs 120.4, 6 CTAs 109.6-114.7; k = 5: 4 CTAs 138.2, 5 CTAs 98.5 GDoF/s)
  if (reg_bound && occ == 5)
    launch_runs_occ<n, Number, reg_bound ? 5 : 4>(p, grid, device, stream);
  else if (reg_bound && occ == 6)
    launch_runs_occ<n, Number, reg_bound ? 6 : 4>(p, grid, device, stream);
  else
    launch_runs_occ<n, Number, 4>(p, grid, device, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("runs kernel launch: ") + cudaGetErrorString(e));
}

template <int n, typename Number>
void launch_runs(const RunsLayout &L, const CellLoopParams &cp, int device, cudaStream_t stream)
{
  if constexpr (runs_supported(n))
    launch_runs_impl<n, Number>(L, cp, device, stream);
  else
    throw std::runtime_error("runs kernel not available for this degree");
}

// Host analysis of the reference index array.  max_gap = unused vector entries tolerated inside a run, min_run = fewest
// cell entries worth a bulk copy, place = choose staging positions against shared-memory bank conflicts.
template <int n, typename Number>
void runs_analyze_impl(RunsHostLayout &L, long long n_cells, long long n_vec, const uint32_t *idx, int max_gap, int min_run, bool place)
{
  using Cfg = PlaneCfg<n, Number>;
  using R   = RunsCfg<n, Number>;
  constexpr int n3 = R::n3, E = R::E, cpw = Cfg::cpw, nn = n * n;
  L.n         = n;
  L.E         = E;
  L.cap       = R::cap;
  L.NT        = R::NT;
  L.RW        = R::RW;
  L.BR        = R::BR;
  L.n_cells   = n_cells;
  L.n_batches = (n_cells + cpw - 1) / cpw;
  const long long nbat = L.n_batches;
  struct Block
  {
    uint32_t first; // aligned
    int size;       // entries, multiple of E
    int count;      // cell entries inside
    int a, b;       // range of the sorted unique list
  };
  struct Sorted
  {
    std::pair<uint32_t, int> s[n3];
    int uq[n3], nu, dup[n3], nd;
  };
  // runs of one cell: blocks (ascending) and single entries (positions in S.s)
  auto cell_runs = [&](const uint32_t *ci, Sorted &S, std::vector<Block> &blocks, std::vector<int> &single) {
    for (int j = 0; j < n3; ++j) S.s[j] = {ci[j], j};
    std::sort(S.s, S.s + n3);
    // unique entries in ascending order; repeated entries of a cell get staging slots of their own
    S.nu = S.nd = 0;
    for (int j = 0; j < n3; ++j)
      if (j > 0 && S.s[j].first == S.s[j - 1].first)
        S.dup[S.nd++] = j;
      else
        S.uq[S.nu++] = j;
    int gap = max_gap;
    for (;;)
      {
        blocks.clear();
        single.assign(S.dup, S.dup + S.nd);
        int a = 0;
        while (a < S.nu)
          {
            int b = a;
            while (b + 1 < S.nu && (long long)S.s[S.uq[b + 1]].first - (long long)S.s[S.uq[b]].first - 1 <= gap) ++b;
            const int count      = b - a + 1;
            const uint32_t first = S.s[S.uq[a]].first & ~(uint32_t)(E - 1);
            const long long end  = ((long long)S.s[S.uq[b]].first + E) / E * E;
            if (count >= min_run && end <= n_vec && end - first <= (long long)R::cap)
              blocks.push_back(Block{first, (int)(end - first), count, a, b});
            else
              for (int i = a; i <= b; ++i) single.push_back(S.uq[i]);
            a = b + 1;
          }
        // too many bulk copies or too much staging: the smallest runs become single entries
        auto total = [&]() {
          int t = (int)single.size();
          for (const Block &B : blocks) t += B.size;
          return t;
        };
        while ((int)blocks.size() > R::maxb || (total() > R::cap && !blocks.empty() && gap == 0))
          {
            size_t w = 0;
            for (size_t i = 1; i < blocks.size(); ++i)
              if (blocks[i].count < blocks[w].count) w = i;
            for (int i = blocks[w].a; i <= blocks[w].b; ++i) single.push_back(S.uq[i]);
            blocks.erase(blocks.begin() + (long)w);
          }
        if (total() <= R::cap) break;
        gap = 0; // retry without tolerated gaps (wastes at most E - 1 entries at either end of a run)
      }
    std::sort(single.begin(), single.end());
  };
  // pass 1: single entries per batch -> rounds of the fixed-stride single-entry rows (covers 7 batches of 8)
  std::vector<int> ns_batch((size_t)std::max<long long>(nbat, 1), 0);
#pragma omp parallel
  {
    Sorted S;
    std::vector<Block> blocks;
    std::vector<int> single;
#pragma omp for schedule(static)
    for (long long c = 0; c < n_cells; ++c)
      {
        cell_runs(idx + c * n3, S, blocks, single);
#pragma omp atomic
        ns_batch[c / cpw] += (int)single.size();
      }
  }
  {
    std::vector<int> sorted_ns(ns_batch.begin(), ns_batch.begin() + nbat);
    std::sort(sorted_ns.begin(), sorted_ns.end());
    const int q = nbat > 0 ? sorted_ns[(size_t)((nbat - 1) * 7 / 8)] : 0;
    L.sr        = (q + 31) / 32;
  }
  const int sr = L.sr;
  L.rec.assign((size_t)std::max<long long>(nbat, 1) * R::RW, 0u);
  L.srow.assign((size_t)std::max<long long>(nbat, 1) * sr * 32, bulk_invalid);
  L.sprow.assign((size_t)std::max<long long>(nbat, 1) * sr * 32, 0);
  std::vector<std::vector<uint32_t>> ov_idx_b((size_t)nbat);
  std::vector<std::vector<uint16_t>> ov_pos_b((size_t)nbat), zpos_b((size_t)nbat);
  std::vector<long long> nblk_b((size_t)std::max<long long>(nbat, 1), 0);
  // pass 2: placement and records
  constexpr int NBK = sizeof(Number) == 8 ? 16 : 32; // lanes that share a shared-memory wavefront = banks of their width
  constexpr int NG  = 32 / NBK;
  const int max_pad = place ? runs_env("MFHN_RUNS_MAXPAD", 6) / E * E : 0, refine = place ? runs_env("MFHN_RUNS_REFINE", 2) : 0;
  struct CellData
  {
    Sorted S;
    std::vector<Block> blocks;
    std::vector<int> single, base, pad, spos;
    int slack;
  };
#pragma omp parallel
  {
    std::vector<CellData> cd((size_t)cpw);
    std::vector<int> order;
#pragma omp for schedule(dynamic, 64)
    for (long long b = 0; b < nbat; ++b)
      {
        uint32_t *rec = &L.rec[(size_t)b * R::RW];
        // lanes per bank of the staging accesses: [plane slot][lane group][bank], and the largest count per access
        // (= its wavefronts)
        unsigned char cnt[nn][NG][NBK], mx[nn][NG];
        std::memset(cnt, 0, sizeof(cnt));
        std::memset(mx, 0, sizeof(mx));
        const int ncell = (int)std::min<long long>(cpw, n_cells - b * cpw);
        auto group_of = [&](int slot, int j) { return (slot * n + j % n) / NBK; };
        auto cost_of  = [&](int slot, int j, int at) {
          const int g = group_of(slot, j), k2 = cnt[j / n][g][(slot * R::cap + at) & (NBK - 1)];
          return (k2 + 1 > mx[j / n][g] ? 32 : 0) + k2; // a new wavefront is expensive, crowded banks are a tie-break
        };
        auto add = [&](int slot, int j, int at) {
          const int g          = group_of(slot, j);
          const unsigned char k2 = ++cnt[j / n][g][(slot * R::cap + at) & (NBK - 1)];
          if (k2 > mx[j / n][g]) mx[j / n][g] = k2;
        };
        auto remove = [&](int slot, int j, int at) {
          const int g = group_of(slot, j);
          --cnt[j / n][g][(slot * R::cap + at) & (NBK - 1)];
          unsigned char m = 0;
          for (int k2 = 0; k2 < NBK; ++k2) m = std::max(m, cnt[j / n][g][k2]);
          mx[j / n][g] = m;
        };
        auto block_cost = [&](int slot, const CellData &C, int bi, int pad) {
          const Block &B = C.blocks[bi];
          int cost       = 0;
          for (int i = B.a; i <= B.b; ++i)
            cost += cost_of(slot, C.S.s[C.S.uq[i]].second, C.base[bi] + pad + (int)(C.S.s[C.S.uq[i]].first - B.first));
          return cost;
        };
        auto block_apply = [&](int slot, const CellData &C, int bi, bool on) {
          const Block &B = C.blocks[bi];
          for (int i = B.a; i <= B.b; ++i)
            {
              const int j = C.S.s[C.S.uq[i]].second, at = C.base[bi] + C.pad[bi] + (int)(C.S.s[C.S.uq[i]].first - B.first);
              on ? add(slot, j, at) : remove(slot, j, at);
            }
        };
        auto block_choose = [&](int slot, CellData &C, int bi) {
          int best_pad = 0, best_cost = 1 << 30;
          for (int pad = 0; pad <= C.slack; pad += E)
            {
              const int cost = block_cost(slot, C, bi, pad);
              if (cost < best_cost)
                {
                  best_cost = cost;
                  best_pad  = pad;
                }
              if (cost == 0) break;
            }
          C.pad[bi] = best_pad;
        };
        // runs: every copied range gets a region of its size + slack, largest first
        for (int slot = 0; slot < ncell; ++slot)
          {
            CellData &C = cd[slot];
            cell_runs(idx + (b * cpw + slot) * n3, C.S, C.blocks, C.single);
            int room = R::cap - (int)C.single.size();
            for (const Block &B : C.blocks) room -= B.size;
            C.slack = C.blocks.empty() ? 0 : std::min(max_pad, room / (int)C.blocks.size() / E * E);
            order.resize(C.blocks.size());
            for (size_t i = 0; i < C.blocks.size(); ++i) order[i] = (int)i;
            std::sort(order.begin(), order.end(), [&](int x, int y) { return C.blocks[x].count > C.blocks[y].count; });
            C.base.assign(C.blocks.size(), 0);
            C.pad.assign(C.blocks.size(), 0);
            int cursor = 0;
            for (const int bi : order)
              {
                C.base[bi] = cursor;
                cursor += C.blocks[bi].size + C.slack;
              }
            for (const int bi : order)
              {
                block_choose(slot, C, bi);
                block_apply(slot, C, bi, true);
              }
          }
        for (int round = 0; round < refine; ++round)
          for (int slot = 0; slot < ncell; ++slot)
            for (size_t bi = 0; bi < cd[slot].blocks.size(); ++bi)
              {
                block_apply(slot, cd[slot], (int)bi, false);
                block_choose(slot, cd[slot], (int)bi);
                block_apply(slot, cd[slot], (int)bi, true);
              }
        // single entries: the free staging slot that costs least
        int nblk = 0, nsgl = 0;
        for (int slot = 0; slot < ncell; ++slot)
          {
            CellData &C = cd[slot];
            bool used[R::cap], mine[R::cap];
            for (int i = 0; i < R::cap; ++i) used[i] = mine[i] = false;
            unsigned char pos[n3];
            for (size_t bi = 0; bi < C.blocks.size(); ++bi)
              {
                const Block &B  = C.blocks[bi];
                const int start = C.base[bi] + C.pad[bi];
                for (int e = 0; e < B.size; ++e) used[start + e] = true;
                for (int i = B.a; i <= B.b; ++i)
                  {
                    const int at                 = start + (int)(C.S.s[C.S.uq[i]].first - B.first);
                    pos[C.S.s[C.S.uq[i]].second] = (unsigned char)at;
                    mine[at]                     = true;
                  }
                for (int e = 0; e < B.size; ++e)
                  if (!mine[start + e]) zpos_b[b].push_back((uint16_t)(slot * R::cap + start + e));
                rec[4 + 2 * nblk]     = B.first;
                rec[4 + 2 * nblk + 1] = (uint32_t)(slot * R::cap + start) | ((uint32_t)(B.size / E) << 16);
                ++nblk;
              }
            C.spos.assign(C.single.size(), -1);
            auto single_choose = [&](size_t si) {
              const int j = C.S.s[C.single[si]].second;
              int best = -1, best_cost = 1 << 30;
              for (int at = 0; at < R::cap; ++at)
                {
                  if (used[at]) continue;
                  const int cost = place ? cost_of(slot, j, at) : 0;
                  if (cost < best_cost)
                    {
                      best_cost = cost;
                      best      = at;
                    }
                  if (cost == 0) break;
                }
              // never runs dry: entries + copied ranges fit the capacity by construction
              if (best < 0) throw std::runtime_error("runs layout: staging area exhausted");
              used[best]  = true;
              C.spos[si]  = best;
              add(slot, j, best);
            };
            for (size_t si = 0; si < C.single.size(); ++si) single_choose(si);
            for (int round = 0; round < refine; ++round)
              for (size_t si = 0; si < C.single.size(); ++si)
                {
                  remove(slot, C.S.s[C.single[si]].second, C.spos[si]);
                  used[C.spos[si]] = false;
                  single_choose(si);
                }
            for (size_t si = 0; si < C.single.size(); ++si)
              {
                const int i       = C.single[si];
                pos[C.S.s[i].second] = (unsigned char)C.spos[si];
                const uint16_t sp = (uint16_t)(slot * R::cap + C.spos[si]);
                if (nsgl < sr * 32)
                  {
                    L.srow[(size_t)b * sr * 32 + nsgl]  = C.S.s[i].first;
                    L.sprow[(size_t)b * sr * 32 + nsgl] = sp;
                  }
                else
                  {
                    ov_idx_b[b].push_back(C.S.s[i].first);
                    ov_pos_b[b].push_back(sp);
                  }
                ++nsgl;
              }
            // kernel axes: thread t = x, plane slot j = y + n z
            uint32_t *tab = rec + 4 + 64 * R::BR;
            for (int t = 0; t < n; ++t)
              for (int j = 0; j < nn; ++j) tab[(j / 4) * 32 + slot * n + t] |= (uint32_t)pos[t + n * j] << (8 * (j % 4));
          }
        uint32_t *tab = rec + 4 + 64 * R::BR;
        for (int lane = Cfg::lanes; lane < 32; ++lane)
          for (int q = 0; q < R::NT; ++q) tab[q * 32 + lane] = tab[q * 32 + lane - R::idle];
        nblk_b[b] = nblk;
      }
  }
  // variable-length parts
  L.ov_idx.clear();
  L.ov_pos.clear();
  L.zpos.clear();
  L.n_blocks = L.n_singles = L.n_zero = 0;
  for (long long b = 0; b < nbat; ++b)
    {
      if (ov_idx_b[b].size() > 0xffff || zpos_b[b].size() > 0xffff || L.ov_idx.size() > 0xffffffffull || L.zpos.size() > 0xffffffffull)
        throw std::runtime_error("runs layout: descriptor counts out of range");
      uint32_t *rec = &L.rec[(size_t)b * R::RW];
      rec[0]        = (uint32_t)L.ov_idx.size();
      rec[1]        = (uint32_t)L.zpos.size();
      rec[2]        = (uint32_t)zpos_b[b].size() | ((uint32_t)ov_idx_b[b].size() << 16);
      L.ov_idx.insert(L.ov_idx.end(), ov_idx_b[b].begin(), ov_idx_b[b].end());
      L.ov_pos.insert(L.ov_pos.end(), ov_pos_b[b].begin(), ov_pos_b[b].end());
      L.zpos.insert(L.zpos.end(), zpos_b[b].begin(), zpos_b[b].end());
      L.n_blocks += nblk_b[b];
      L.n_singles += ns_batch[b];
      L.n_zero += (long long)zpos_b[b].size();
    }
}

// Emulation of the kernel's gather through the layout with src[i] = i: returns the number of cell entries that do
// not come out as the reference index array says, plus the number of copied entries outside the cell that the
// zero list misses (0 = the layout reproduces the index array and scatters nothing but the cell's entries).
// wavefronts (optional): shared-memory wavefronts of the staging reads of all batches (2 per plane slot = no conflict)
template <int n, typename Number>
long long runs_verify_impl(const RunsHostLayout &L, const uint32_t *idx, long long *wavefronts)
{
  using Cfg = PlaneCfg<n, Number>;
  using R   = RunsCfg<n, Number>;
  constexpr int n3 = R::n3, E = R::E, cpw = Cfg::cpw, nn = n * n;
  long long bad = 0, wf = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad, wf)
  for (long long b = 0; b < L.n_batches; ++b)
    {
      std::vector<long long> S((size_t)cpw * R::cap, -1);
      std::vector<char> zero((size_t)cpw * R::cap, 0), mine((size_t)cpw * R::cap, 0);
      const uint32_t *rec = &L.rec[(size_t)b * R::RW];
      const int nz = (int)(rec[2] & 0xffff), nov = (int)(rec[2] >> 16);
      for (int i = 0; i < 32 * R::BR; ++i)
        {
          const uint32_t first = rec[4 + 2 * i], y = rec[4 + 2 * i + 1];
          if ((y >> 16) == 0) continue;
          if (first % E || (y & 0xffff) % E) ++bad;
          for (uint32_t e = 0; e < (y >> 16) * E; ++e)
            {
              if ((y & 0xffff) + e >= S.size() || S[(y & 0xffff) + e] != -1) ++bad; // outside / overlapping staging ranges
              else S[(y & 0xffff) + e] = (long long)first + e;
            }
        }
      auto single = [&](uint32_t g, uint16_t sp) {
        if (sp >= S.size() || S[sp] != -1) ++bad;
        else
          {
            S[sp]    = g;
            mine[sp] = 1;
          }
      };
      for (int i = 0; i < L.sr * 32; ++i)
        if (L.srow[(size_t)b * L.sr * 32 + i] != bulk_invalid) single(L.srow[(size_t)b * L.sr * 32 + i], L.sprow[(size_t)b * L.sr * 32 + i]);
      for (int i = 0; i < nov; ++i) single(L.ov_idx[(size_t)rec[0] + i], L.ov_pos[(size_t)rec[0] + i]);
      for (int i = 0; i < nz; ++i) zero[L.zpos[(size_t)rec[1] + i]] = 1;
      const uint32_t *tab = rec + 4 + 64 * R::BR;
      for (int slot = 0; slot < cpw; ++slot)
        {
          const long long c = b * cpw + slot;
          if (c >= L.n_cells) continue;
          for (int t = 0; t < n; ++t)
            for (int j = 0; j < nn; ++j)
              {
                const int pos = (int)((tab[(j / 4) * 32 + slot * n + t] >> (8 * (j % 4))) & 0xff);
                if (pos >= R::cap || S[(size_t)slot * R::cap + pos] != (long long)idx[c * n3 + t + n * j]) ++bad;
                if (mine[(size_t)slot * R::cap + pos] > 1 || zero[(size_t)slot * R::cap + pos]) ++bad;
                mine[(size_t)slot * R::cap + pos] = 2; // written by exactly one plane slot
              }
        }
      // every copied entry is the cell's (written once) or zeroed; every single entry is written
      for (size_t i = 0; i < S.size(); ++i)
        if (S[i] != -1 && ((mine[i] == 2) == (zero[i] == 1))) ++bad;
      // bank model of the staging reads (lanes of existing cells): the lanes that share a wavefront (16 for 8-byte,
      // 32 for 4-byte entries) hit as many banks of their width
      constexpr int NBK = sizeof(Number) == 8 ? 16 : 32;
      for (int j = 0; j < nn; ++j)
        for (int h = 0; h < 32 / NBK; ++h)
          {
            int cnt[NBK] = {};
            int addr[NBK][NBK];
            for (int lane = NBK * h; lane < NBK * h + NBK; ++lane)
              {
                const int ml = lane < Cfg::lanes ? lane : lane - R::idle, slot = ml / n;
                if (b * cpw + slot >= L.n_cells) continue;
                const int at = slot * R::cap + (int)((tab[(j / 4) * 32 + lane] >> (8 * (j % 4))) & 0xff);
                const int bk = at & (NBK - 1);
                bool seen    = false;
                for (int q = 0; q < cnt[bk]; ++q) seen |= addr[bk][q] == at;
                if (!seen) addr[bk][cnt[bk]++] = at;
              }
            int m = 0;
            for (int k2 = 0; k2 < NBK; ++k2) m = std::max(m, cnt[k2]);
            wf += m;
          }
    }
  if (wavefronts) *wavefronts = wf;
  return bad;
}
} // namespace mfhn
