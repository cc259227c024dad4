#include "matrix_free.hpp"

#include "octree.hpp"

#include <omp.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <stdexcept>

namespace mfhn
{
namespace
{
constexpr uint32_t ghost_placeholder = 0xffffffffu;
}

void MatrixFreeData::reinit(const DoFHandler &dh, const Octree &tree, const MatrixFreeOptions &opt)
{
  if (opt.rank < 0 || opt.rank >= dh.n_ranks()) throw std::invalid_argument("rank out of range");
  if (opt.window < 1 || opt.batch_alignment < 1) throw std::invalid_argument("window and batch_alignment must be positive");
  degree  = dh.degree();
  rank    = opt.rank;
  n_ranks = dh.n_ranks();
  dh.owned_range(rank, owned_begin, owned_end);
  n_owned = owned_end - owned_begin;
  rank_begin.resize(n_ranks + 1);
  for (int r = 0; r < n_ranks; ++r)
    {
      int64_t b, e;
      dh.owned_range(r, b, e);
      rank_begin[r]     = b;
      rank_begin[r + 1] = e;
    }
  const int np     = degree + 1;
  const int64_t n3 = (int64_t)np * np * np;

  // 1. the rank's cells along the Morton curve (MatrixFree is free to order its cell batches)
  std::vector<int64_t> cells = dh.cells_of_rank(rank);
  n_cells                    = (int64_t)cells.size();
  {
    const std::vector<int64_t> order = tree.morton_order();
    std::vector<int64_t> pos(order.size());
    for (size_t p = 0; p < order.size(); ++p) pos[order[p]] = (int64_t)p;
    std::stable_sort(cells.begin(), cells.end(), [&](int64_t a, int64_t b) { return pos[a] < pos[b]; });
  }
  std::vector<uint8_t> m0(n_cells);
  std::vector<uint8_t> touches(n_cells, 0);
  bool failed = false;
#pragma omp parallel
  {
    std::vector<uint64_t> buf(n3);
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_cells; ++i)
      {
        try
          {
            const uint16_t kind = dh.kind(cells[i]);
            m0[i]               = compress_kind(kind);
            if (n_ranks > 1)
              {
                // cells that touch ghost entries need the ghost import: they form the last partition
                dh.substituted_indices(cells[i], kind, buf.data());
                bool t = false;
                for (int64_t j = 0; j < n3; ++j) t |= (int64_t)buf[j] < owned_begin || (int64_t)buf[j] >= owned_end;
                touches[i] = t;
              }
          }
        catch (...)
          {
            failed = true;
          }
      }
  }
  if (failed) throw std::logic_error("DoF setup failed (mesh not balanced?)");

  // 2. [interior | boundary], partition boundaries on whole warp batches of the cell kernels
  std::vector<int64_t> order;
  order.reserve(n_cells);
  for (int64_t i = 0; i < n_cells; ++i)
    if (!touches[i]) order.push_back(i);
  n_interior = (int64_t)order.size();
  for (int64_t i = 0; i < n_cells; ++i)
    if (touches[i]) order.push_back(i);
  if (n_ranks > 1) n_interior -= n_interior % opt.batch_alignment; // the last few interior cells join the boundary partition

  // 3. categorisation (cell_vectorization_category = constraint mask, benchmark_01.h:258-284) inside windows of the
  //    Morton order: fewer warps pay for the interpolation and a warp holds few different constraint kinds
  if (opt.categorize)
    {
      std::vector<int64_t> key(n_cells), perm(n_cells);
      for (int64_t p = 0; p < n_cells; ++p)
        {
          const int64_t seg    = p >= n_interior;
          const int64_t window = (seg ? p - n_interior : p) / opt.window;
          const int64_t kind   = opt.categorize == 1 ? m0[order[p]] : (m0[order[p]] != 0);
          key[p]               = (seg << 62) | (window << 16) | kind; // sort by (segment, window, kind), stable in the position
          perm[p]              = p;
        }
      std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
      std::vector<int64_t> sorted(n_cells);
      for (int64_t p = 0; p < n_cells; ++p) sorted[p] = order[perm[p]];
      order.swap(sorted);
    }
  // deal.II's cell_loop overlaps the two ghost exchanges with two interior partitions
  n_interior_a = n_ranks > 1 ? (n_interior / 2) / opt.batch_alignment * opt.batch_alignment : n_interior;

  // 4. rank-local numbering: owned entries [0, n_owned), ghosts behind them sorted by global index
  cell_ids.resize(n_cells);
  masks.resize(n_cells);
  h.resize(n_cells);
  dof_indices.assign((size_t)(n_cells * n3), 0u);
  struct GhostRef
  {
    int64_t pos, global;
  };
  std::vector<std::vector<GhostRef>> refs(omp_get_max_threads());
#pragma omp parallel
  {
    std::vector<uint64_t> buf(n3);
    std::vector<GhostRef> &mine = refs[omp_get_thread_num()];
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_cells; ++i)
      {
        try
          {
            const int64_t c     = cells[order[i]];
            const uint16_t kind = dh.kind(c);
            cell_ids[i]         = c;
            masks[i]            = compress_kind(kind);
            h[i]                = dh.h(c);
            dh.substituted_indices(c, kind, buf.data());
            uint32_t *out = dof_indices.data() + i * n3;
            for (int64_t j = 0; j < n3; ++j)
              {
                const int64_t g = (int64_t)buf[j];
                if (g >= owned_begin && g < owned_end)
                  out[j] = (uint32_t)(g - owned_begin);
                else
                  {
                    out[j] = ghost_placeholder;
                    mine.push_back(GhostRef{i * n3 + j, g});
                  }
              }
          }
        catch (...)
          {
            failed = true;
          }
      }
  }
  if (failed) throw std::logic_error("DoF setup failed (mesh not balanced?)");
  ghost_global.clear();
  for (const auto &v : refs)
    for (const GhostRef &r : v) ghost_global.push_back(r.global);
  std::sort(ghost_global.begin(), ghost_global.end());
  ghost_global.erase(std::unique(ghost_global.begin(), ghost_global.end()), ghost_global.end());
  n_ghost = (int64_t)ghost_global.size();
  if (n_owned + n_ghost >= (int64_t)ghost_placeholder) throw std::invalid_argument("more than 2^32 - 1 local vector entries");
  for (const auto &v : refs)
    for (const GhostRef &r : v)
      dof_indices[r.pos] = (uint32_t)(n_owned + (std::lower_bound(ghost_global.begin(), ghost_global.end(), r.global) - ghost_global.begin()));

  // 5. partitioner: ghosts are sorted by global index and the owners' ranges ascend => one contiguous range per owner
  ghost_owner.resize(n_ghost);
  ghost_peers.clear();
  ghost_begin.clear();
  ghost_end.clear();
  for (int64_t g = 0; g < n_ghost; ++g)
    {
      const int o    = (int)(std::upper_bound(rank_begin.begin(), rank_begin.end(), ghost_global[g]) - rank_begin.begin()) - 1;
      ghost_owner[g] = o;
      if (ghost_peers.empty() || ghost_peers.back() != o)
        {
          ghost_peers.push_back(o);
          ghost_begin.push_back(g);
          ghost_end.push_back(g);
        }
      ghost_end.back() = g + 1;
    }
  import_peers.clear();
  import_offsets.assign(1, 0);
  import_indices.clear();
}

void MatrixFreeData::set_imports(int peer, const int64_t *global_indices, int64_t n)
{
  if (peer < 0 || peer >= n_ranks || peer == rank) throw std::invalid_argument("bad import peer");
  for (int64_t i = 0; i < n; ++i)
    if (global_indices[i] < owned_begin || global_indices[i] >= owned_end) throw std::invalid_argument("import index is not owned by this rank");
  // keep the lists grouped by ascending peer
  std::map<int, std::vector<int32_t>> lists;
  for (size_t p = 0; p < import_peers.size(); ++p)
    lists[import_peers[p]].assign(import_indices.begin() + import_offsets[p], import_indices.begin() + import_offsets[p + 1]);
  std::vector<int32_t> &mine = lists[peer];
  mine.resize(n);
  for (int64_t i = 0; i < n; ++i) mine[i] = (int32_t)(global_indices[i] - owned_begin);
  if (n == 0) lists.erase(peer);
  import_peers.clear();
  import_offsets.assign(1, 0);
  import_indices.clear();
  for (const auto &kv : lists)
    {
      import_peers.push_back(kv.first);
      import_indices.insert(import_indices.end(), kv.second.begin(), kv.second.end());
      import_offsets.push_back((int64_t)import_indices.size());
    }
}

int64_t MatrixFreeData::n_cells_hn() const
{
  int64_t n = 0;
  for (uint8_t m : masks) n += m != 0;
  return n;
}
} // namespace mfhn
