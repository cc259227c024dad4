// C ABI of the mesh and DoF layers (include/mfhn.h).
#include "../../include/mfhn.h"
#include "dof_handler.hpp"
#include "error.hpp"
#include "matrix_free.hpp"
#include "octree.hpp"

#include <memory>
#include <string>
#include <vector>

namespace mfhn
{
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }

struct MeshHandle
{
  Octree tree;
  std::vector<uint16_t> kinds; // lazily computed
  explicit MeshHandle(Octree &&t)
    : tree(std::move(t))
  {}
  const std::vector<uint16_t> &get_kinds()
  {
    if (kinds.empty() && !tree.cells().empty())
      {
        kinds.resize(tree.cells().size());
        const int64_t nc = (int64_t)kinds.size();
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < nc; ++s) kinds[s] = tree.constraint_kind(tree.cells()[s]);
      }
    return kinds;
  }
};
struct DofsHandle
{
  MeshHandle *mesh;
  DoFHandler dh;
  DofsHandle(MeshHandle *m, int degree, int n_ranks, const int32_t *rank)
    : mesh(m)
    , dh(m->tree, degree, n_ranks, rank)
  {}
};
} // namespace mfhn

using namespace mfhn;

extern "C" {
const char *mfhn_last_error(void) { return g_last_error.c_str(); }
const char *mfhn_version(void) { return "mfhn 0.1 (sm_100a)"; }

int mfhn_mesh_create(const char *geometry, int n_refinements, int flavour, mfhn_mesh *out)
{
  return guard([&] {
    if (!geometry || !out) throw InvalidArgument("null argument");
    *out = reinterpret_cast<mfhn_mesh>(new MeshHandle(Octree::create(geometry, n_refinements, flavour)));
  });
}
void mfhn_mesh_destroy(mfhn_mesh m) { delete reinterpret_cast<MeshHandle *>(m); }
int64_t mfhn_mesh_n_cells(mfhn_mesh m) { return m ? (int64_t) reinterpret_cast<MeshHandle *>(m)->tree.cells().size() : -1; }
int mfhn_mesh_n_levels(mfhn_mesh m) { return m ? reinterpret_cast<MeshHandle *>(m)->tree.n_levels() : -1; }
int mfhn_mesh_cells(mfhn_mesh m, int32_t *out)
{
  return guard([&] {
    if (!m || !out) throw InvalidArgument("null argument");
    const Octree &t = reinterpret_cast<MeshHandle *>(m)->tree;
    int64_t s       = 0;
    for (int32_t id : t.cells())
      {
        const Node &nd = t.nodes()[id];
        out[4 * s + 0] = nd.level;
        out[4 * s + 1] = nd.c[0];
        out[4 * s + 2] = nd.c[1];
        out[4 * s + 3] = nd.c[2];
        ++s;
      }
  });
}
int64_t mfhn_mesh_n_cells_hn(mfhn_mesh m)
{
  if (!m) return -1;
  int64_t n = 0;
  for (uint16_t k : reinterpret_cast<MeshHandle *>(m)->get_kinds()) n += k != 0;
  return n;
}
int mfhn_mesh_morton_position(mfhn_mesh m, int64_t *pos)
{
  return guard([&] {
    if (!m || !pos) throw InvalidArgument("null argument");
    const auto order = reinterpret_cast<MeshHandle *>(m)->tree.morton_order();
    for (size_t p = 0; p < order.size(); ++p) pos[order[p]] = (int64_t)p;
  });
}
int mfhn_mesh_partition(mfhn_mesh m, int n_ranks, double hn_weight, int32_t *rank_of_cell)
{
  return guard([&] {
    if (!m || !rank_of_cell) throw InvalidArgument("null argument");
    if (n_ranks < 1) throw InvalidArgument("n_ranks must be >= 1");
    MeshHandle &mh    = *reinterpret_cast<MeshHandle *>(m);
    const auto order  = mh.tree.morton_order();
    const auto &kinds = mh.get_kinds();
    // weights of benchmark_02.cc:24-33: 1 + 10 w for cells with hanging nodes, 1 + 10 otherwise
    std::vector<double> w(order.size());
    double total = 0;
    for (size_t p = 0; p < order.size(); ++p)
      {
        // (the reference's weight callback returns unsigned int: 1 + 10 w is truncated, benchmark_02.cc:19-33)
        w[p] = kinds[order[p]] != 0 ? (double)(unsigned)(1.0 + 10.0 * hn_weight) : 11.0;
        total += w[p];
      }
    double prefix = 0;
    for (size_t p = 0; p < order.size(); ++p)
      {
        // a cell belongs to the rank in whose weight interval its midpoint falls
        int r = (int)((prefix + 0.5 * w[p]) * n_ranks / total);
        if (r >= n_ranks) r = n_ranks - 1;
        rank_of_cell[order[p]] = r;
        prefix += w[p];
      }
  });
}

int mfhn_dofs_create(mfhn_mesh m, int degree, int n_ranks, const int32_t *rank_of_cell, mfhn_dofs *out)
{
  return guard([&] {
    if (!m || !out) throw InvalidArgument("null argument");
    if (degree < 1 || degree > 8) throw InvalidArgument("degree must be in 1..8");
    *out = reinterpret_cast<mfhn_dofs>(new DofsHandle(reinterpret_cast<MeshHandle *>(m), degree, n_ranks, rank_of_cell));
  });
}
void mfhn_dofs_destroy(mfhn_dofs d) { delete reinterpret_cast<DofsHandle *>(d); }
int64_t mfhn_dofs_n_dofs(mfhn_dofs d) { return d ? reinterpret_cast<DofsHandle *>(d)->dh.n_dofs() : -1; }
int mfhn_dofs_owned_range(mfhn_dofs d, int rank, int64_t *begin, int64_t *end)
{
  return guard([&] {
    if (!d || !begin || !end) throw InvalidArgument("null argument");
    const DoFHandler &dh = reinterpret_cast<DofsHandle *>(d)->dh;
    if (rank < 0 || rank >= dh.n_ranks()) throw InvalidArgument("rank out of range");
    dh.owned_range(rank, *begin, *end);
  });
}
int64_t mfhn_dofs_n_cells_of_rank(mfhn_dofs d, int rank)
{
  if (!d) return -1;
  const DoFHandler &dh = reinterpret_cast<DofsHandle *>(d)->dh;
  if (rank < 0 || rank >= dh.n_ranks()) return -1;
  return (int64_t)dh.cells_of_rank(rank).size();
}
int mfhn_dofs_cells_of_rank(mfhn_dofs d, int rank, int64_t *cell_ids)
{
  return guard([&] {
    if (!d || !cell_ids) throw InvalidArgument("null argument");
    const DoFHandler &dh = reinterpret_cast<DofsHandle *>(d)->dh;
    if (rank < 0 || rank >= dh.n_ranks()) throw InvalidArgument("rank out of range");
    const auto &c = dh.cells_of_rank(rank);
    for (size_t i = 0; i < c.size(); ++i) cell_ids[i] = c[i];
  });
}
int mfhn_dofs_fill(mfhn_dofs d, int64_t n, const int64_t *cell_ids, uint64_t *raw, uint64_t *sub, uint8_t *masks, double *h)
{
  return guard([&] {
    if (!d || (n > 0 && !cell_ids)) throw InvalidArgument("null argument");
    DofsHandle &h_       = *reinterpret_cast<DofsHandle *>(d);
    const DoFHandler &dh = h_.dh;
    const int64_t nc     = (int64_t)h_.mesh->tree.cells().size();
    const int np         = dh.degree() + 1;
    const int64_t n3     = (int64_t)np * np * np;
    for (int64_t i = 0; i < n; ++i)
      if (cell_ids[i] < 0 || cell_ids[i] >= nc) throw InvalidArgument("cell id out of range");
    bool failed = false;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i)
      {
        try
          {
            const int64_t c = cell_ids[i];
            const uint16_t k = dh.kind(c);
            if (raw) dh.raw_indices(c, raw + i * n3);
            if (sub) dh.substituted_indices(c, k, sub + i * n3);
            if (masks) masks[i] = compress_kind(k);
            if (h) h[i] = dh.h(c);
          }
        catch (...)
          {
            failed = true;
          }
      }
    if (failed) throw std::logic_error("DoF setup failed (mesh not balanced?)");
  });
}
int mfhn_dofs_support_points(mfhn_dofs d, int64_t begin, int64_t end, double *xyz)
{
  return guard([&] {
    if (!d || !xyz) throw InvalidArgument("null argument");
    const DoFHandler &dh = reinterpret_cast<DofsHandle *>(d)->dh;
    if (begin < 0 || end > dh.n_dofs() || begin > end) throw InvalidArgument("range out of bounds");
    dh.support_points(begin, end, xyz);
  });
}

// ---- MatrixFree::reinit ---------------------------------------------------------------------
int mfhn_mf_create(mfhn_dofs d, const mfhn_mf_options *options, mfhn_mf *out)
{
  return guard([&] {
    if (!d || !out) throw InvalidArgument("null argument");
    DofsHandle &dh = *reinterpret_cast<DofsHandle *>(d);
    MatrixFreeOptions opt;
    if (options)
      {
        opt.rank       = options->rank;
        opt.categorize = options->categorize;
        if (options->window > 0) opt.window = options->window;
        if (options->batch_alignment > 0) opt.batch_alignment = options->batch_alignment;
      }
    if (opt.categorize < 0 || opt.categorize > 2) throw InvalidArgument("categorize must be 0, 1 or 2");
    std::unique_ptr<MatrixFreeData> mf(new MatrixFreeData);
    mf->reinit(dh.dh, dh.mesh->tree, opt);
    *out = reinterpret_cast<mfhn_mf>(mf.release());
  });
}
void mfhn_mf_destroy(mfhn_mf m) { delete reinterpret_cast<MatrixFreeData *>(m); }
int mfhn_mf_info(mfhn_mf m, mfhn_mf_sizes *out)
{
  return guard([&] {
    if (!m || !out) throw InvalidArgument("null argument");
    const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(m);
    out->degree         = mf.degree;
    out->rank           = mf.rank;
    out->n_ranks        = mf.n_ranks;
    out->n_cells        = mf.n_cells;
    out->n_cells_hn     = mf.n_cells_hn();
    out->n_owned        = mf.n_owned;
    out->n_ghost        = mf.n_ghost;
    out->owned_begin    = mf.owned_begin;
    out->n_interior_a   = mf.n_interior_a;
    out->n_interior     = mf.n_interior;
    out->n_ghost_peers  = (int)mf.ghost_peers.size();
    out->n_import_peers = (int)mf.import_peers.size();
    out->n_import       = (int64_t)mf.import_indices.size();
  });
}
int mfhn_mf_arrays(mfhn_mf m, const int64_t **cell_ids, const uint32_t **dof_indices, const uint8_t **masks, const double **h,
                   const int64_t **ghost_global, const int32_t **ghost_owner)
{
  return guard([&] {
    if (!m) throw InvalidArgument("null argument");
    const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(m);
    if (cell_ids) *cell_ids = mf.cell_ids.data();
    if (dof_indices) *dof_indices = mf.dof_indices.data();
    if (masks) *masks = mf.masks.data();
    if (h) *h = mf.h.data();
    if (ghost_global) *ghost_global = mf.ghost_global.data();
    if (ghost_owner) *ghost_owner = mf.ghost_owner.data();
  });
}
int mfhn_mf_partitioner(mfhn_mf m, const int32_t **ghost_peers, const int64_t **ghost_begin, const int64_t **ghost_end,
                        const int32_t **import_peers, const int64_t **import_offsets, const int32_t **import_indices, const int64_t **rank_begin)
{
  return guard([&] {
    if (!m) throw InvalidArgument("null argument");
    const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(m);
    if (ghost_peers) *ghost_peers = mf.ghost_peers.data();
    if (ghost_begin) *ghost_begin = mf.ghost_begin.data();
    if (ghost_end) *ghost_end = mf.ghost_end.data();
    if (import_peers) *import_peers = mf.import_peers.data();
    if (import_offsets) *import_offsets = mf.import_offsets.data();
    if (import_indices) *import_indices = mf.import_indices.data();
    if (rank_begin) *rank_begin = mf.rank_begin.data();
  });
}
int mfhn_mf_set_imports(mfhn_mf m, int peer, const int64_t *global_indices, int64_t n)
{
  return guard([&] {
    if (!m || (n > 0 && !global_indices)) throw InvalidArgument("null argument");
    reinterpret_cast<MatrixFreeData *>(m)->set_imports(peer, global_indices, n);
  });
}
int mfhn_mf_exchange_local(mfhn_mf *all, int n_ranks)
{
  return guard([&] {
    if (!all) throw InvalidArgument("null argument");
    for (int r = 0; r < n_ranks; ++r)
      {
        if (!all[r]) throw InvalidArgument("null argument");
        const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(all[r]);
        if (mf.rank != r || mf.n_ranks != n_ranks) throw InvalidArgument("handles must be ordered by rank");
      }
    // what rank r ghosts from owner o is what o imports for r
    for (int r = 0; r < n_ranks; ++r)
      {
        const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(all[r]);
        for (size_t p = 0; p < mf.ghost_peers.size(); ++p)
          reinterpret_cast<MatrixFreeData *>(all[mf.ghost_peers[p]])
            ->set_imports(r, mf.ghost_global.data() + mf.ghost_begin[p], mf.ghost_end[p] - mf.ghost_begin[p]);
      }
  });
}

uint8_t mfhn_compress(uint16_t kind) { return compress_kind(kind); }
uint16_t mfhn_decompress(uint8_t c) { return decompress_kind(c); }
int mfhn_check_kind(uint16_t kind) { return check_kind(kind) ? 1 : 0; }
}
