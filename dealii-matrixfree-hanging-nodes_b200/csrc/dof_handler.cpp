#include "dof_handler.hpp"

#include "fe1d.hpp"

#include <stdexcept>

namespace mfhn
{
namespace
{
// deal.II GeometryInfo<3>: vertex v = vx + 2 vy + 4 vz; lines 0-3 on z=0
// {x=0 along y, x=1 along y, y=0 along x, y=1 along x}, 4-7 the same on z=1,
// 8-11 along z at (x,y) = (0,0),(1,0),(0,1),(1,1)  (cf. constraint_helper.h:21-32);
// quads x-,x+,y-,y+,z-,z+ with face-local axes (y,z), (z,x), (x,y).
const int line_dir[12]     = {1, 1, 0, 0, 1, 1, 0, 0, 2, 2, 2, 2};
const int line_side[12][3] = {{0, 0, 0}, {1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1},
                              {0, 0, 1}, {0, 1, 1}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}};
const int quad_fast[3]     = {1, 2, 0};
const int quad_slow[3]     = {2, 0, 1};

int line_id(int d, const int side[3])
{
  for (int l = 0; l < 12; ++l)
    {
      if (line_dir[l] != d) continue;
      bool ok = true;
      for (int t = 0; t < 3; ++t)
        if (t != d && line_side[l][t] != side[t]) ok = false;
      if (ok) return l;
    }
  return -1;
}
inline int popc(uint32_t x) { return __builtin_popcount(x); }
} // namespace

DoFHandler::DoFHandler(const Octree &tree, int degree, int n_ranks, const int32_t *rank_of_cell)
  : tree_(tree)
  , k_(degree)
  , n_ranks_(n_ranks < 1 ? 1 : n_ranks)
{
  if (degree < 1) throw std::invalid_argument("degree must be >= 1");
  const int64_t nc = (int64_t)tree.cells().size();
  walk_pos_.resize(nc);
  cells_of_rank_.assign(n_ranks_, {});
  if (n_ranks_ == 1 || rank_of_cell == nullptr)
    {
      n_ranks_ = 1;
      cells_of_rank_.assign(1, {});
      cells_of_rank_[0].resize(nc);
      for (int64_t s = 0; s < nc; ++s) walk_pos_[s] = cells_of_rank_[0][s] = s;
    }
  else
    {
      for (int64_t s = 0; s < nc; ++s)
        {
          if (rank_of_cell[s] < 0 || rank_of_cell[s] >= n_ranks_) throw std::invalid_argument("rank_of_cell out of range");
          cells_of_rank_[rank_of_cell[s]].push_back(s);
        }
      int64_t pos = 0;
      for (int r = 0; r < n_ranks_; ++r)
        for (int64_t s : cells_of_rank_[r]) walk_pos_[s] = pos++;
    }
  // which objects does each cell see first?
  own_mask_.assign(nc, 0);
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < nc; ++s)
    {
      uint32_t m = 0;
      for (int obj = 0; obj < 26; ++obj)
        if (owner_of(s, obj).cell == s) m |= 1u << obj;
      own_mask_[s] = m;
    }
  // prefix sum of the per-cell new-index counts in walk order
  const int64_t km1 = k_ - 1;
  base_.assign(nc, 0);
  rank_begin_.assign(n_ranks_ + 1, 0);
  int64_t next = 0;
  for (int r = 0; r < n_ranks_; ++r)
    {
      rank_begin_[r] = next;
      for (int64_t s : cells_of_rank_[r])
        {
          base_[s]         = next;
          const uint32_t m = own_mask_[s];
          next += popc(m & 0xffu) + km1 * popc((m >> 8) & 0xfffu) + km1 * km1 * popc((m >> 20) & 0x3fu) + km1 * km1 * km1;
        }
    }
  rank_begin_[n_ranks_] = next;
  n_dofs_               = next;
}

DoFHandler::Owner DoFHandler::owner_of(int64_t cell, int obj) const
{
  const auto &nodes = tree_.nodes();
  const Node &nd    = nodes[tree_.cells()[cell]];
  const int l       = nd.level;
  Owner best{cell, obj};
  int64_t best_pos = walk_pos_[cell];
  auto consider    = [&](int32_t nb, int o) {
    const int64_t c = tree_.cell_index(nb);
    if (walk_pos_[c] < best_pos)
      {
        best_pos = walk_pos_[c];
        best     = Owner{c, o};
      }
  };
  if (obj < 8)
    {
      const int b[3] = {obj & 1, (obj >> 1) & 1, (obj >> 2) & 1};
      const int P[3] = {nd.c[0] + b[0], nd.c[1] + b[1], nd.c[2] + b[2]};
      for (int oct = 0; oct < 8; ++oct)
        {
          const int o[3] = {-(oct & 1), -((oct >> 1) & 1), -((oct >> 2) & 1)};
          const int r[3] = {P[0] + o[0], P[1] + o[1], P[2] + o[2]};
          if (r[0] == nd.c[0] && r[1] == nd.c[1] && r[2] == nd.c[2]) continue;
          if (!tree_.inside(l, r[0], r[1], r[2])) continue;
          int32_t nb = tree_.find(l, r[0], r[1], r[2]);
          int vb[3];
          if (nodes[nb].first_child < 0)
            {
              const int sh = l - nodes[nb].level;
              const int mk = (1 << sh) - 1;
              if ((P[0] & mk) || (P[1] & mk) || (P[2] & mk)) continue; // P is not a corner of the coarser leaf
              for (int d = 0; d < 3; ++d) vb[d] = (P[d] >> sh) - nodes[nb].c[d];
            }
          else
            {
              for (int d = 0; d < 3; ++d) vb[d] = (o[d] == -1) ? 1 : 0;
              while (nodes[nb].first_child >= 0) nb = nodes[nb].first_child + vb[0] + 2 * vb[1] + 4 * vb[2];
            }
          consider(nb, vb[0] + 2 * vb[1] + 4 * vb[2]);
        }
    }
  else if (obj < 20)
    {
      const int ln = obj - 8, d = line_dir[ln], a = (d + 1) % 3, b = (d + 2) % 3;
      for (int q = 0; q < 4; ++q)
        {
          const int oa = -(q & 1), ob = -((q >> 1) & 1);
          int r[3];
          r[d] = nd.c[d];
          r[a] = nd.c[a] + line_side[ln][a] + oa;
          r[b] = nd.c[b] + line_side[ln][b] + ob;
          if (r[a] == nd.c[a] && r[b] == nd.c[b]) continue;
          if (!tree_.inside(l, r[0], r[1], r[2])) continue;
          const int32_t nb = tree_.find(l, r[0], r[1], r[2]);
          if (nodes[nb].level != l || nodes[nb].first_child >= 0) continue;
          int side[3] = {0, 0, 0};
          side[a]     = (oa == -1) ? 1 : 0;
          side[b]     = (ob == -1) ? 1 : 0;
          consider(nb, 8 + line_id(d, side));
        }
    }
  else
    {
      const int q = obj - 20, ndir = q >> 1, s = q & 1;
      int r[3] = {nd.c[0], nd.c[1], nd.c[2]};
      r[ndir] += s ? 1 : -1;
      if (tree_.inside(l, r[0], r[1], r[2]))
        {
          const int32_t nb = tree_.find(l, r[0], r[1], r[2]);
          if (nodes[nb].level == l && nodes[nb].first_child < 0) consider(nb, 20 + 2 * ndir + (1 - s));
        }
    }
  return best;
}

int64_t DoFHandler::object_base(int64_t cell, int obj) const
{
  const uint32_t m  = own_mask_[cell];
  const int64_t km1 = k_ - 1;
  const int64_t nv = popc(m & 0xffu), nl = popc((m >> 8) & 0xfffu);
  if (obj < 8) return base_[cell] + popc(m & ((1u << obj) - 1u));
  if (obj < 20) return base_[cell] + nv + km1 * popc((m >> 8) & ((1u << (obj - 8)) - 1u));
  return base_[cell] + nv + km1 * nl + km1 * km1 * popc((m >> 20) & ((1u << (obj - 20)) - 1u));
}

void DoFHandler::raw_indices(int64_t cell, uint64_t *out) const
{
  const int k = k_, n = k + 1, km1 = k - 1;
  auto lex = [n](int ax, int ay, int az) { return ax + n * (ay + n * az); };
  for (int v = 0; v < 8; ++v)
    {
      const Owner ow = owner_of(cell, v);
      out[lex((v & 1) * k, ((v >> 1) & 1) * k, ((v >> 2) & 1) * k)] = (uint64_t)object_base(ow.cell, ow.obj);
    }
  if (km1 == 0) return;
  for (int ln = 0; ln < 12; ++ln)
    {
      const Owner ow  = owner_of(cell, 8 + ln);
      const int64_t g = object_base(ow.cell, ow.obj);
      const int d     = line_dir[ln];
      int a[3]        = {line_side[ln][0] * k, line_side[ln][1] * k, line_side[ln][2] * k};
      for (int m = 0; m < km1; ++m)
        {
          a[d]                     = m + 1;
          out[lex(a[0], a[1], a[2])] = (uint64_t)(g + m);
        }
    }
  for (int q = 0; q < 6; ++q)
    {
      const Owner ow  = owner_of(cell, 20 + q);
      const int64_t g = object_base(ow.cell, ow.obj);
      const int ndir = q >> 1, s = q & 1, fast = quad_fast[ndir], slow = quad_slow[ndir];
      int a[3];
      a[ndir] = s * k;
      for (int ms = 0; ms < km1; ++ms)
        for (int mf = 0; mf < km1; ++mf)
          {
            a[slow]                  = ms + 1;
            a[fast]                  = mf + 1;
            out[lex(a[0], a[1], a[2])] = (uint64_t)(g + mf + km1 * ms);
          }
    }
  const uint32_t m = own_mask_[cell];
  int64_t g        = base_[cell] + popc(m & 0xffu) + (int64_t)km1 * popc((m >> 8) & 0xfffu) + (int64_t)km1 * km1 * popc((m >> 20) & 0x3fu);
  for (int az = 1; az < k; ++az)
    for (int ay = 1; ay < k; ++ay)
      for (int ax = 1; ax < k; ++ax) out[lex(ax, ay, az)] = (uint64_t)g++;
}

void DoFHandler::substituted_indices(int64_t cell, uint16_t kind, uint64_t *out) const
{
  raw_indices(cell, out);
  if (kind == 0) return;
  const int k = k_, n = k + 1;
  const Node &nd = tree_.nodes()[tree_.cells()[cell]];
  const int l    = nd.level;
  int b[3];
  for (int d = 0; d < 3; ++d) b[d] = nd.c[d] & 1;
  std::vector<uint64_t> nbr((size_t)n * n * n);
  auto lex = [n](const int *a) { return a[0] + n * (a[1] + n * a[2]); };
  auto coarse_cell = [&](const int p[3]) {
    const int32_t nb = tree_.find(l - 1, p[0], p[1], p[2]);
    if (tree_.nodes()[nb].level != l - 1 || tree_.nodes()[nb].first_child >= 0) throw std::logic_error("coarse neighbour not found");
    return tree_.cell_index(nb);
  };
  for (int d = 0; d < 3; ++d)
    if (kind & (1u << (3 + d)))
      {
        int p[3] = {nd.c[0] >> 1, nd.c[1] >> 1, nd.c[2] >> 1};
        p[d] += 2 * b[d] - 1;
        raw_indices(coarse_cell(p), nbr.data());
        const int t0 = (d + 1) % 3, t1 = (d + 2) % 3;
        int a[3], an[3];
        a[d]  = b[d] * k;
        an[d] = (1 - b[d]) * k;
        for (int m1 = 0; m1 < n; ++m1)
          for (int m0 = 0; m0 < n; ++m0)
            {
              a[t0] = an[t0] = m0;
              a[t1] = an[t1] = m1;
              out[lex(a)]    = nbr[lex(an)];
            }
      }
  for (int d = 0; d < 3; ++d)
    if (kind & (1u << (6 + d)))
      {
        const int t0 = (d + 1) % 3, t1 = (d + 2) % 3;
        int p[3] = {nd.c[0] >> 1, nd.c[1] >> 1, nd.c[2] >> 1};
        p[t0] += 2 * b[t0] - 1;
        p[t1] += 2 * b[t1] - 1;
        raw_indices(coarse_cell(p), nbr.data());
        int a[3], an[3];
        a[t0]  = b[t0] * k;
        a[t1]  = b[t1] * k;
        an[t0] = (1 - b[t0]) * k;
        an[t1] = (1 - b[t1]) * k;
        for (int m = 0; m < n; ++m)
          {
            a[d] = an[d] = m;
            out[lex(a)]  = nbr[lex(an)];
          }
      }
}

void DoFHandler::support_points(int64_t begin, int64_t end, double *xyz) const
{
  const Shape1D sh = make_shape(k_);
  const int k = k_, n = k + 1;
  const int64_t nc = (int64_t)tree_.cells().size();
#pragma omp parallel
  {
    std::vector<uint64_t> idx((size_t)n * n * n);
#pragma omp for schedule(static)
    for (int64_t s = 0; s < nc; ++s)
      {
        // only the cell's own block of new indices can fall into the range
        const int64_t b0 = base_[s];
        if (b0 >= end) continue;
        raw_indices(s, idx.data());
        const Node &nd = tree_.nodes()[tree_.cells()[s]];
        const double hh = 2.0 / (double)(1 << nd.level);
        for (int az = 0; az < n; ++az)
          for (int ay = 0; ay < n; ++ay)
            for (int ax = 0; ax < n; ++ax)
              {
                const int64_t g = (int64_t)idx[ax + n * (ay + n * az)];
                if (g < b0 || g < begin || g >= end) continue; // numbered by an earlier cell
                double *p = xyz + 3 * (g - begin);
                p[0]      = -1.0 + (nd.c[0] + sh.nodes[ax]) * hh;
                p[1]      = -1.0 + (nd.c[1] + sh.nodes[ay]) * hh;
                p[2]      = -1.0 + (nd.c[2] + sh.nodes[az]) * hh;
              }
      }
  }
}
} // namespace mfhn
