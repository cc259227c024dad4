#include "octree.hpp"

#include <cmath>
#include <stdexcept>

namespace mfhn
{
Octree::Octree(int flavour)
  : flavour_(flavour)
{
  nodes_.push_back(Node{-1, 0, {0, 0, 0}});
  levels_.push_back({0});
  // 2:1 balance neighbourhood: faces + edges (deal.II serial Triangulation in
  // 3D), plus corners for p4est's full balance.
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx)
        {
          const int s = std::abs(dx) + std::abs(dy) + std::abs(dz);
          if (s == 0 || (s == 3 && flavour_ == 0)) continue;
          offsets_.push_back((int8_t)dx);
          offsets_.push_back((int8_t)dy);
          offsets_.push_back((int8_t)dz);
        }
}

void Octree::refine(std::vector<uint8_t> &flag)
{
  // close the flags under the 2:1 rule (prepare_coarsening_and_refinement)
  std::vector<int32_t> queue;
  for (int32_t id = 0; id < (int32_t)nodes_.size(); ++id)
    if (flag[id]) queue.push_back(id);
  while (!queue.empty())
    {
      const int32_t id = queue.back();
      queue.pop_back();
      const Node nd = nodes_[id];
      if (nd.level == 0) continue;
      for (size_t o = 0; o < offsets_.size(); o += 3)
        {
          const int ni = nd.c[0] + offsets_[o], nj = nd.c[1] + offsets_[o + 1], nk = nd.c[2] + offsets_[o + 2];
          if (!inside(nd.level, ni, nj, nk)) continue;
          const int32_t nb = find(nd.level, ni, nj, nk);
          if (nodes_[nb].level < nd.level)
            {
              if (nodes_[nb].level != nd.level - 1) throw std::logic_error("mesh was not 2:1 balanced");
              if (!flag[nb])
                {
                  flag[nb] = 1;
                  queue.push_back(nb);
                }
            }
        }
    }
  // execute_refinement: level by level, cells in index order, children appended
  const int nl = (int)levels_.size();
  for (int l = 0; l < nl; ++l)
    {
      const size_t cnt = levels_[l].size();
      for (size_t t = 0; t < cnt; ++t)
        {
          const int32_t id = levels_[l][t];
          if (id >= (int32_t)flag.size() || !flag[id]) continue;
          if ((int)levels_.size() == l + 1) levels_.emplace_back();
          const int32_t fc = (int32_t)nodes_.size();
          nodes_[id].first_child = fc;
          const Node p = nodes_[id];
          for (int ch = 0; ch < 8; ++ch)
            {
              nodes_.push_back(Node{-1, l + 1, {2 * p.c[0] + (ch & 1), 2 * p.c[1] + ((ch >> 1) & 1), 2 * p.c[2] + ((ch >> 2) & 1)}});
              levels_[l + 1].push_back(fc + ch);
            }
        }
    }
}

void Octree::refine_global(int times)
{
  for (int t = 0; t < times; ++t)
    {
      std::vector<uint8_t> flag(nodes_.size(), 0);
      for (size_t id = 0; id < nodes_.size(); ++id)
        if (nodes_[id].first_child < 0) flag[id] = 1;
      refine(flag);
    }
}

void Octree::refine_if(const std::function<bool(const double *)> &pred)
{
  std::vector<uint8_t> flag(nodes_.size(), 0);
  for (size_t id = 0; id < nodes_.size(); ++id)
    if (nodes_[id].first_child < 0)
      {
        const Node &nd = nodes_[id];
        const double h = 2.0 / (double)(1 << nd.level);
        const double c[3] = {-1.0 + (nd.c[0] + 0.5) * h, -1.0 + (nd.c[1] + 0.5) * h, -1.0 + (nd.c[2] + 0.5) * h};
        if (pred(c)) flag[id] = 1;
      }
  refine(flag);
}

void Octree::finalize()
{
  cells_.clear();
  cell_of_node_.assign(nodes_.size(), -1);
  for (const auto &lev : levels_)
    for (int32_t id : lev)
      if (nodes_[id].first_child < 0)
        {
          cell_of_node_[id] = (int64_t)cells_.size();
          cells_.push_back(id);
        }
}

std::vector<int64_t> Octree::morton_order() const
{
  std::vector<int64_t> out;
  out.reserve(cells_.size());
  std::vector<int32_t> stack{0};
  while (!stack.empty())
    {
      const int32_t id = stack.back();
      stack.pop_back();
      const int32_t fc = nodes_[id].first_child;
      if (fc < 0)
        out.push_back(cell_of_node_[id]);
      else
        for (int ch = 7; ch >= 0; --ch) stack.push_back(fc + ch);
    }
  return out;
}

uint16_t Octree::constraint_kind(int32_t id) const
{
  // Detection of coarser face / edge neighbours: the cell-level twin of
  // Helper::is_constrained (constraint_helper.h:89-125).  Constraints can only
  // sit on the parent's outer faces / edges.
  const Node &nd = nodes_[id];
  const int l = nd.level;
  if (l == 0) return 0;
  int b[3], out[3];
  for (int d = 0; d < 3; ++d)
    {
      b[d]   = nd.c[d] & 1;
      out[d] = nd.c[d] + 2 * b[d] - 1;
    }
  auto coarser = [&](const int p[3]) {
    if (!inside(l, p[0], p[1], p[2])) return false;
    return nodes_[find(l, p[0], p[1], p[2])].level < l;
  };
  bool face[3], edge[3] = {false, false, false};
  for (int d = 0; d < 3; ++d)
    {
      int p[3] = {nd.c[0], nd.c[1], nd.c[2]};
      p[d]     = out[d];
      face[d]  = coarser(p);
    }
  for (int d = 0; d < 3; ++d)
    {
      const int a = (d + 1) % 3, bb = (d + 2) % 3;
      if (face[a] || face[bb]) continue;
      int p[3] = {nd.c[0], nd.c[1], nd.c[2]};
      p[a]     = out[a];
      p[bb]    = out[bb];
      edge[d]  = coarser(p);
    }
  if (!(face[0] || face[1] || face[2] || edge[0] || edge[1] || edge[2])) return 0;
  uint16_t kind = 0;
  for (int d = 0; d < 3; ++d)
    kind |= (uint16_t)(((1 - b[d]) << d) | ((int)face[d] << (3 + d)) | ((int)edge[d] << (6 + d)));
  return kind;
}

Octree Octree::create(const std::string &geometry, int L, int flavour)
{
  if (flavour != 0 && flavour != 1) throw std::invalid_argument("Unknown mesh flavour!");
  if (L < 0 || L > 14) throw std::invalid_argument("n_refinements out of range [0,14]");
  Octree t(flavour);
  auto all_neg = [](const double *c) { return c[0] <= 0.0 && c[1] <= 0.0 && c[2] <= 0.0; };
  if (geometry == "quadrant" || geometry == "step")
    {
      // benchmark.h:38-69 (create_quadrant), :7-34 (create_step)
      if (L > 0)
        {
          t.refine_global(1);
          for (int i = 1; i < L; ++i)
            if (geometry == "quadrant")
              t.refine_if(all_neg);
            else
              t.refine_if([](const double *c) { return c[0] <= 0.0; });
          if (t.n_levels() - 1 != L) throw std::logic_error("n_global_levels-1 != n_refinements");
        }
    }
  else if (geometry == "quadrant_flexible")
    {
      // benchmark.h:73-96 with n_ref_local = 1
      t.refine_global(L);
      t.refine_if(all_neg);
    }
  else if (geometry == "annulus")
    {
      // benchmark.h:100-144
      if (L > 0)
        {
          for (int i = 0; i < L - 3; ++i) t.refine_global(1);
          auto norm = [](const double *c) { return std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]); };
          if (L >= 1) t.refine_if([&](const double *c) { return norm(c) < 0.55; });
          if (L >= 2) t.refine_if([&](const double *c) { return 0.3 <= norm(c) && norm(c) <= 0.43; });
          if (L >= 3) t.refine_if([&](const double *c) { return 0.335 <= norm(c) && norm(c) <= 0.39; });
        }
    }
  else
    throw std::invalid_argument("Unknown geometry type!"); // benchmark_01.h:217, benchmark_03.h:404
  t.finalize();
  return t;
}
} // namespace mfhn
