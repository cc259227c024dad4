// Single-tree octree on hyper_cube(-1,1)^3 with deal.II / p4est refinement
// semantics.  Stands in for the Triangulation layer the reference gets from
// deal.II (benchmark.h:7-144, benchmark_03.h:26-104, 397).  Pointer-based
// linear octree: children of a node are 8 consecutive entries, lookups walk
// down from the root (depth <= 15), no hashing.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace mfhn
{
struct Node
{
  int32_t first_child; // -1: active leaf
  int32_t level;
  int32_t c[3];        // integer coordinates at its level
};

class Octree
{
public:
  explicit Octree(int flavour);

  // geometry generators of the reference
  static Octree create(const std::string &geometry, int n_refinements, int flavour);

  int n_levels() const { return (int)levels_.size(); }
  const std::vector<Node> &nodes() const { return nodes_; }
  // active cells in deal.II iteration order (level, index)
  const std::vector<int32_t> &cells() const { return cells_; }
  // storage index of an active node, -1 otherwise
  int64_t cell_index(int32_t node) const { return cell_of_node_[node]; }
  // active cells (storage indices) along the Morton curve
  std::vector<int64_t> morton_order() const;

  // deepest existing node covering region (level; i,j,k); its level is <= level
  int32_t find(int level, int i, int j, int k) const
  {
    int32_t node = 0;
    for (int l = 1; l <= level; ++l)
      {
        const int32_t fc = nodes_[node].first_child;
        if (fc < 0) return node;
        const int sh = level - l;
        node = fc + ((i >> sh) & 1) + 2 * ((j >> sh) & 1) + 4 * ((k >> sh) & 1);
      }
    return node;
  }
  bool inside(int level, int i, int j, int k) const
  {
    const int n = 1 << level;
    return i >= 0 && j >= 0 && k >= 0 && i < n && j < n && k < n;
  }

  void refine_global(int times);
  void refine_if(const std::function<bool(const double *)> &pred);
  void finalize(); // build cells_ / cell_of_node_

  // ConstraintKinds of an active node (0 = unconstrained)
  uint16_t constraint_kind(int32_t node) const;

  int flavour() const { return flavour_; }

private:
  void refine(std::vector<uint8_t> &flag);

  int flavour_;
  std::vector<Node> nodes_;
  std::vector<std::vector<int32_t>> levels_;
  std::vector<int32_t> cells_;
  std::vector<int64_t> cell_of_node_;
  std::vector<int8_t> offsets_; // balance neighbourhood (dx,dy,dz) triples
};

inline uint8_t compress_kind(uint16_t kind)
{
  const unsigned subcell = kind & 7u, face = (kind >> 3) & 7u, edge = (kind >> 6) & 7u;
  return (uint8_t)(subcell + ((face > 0) << 3) + ((edge > 0) << 4) + ((face > edge ? face : edge) << 5));
}
inline uint16_t decompress_kind(uint8_t b)
{
  const unsigned subcell = b & 7u, flag0 = (b >> 3) & 3u, flag1 = (b >> 5) & 7u;
  return (uint16_t)(subcell + (((flag0 & 1u) ? flag1 : 0u) << 3) + (((flag0 & 2u) ? flag1 : 0u) << 6));
}
inline bool check_kind(uint16_t kind)
{
  if (kind == 0) return true;
  if (kind >> 9) return false;
  const unsigned face = (kind >> 3) & 7u, edge = (kind >> 6) & 7u;
  if (face == 0 && edge == 0) return false;
  if (face && edge) return face == edge && (face == 1 || face == 2 || face == 4);
  return true;
}
} // namespace mfhn
