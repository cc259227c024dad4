// FE_Q(k) DoF enumeration + hanging-node masks + coarse-index substitution on
// an Octree.  Stands in for DoFHandler::distribute_dofs (benchmark_01.h:247,
// benchmark_03.h:438-439) and the index / mask part of MatrixFree::reinit
// (HangingNodes::setup_constraints inside deal.II).
//
// Hash-free and parallel: every geometric object (vertex / line / quad) is
// numbered by the first active cell in walk order that has it, exactly the
// outcome of deal.II's sequential "number what is not numbered yet" walk; a
// prefix sum over the per-cell counts of first-seen objects gives each cell's
// block of new indices.
#pragma once
#include "octree.hpp"

#include <cstdint>
#include <vector>

namespace mfhn
{
class DoFHandler
{
public:
  DoFHandler(const Octree &tree, int degree, int n_ranks, const int32_t *rank_of_cell);

  int degree() const { return k_; }
  int64_t n_dofs() const { return n_dofs_; }
  int n_ranks() const { return n_ranks_; }
  void owned_range(int rank, int64_t &begin, int64_t &end) const
  {
    begin = rank_begin_[rank];
    end   = rank_begin_[rank + 1];
  }
  // storage indices of the cells of a rank, in storage order
  const std::vector<int64_t> &cells_of_rank(int rank) const { return cells_of_rank_[rank]; }

  void raw_indices(int64_t cell, uint64_t *out) const;         // (k+1)^3 lexicographic
  void substituted_indices(int64_t cell, uint16_t kind, uint64_t *out) const;
  uint16_t kind(int64_t cell) const { return tree_.constraint_kind(tree_.cells()[cell]); }
  double h(int64_t cell) const { return 2.0 / (double)(1 << tree_.nodes()[tree_.cells()[cell]].level); }
  void support_points(int64_t begin, int64_t end, double *xyz) const;

private:
  struct Owner
  {
    int64_t cell; // storage index of the owning cell
    int obj;      // local object number in the owner (0-7 vertex, 8-19 line, 20-25 quad)
  };
  Owner owner_of(int64_t cell, int obj) const;
  int64_t object_base(int64_t cell, int obj) const;

  const Octree &tree_;
  int k_, n_ranks_;
  int64_t n_dofs_ = 0;
  std::vector<int64_t> walk_pos_;  // storage index -> position in the numbering walk
  std::vector<uint32_t> own_mask_; // 26 bits per cell: objects first seen by this cell
  std::vector<int64_t> base_;      // first new index of each cell
  std::vector<int64_t> rank_begin_;
  std::vector<std::vector<int64_t>> cells_of_rank_;
};
} // namespace mfhn
