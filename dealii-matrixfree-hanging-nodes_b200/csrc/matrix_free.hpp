// Rank-local product of MatrixFree::reinit (benchmark_03.h:326-340, benchmark_01.h:251-284): the order in which the
// cell loop visits the cells (Morton curve, interior cells first, grouped by constraint mask inside windows -- the
// reference's cell_vectorization_category / Categorize option), rank-local DoF numbering (owned range, then ghosts
// sorted by global index), compressed constraint masks, Cartesian geometry, and the Utilities::MPI::Partitioner data
// (ghost ranges per owner, import lists per reader).  Everything the operator (op.cu) and the partitioned operator
// need; host memory only.
#pragma once
#include "dof_handler.hpp"

#include <cstdint>
#include <vector>

namespace mfhn
{
struct MatrixFreeOptions
{
  int rank             = 0;
  int categorize       = 1;    // 0: plain Morton order, 1: group by constraint mask, 2: group by constrained / unconstrained
  int window           = 3840; // cells per categorisation window
  int batch_alignment  = 240;  // partition boundaries on multiples of this many cells (lcm of the warp batch sizes)
};

struct MatrixFreeData
{
  int degree = 0, rank = 0, n_ranks = 1;
  int64_t n_cells = 0, n_owned = 0, n_ghost = 0, owned_begin = 0, owned_end = 0;
  int64_t n_interior = 0, n_interior_a = 0; // cells [0, n_interior_a) | [n_interior_a, n_interior) | boundary cells
  std::vector<int64_t> cell_ids;            // storage indices of the cells in loop order
  std::vector<uint32_t> dof_indices;        // [n_cells][(k+1)^3], rank-local, lexicographic, coarse-substituted
  std::vector<uint8_t> masks;
  std::vector<double> h;
  std::vector<int64_t> rank_begin;          // [n_ranks + 1] owned ranges of all ranks
  std::vector<int64_t> ghost_global;        // sorted global indices of the ghost entries
  std::vector<int32_t> ghost_owner;         // per ghost entry
  std::vector<int32_t> ghost_peers;         // owners of this rank's ghosts, ascending
  std::vector<int64_t> ghost_begin, ghost_end; // per peer: contiguous range inside the ghost section
  std::vector<int32_t> import_peers;        // ranks that ghost entries owned here, ascending
  std::vector<int64_t> import_offsets;      // [n_import_peers + 1]
  std::vector<int32_t> import_indices;      // local owned indices, grouped by peer

  void reinit(const DoFHandler &dh, const Octree &tree, const MatrixFreeOptions &opt);
  // peer -> global indices (owned by this rank) that the peer ghosts; replaces the list of that peer
  void set_imports(int peer, const int64_t *global_indices, int64_t n);
  int64_t n_cells_hn() const;
};
} // namespace mfhn
