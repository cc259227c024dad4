// C++ host-side wrapper over the C ABI (mfhn.h), shaped like the reference's
// operator surface so that benchmark_03.h's `run` reads the same:
//
//   reference (deal.II)                              here
//   -----------------------------------------------  -----------------------------------
//   parallel::distributed::Triangulation<3> tria     mfhn::Triangulation tria(geometry, L)
//   GridGenerator::create_annulus(tria, L)             (benchmark_03.h:397-404)
//   DoFHandler<3> dof_handler(tria);                 mfhn::DoFHandler dof_handler(tria, degree)
//   dof_handler.distribute_dofs(FE_Q<3>(degree))       (benchmark_03.h:438-439)
//   MatrixFree<3,Number> matrix_free;                mfhn::MatrixFree matrix_free(dof_handler, rank)
//     matrix_free.reinit(mapping, dof_handler, ...)      (benchmark_03.h:326-340; cell order, categorisation,
//                                                         rank-local numbering and partitioner come from the library)
//   LaplaceOperator<3,degree,Number,MemorySpace::CUDA> mfhn::LaplaceOperator<3, degree, Number>
//     op(mapping, dof_handler, constraints, quad, ac)    op(matrix_free, apply_constraints)
//   op.initialize_dof_vector(v); op.vmult(dst, src)  identical                (benchmark_03.h:342-353)
//   (ghost exchange inside cell_loop)                op.attach_communicator(unique_id): NCCL import / compress
//                                                         overlapped with the interior cells (mfhn_dist_vmult)
//
// Non-zero C status codes become exceptions, mirroring the reference's
// AssertThrow(..., ExcMessage / ExcNotImplemented()) convention.
#pragma once
#include "mfhn.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace mfhn
{
struct ExcMessage : std::runtime_error
{
  using std::runtime_error::runtime_error;
};
struct ExcNotImplemented : std::runtime_error
{
  using std::runtime_error::runtime_error;
};
inline void check(int status)
{
  if (status == MFHN_OK) return;
  const std::string msg = mfhn_last_error();
  if (status == MFHN_ERR_NOT_IMPL) throw ExcNotImplemented(msg);
  throw ExcMessage(msg);
}

class Triangulation
{
public:
  Triangulation(const std::string &geometry_type, unsigned n_refinements, bool p4est_balance = true)
  {
    check(mfhn_mesh_create(geometry_type.c_str(), (int)n_refinements, p4est_balance ? MFHN_P4EST : MFHN_SERIAL, &h_));
  }
  ~Triangulation() { mfhn_mesh_destroy(h_); }
  Triangulation(const Triangulation &) = delete;
  Triangulation &operator=(const Triangulation &) = delete;
  int64_t n_global_active_cells() const { return mfhn_mesh_n_cells(h_); }
  int n_global_levels() const { return mfhn_mesh_n_levels(h_); }
  // Helper<dim>::is_constrained count (constraint_helper.h:89-125)
  int64_t n_cells_with_hanging_nodes() const { return mfhn_mesh_n_cells_hn(h_); }
  std::vector<int64_t> morton_position() const
  {
    std::vector<int64_t> p(n_global_active_cells());
    check(mfhn_mesh_morton_position(h_, p.data()));
    return p;
  }
  mfhn_mesh handle() const { return h_; }

private:
  mfhn_mesh h_ = nullptr;
};

class DoFHandler
{
public:
  // n_ranks > 1: rank-major numbering of the Morton (p4est) partition, optionally weighted (benchmark_02.cc:15-37)
  DoFHandler(const Triangulation &tria, int fe_degree, int n_ranks = 1, double hn_weight = 1.0)
    : tria_(tria)
    , degree_(fe_degree)
    , n_ranks_(n_ranks)
  {
    if (n_ranks > 1)
      {
        std::vector<int32_t> rank_of_cell(tria.n_global_active_cells());
        check(mfhn_mesh_partition(tria.handle(), n_ranks, hn_weight, rank_of_cell.data()));
        check(mfhn_dofs_create(tria.handle(), fe_degree, n_ranks, rank_of_cell.data(), &h_));
      }
    else
      check(mfhn_dofs_create(tria.handle(), fe_degree, 1, nullptr, &h_));
  }
  ~DoFHandler() { mfhn_dofs_destroy(h_); }
  DoFHandler(const DoFHandler &) = delete;
  DoFHandler &operator=(const DoFHandler &) = delete;
  int64_t n_dofs() const { return mfhn_dofs_n_dofs(h_); }
  int degree() const { return degree_; }
  int n_ranks() const { return n_ranks_; }
  const Triangulation &get_triangulation() const { return tria_; }
  mfhn_dofs handle() const { return h_; }

private:
  const Triangulation &tria_;
  int degree_, n_ranks_;
  mfhn_dofs h_ = nullptr;
};

// MatrixFree::reinit for one rank (mfhn_mf_create): benchmark_03.h:326-340, benchmark_01.h:251-284
class MatrixFree
{
public:
  explicit MatrixFree(const DoFHandler &dof_handler, int rank = 0, bool categorize = true)
    : dof_handler_(dof_handler)
  {
    mfhn_mf_options opt{};
    opt.rank       = rank;
    opt.categorize = categorize ? 1 : 0;
    check(mfhn_mf_create(dof_handler.handle(), &opt, &h_));
    check(mfhn_mf_info(h_, &sizes_));
  }
  ~MatrixFree() { mfhn_mf_destroy(h_); }
  MatrixFree(const MatrixFree &) = delete;
  MatrixFree &operator=(const MatrixFree &) = delete;
  const mfhn_mf_sizes &sizes() const { return sizes_; }
  const DoFHandler &get_dof_handler() const { return dof_handler_; }
  // global indices this rank ghosts from `peer` (peer = ghost peer number p < n_ghost_peers): rank and index list
  int ghost_peer(int p, const int64_t *&indices, int64_t &n) const
  {
    const int32_t *peers;
    const int64_t *gb, *ge, *ghost_global;
    check(mfhn_mf_partitioner(h_, &peers, &gb, &ge, nullptr, nullptr, nullptr, nullptr));
    check(mfhn_mf_arrays(h_, nullptr, nullptr, nullptr, nullptr, &ghost_global, nullptr));
    indices = ghost_global + gb[p];
    n       = ge[p] - gb[p];
    return peers[p];
  }
  void set_imports(int peer, const int64_t *global_indices, int64_t n)
  {
    check(mfhn_mf_set_imports(h_, peer, global_indices, n));
    check(mfhn_mf_info(h_, &sizes_));
  }
  // support points of the locally owned DoFs (VectorTools::interpolate, benchmark_03.h:455-468)
  std::vector<double> owned_support_points() const
  {
    std::vector<double> xyz(3 * (size_t)sizes_.n_owned);
    check(mfhn_dofs_support_points(dof_handler_.handle(), sizes_.owned_begin, sizes_.owned_begin + sizes_.n_owned, xyz.data()));
    return xyz;
  }
  mfhn_mf handle() const { return h_; }

private:
  const DoFHandler &dof_handler_;
  mfhn_mf h_ = nullptr;
  mfhn_mf_sizes sizes_{};
};

// LinearAlgebra::distributed::Vector<Number, MemorySpace::CUDA> for one rank
template <typename Number>
class Vector
{
public:
  Vector() = default;
  ~Vector() { cudaFree(d_); }
  Vector(const Vector &) = delete;
  Vector &operator=(const Vector &) = delete;
  void reinit(int64_t n)
  {
    cudaFree(d_);
    d_ = nullptr;
    n_ = n;
    // MFHN_VECTOR_PADDING spare (zero) entries behind the vector: see mfhn_op_desc::vector_padding
    if (cudaMalloc(&d_, sizeof(Number) * (n + MFHN_VECTOR_PADDING)) != cudaSuccess) throw ExcMessage("cudaMalloc failed");
    cudaMemset(d_, 0, sizeof(Number) * (n + MFHN_VECTOR_PADDING));
  }
  Vector &operator=(Number v)
  {
    if (v == Number(0))
      cudaMemset(d_, 0, sizeof(Number) * n_);
    else
      import_from_host(std::vector<Number>((size_t)n_, v));
    return *this;
  }
  void import_from_host(const std::vector<Number> &h) { cudaMemcpy(d_, h.data(), sizeof(Number) * n_, cudaMemcpyHostToDevice); }
  std::vector<Number> to_host() const
  {
    std::vector<Number> h(n_);
    cudaMemcpy(h.data(), d_, sizeof(Number) * n_, cudaMemcpyDeviceToHost);
    return h;
  }
  int64_t size() const { return n_; }
  Number *data() { return d_; }
  const Number *data() const { return d_; }

private:
  Number *d_ = nullptr;
  int64_t n_ = 0;
};

template <int dim, int fe_degree, typename Number>
class LaplaceOperator
{
  static_assert(dim == 3, "the engine covers the reference's 3D path");

public:
  using VectorType = Vector<Number>;

  // benchmark_03.h:326-340.  Mapping is MappingQ1 on Cartesian cells, constraints are empty
  // and the quadrature is QGauss<1>(fe_degree + 1), exactly as in the reference driver.
  LaplaceOperator(const MatrixFree &matrix_free, const bool apply_constraints, const int kernel = MFHN_KERNEL_AUTO)
    : matrix_free_(matrix_free)
  {
    if (matrix_free.sizes().degree != fe_degree) throw ExcMessage("Degrees do not match!"); // benchmark_01.h:204-206
    check(mfhn_op_create_mf_padded(matrix_free.handle(), sizeof(Number) == 8 ? MFHN_F64 : MFHN_F32, kernel, apply_constraints, -1,
                                   MFHN_VECTOR_PADDING, &op_)); // Vector::reinit pads its allocation
  }
  ~LaplaceOperator()
  {
    if (dist_) mfhn_dist_destroy(dist_);
    mfhn_op_destroy(op_);
  }
  // partitioned operator: unique_id = 128 bytes from mfhn_dist_unique_id on rank 0, the same on every rank; the
  // import lists of the MatrixFree must be set (MatrixFree::set_imports)
  void attach_communicator(const void *unique_id) { check(mfhn_dist_create_mf(op_, matrix_free_.handle(), unique_id, &dist_)); }
  LaplaceOperator(const LaplaceOperator &) = delete;
  LaplaceOperator &operator=(const LaplaceOperator &) = delete;

  // owned entries first, ghost entries behind them (LinearAlgebra::distributed::Vector layout)
  void initialize_dof_vector(VectorType &vec) const { vec.reinit(matrix_free_.sizes().n_owned + matrix_free_.sizes().n_ghost); }

  // accumulates into dst like cell_loop(local_operator, src, dst) (benchmark_03.h:348-353); asynchronous.
  // With a communicator: ghost import, three cell partitions and ghost compress in one call.
  void vmult(VectorType &dst, const VectorType &src, cudaStream_t stream = nullptr) const
  {
    if (dist_)
      check(mfhn_dist_vmult(dist_, dst.data(), src.data(), stream, 0));
    else
      check(mfhn_op_vmult(op_, dst.data(), src.data(), stream, 0));
  }
  double query(const char *what) const
  {
    double v;
    check(mfhn_op_query(op_, what, &v));
    return v;
  }
  // switches of the reference's stage decomposition (benchmark_01.cc:70-116, 179-234)
  void set_apply_constraints(const bool flag) { check(mfhn_op_set_apply_constraints(op_, flag)); } // do_apply_constraints
  void set_kernel(const int kernel) { check(mfhn_op_set_kernel(op_, kernel)); }                   // e.g. MFHN_KERNEL_QPOINT_ROWS: general-purpose algorithm
  void set_hn_strategy(const int strategy) { check(mfhn_op_set_hn_strategy(op_, strategy)); }     // MFHN_HN_BRANCH / MFHN_HN_MASK
  // "DG (C)": dst_cells += [W^T W] src_cells on cell-local arrays [n_cells][(fe_degree+1)^3] (device pointers)
  void dg_copy(Number *dst_cells, const Number *src_cells, cudaStream_t stream = nullptr) const { check(mfhn_op_dg_copy(op_, dst_cells, src_cells, stream)); }

  // Extension (BASELINE.json config 5; the reference has no solver): 1 / diag(A), and CG with point-Jacobi -- one
  // (partitioned) vmult, two fused vector kernels and one batched all-reduce per iteration, all inside the library
  void compute_inverse_diagonal(VectorType &inv_diag, cudaStream_t stream = nullptr) const
  {
    initialize_dof_vector(inv_diag);
    check(mfhn_op_inverse_diagonal(op_, dist_, inv_diag.data(), stream));
  }
  mfhn_cg_result solve_cg(VectorType &x, const VectorType &b, const VectorType *inv_diag, const double rel_tol = 1e-8, const int max_iter = 1000,
                          const bool timings = false, cudaStream_t stream = nullptr) const
  {
    mfhn_cg_options opt{};
    opt.max_iter    = max_iter;
    opt.rel_tol     = rel_tol;
    opt.check_every = 10;
    opt.timings     = timings;
    mfhn_cg_result res{};
    check(mfhn_cg_solve(op_, dist_, x.data(), b.data(), inv_diag ? inv_diag->data() : nullptr, &opt, &res, nullptr, stream));
    return res;
  }

private:
  const MatrixFree &matrix_free_;
  mfhn_op op_     = nullptr;
  mfhn_dist dist_ = nullptr;
};
} // namespace mfhn
