// C++ host-side wrapper over the C ABI (mfhn.h), shaped like the reference's
// operator surface so that benchmark_03.h's `run` reads the same:
//
//   reference (deal.II)                              here
//   -----------------------------------------------  -----------------------------------
//   parallel::distributed::Triangulation<3> tria     mfhn::Triangulation tria(geometry, L)
//   GridGenerator::create_annulus(tria, L)             (benchmark_03.h:397-404)
//   DoFHandler<3> dof_handler(tria);                 mfhn::DoFHandler dof_handler(tria, degree)
//   dof_handler.distribute_dofs(FE_Q<3>(degree))       (benchmark_03.h:438-439)
//   LaplaceOperator<3,degree,Number,MemorySpace::CUDA> mfhn::LaplaceOperator<3, degree, Number>
//     op(mapping, dof_handler, constraints, quad, ac)    op(dof_handler, apply_constraints)
//   op.initialize_dof_vector(v); op.vmult(dst, src)  identical                (benchmark_03.h:342-353)
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
  DoFHandler(const Triangulation &tria, int fe_degree)
    : tria_(tria)
    , degree_(fe_degree)
  {
    check(mfhn_dofs_create(tria.handle(), fe_degree, 1, nullptr, &h_));
  }
  ~DoFHandler() { mfhn_dofs_destroy(h_); }
  DoFHandler(const DoFHandler &) = delete;
  DoFHandler &operator=(const DoFHandler &) = delete;
  int64_t n_dofs() const { return mfhn_dofs_n_dofs(h_); }
  int degree() const { return degree_; }
  const Triangulation &get_triangulation() const { return tria_; }
  mfhn_dofs handle() const { return h_; }

private:
  const Triangulation &tria_;
  int degree_;
  mfhn_dofs h_ = nullptr;
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
    if (cudaMalloc(&d_, sizeof(Number) * (n > 0 ? n : 1)) != cudaSuccess) throw ExcMessage("cudaMalloc failed");
    *this = Number(0);
  }
  Vector &operator=(Number v)
  {
    if (v != Number(0)) throw ExcNotImplemented("only zero assignment");
    cudaMemset(d_, 0, sizeof(Number) * n_);
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
  LaplaceOperator(const DoFHandler &dof_handler, const bool apply_constraints, const int kernel = MFHN_KERNEL_AUTO)
  {
    if (dof_handler.degree() != fe_degree) throw ExcMessage("Degrees do not match!"); // benchmark_01.h:204-206
    const Triangulation &tria = dof_handler.get_triangulation();
    const int64_t n_cells     = tria.n_global_active_cells();
    // MatrixFree reorders its cell batches: visit the cells along the Morton curve
    const std::vector<int64_t> pos = tria.morton_position();
    std::vector<int64_t> cells(n_cells);
    for (int64_t c = 0; c < n_cells; ++c) cells[pos[c]] = c;
    const int64_t n3 = (int64_t)(fe_degree + 1) * (fe_degree + 1) * (fe_degree + 1);
    std::vector<uint64_t> global((size_t)n_cells * n3);
    std::vector<uint8_t> masks(n_cells);
    std::vector<double> h(n_cells);
    check(mfhn_dofs_fill(dof_handler.handle(), n_cells, cells.data(), nullptr, global.data(), masks.data(), h.data()));
    std::vector<uint32_t> local(global.begin(), global.end());
    n_dofs_ = dof_handler.n_dofs();
    mfhn_op_desc d{};
    d.degree            = fe_degree;
    d.number            = sizeof(Number) == 8 ? MFHN_F64 : MFHN_F32;
    d.n_cells           = n_cells;
    d.n_owned           = n_dofs_;
    d.n_ghost           = 0;
    d.dof_indices       = local.data();
    d.masks             = masks.data();
    d.geometry_type     = MFHN_GEOM_CARTESIAN;
    d.geometry          = h.data();
    d.apply_constraints = apply_constraints;
    d.kernel            = kernel;
    d.device            = -1;
    check(mfhn_op_create(&d, &op_));
  }
  ~LaplaceOperator() { mfhn_op_destroy(op_); }
  LaplaceOperator(const LaplaceOperator &) = delete;
  LaplaceOperator &operator=(const LaplaceOperator &) = delete;

  void initialize_dof_vector(VectorType &vec) const { vec.reinit(n_dofs_); }

  // accumulates into dst like cell_loop(local_operator, src, dst) (benchmark_03.h:348-353); asynchronous
  void vmult(VectorType &dst, const VectorType &src, cudaStream_t stream = nullptr) const
  {
    check(mfhn_op_vmult(op_, dst.data(), src.data(), stream, 0));
  }
  double query(const char *what) const
  {
    double v;
    check(mfhn_op_query(op_, what, &v));
    return v;
  }

private:
  mfhn_op op_ = nullptr;
  int64_t n_dofs_ = 0;
};
} // namespace mfhn
