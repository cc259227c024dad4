// deal.II-side shim: LaplaceOperator<3, fe_degree, Number, MemorySpace::CUDA> of the reference
// (/root/reference/benchmark_03.h:319-357) on top of libmfhn.so (include/mfhn.h).
//
// Usage in the reference: include this header in benchmark_03.h instead of the CUDAWrappers::MatrixFree based
// specialisation (benchmark_03.h:319-357), link -lmfhn.  Nothing else in the driver changes: the constructor arguments,
// initialize_dof_vector() and vmult(dst, src) are the reference's.  The host MatrixFree (fast hanging-node algorithm,
// benchmark_01.h:286-293) provides what the engine consumes: dof_info.dof_indices (rank-local, owned then ghosts,
// lexicographic, coarse-substituted), dof_info.hanging_node_constraint_masks (compressed_constraint_kind) and the vector
// partitioner.  For the partitioned path the ghost exchange stays deal.II's (update_ghost_values / compress(add) of
// LinearAlgebra::distributed::Vector); INTEGRATION.md section 3 shows the NCCL variant (mfhn_dist_*).
//
// This header needs deal.II (not available in this repository's build environment, so it is not compiled here);
// tools/diff_setup.py compares the arrays it hands over -- written with MFHN_DEALII_DUMP_SETUP -- with this engine's
// own setup layer, bit for bit.
#ifndef MFHN_DEALII_H
#define MFHN_DEALII_H

#include <deal.II/base/exceptions.h>
#include <deal.II/base/quadrature.h>
#include <deal.II/base/utilities.h>
#include <deal.II/base/vectorization.h>

#include <deal.II/dofs/dof_handler.h>

#include <deal.II/fe/mapping.h>

#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/la_parallel_vector.h>

#include <deal.II/matrix_free/matrix_free.h>

#include <mfhn.h>

#include <cstdint>
#include <fstream>
#include <type_traits>
#include <vector>

namespace mfhn_dealii
{
  using namespace dealii;

  // The three arrays of mfhn_op_desc from a host MatrixFree: one row per active cell, in MatrixFree's own cell order.
  template <int fe_degree, typename Number>
  struct SetupArrays
  {
    std::vector<std::uint32_t> dof_indices; // [n_cells][(fe_degree+1)^3]
    std::vector<std::uint8_t>  masks;       // [n_cells]
    std::vector<double>        h;           // [n_cells] Cartesian edge length

    explicit SetupArrays(const MatrixFree<3, Number> &matrix_free)
    {
      constexpr unsigned int n3 = Utilities::pow(fe_degree + 1, 3);
      const auto &           di = matrix_free.get_dof_info();
      for (unsigned int b = 0; b < matrix_free.n_cell_batches(); ++b)
        for (unsigned int v = 0; v < matrix_free.n_active_entries_per_cell_batch(b); ++v)
          {
            const unsigned int  c = b * VectorizedArray<Number>::size() + v;
            const unsigned int *p = di.dof_indices.data() + di.row_starts[c].first; // plain (unconstrained) layout
            dof_indices.insert(dof_indices.end(), p, p + n3);
            masks.push_back(static_cast<std::uint8_t>(di.hanging_node_constraint_masks[c])); // benchmark_01.h:335
            h.push_back(matrix_free.get_cell_iterator(b, v)->extent_in_direction(0));
          }
    }

    // raw little-endian dump for tools/diff_setup.py: int64 n_cells, int64 n3, uint32 indices, uint8 masks, float64 h
    void dump(const std::string &file) const
    {
      std::ofstream      out(file, std::ios::binary);
      const std::int64_t n_cells = masks.size(), n3 = n_cells ? dof_indices.size() / n_cells : 0;
      out.write(reinterpret_cast<const char *>(&n_cells), 8);
      out.write(reinterpret_cast<const char *>(&n3), 8);
      out.write(reinterpret_cast<const char *>(dof_indices.data()), dof_indices.size() * 4);
      out.write(reinterpret_cast<const char *>(masks.data()), masks.size());
      out.write(reinterpret_cast<const char *>(h.data()), h.size() * 8);
    }
  };
} // namespace mfhn_dealii

template <int fe_degree, typename Number>
class LaplaceOperator<3, fe_degree, Number, dealii::MemorySpace::CUDA>
{
public:
  using VectorType = dealii::LinearAlgebra::distributed::Vector<Number, dealii::MemorySpace::CUDA>;

  LaplaceOperator(const dealii::Mapping<3> &mapping, const dealii::DoFHandler<3> &dof_handler,
                  const dealii::AffineConstraints<Number> &constraints, const dealii::Quadrature<1> &quadrature,
                  const bool apply_constraints)
  {
    typename dealii::MatrixFree<3, Number>::AdditionalData ad;
    ad.mapping_update_flags = dealii::update_gradients;
    matrix_free.reinit(mapping, dof_handler, constraints, quadrature, ad);
    const mfhn_dealii::SetupArrays<fe_degree, Number> arrays(matrix_free);
#ifdef MFHN_DEALII_DUMP_SETUP
    arrays.dump(MFHN_DEALII_DUMP_SETUP);
#endif
    const auto &partitioner = *matrix_free.get_dof_info().vector_partitioner;
    mfhn_op_desc d{};
    d.degree            = fe_degree;
    d.number            = std::is_same<Number, double>::value ? MFHN_F64 : MFHN_F32;
    d.n_cells           = arrays.masks.size();
    d.n_owned           = partitioner.locally_owned_size();
    d.n_ghost           = partitioner.n_ghost_indices();
    d.dof_indices       = arrays.dof_indices.data();
    d.masks             = arrays.masks.data();
    d.geometry_type     = MFHN_GEOM_CARTESIAN;
    d.geometry          = arrays.h.data();
    d.apply_constraints = apply_constraints;
    d.kernel            = MFHN_KERNEL_AUTO;
    d.device            = -1;
    d.vector_padding    = 0; // deal.II allocates exactly n_owned + n_ghost entries: runs that would reach past the end become single entries
    AssertThrow(mfhn_op_create(&d, &op) == MFHN_OK, dealii::ExcMessage(mfhn_last_error()));
  }
  ~LaplaceOperator()
  {
    mfhn_op_destroy(op);
  }
  LaplaceOperator(const LaplaceOperator &) = delete;
  LaplaceOperator &operator=(const LaplaceOperator &) = delete;

  void
  initialize_dof_vector(VectorType &vec) const
  {
    matrix_free.initialize_dof_vector(vec);
  }

  // dst += A src like cell_loop(LaplaceOperatorLocal, src, dst) (benchmark_03.h:348-353)
  void
  vmult(VectorType &dst, const VectorType &src) const
  {
    src.update_ghost_values();
    const int status = mfhn_op_vmult(op, dst.get_values(), src.get_values(), /*stream*/ nullptr, /*zero_dst*/ 0);
    AssertThrow(status != MFHN_ERR_NOT_IMPL, dealii::ExcNotImplemented());
    AssertThrow(status == MFHN_OK, dealii::ExcMessage(mfhn_last_error()));
    dst.compress(dealii::VectorOperation::add);
    src.zero_out_ghost_values();
  }

private:
  dealii::MatrixFree<3, Number> matrix_free;
  mfhn_op                       op = nullptr;
};

#endif // MFHN_DEALII_H
