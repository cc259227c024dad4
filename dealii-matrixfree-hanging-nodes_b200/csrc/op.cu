// Operator object behind the C ABI: device-resident MatrixFree data + kernel
// dispatch.  Stands in for CUDAWrappers::MatrixFree::reinit / cell_loop as used
// by LaplaceOperator<..., MemorySpace::CUDA> (benchmark_03.h:319-357).
#include "../../include/mfhn.h"
#include "error.hpp"
#include "fe1d.hpp"
#include "layouts.hpp"
#include "matrix_free.hpp"
#include "dist.cuh"
#include "octree.hpp"

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

namespace mfhn
{

#define CUDA_CHECK(x)                                                                            \
  do                                                                                             \
    {                                                                                            \
      cudaError_t e_ = (x);                                                                      \
      if (e_ != cudaSuccess)                                                                     \
        throw CudaError(std::string(#x) + ": " + cudaGetErrorString(e_));                        \
    }                                                                                            \
  while (0)

namespace
{
template <typename T>
T *to_device(const std::vector<T> &h)
{
  T *d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
} // namespace

void PlaneLayout::build(int n_, long long n_cells_, const uint32_t *idx)
{
  n               = n_;
  n_cells         = n_cells_;
  const int cpw   = 32 / n;
  n_batches       = (n_cells + cpw - 1) / cpw;
  const int n2    = n * n;
  const long long n3 = (long long)n2 * n;
  std::vector<uint32_t> p((size_t)std::max<long long>(n_batches, 1) * n2 * 32, 0xffffffffu);
#pragma omp parallel for schedule(static)
  for (long long c = 0; c < n_cells; ++c)
    {
      const long long batch = c / cpw;
      const int slot        = (int)(c % cpw);
      // kernel axes (X,Y,Z) = physical (y,z,x): thread t = x, plane slot j = y + n z
      for (int t = 0; t < n; ++t)
        for (int j = 0; j < n2; ++j) p[((size_t)batch * n2 + j) * 32 + slot * n + t] = idx[c * n3 + t + (long long)n * j];
    }
  d_pidx = to_device(p);
}

// ---- bulk-copy layout -------------------------------------------------------------
// more irregular cells than this: the layout does not fit the numbering, the plane kernel runs everything
constexpr long long bulk_max_irregular = 32;

static void bulk_build(BulkLayout &B, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx)
{
  BulkHostLayout L;
  bulk_analyze(L, n, number, n_cells, n_vec, idx);
  B.n         = n;
  B.n_cells   = n_cells;
  B.n_batches = L.n_batches;
  B.irregular = L.irregular;
  B.usable    = (long long)L.irregular.size() <= bulk_max_irregular;
  if (!B.usable) return;
  B.d_bidx  = to_device(L.bidx);
  B.d_lvidx = to_device(L.lvidx);
  B.d_cinfo = to_device(L.cinfo);
}

// ---- run-wise layout ------------------------------------------------------------------
// run detection parameters (development switches MFHN_RUNS_GAP / MFHN_RUNS_MIN): unused vector entries tolerated
// inside a run, fewest cell entries worth a bulk copy
static int env_int(const char *name, int fallback)
{
  const char *e = std::getenv(name);
  return e && *e ? std::atoi(e) : fallback;
}
static void runs_build(RunsLayout &B, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx)
{
  RunsHostLayout L;
  runs_analyze(L, n, number, n_cells, n_vec, idx, env_int("MFHN_RUNS_GAP", 10), env_int("MFHN_RUNS_MIN", 6), env_int("MFHN_RUNS_PLACE", 1) != 0);
  B.n         = n;
  B.sr        = L.sr;
  B.n_cells   = n_cells;
  B.n_batches = L.n_batches;
  B.n_blocks  = L.n_blocks;
  B.n_singles = L.n_singles;
  B.n_zero    = L.n_zero;
  if (env_int("MFHN_RUNS_STATS", 0)) runs_verify(L, number, idx, &B.staging_wavefronts);
  B.d_rec     = to_device(L.rec);
  B.d_srow    = to_device(L.srow);
  B.d_sprow   = to_device(L.sprow);
  B.d_ov_idx  = to_device(L.ov_idx);
  B.d_ov_pos  = to_device(L.ov_pos);
  B.d_zpos    = to_device(L.zpos);
  B.usable    = true;
}

struct Operator
{
  int degree = 0, number = 0, device = 0, geometry_type = 0;
  int apply_constraints = 1, kernel = MFHN_KERNEL_AUTO;
  long long n_cells = 0, n_owned = 0, n_ghost = 0, n_cells_hn = 0;
  uint32_t *d_idx = nullptr;    // reference layout [cell][lexicographic]
  uint8_t *d_masks = nullptr;
  int32_t *d_hn_cells = nullptr; // cells with a non-zero mask, ascending (DG (C) stage)
  void *d_geom = nullptr;       // Number h[cell], Number G[cell][6] or Number G[cell][6][q]
  PlaneLayout plane;            // warp-interleaved layout of the plane kernels
  BulkLayout bulk;              // block descriptors of the bulk-copy kernel (degrees 3..5)
  RunsLayout runs;              // run descriptors of the run-wise bulk-copy kernel (degrees 1..5)
  ConstraintRows rows;          // weighted constraint rows per distinct mask (general-purpose algorithm, built on first use)
  BaselineArrays baseline;      // padded deal.II-CUDA-style arrays of the baseline kernel (built on first use)
  std::vector<long long> segments;
  long long launches = 0;
  int hn_strategy = 0;    // MFHN_HN_BRANCH / MFHN_HN_MASK
  int vector_padding = 0; // valid entries behind n_owned + n_ghost in every vector (promise of the caller)
  void *d_stage_src[2] = {nullptr, nullptr}, *d_stage_dst[2] = {nullptr, nullptr}; // device staging of the host-vector entry point (2 slots)

  ~Operator()
  {
    cudaFree(d_idx);
    cudaFree(d_masks);
    cudaFree(d_hn_cells);
    cudaFree(d_geom);
    for (int i = 0; i < 2; ++i)
      {
        cudaFree(d_stage_src[i]);
        cudaFree(d_stage_dst[i]);
      }
    cudaFree(baseline.l2g);
    cudaFree(baseline.invjac);
    cudaFree(baseline.jxw);
    plane.free();
    bulk.free();
    runs.free();
    rows.free();
  }
};

// ---------------------------------------------------------------------------
int generic_variant(const Operator &op, int kernel)
{
  if (op.geometry_type == MFHN_GEOM_AFFINE) return GV_QPOINT_METRIC;
  if (op.geometry_type == MFHN_GEOM_GENERAL) return GV_QPOINT_GENERAL;
  return kernel == MFHN_KERNEL_SEPARABLE ? GV_SEPARABLE : GV_QPOINT_CARTESIAN;
}

// diagonal of the operator (for point-Jacobi): Cartesian cells through the separable form, affine / general through the q-point form
void launch_diagonal(Operator &op, const CellLoopParams &p, cudaStream_t stream)
{
  run_generic(op.degree, op.number, op.geometry_type == MFHN_GEOM_CARTESIAN ? GV_SEPARABLE : generic_variant(op, MFHN_KERNEL_QPOINT), true, p, op.device, stream);
  ++op.launches;
}

// The run-wise layout is built on first use (its placement pass takes seconds on 10^6 cells): from the operator's
// device copy of the index array.
void ensure_runs(Operator &op)
{
  if (op.runs.usable) return;
  const int n = op.degree + 1;
  std::vector<uint32_t> idx((size_t)op.n_cells * n * n * n);
  if (!idx.empty()) CUDA_CHECK(cudaMemcpy(idx.data(), op.d_idx, idx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  runs_build(op.runs, n, op.number, op.n_cells, op.n_owned + op.n_ghost + op.vector_padding, idx.data());
}

// General-purpose constraint algorithm: one sparse interpolation matrix W per distinct mask, obtained by sending the
// (k+1)^3 unit vectors through the interpolation kernel (double precision), stored row-wise with local column indices.
void ensure_rows(Operator &op)
{
  if (op.rows.built) return;
  const int n = op.degree + 1, n3 = n * n * n;
  std::vector<uint8_t> masks((size_t)std::max<long long>(op.n_cells, 1), 0);
  if (op.n_cells > 0) CUDA_CHECK(cudaMemcpy(masks.data(), op.d_masks, (size_t)op.n_cells, cudaMemcpyDeviceToHost));
  int id_of_mask[256];
  for (int &v : id_of_mask) v = 0;
  std::vector<uint8_t> distinct;
  std::vector<int32_t> kind((size_t)std::max<long long>(op.n_cells, 1), 0);
  for (long long c = 0; c < op.n_cells; ++c)
    if (masks[c] != 0)
      {
        if (id_of_mask[masks[c]] == 0)
          {
            distinct.push_back(masks[c]);
            id_of_mask[masks[c]] = (int)distinct.size();
          }
        kind[c] = id_of_mask[masks[c]];
      }
  std::vector<int32_t> ptr;
  std::vector<uint16_t> col;
  std::vector<double> val;
  std::vector<double> unit((size_t)n3 * n3), W((size_t)n3 * n3);
  double *d_unit = nullptr;
  uint8_t *d_m   = nullptr;
  CUDA_CHECK(cudaMalloc(&d_unit, unit.size() * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&d_m, (size_t)n3));
  for (const uint8_t m : distinct)
    {
      std::fill(unit.begin(), unit.end(), 0.0);
      for (int c = 0; c < n3; ++c) unit[(size_t)c * n3 + c] = 1.0; // "cell" c holds the unit vector e_c
      CUDA_CHECK(cudaMemcpy(d_unit, unit.data(), unit.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemset(d_m, m, (size_t)n3));
      run_hn_only(op.degree, MFHN_F64, d_unit, d_m, n3, 0, nullptr);
      CUDA_CHECK(cudaMemcpy(W.data(), d_unit, W.size() * sizeof(double), cudaMemcpyDeviceToHost)); // W[c][i] = (W e_c)_i
      for (int i = 0; i < n3; ++i)
        {
          ptr.push_back((int32_t)col.size());
          for (int c = 0; c < n3; ++c)
            if (std::abs(W[(size_t)c * n3 + i]) > 1e-14)
              {
                col.push_back((uint16_t)c);
                val.push_back(W[(size_t)c * n3 + i]);
              }
        }
      ptr.push_back((int32_t)col.size());
    }
  cudaFree(d_unit);
  cudaFree(d_m);
  op.rows.n_kinds   = (long long)distinct.size();
  op.rows.n_entries = (long long)col.size();
  op.rows.d_kind    = to_device(kind);
  op.rows.d_ptr     = to_device(ptr);
  op.rows.d_col     = to_device(col);
  if (op.number == MFHN_F64)
    op.rows.d_val = to_device(val);
  else
    {
      std::vector<float> vf(val.begin(), val.end());
      op.rows.d_val = to_device(vf);
    }
  op.rows.built = true;
}

void launch_kernel(Operator &op, int kernel, const CellLoopParams &p, cudaStream_t stream)
{
  if (kernel == MFHN_KERNEL_QPOINT_ROWS)
    {
      ensure_rows(op);
      CellLoopParams q = p;
      q.row_kind       = op.rows.d_kind;
      q.row_ptr        = op.rows.d_ptr;
      q.row_col        = op.rows.d_col;
      q.row_val        = op.rows.d_val;
      run_generic(op.degree, op.number, GV_QPOINT_ROWS, false, q, op.device, stream);
      ++op.launches;
      return;
    }
  if (kernel == MFHN_KERNEL_BASELINE)
    {
      run_baseline(op.degree, op.number, op.baseline, op.d_idx, op.d_geom, op.n_cells, p, op.device, stream);
      ++op.launches;
      return;
    }
  if (kernel == MFHN_KERNEL_BULK)
    {
      // the bulk-copy kernel takes whole warp batches; the (at most cpw - 1) cells at either end of an unaligned
      // range and the cells that do not show the block pattern (at most bulk_max_irregular) go to the plane kernel
      const long long cpw = 32 / (op.degree + 1);
      // (the last batch of the mesh may be incomplete: its missing cells are marked in the layout)
      const long long b0 = (p.cell_begin + cpw - 1) / cpw * cpw;
      const long long b1 = p.cell_end == op.n_cells ? (op.n_cells + cpw - 1) / cpw * cpw : p.cell_end / cpw * cpw;
      auto plane_part = [&](const long long cb, const long long ce) {
        if (ce <= cb) return;
        CellLoopParams q = p;
        q.cell_begin     = cb;
        q.cell_end       = ce;
        run_plane(op.degree, op.number, op.plane, q, op.device, stream, nullptr);
        ++op.launches;
      };
      if (b1 <= b0)
        {
          plane_part(p.cell_begin, p.cell_end);
          return;
        }
      plane_part(p.cell_begin, b0);
      {
        CellLoopParams q = p;
        q.cell_begin     = b0;
        q.cell_end       = b1;
        run_bulk(op.degree, op.number, op.bulk, q, op.device, stream);
      }
      plane_part(b1, p.cell_end);
      for (const long long c : op.bulk.irregular)
        if (c >= b0 && c < std::min(b1, p.cell_end)) plane_part(c, c + 1);
    }
  else if (kernel == MFHN_KERNEL_RUNS)
    {
      ensure_runs(op);
      // whole warp batches; the (at most cpw - 1) cells in front of an unaligned range go to the plane kernel
      const long long cpw = 32 / (op.degree + 1);
      const long long b0  = std::min((p.cell_begin + cpw - 1) / cpw * cpw, p.cell_end);
      const long long b1  = p.cell_end == op.n_cells ? p.cell_end : std::max(b0, p.cell_end / cpw * cpw);
      auto part = [&](const long long cb, const long long ce, const bool runs) {
        if (ce <= cb) return;
        CellLoopParams q = p;
        q.cell_begin     = cb;
        q.cell_end       = ce;
        if (runs)
          run_runs(op.degree, op.number, op.runs, q, op.device, stream);
        else
          run_plane(op.degree, op.number, op.plane, q, op.device, stream, nullptr);
        ++op.launches;
      };
      part(p.cell_begin, b0, false);
      part(b0, b1, true);
      part(b1, p.cell_end, false);
      return;
    }
  else if (kernel == MFHN_KERNEL_PLANE)
    run_plane(op.degree, op.number, op.plane, p, op.device, stream, nullptr);
  else
    run_generic(op.degree, op.number, generic_variant(op, kernel), false, p, op.device, stream);
  ++op.launches;
}

int resolve_kernel(const Operator &op)
{
  int kernel = op.kernel;
  if (kernel == MFHN_KERNEL_AUTO)
    {
      kernel = op.geometry_type == MFHN_GEOM_CARTESIAN ? MFHN_KERNEL_PLANE : MFHN_KERNEL_QPOINT;
      // measured on B200 (profiles/r2_kernel_choice.jsonl, annulus L=9 / L=8, GDoF/s double | float):
      //   k=3: plane 82.6 | 99.8, runs 66.7 | 103.0;  k=4: plane 83.6 | 125.1, bulk 109.0 | 128.8, runs 122.2 | 132.8;
      //   k=5: plane 93.6 | 139.2, bulk 120.1 | 141.1, runs 113.2 | 124.6;  k=1,2: plane ahead of runs by 20-50 %
      if (kernel == MFHN_KERNEL_PLANE && op.degree == 4) kernel = MFHN_KERNEL_RUNS;
      if (kernel == MFHN_KERNEL_PLANE && op.degree == 3 && op.number == MFHN_F32) kernel = MFHN_KERNEL_RUNS;
      //   k=5 on the L=9 mesh (profiles/r2_kernel_choice_k5_L9.jsonl): runs 138.2 (4 CTAs), bulk 135.6
      if (kernel == MFHN_KERNEL_PLANE && op.degree == 5 && op.number == MFHN_F64) kernel = MFHN_KERNEL_RUNS;
      if (kernel == MFHN_KERNEL_PLANE && op.degree == 5 && op.bulk.usable) kernel = MFHN_KERNEL_BULK;
    }
  if (op.geometry_type != MFHN_GEOM_CARTESIAN && kernel != MFHN_KERNEL_QPOINT && kernel != MFHN_KERNEL_QPOINT_ROWS)
    throw InvalidArgument("affine / general geometry requires MFHN_KERNEL_QPOINT");
  if (kernel == MFHN_KERNEL_BULK && !bulk_supported(op.degree + 1)) throw NotImplemented("MFHN_KERNEL_BULK is available for degrees 3..5");
  if (kernel == MFHN_KERNEL_BULK && op.geometry_type == MFHN_GEOM_CARTESIAN && !op.bulk.usable)
    throw InvalidArgument("MFHN_KERNEL_BULK: the DoF numbering does not show contiguous cell-interior / face blocks");
  if (kernel == MFHN_KERNEL_RUNS && !runs_supported(op.degree + 1)) throw NotImplemented("MFHN_KERNEL_RUNS is available for degrees 1..5");
  if (kernel == MFHN_KERNEL_PATCH)
    throw NotImplemented("MFHN_KERNEL_PATCH (sorted-unique patch gather, round 1) was measured slower than the plane kernel and has been removed");
  if ((kernel == MFHN_KERNEL_PLANE || kernel == MFHN_KERNEL_BULK || kernel == MFHN_KERNEL_RUNS || kernel == MFHN_KERNEL_SEPARABLE ||
       kernel == MFHN_KERNEL_QPOINT_ROWS) && op.geometry_type != MFHN_GEOM_CARTESIAN)
    throw InvalidArgument("this kernel requires Cartesian geometry");
  if (kernel == MFHN_KERNEL_BASELINE && op.geometry_type != MFHN_GEOM_CARTESIAN)
    throw InvalidArgument("the baseline kernel is set up for Cartesian cells");
  return kernel;
}

void op_vmult_range(Operator &op, void *dst, const void *src, cudaStream_t stream, long long cb, long long ce)
{
  if (ce < 0) ce = op.n_cells;
  if (cb < 0 || ce > op.n_cells || cb > ce) throw InvalidArgument("cell range out of bounds");
  CellLoopParams p;
  p.idx               = op.d_idx;
  p.masks             = op.d_masks;
  p.geom              = op.d_geom;
  p.src               = src;
  p.dst               = dst;
  p.cell_begin        = cb;
  p.cell_end          = ce;
  p.apply_constraints = op.apply_constraints;
  p.hn_mask_strategy  = op.hn_strategy == MFHN_HN_MASK;
  int kernel          = resolve_kernel(op);
  if ((kernel == MFHN_KERNEL_BULK || kernel == MFHN_KERNEL_RUNS) && (((uintptr_t)src | (uintptr_t)dst) & 15u))
    {
      if (op.kernel != MFHN_KERNEL_AUTO)
        throw InvalidArgument("MFHN_KERNEL_BULK / MFHN_KERNEL_RUNS need 16-byte aligned vectors (use MFHN_KERNEL_PLANE for unaligned views)");
      kernel = MFHN_KERNEL_PLANE; // AUTO: unaligned views take the plane kernel
    }
  launch_kernel(op, kernel, p, stream);
}

Operator *op_create(const mfhn_op_desc &d)
{
  if (d.degree < 1 || d.degree > 8) throw InvalidArgument("degree must be in 1..8");
  if (d.number != MFHN_F64 && d.number != MFHN_F32) throw InvalidArgument("number must be MFHN_F64 or MFHN_F32");
  if (d.n_cells < 0 || d.n_owned < 0 || d.n_ghost < 0) throw InvalidArgument("negative size");
  if (d.n_cells > 0 && (!d.dof_indices || !d.masks || !d.geometry)) throw InvalidArgument("null array");
  if (d.geometry_type != MFHN_GEOM_CARTESIAN && d.geometry_type != MFHN_GEOM_AFFINE && d.geometry_type != MFHN_GEOM_GENERAL)
    throw InvalidArgument("unknown geometry type");
  if (d.kernel < MFHN_KERNEL_AUTO || d.kernel > MFHN_KERNEL_QPOINT_ROWS) throw InvalidArgument("unknown kernel");
  int device = d.device;
  if (device < 0)
    CUDA_CHECK(cudaGetDevice(&device));
  else
    CUDA_CHECK(cudaSetDevice(device));
  std::unique_ptr<Operator> op(new Operator);
  op->degree            = d.degree;
  op->number            = d.number;
  op->device            = device;
  op->geometry_type     = d.geometry_type;
  op->apply_constraints = d.apply_constraints;
  op->kernel            = d.kernel;
  op->n_cells           = d.n_cells;
  op->n_owned           = d.n_owned;
  op->n_ghost           = d.n_ghost;
  const int n           = d.degree + 1;
  const long long n3    = (long long)n * n * n;
  const long long nvec  = d.n_owned + d.n_ghost;
  for (long long i = 0; i < d.n_cells * n3; ++i)
    if (d.dof_indices[i] >= (unsigned long long)nvec) throw InvalidArgument("dof index out of range");
  if (d.n_cells > 0x7fffffffll) throw InvalidArgument("more than 2^31 cells on one rank");
  std::vector<int32_t> hn_cells;
  for (long long c = 0; c < d.n_cells; ++c)
    {
      if (!check_kind(decompress_kind(d.masks[c])) || compress_kind(decompress_kind(d.masks[c])) != d.masks[c])
        throw InvalidArgument("invalid compressed constraint mask");
      if (d.masks[c] != 0) hn_cells.push_back((int32_t)c);
    }
  op->n_cells_hn = (long long)hn_cells.size();
  op->d_hn_cells = to_device(hn_cells);
  CUDA_CHECK(cudaMalloc(&op->d_idx, std::max<size_t>(1, d.n_cells * n3) * sizeof(uint32_t)));
  CUDA_CHECK(cudaMemcpy(op->d_idx, d.dof_indices, d.n_cells * n3 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&op->d_masks, std::max<size_t>(1, d.n_cells)));
  CUDA_CHECK(cudaMemcpy(op->d_masks, d.masks, d.n_cells, cudaMemcpyHostToDevice));
  // geometry -> per-cell factors in the operator's Number type
  std::vector<double> g;
  if (d.geometry_type == MFHN_GEOM_CARTESIAN)
    g.assign(d.geometry, d.geometry + d.n_cells);
  else if (d.geometry_type == MFHN_GEOM_GENERAL)
    g.assign(d.geometry, d.geometry + d.n_cells * 6 * n3); // [cell][6][q], JxW J^-1 J^-T per quadrature point
  else
    {
      g.resize(d.n_cells * 6);
      for (long long c = 0; c < d.n_cells; ++c)
        {
          const double *J = d.geometry + 9 * c;
          const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
          if (!(det > 0)) throw InvalidArgument("non-positive Jacobian determinant");
          double inv[9] = {(J[4] * J[8] - J[5] * J[7]) / det, (J[2] * J[7] - J[1] * J[8]) / det, (J[1] * J[5] - J[2] * J[4]) / det,
                           (J[5] * J[6] - J[3] * J[8]) / det, (J[0] * J[8] - J[2] * J[6]) / det, (J[2] * J[3] - J[0] * J[5]) / det,
                           (J[3] * J[7] - J[4] * J[6]) / det, (J[1] * J[6] - J[0] * J[7]) / det, (J[0] * J[4] - J[1] * J[3]) / det};
          // G = det * inv * inv^T  (symmetric): rows of inv are gradients of xi_r
          int t = 0;
          for (int r = 0; r < 3; ++r)
            for (int s = r; s < 3; ++s)
              g[6 * c + t++] = det * (inv[3 * r] * inv[3 * s] + inv[3 * r + 1] * inv[3 * s + 1] + inv[3 * r + 2] * inv[3 * s + 2]);
        }
    }
  if (d.number == MFHN_F64)
    op->d_geom = to_device(g);
  else
    {
      std::vector<float> gf(g.begin(), g.end());
      op->d_geom = to_device(gf);
    }
  if (d.geometry_type == MFHN_GEOM_CARTESIAN)
    {
      op->plane.build(n, d.n_cells, d.dof_indices);
      op->segments.assign(1, 0);
      if (d.segments)
        {
          if (d.n_segments < 1 || d.segments[0] != 0) throw InvalidArgument("segments must start at cell 0");
          op->segments.assign(d.segments, d.segments + d.n_segments);
          for (int i = 1; i < d.n_segments; ++i)
            if (d.segments[i] < d.segments[i - 1] || d.segments[i] > d.n_cells) throw InvalidArgument("segments must be ascending");
        }
      if (d.vector_padding < 0) throw InvalidArgument("negative vector padding");
      op->vector_padding = d.vector_padding;
      if (bulk_supported(n)) bulk_build(op->bulk, n, d.number, d.n_cells, nvec + d.vector_padding, d.dof_indices);
      // (the run-wise layout is built when MFHN_KERNEL_RUNS is first used: ensure_runs)
    }
  resolve_kernel(*op);
  return op.release();
}

// ---------------------------------------------------------------------------
// partitioned operator: ghost import / compress over NCCL, overlapped with the cell partitions
struct Dist
{
  Operator *op = nullptr;
  void *comm   = nullptr;
  int rank = 0, world = 1;
  std::vector<int> import_peers, ghost_peers;
  std::vector<long long> import_off, ghost_begin, ghost_end;
  long long n_import = 0;
  long long seg[4]   = {0, 0, 0, 0};
  int32_t *d_import_idx = nullptr;
  void *d_send = nullptr, *d_recv = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev[4]        = {nullptr, nullptr, nullptr, nullptr};
  long long launches       = 0;
  // peer mode: registered vector pair + per-ghost addresses inside the owners' vectors
  void *peer_src_local = nullptr, *peer_dst_local = nullptr;
  void **d_ghost_src = nullptr, **d_ghost_dst = nullptr;
  int *d_barrier = nullptr;
  // barriers of the peer path through flags in each other's memory instead of NCCL all-reduces
  unsigned *flags_local = nullptr, **d_peer_flags = nullptr;

  ~Dist()
  {
    cudaFree(d_ghost_src);
    cudaFree(d_ghost_dst);
    cudaFree(d_barrier);
    cudaFree(d_peer_flags);
    if (comm) NcclApi::get().CommDestroy(comm);
    cudaFree(d_import_idx);
    cudaFree(d_send);
    cudaFree(d_recv);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    for (auto e : ev)
      if (e) cudaEventDestroy(e);
  }
};

#define NCCL_CHECK(x)                                                                            \
  do                                                                                             \
    {                                                                                            \
      int r_ = (x);                                                                              \
      if (r_ != 0) throw CudaError(std::string(#x) + ": " + NcclApi::get().GetErrorString(r_)); \
    }                                                                                            \
  while (0)

// ncclGroupStart ... ncclGroupEnd around f: an error inside still closes the group before it propagates
template <typename F>
void nccl_group(NcclApi &nccl, F &&f)
{
  NCCL_CHECK(nccl.GroupStart());
  try
    {
      f();
    }
  catch (...)
    {
      nccl.GroupEnd();
      throw;
    }
  NCCL_CHECK(nccl.GroupEnd());
}

// One partitioned vmult.  The whole exchange chain runs on the high-priority communication stream -- pack, import
// (update_ghost_values), the boundary cells, compress (add), unpack -- while ONE launch covers all interior cells on the
// caller's stream: the scatter is atomic, so the order in which interior cells, boundary cells and imported
// contributions reach dst does not matter, and the chain (about 60 us at 8 ranks) hides behind the interior cells
// instead of cutting their launch in two.  That holds for two ranks (measured 232.9 vs 223.0 GDoF/s).  With more peers
// the second NCCL kernel (compress) is launched while the interior kernel fills every SM; its large CTAs find no room until
// the interior cells are through (8 ranks: 544.6 vs 650.8 GDoF/s), so from three ranks on the interior cells are cut in
// two launches -- interior A | boundary | interior B on the caller's stream -- and each NCCL kernel is launched in front
// of one of them.  MFHN_DIST_SPLIT=0/1 forces either schedule.
void dist_vmult(Dist &d, void *dst, const void *src, cudaStream_t main, int zero_dst)
{
  Operator &op  = *d.op;
  NcclApi &nccl = NcclApi::get();
  const size_t s = op.number == MFHN_F64 ? 8 : 4;
  char *dstb = static_cast<char *>(dst);
  char *srcb = static_cast<char *>(const_cast<void *>(src));
  const bool split = env_int("MFHN_DIST_SPLIT", d.world > 2 ? 1 : 0) != 0;
  cudaStream_t cs = d.comm_stream;
  if (zero_dst) CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)(op.n_owned + op.n_ghost) * s, main));
  CUDA_CHECK(cudaEventRecord(d.ev[0], main)); // src is final, dst zeroed
  CUDA_CHECK(cudaStreamWaitEvent(cs, d.ev[0], 0));
  // pack the entries the peers ghost
  if (d.n_import > 0)
    {
      run_pack(op.number, d.d_send, src, d.d_import_idx, d.n_import, cs);
      ++d.launches;
    }
  // owners -> ghosts (update_ghost_values)
  nccl_group(nccl, [&] {
    for (size_t i = 0; i < d.ghost_peers.size(); ++i)
      NCCL_CHECK(nccl.Recv(srcb + (size_t)(op.n_owned + d.ghost_begin[i]) * s, (size_t)(d.ghost_end[i] - d.ghost_begin[i]) * s, 0,
                           d.ghost_peers[i], d.comm, cs));
    for (size_t i = 0; i < d.import_peers.size(); ++i)
      NCCL_CHECK(nccl.Send(static_cast<char *>(d.d_send) + (size_t)d.import_off[i] * s, (size_t)(d.import_off[i + 1] - d.import_off[i]) * s, 0,
                           d.import_peers[i], d.comm, cs));
  });
  cudaStream_t bs = cs; // stream of the boundary cells
  if (split)
    {
      CUDA_CHECK(cudaEventRecord(d.ev[1], cs));
      if (d.seg[1] > d.seg[0]) op_vmult_range(op, dst, src, main, d.seg[0], d.seg[1]); // interior A overlaps the import
      CUDA_CHECK(cudaStreamWaitEvent(main, d.ev[1], 0));
      bs = main;
    }
  if (d.seg[3] > d.seg[2]) op_vmult_range(op, dst, src, bs, d.seg[2], d.seg[3]); // boundary cells need the ghosts
  // zero_out_ghost_values: the imported entries are scratch, a later use of src as dst must not send them to the owners
  if (op.n_ghost > 0) CUDA_CHECK(cudaMemsetAsync(srcb + (size_t)op.n_owned * s, 0, (size_t)op.n_ghost * s, bs));
  if (split)
    {
      CUDA_CHECK(cudaEventRecord(d.ev[2], main));
      CUDA_CHECK(cudaStreamWaitEvent(cs, d.ev[2], 0));
    }
  // ghosts -> owners (compress, add)
  nccl_group(nccl, [&] {
    for (size_t i = 0; i < d.import_peers.size(); ++i)
      NCCL_CHECK(nccl.Recv(static_cast<char *>(d.d_recv) + (size_t)d.import_off[i] * s, (size_t)(d.import_off[i + 1] - d.import_off[i]) * s, 0,
                           d.import_peers[i], d.comm, cs));
    for (size_t i = 0; i < d.ghost_peers.size(); ++i)
      NCCL_CHECK(nccl.Send(dstb + (size_t)(op.n_owned + d.ghost_begin[i]) * s, (size_t)(d.ghost_end[i] - d.ghost_begin[i]) * s, 0,
                           d.ghost_peers[i], d.comm, cs));
  });
  if (d.n_import > 0)
    {
      run_unpack_add(op.number, dst, d.d_recv, d.d_import_idx, d.n_import, true, cs); // atomic: interior cells add to dst at the same time
      ++d.launches;
    }
  CUDA_CHECK(cudaEventRecord(d.ev[3], cs));
  // interior cells: one launch beside the chain (split schedule: the second half)
  const long long ib = split ? d.seg[1] : d.seg[0];
  if (d.seg[2] > ib) op_vmult_range(op, dst, src, main, ib, d.seg[2]);
  CUDA_CHECK(cudaStreamWaitEvent(main, d.ev[3], 0));
  if (op.n_ghost > 0) CUDA_CHECK(cudaMemsetAsync(dstb + (size_t)op.n_owned * s, 0, (size_t)op.n_ghost * s, main));
}

// Peer-memory variant (fused compute + exchange): no pack / unpack and no data-path collective.  After a barrier the
// boundary cells read the owners' src entries and add into the owners' dst entries directly over NVLink; a second
// barrier makes the remote contributions visible before anyone consumes dst.  The barriers are flag exchanges in each
// other's memory (mfhn_dist_enable_peer_flags; pure CUDA, so the whole vmult can be captured in a CUDA graph) or,
// without flags, 4-byte NCCL all-reduces.
void dist_vmult_peer_n(Dist &d, cudaStream_t main)
{
  Operator &op = *d.op;
  CellLoopParams p;
  p.idx               = op.d_idx;
  p.masks             = op.d_masks;
  p.geom              = op.d_geom;
  p.src               = d.peer_src_local;
  p.dst               = d.peer_dst_local;
  p.apply_constraints = op.apply_constraints;
  PeerTables pt;
  pt.n_owned   = op.n_owned;
  pt.ghost_src = d.d_ghost_src;
  pt.ghost_dst = d.d_ghost_dst;
  auto barrier = [&]() {
    if (d.flags_local)
      run_peer_barrier(d.flags_local, d.d_peer_flags, d.rank, d.world, d.comm_stream);
    else
      NCCL_CHECK(NcclApi::get().AllReduce(d.d_barrier, d.d_barrier, 1, 2 /*ncclInt32*/, 0 /*ncclSum*/, d.comm, d.comm_stream));
    ++d.launches;
  };
  // high-priority stream: barrier (peers' src final, peers' dst zeroed) -> boundary cells with remote
  // entries over NVLink -> barrier (remote contributions landed); the interior cells (local entries
  // only) run concurrently on the compute stream and hide the whole chain
  CUDA_CHECK(cudaEventRecord(d.ev[0], main));
  CUDA_CHECK(cudaStreamWaitEvent(d.comm_stream, d.ev[0], 0));
  barrier();
  if (d.seg[3] > d.seg[2])
    {
      // boundary cells: plane kernel with peer tables (register-tiled for k <= 5, plane in shared memory above)
      p.cell_begin = d.seg[2];
      p.cell_end   = d.seg[3];
      run_plane(op.degree, op.number, op.plane, p, op.device, d.comm_stream, &pt);
      ++op.launches;
      ++d.launches;
    }
  barrier();
  CUDA_CHECK(cudaEventRecord(d.ev[3], d.comm_stream));
  if (d.seg[2] > d.seg[0])
    {
      const long long before = op.launches;
      op_vmult_range(op, d.peer_dst_local, d.peer_src_local, main, d.seg[0], d.seg[2]); // the operator's own kernel (bulk copies at k = 4, 5)
      d.launches += op.launches - before;
    }
  CUDA_CHECK(cudaStreamWaitEvent(main, d.ev[3], 0));
}
} // namespace mfhn

using namespace mfhn;

extern "C" {
int mfhn_op_create(const mfhn_op_desc *desc, mfhn_op *out)
{
  return guard([&] {
    if (!desc || !out) throw InvalidArgument("null argument");
    *out = reinterpret_cast<mfhn_op>(op_create(*desc));
  });
}
int mfhn_op_create_mf(mfhn_mf m, int number, int kernel, int apply_constraints, int device, mfhn_op *out)
{
  return mfhn_op_create_mf_padded(m, number, kernel, apply_constraints, device, 0, out);
}
int mfhn_op_create_mf_padded(mfhn_mf m, int number, int kernel, int apply_constraints, int device, int vector_padding, mfhn_op *out)
{
  return guard([&] {
    if (!m || !out) throw InvalidArgument("null argument");
    const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(m);
    std::vector<int64_t> seg;
    for (int64_t s : {(int64_t)0, mf.n_interior_a, mf.n_interior})
      if (s < mf.n_cells && (seg.empty() || s > seg.back())) seg.push_back(s);
    if (seg.empty()) seg.push_back(0);
    mfhn_op_desc d{};
    d.degree            = mf.degree;
    d.number            = number;
    d.n_cells           = mf.n_cells;
    d.n_owned           = mf.n_owned;
    d.n_ghost           = mf.n_ghost;
    d.dof_indices       = mf.dof_indices.data();
    d.masks             = mf.masks.data();
    d.geometry_type     = MFHN_GEOM_CARTESIAN;
    d.geometry          = mf.h.data();
    d.apply_constraints = apply_constraints;
    d.kernel            = kernel;
    d.device            = device;
    d.segments          = seg.data();
    d.n_segments        = (int)seg.size();
    d.vector_padding    = vector_padding;
    *out                = reinterpret_cast<mfhn_op>(op_create(d));
  });
}
void mfhn_op_destroy(mfhn_op op) { delete reinterpret_cast<Operator *>(op); }

int mfhn_op_vmult(mfhn_op h, void *dst, const void *src, void *stream, int zero_dst)
{
  return guard([&] {
    if (!h || !dst || !src) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (zero_dst)
      CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)(op.n_owned + op.n_ghost) * (op.number == MFHN_F64 ? 8 : 4), st));
    op_vmult_range(op, dst, src, st, 0, op.n_cells);
  });
}
int mfhn_op_vmult_range(mfhn_op h, void *dst, const void *src, void *stream, int64_t cb, int64_t ce)
{
  return guard([&] {
    if (!h || !dst || !src) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    op_vmult_range(op, dst, src, static_cast<cudaStream_t>(stream), cb, ce);
  });
}
int mfhn_op_vmult_host_slot(mfhn_op h, void *dst_host, const void *src_host, void *stream, int zero_dst, int slot)
{
  return guard([&] {
    if (!h || !dst_host || !src_host) throw InvalidArgument("null argument");
    if (slot < 0 || slot > 1) throw InvalidArgument("slot must be 0 or 1");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t st    = static_cast<cudaStream_t>(stream);
    const size_t es = op.number == MFHN_F64 ? 8 : 4, bytes = (size_t)(op.n_owned + op.n_ghost) * es;
    if (!op.d_stage_src[slot])
      {
        const size_t padded = std::max<size_t>(bytes, 8) + (size_t)op.vector_padding * es;
        CUDA_CHECK(cudaMalloc(&op.d_stage_src[slot], padded));
        CUDA_CHECK(cudaMalloc(&op.d_stage_dst[slot], padded));
        CUDA_CHECK(cudaMemset(op.d_stage_src[slot], 0, padded));
        CUDA_CHECK(cudaMemset(op.d_stage_dst[slot], 0, padded));
      }
    CUDA_CHECK(cudaMemcpyAsync(op.d_stage_src[slot], src_host, bytes, cudaMemcpyHostToDevice, st));
    if (zero_dst)
      CUDA_CHECK(cudaMemsetAsync(op.d_stage_dst[slot], 0, bytes, st));
    else
      CUDA_CHECK(cudaMemcpyAsync(op.d_stage_dst[slot], dst_host, bytes, cudaMemcpyHostToDevice, st));
    op_vmult_range(op, op.d_stage_dst[slot], op.d_stage_src[slot], st, 0, op.n_cells);
    CUDA_CHECK(cudaMemcpyAsync(dst_host, op.d_stage_dst[slot], bytes, cudaMemcpyDeviceToHost, st));
  });
}
int mfhn_op_vmult_host(mfhn_op h, void *dst_host, const void *src_host, void *stream, int zero_dst)
{
  return mfhn_op_vmult_host_slot(h, dst_host, src_host, stream, zero_dst, 0);
}
int mfhn_op_diagonal(mfhn_op h, void *diag, void *stream)
{
  return guard([&] {
    if (!h || !diag) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    CellLoopParams p;
    p.idx               = op.d_idx;
    p.masks             = op.d_masks;
    p.geom              = op.d_geom;
    p.src               = diag; // unused
    p.dst               = diag;
    p.cell_begin        = 0;
    p.cell_end          = op.n_cells;
    p.apply_constraints = op.apply_constraints;
    launch_diagonal(op, p, static_cast<cudaStream_t>(stream));
  });
}
int mfhn_op_set_apply_constraints(mfhn_op h, int v)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    reinterpret_cast<Operator *>(h)->apply_constraints = v != 0;
  });
}
int mfhn_op_set_kernel(mfhn_op h, int kernel)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    Operator &op  = *reinterpret_cast<Operator *>(h);
    const int old = op.kernel;
    if (kernel < MFHN_KERNEL_AUTO || kernel > MFHN_KERNEL_QPOINT_ROWS) throw InvalidArgument("unknown kernel");
    op.kernel = kernel;
    try
      {
        resolve_kernel(op);
      }
    catch (...)
      {
        op.kernel = old;
        throw;
      }
  });
}
int mfhn_op_apply_hn(mfhn_op h, void *values, int transpose, void *stream)
{
  return guard([&] {
    if (!h || !values) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    run_hn_only(op.degree, op.number, values, op.d_masks, op.n_cells, transpose, static_cast<cudaStream_t>(stream));
    ++op.launches;
  });
}
int mfhn_op_set_hn_strategy(mfhn_op h, int strategy)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    if (strategy != MFHN_HN_BRANCH && strategy != MFHN_HN_MASK) throw InvalidArgument("unknown hanging-node strategy");
    reinterpret_cast<Operator *>(h)->hn_strategy = strategy;
  });
}
int mfhn_op_dg_copy(mfhn_op h, void *dst_cells, const void *src_cells, void *stream)
{
  return guard([&] {
    if (!h || !dst_cells || !src_cells) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    run_dg_copy(op.degree, op.number, dst_cells, src_cells, op.d_masks, op.n_cells, op.d_hn_cells, op.n_cells_hn, op.apply_constraints,
                static_cast<cudaStream_t>(stream));
    op.launches += 1 + (op.apply_constraints && op.n_cells_hn > 0);
  });
}
int mfhn_op_query(mfhn_op h, const char *what, double *value)
{
  return guard([&] {
    if (!h || !what || !value) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    const double n = op.degree + 1, n3 = n * n * n, s = op.number == MFHN_F64 ? 8 : 4;
    const double nvec = (double)(op.n_owned + op.n_ghost);
    const std::string w(what);
    if (w == "n_cells")
      *value = (double)op.n_cells;
    else if (w == "n_cells_hn")
      *value = (double)op.n_cells_hn;
    else if (w == "algorithmic_bytes") // 2 s n_dofs + n_cells (4 (k+1)^3 + 1 + G)
      *value = 2 * s * nvec + (double)op.n_cells * (4 * n3 + 1 + (op.geometry_type == MFHN_GEOM_CARTESIAN ? 3 * s : op.geometry_type == MFHN_GEOM_AFFINE ? 10 * s : 6 * s * n3));
    else if (w == "algorithmic_bytes_accumulate") // + s n_dofs: dst is read as well when vmult accumulates
      *value = 3 * s * nvec + (double)op.n_cells * (4 * n3 + 1 + (op.geometry_type == MFHN_GEOM_CARTESIAN ? 3 * s : op.geometry_type == MFHN_GEOM_AFFINE ? 10 * s : 6 * s * n3));
    else if (w == "algorithmic_flops") // even-odd sum factorisation count of SURVEY 8d, without HN terms
      *value = (double)op.n_cells * (12 * n * n * (n * n + 2 * n) + 3 * n3);
    else if (w == "kernel")
      *value = (double)resolve_kernel(op);
    else if (w.rfind("runs_", 0) == 0 && !runs_supported(op.degree + 1))
      *value = -1.0;
    else if (w == "runs_bulk_copies") // bulk copies / single entries / zeroed entries of the run-wise layout, whole mesh
      {
        ensure_runs(op);
        *value = (double)op.runs.n_blocks;
      }
    else if (w == "runs_single_entries" || w == "runs_zero_entries" || w == "runs_single_rounds")
      {
        ensure_runs(op);
        *value = w == "runs_single_entries" ? (double)op.runs.n_singles : w == "runs_zero_entries" ? (double)op.runs.n_zero : (double)op.runs.sr; // sr: rounds of the fixed-stride single-entry rows
      }
    else if (w == "runs_staging_wavefronts") // bank model of the staging reads (MFHN_RUNS_STATS=1 at creation), 2 per plane slot and batch = no conflict
      *value = (double)op.runs.staging_wavefronts;
    else if (w == "constraint_row_entries") // general-purpose constraint algorithm: non-zeros of the interpolation matrices of all distinct masks
      {
        ensure_rows(op);
        *value = (double)op.rows.n_entries;
      }
    else if (w == "bulk_irregular_cells") // cells the bulk-copy kernel leaves to the plane kernel (-1: layout not usable)
      *value = op.bulk.usable ? (double)op.bulk.irregular.size() : -1.0;
    else
      throw InvalidArgument("unknown query '" + w + "'");
  });
}
int64_t mfhn_op_launch_count(mfhn_op h) { return h ? reinterpret_cast<Operator *>(h)->launches : 0; }

int mfhn_bulk_layout_check(int degree, int number, int64_t n_cells, int64_t n_vec, const uint32_t *dof_indices, int64_t *n_irregular,
                           int64_t *n_mismatch)
{
  return guard([&] {
    if (n_cells > 0 && !dof_indices) throw InvalidArgument("null argument");
    if (number != MFHN_F64 && number != MFHN_F32) throw InvalidArgument("number must be MFHN_F64 or MFHN_F32");
    BulkHostLayout L;
    bulk_analyze(L, degree + 1, number, n_cells, n_vec, dof_indices);
    if (n_irregular) *n_irregular = (int64_t)L.irregular.size();
    if (n_mismatch) *n_mismatch = bulk_verify(L, number, dof_indices);
  });
}

int mfhn_runs_layout_check(int degree, int number, int64_t n_cells, int64_t n_vec, const uint32_t *dof_indices, int max_gap, int min_run, int place,
                           int64_t *n_bulk_copies, int64_t *n_single_entries, int64_t *n_mismatch, int64_t *staging_wavefronts)
{
  return guard([&] {
    if (n_cells > 0 && !dof_indices) throw InvalidArgument("null argument");
    if (number != MFHN_F64 && number != MFHN_F32) throw InvalidArgument("number must be MFHN_F64 or MFHN_F32");
    if (max_gap < 0 || min_run < 1) throw InvalidArgument("max_gap must be >= 0 and min_run >= 1");
    RunsHostLayout L;
    runs_analyze(L, degree + 1, number, n_cells, n_vec, dof_indices, max_gap, min_run, place != 0);
    if (n_bulk_copies) *n_bulk_copies = L.n_blocks;
    if (n_single_entries) *n_single_entries = L.n_singles;
    long long wf = 0;
    const long long bad = runs_verify(L, number, dof_indices, &wf);
    if (n_mismatch) *n_mismatch = bad;
    if (staging_wavefronts) *staging_wavefronts = wf;
  });
}

int mfhn_pack(int number, void *buffer, const void *vec, const int32_t *idx, int64_t n, void *stream)
{
  return guard([&] { run_pack(number, buffer, vec, idx, n, static_cast<cudaStream_t>(stream)); });
}
int mfhn_unpack_add(int number, void *vec, const void *buffer, const int32_t *idx, int64_t n, void *stream)
{
  return guard([&] { run_unpack_add(number, vec, buffer, idx, n, false, static_cast<cudaStream_t>(stream)); });
}
int mfhn_bench_dfma(int number, int iters, double *tflops)
{
  return guard([&] {
    if (!tflops) throw InvalidArgument("null argument");
    *tflops = run_fma_bench(number, iters);
  });
}

int mfhn_dist_unique_id(void *id128)
{
  return guard([&] {
    if (!id128) throw InvalidArgument("null argument");
    NcclApi &nccl = NcclApi::get();
    if (!nccl.ok) throw CudaError(nccl.error);
    NCCL_CHECK(nccl.GetUniqueId(id128));
  });
}
int mfhn_dist_create(mfhn_op h, const mfhn_dist_desc *dd, mfhn_dist *out)
{
  return guard([&] {
    if (!h || !dd || !out || !dd->unique_id) throw InvalidArgument("null argument");
    NcclApi &nccl = NcclApi::get();
    if (!nccl.ok) throw CudaError(nccl.error);
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    std::unique_ptr<Dist> d(new Dist);
    d->op    = &op;
    d->rank  = dd->rank;
    d->world = dd->world;
    for (int i = 0; i < 4; ++i) d->seg[i] = dd->segments[i];
    if (!(0 == d->seg[0] && d->seg[0] <= d->seg[1] && d->seg[1] <= d->seg[2] && d->seg[2] <= d->seg[3] && d->seg[3] == op.n_cells))
      throw InvalidArgument("segments must be 0 <= a <= b <= n_cells");
    d->import_peers.assign(dd->import_peers, dd->import_peers + dd->n_import_peers);
    d->import_off.assign(dd->import_offsets, dd->import_offsets + dd->n_import_peers + 1);
    d->ghost_peers.assign(dd->ghost_peers, dd->ghost_peers + dd->n_ghost_peers);
    d->ghost_begin.assign(dd->ghost_begin, dd->ghost_begin + dd->n_ghost_peers);
    d->ghost_end.assign(dd->ghost_end, dd->ghost_end + dd->n_ghost_peers);
    if (dd->world < 1 || dd->rank < 0 || dd->rank >= dd->world) throw InvalidArgument("rank / world out of range");
    if (d->import_off[0] != 0) throw InvalidArgument("import offsets must start at 0");
    for (size_t i = 0; i < d->import_peers.size(); ++i)
      {
        if (d->import_peers[i] < 0 || d->import_peers[i] >= dd->world || d->import_peers[i] == dd->rank) throw InvalidArgument("import peer out of range");
        if (d->import_off[i + 1] < d->import_off[i]) throw InvalidArgument("import offsets must be ascending");
      }
    for (const int pr : d->ghost_peers)
      if (pr < 0 || pr >= dd->world || pr == dd->rank) throw InvalidArgument("ghost peer out of range");
    d->n_import = d->import_off.back();
    for (long long i = 0; i < d->n_import; ++i)
      if (dd->import_indices[i] < 0 || dd->import_indices[i] >= op.n_owned) throw InvalidArgument("import index out of range");
    for (size_t i = 0; i < d->ghost_peers.size(); ++i)
      if (d->ghost_begin[i] < 0 || d->ghost_end[i] > op.n_ghost || d->ghost_begin[i] > d->ghost_end[i]) throw InvalidArgument("ghost range out of bounds");
    const size_t s = op.number == MFHN_F64 ? 8 : 4;
    CUDA_CHECK(cudaMalloc(&d->d_import_idx, std::max<size_t>(1, d->n_import) * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(d->d_import_idx, dd->import_indices, d->n_import * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&d->d_send, std::max<size_t>(8, d->n_import * s)));
    CUDA_CHECK(cudaMalloc(&d->d_recv, std::max<size_t>(8, d->n_import * s)));
    {
      // highest priority: the NCCL kernels must not queue behind the blocks of an interior cell partition
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, hi));
    }
    for (auto &e : d->ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    NcclApi::Id id;
    std::memcpy(&id, dd->unique_id, sizeof(id));
    NCCL_CHECK(nccl.CommInitRank(&d->comm, dd->world, id, dd->rank));
    *out = reinterpret_cast<mfhn_dist>(d.release());
  });
}
int mfhn_dist_create_mf(mfhn_op h, mfhn_mf m, const void *unique_id, mfhn_dist *out)
{
  if (!m)
    {
      set_last_error("null argument");
      return MFHN_ERR_INVALID;
    }
  const MatrixFreeData &mf = *reinterpret_cast<MatrixFreeData *>(m);
  mfhn_dist_desc dd{};
  dd.rank           = mf.rank;
  dd.world          = mf.n_ranks;
  dd.unique_id      = unique_id;
  dd.n_import_peers = (int)mf.import_peers.size();
  dd.import_peers   = mf.import_peers.data();
  dd.import_offsets = mf.import_offsets.data();
  dd.import_indices = mf.import_indices.data();
  dd.n_ghost_peers  = (int)mf.ghost_peers.size();
  dd.ghost_peers    = mf.ghost_peers.data();
  dd.ghost_begin    = mf.ghost_begin.data();
  dd.ghost_end      = mf.ghost_end.data();
  dd.segments[0]    = 0;
  dd.segments[1]    = mf.n_interior_a;
  dd.segments[2]    = mf.n_interior;
  dd.segments[3]    = mf.n_cells;
  return mfhn_dist_create(h, &dd, out);
}
void mfhn_dist_destroy(mfhn_dist d) { delete reinterpret_cast<Dist *>(d); }
int mfhn_dist_vmult(mfhn_dist h, void *dst, const void *src, void *stream, int zero_dst)
{
  return guard([&] {
    if (!h || !dst || !src) throw InvalidArgument("null argument");
    Dist &d = *reinterpret_cast<Dist *>(h);
    CUDA_CHECK(cudaSetDevice(d.op->device));
    dist_vmult(d, dst, src, static_cast<cudaStream_t>(stream), zero_dst);
  });
}
int64_t mfhn_dist_launch_count(mfhn_dist h) { return h ? reinterpret_cast<Dist *>(h)->launches : 0; }

// ---- CG with point-Jacobi (extension, BASELINE.json config 5) ---------------------------------------------------
int mfhn_op_inverse_diagonal(mfhn_op h, mfhn_dist dh, void *inv_diag, void *stream)
{
  return guard([&] {
    if (!h || !inv_diag) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t s  = op.number == MFHN_F64 ? 8 : 4;
    CUDA_CHECK(cudaMemsetAsync(inv_diag, 0, (size_t)(op.n_owned + op.n_ghost) * s, st));
    if (int rc = mfhn_op_diagonal(h, inv_diag, stream)) throw CudaError(mfhn_last_error() + std::string(" (status ") + std::to_string(rc) + ")");
    if (dh)
      {
        // compress(add) of the ghost contributions: the exchange of a vmult with no cells
        Dist &d       = *reinterpret_cast<Dist *>(dh);
        NcclApi &nccl = NcclApi::get();
        char *db      = static_cast<char *>(inv_diag);
        nccl_group(nccl, [&] {
          for (size_t i = 0; i < d.import_peers.size(); ++i)
            NCCL_CHECK(nccl.Recv(static_cast<char *>(d.d_recv) + (size_t)d.import_off[i] * s, (size_t)(d.import_off[i + 1] - d.import_off[i]) * s, 0,
                                 d.import_peers[i], d.comm, st));
          for (size_t i = 0; i < d.ghost_peers.size(); ++i)
            NCCL_CHECK(nccl.Send(db + (size_t)(op.n_owned + d.ghost_begin[i]) * s, (size_t)(d.ghost_end[i] - d.ghost_begin[i]) * s, 0, d.ghost_peers[i], d.comm, st));
        });
        run_unpack_add(op.number, inv_diag, d.d_recv, d.d_import_idx, d.n_import, true, st);
        if (op.n_ghost > 0) CUDA_CHECK(cudaMemsetAsync(db + (size_t)op.n_owned * s, 0, (size_t)op.n_ghost * s, st));
      }
    run_invert_diagonal(op.number, inv_diag, op.n_owned, st);
  });
}

int mfhn_cg_solve(mfhn_op h, mfhn_dist dh, void *x, const void *b, const void *inv_diag, const mfhn_cg_options *options, mfhn_cg_result *result,
                  double *residual_history, void *stream)
{
  return guard([&] {
    if (!h || !x || !b || !options) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    Dist *d      = reinterpret_cast<Dist *>(dh);
    if (d && d->op != &op) throw InvalidArgument("the partitioned operator belongs to another operator");
    if (options->max_iter < 1 || options->check_every < 1) throw InvalidArgument("max_iter and check_every must be positive");
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t st    = static_cast<cudaStream_t>(stream);
    const size_t es    = op.number == MFHN_F64 ? 8 : 4;
    const long long n  = op.n_owned, nvec = op.n_owned + op.n_ghost;
    const size_t bytes = (size_t)(nvec + op.vector_padding) * es;
    const int max_iter = options->max_iter;
    // work vectors r, u, w, p, s; scalars and residual history on the device
    char *work = nullptr;
    CUDA_CHECK(cudaMalloc(&work, 5 * bytes + cg_scalars_bytes() + sizeof(double) * (size_t)(max_iter + 1) + 64));
    struct Free
    {
      char *p;
      ~Free() { cudaFree(p); }
    } guard_work{work};
    CUDA_CHECK(cudaMemsetAsync(work, 0, 5 * bytes + cg_scalars_bytes() + sizeof(double) * (size_t)(max_iter + 1) + 64, st));
    void *r = work, *u = work + bytes, *w = work + 2 * bytes, *p = work + 3 * bytes, *s = work + 4 * bytes;
    void *sc      = work + 5 * bytes;
    double *hist  = reinterpret_cast<double *>(work + 5 * bytes + ((cg_scalars_bytes() + 15) / 16) * 16);
    const bool timed = options->timings != 0;
    std::vector<cudaEvent_t> ev;
    auto mark = [&]() {
      if (!timed) return;
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      CUDA_CHECK(cudaEventRecord(e, st));
      ev.push_back(e);
    };
    auto apply = [&](void *dst, const void *src) {
      if (d)
        dist_vmult(*d, dst, src, st, 1);
      else
        {
          CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)nvec * es, st));
          op_vmult_range(op, dst, src, st, 0, op.n_cells);
        }
    };
    auto reduce = [&]() {
      if (!d) return;
      // one batched all-reduce of (r.u, w.u, r.r) per iteration
      NCCL_CHECK(NcclApi::get().AllReduce(sc, sc, 3, 8 /*ncclFloat64*/, 0 /*ncclSum*/, d->comm, st));
    };
    apply(r, x);
    run_cg_residual(op.number, r, u, b, inv_diag, n, st);
    mark(); // [0]
    apply(w, u);
    mark();
    run_cg_dots(op.number, r, u, w, n, sc, st);
    mark();
    reduce();
    mark();
    run_cg_scalars(sc, hist, st);
    double r0 = 0, res = 0;
    CUDA_CHECK(cudaMemcpyAsync(&r0, hist, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    int it = 0;
    res    = r0;
    // per iteration four marks: after update, vmult, dots, all-reduce
    while (it < max_iter && r0 > 0.0)
      {
        run_cg_update(op.number, p, s, x, r, u, w, inv_diag, n, sc, st);
        mark();
        apply(w, u);
        mark();
        run_cg_dots(op.number, r, u, w, n, sc, st);
        mark();
        reduce();
        mark();
        run_cg_scalars(sc, hist, st);
        ++it;
        if (it % options->check_every == 0 || it == max_iter)
          {
            CUDA_CHECK(cudaMemcpyAsync(&res, hist + it, sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            if (res <= options->rel_tol * r0) break;
          }
      }
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (residual_history) CUDA_CHECK(cudaMemcpy(residual_history, hist, sizeof(double) * (size_t)(it + 1), cudaMemcpyDeviceToHost));
    if (result)
      {
        result->iterations       = it;
        result->initial_residual = r0;
        result->final_residual   = res;
        result->ms_vmult = result->ms_vector_ops = result->ms_allreduce = result->ms_total = 0;
        if (timed && ev.size() >= 4)
          {
            auto ms = [&](size_t a, size_t b2) {
              float t = 0;
              cudaEventElapsedTime(&t, ev[a], ev[b2]);
              return (double)t;
            };
            // marks: 0 residual | 1 vmult | 2 dots | 3 all-reduce | then per iteration: update, vmult, dots, all-reduce
            for (size_t i = 4; i + 3 < ev.size(); i += 4)
              {
                result->ms_vector_ops += ms(i - 1, i) + ms(i + 1, i + 2); // scalars + update, dots
                result->ms_vmult += ms(i, i + 1);
                result->ms_allreduce += ms(i + 2, i + 3);
              }
            result->ms_total = ms(3, ev.size() - 1);
          }
      }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  });
}

int mfhn_vec_alloc(int64_t bytes, void **ptr)
{
  return guard([&] {
    if (!ptr || bytes < 0) throw InvalidArgument("bad argument");
    // These vectors are exported with cudaIpcGetMemHandle.  cudaMalloc carves small requests out of shared 2 MiB
    // blocks and the handle maps the BLOCK: a peer that opens it gets the block's base, not this pointer.  Whole
    // 2 MiB multiples give every vector blocks of its own, so the opened pointer is the vector.
    const size_t granule = (size_t)2 << 20, size = (std::max<size_t>((size_t)bytes, 8) + granule - 1) / granule * granule;
    CUDA_CHECK(cudaMalloc(ptr, size));
    CUDA_CHECK(cudaMemset(*ptr, 0, size));
  });
}
int mfhn_vec_free(void *ptr)
{
  return guard([&] { CUDA_CHECK(cudaFree(ptr)); });
}
int mfhn_ipc_get_handle(void *ptr, void *handle64)
{
  return guard([&] {
    if (!ptr || !handle64) throw InvalidArgument("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CUDA_CHECK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64), ptr));
  });
}
int mfhn_ipc_open_handle(const void *handle64, void **ptr)
{
  return guard([&] {
    if (!ptr || !handle64) throw InvalidArgument("null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  });
}
int mfhn_ipc_close_handle(void *ptr)
{
  return guard([&] { CUDA_CHECK(cudaIpcCloseMemHandle(ptr)); });
}
int mfhn_dist_enable_peer(mfhn_dist h, void *src_local, void *dst_local, void *const *peer_src, void *const *peer_dst,
                          const int32_t *ghost_owner, const int64_t *ghost_remote_index)
{
  return guard([&] {
    if (!h || !src_local || !dst_local || !peer_src || !peer_dst) throw InvalidArgument("null argument");
    Dist &d      = *reinterpret_cast<Dist *>(h);
    Operator &op = *d.op;
    if (op.geometry_type != MFHN_GEOM_CARTESIAN) throw NotImplemented("peer mode covers Cartesian cells");
    CUDA_CHECK(cudaSetDevice(op.device));
    const size_t s = op.number == MFHN_F64 ? 8 : 4;
    std::vector<void *> gs((size_t)std::max<long long>(op.n_ghost, 1)), gd(gs.size());
    for (long long g = 0; g < op.n_ghost; ++g)
      {
        if (!ghost_owner || !ghost_remote_index) throw InvalidArgument("null argument");
        const int o = ghost_owner[g];
        if (o < 0 || o >= d.world || o == d.rank || !peer_src[o] || !peer_dst[o]) throw InvalidArgument("bad ghost owner");
        gs[g] = static_cast<char *>(peer_src[o]) + (size_t)ghost_remote_index[g] * s;
        gd[g] = static_cast<char *>(peer_dst[o]) + (size_t)ghost_remote_index[g] * s;
      }
    cudaFree(d.d_ghost_src);
    cudaFree(d.d_ghost_dst);
    d.d_ghost_src = to_device(gs);
    d.d_ghost_dst = to_device(gd);
    if (!d.d_barrier)
      {
        CUDA_CHECK(cudaMalloc(&d.d_barrier, sizeof(int)));
        CUDA_CHECK(cudaMemset(d.d_barrier, 0, sizeof(int)));
      }
    d.peer_src_local = src_local;
    d.peer_dst_local = dst_local;
  });
}
int mfhn_dist_enable_peer_flags(mfhn_dist h, void *flags_local, void *const *peer_flags)
{
  return guard([&] {
    if (!h || !flags_local || !peer_flags) throw InvalidArgument("null argument");
    Dist &d = *reinterpret_cast<Dist *>(h);
    CUDA_CHECK(cudaSetDevice(d.op->device));
    std::vector<unsigned *> pf((size_t)d.world, nullptr);
    for (int r = 0; r < d.world; ++r)
      {
        if (r != d.rank && !peer_flags[r]) throw InvalidArgument("flag arrays of all peers are required");
        pf[r] = static_cast<unsigned *>(r == d.rank ? flags_local : peer_flags[r]);
      }
    cudaFree(d.d_peer_flags);
    d.d_peer_flags = to_device(pf);
    d.flags_local  = static_cast<unsigned *>(flags_local);
  });
}
int mfhn_dist_vmult_peer(mfhn_dist h, void *stream, int zero_dst)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    Dist &d = *reinterpret_cast<Dist *>(h);
    if (!d.peer_src_local) throw InvalidArgument("mfhn_dist_enable_peer has not been called");
    Operator &op = *d.op;
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t main = static_cast<cudaStream_t>(stream);
    if (zero_dst) CUDA_CHECK(cudaMemsetAsync(d.peer_dst_local, 0, (size_t)(op.n_owned + op.n_ghost) * (op.number == MFHN_F64 ? 8 : 4), main));
    dist_vmult_peer_n(d, main);
  });
}
}
