// Host-side descriptions shared by the kernel translation units and the operator object (op.cu): cell-loop
// parameters, the device layouts of the cell kernels and the non-template launch entry points.  Every kernel
// family lives in its own translation unit (k_*.cu) so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace mfhn
{
struct CellLoopParams
{
  const uint32_t *idx;  // [n_cells][(k+1)^3] lexicographic
  const uint8_t *masks; // [n_cells]
  const void *geom;     // Number[n_cells] (h), Number[n_cells][6] (metric) or Number[n_cells][6][(k+1)^3]
  const void *src;
  void *dst;
  long long cell_begin, cell_end;
  int apply_constraints;
  int hn_mask_strategy = 0; // every warp takes the interpolation passes (MFHN_HN_MASK)
  // GV_QPOINT_ROWS: constraint rows per distinct mask (local columns), see ConstraintRows
  const int32_t *row_kind = nullptr; // [n_cells] 0 = unconstrained, else 1 + index of the cell's mask among the distinct masks
  const int32_t *row_ptr  = nullptr; // [n_kinds][(k+1)^3 + 1] offsets into row_col / row_val
  const uint16_t *row_col = nullptr; // local DoF of the cell
  const void *row_val     = nullptr; // Number weight
};

// general-purpose constraint algorithm (MFHN_KERNEL_QPOINT_ROWS): one sparse interpolation matrix per distinct mask
struct ConstraintRows
{
  int32_t *d_kind = nullptr, *d_ptr = nullptr;
  uint16_t *d_col = nullptr;
  void *d_val     = nullptr;
  long long n_kinds = 0, n_entries = 0;
  bool built = false;

  void free()
  {
    cudaFree(d_kind);
    cudaFree(d_ptr);
    cudaFree(d_col);
    cudaFree(d_val);
    d_kind = d_ptr = nullptr;
    d_col  = nullptr;
    d_val  = nullptr;
    built  = false;
  }
};

enum GenericVariant
{
  GV_QPOINT_CARTESIAN = 0, // collocation gradients, diagonal q-point factor w_q h
  GV_QPOINT_METRIC    = 1, // collocation gradients, symmetric 3x3 metric per cell
  GV_SEPARABLE        = 2, // h (K x M x M + M x K x M + M x M x K)
  GV_QPOINT_GENERAL   = 3, // collocation gradients, symmetric 3x3 coefficient per QUADRATURE POINT
                           // (JxW J^-1 J^-T, [cell][6][q]): curved cells / high-order mappings
  GV_QPOINT_ROWS      = 4  // GV_QPOINT_CARTESIAN with the GENERAL-PURPOSE constraint algorithm: hanging-node constraints
                           // resolved entry by entry through weighted rows in the gather / scatter instead of the
                           // interpolation passes (use_fast_hanging_node_algorithm = false, benchmark_01.h:286-293)
};

// warp-interleaved index layout of the plane kernels: [n_batches][n*n][32]
struct PlaneLayout
{
  int n = 0;
  long long n_cells = 0, n_batches = 0;
  uint32_t *d_pidx = nullptr;

  void free()
  {
    cudaFree(d_pidx);
    d_pidx = nullptr;
  }
  void build(int n_, long long n_cells_, const uint32_t *idx); // op.cu
};

// peer mode (boundary cells of a partitioned operator): ghost entries are read from / added to the OWNER's vectors
struct PeerTables
{
  long long n_owned           = 0;
  const void *const *ghost_src = nullptr;
  void *const *ghost_dst       = nullptr;
};

struct BulkHostLayout
{
  int n = 0, E = 2;
  long long n_cells = 0, n_batches = 0;
  std::vector<uint32_t> bidx, lvidx, cinfo;
  std::vector<long long> irregular; // cells left to the plane kernel, ascending
};

struct BulkLayout
{
  int n = 0;
  long long n_cells = 0, n_batches = 0;
  uint32_t *d_bidx = nullptr, *d_lvidx = nullptr, *d_cinfo = nullptr;
  std::vector<long long> irregular;
  bool usable = false; // built and few enough irregular cells

  void free()
  {
    cudaFree(d_bidx);
    cudaFree(d_lvidx);
    cudaFree(d_cinfo);
    d_bidx = d_lvidx = d_cinfo = nullptr;
    usable = false;
  }
};

// run-wise layout of MFHN_KERNEL_RUNS (kernels_runs.cuh)
struct RunsHostLayout
{
  int n = 0, E = 2, cap = 0, NT = 0, RW = 0, BR = 0, sr = 0;
  long long n_cells = 0, n_batches = 0, n_blocks = 0, n_singles = 0, n_zero = 0;
  std::vector<uint32_t> rec, srow, ov_idx; // rec: RW words per batch
  std::vector<uint16_t> sprow, ov_pos, zpos;
};

struct RunsLayout
{
  int n = 0, sr = 0;
  long long n_cells = 0, n_batches = 0, n_blocks = 0, n_singles = 0, n_zero = 0, staging_wavefronts = 0;
  uint32_t *d_rec = nullptr, *d_srow = nullptr, *d_ov_idx = nullptr;
  uint16_t *d_sprow = nullptr, *d_ov_pos = nullptr, *d_zpos = nullptr;
  bool usable = false;

  void free()
  {
    cudaFree(d_rec);
    cudaFree(d_srow);
    cudaFree(d_ov_idx);
    cudaFree(d_sprow);
    cudaFree(d_ov_pos);
    cudaFree(d_zpos);
    d_rec = d_srow = d_ov_idx = nullptr;
    d_sprow = d_ov_pos = d_zpos = nullptr;
    usable = false;
  }
};

struct BaselineArrays
{
  uint32_t *l2g = nullptr;
  void *invjac = nullptr, *jxw = nullptr;
};

constexpr bool plane_supported(int n) { return n >= 2 && n <= 6; }      // register-tiled plane kernel
constexpr bool plane_smem_supported(int n) { return n >= 2 && n <= 9; } // plane in shared memory
constexpr bool bulk_supported(int n) { return n >= 4 && n <= 6; }
constexpr bool runs_supported(int n) { return n >= 2 && n <= 6; } // staging positions are bytes: (k+1)^3 <= 256

// ---- launch entry points (degree = k, number = MFHN_F64 / MFHN_F32); they throw std::runtime_error ----
// k_generic_*.cu
void run_generic(int degree, int number, int variant, bool diag, const CellLoopParams &p, int device, cudaStream_t stream);
// k_plane.cu: register-tiled kernel for k <= 5, shared-memory plane kernel above; peer != nullptr: ghost entries are
// read from / added to the owners' vectors through peer-mapped pointers
void run_plane(int degree, int number, const PlaneLayout &L, const CellLoopParams &p, int device, cudaStream_t stream, const PeerTables *peer);
// k_bulk.cu
void run_bulk(int degree, int number, const BulkLayout &L, const CellLoopParams &p, int device, cudaStream_t stream);
void bulk_analyze(BulkHostLayout &L, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx);
long long bulk_verify(const BulkHostLayout &L, int number, const uint32_t *idx);
// k_runs.cu
void run_runs(int degree, int number, const RunsLayout &L, const CellLoopParams &p, int device, cudaStream_t stream);
void runs_analyze(RunsHostLayout &L, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx, int max_gap, int min_run, bool place);
long long runs_verify(const RunsHostLayout &L, int number, const uint32_t *idx, long long *staging_wavefronts);
// k_misc.cu
void run_baseline(int degree, int number, BaselineArrays &arrays, const uint32_t *d_idx, const void *d_h, long long n_cells, const CellLoopParams &p,
                  int device, cudaStream_t stream);
void run_hn_only(int degree, int number, void *values, const uint8_t *d_masks, long long n_cells, int transpose, cudaStream_t stream);
void run_dg_copy(int degree, int number, void *dst, const void *src, const uint8_t *d_masks, long long n_cells, const int32_t *d_hn_cells, long long n_hn,
                 int apply_constraints, cudaStream_t stream);
double run_fma_bench(int number, int iters);
void run_pack(int number, void *buffer, const void *vec, const int32_t *idx, long long n, cudaStream_t stream);
void run_unpack_add(int number, void *vec, const void *buffer, const int32_t *idx, long long n, bool atomic, cudaStream_t stream);
// barrier between the ranks of a peer-memory operator: flags in each other's (IPC-mapped) memory, see k_misc.cu
void run_peer_barrier(unsigned *flags_local, unsigned *const *d_peer_flags, int rank, int world, cudaStream_t stream);
// k_cg.cu: fused vector kernels of the Chronopoulos / Gear CG iteration
size_t cg_scalars_bytes();
void run_cg_dots(int number, const void *r, const void *u, const void *w, long long n, void *scalars, cudaStream_t stream);
void run_cg_scalars(void *scalars, double *history, cudaStream_t stream);
void run_cg_update(int number, void *p, void *s, void *x, void *r, void *u, const void *w, const void *inv_diag, long long n, const void *scalars,
                   cudaStream_t stream);
void run_cg_residual(int number, void *r, void *u, const void *b, const void *inv_diag, long long n, cudaStream_t stream);
void run_invert_diagonal(int number, void *d, long long n, cudaStream_t stream);
} // namespace mfhn
