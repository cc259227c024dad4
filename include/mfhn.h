/*
 * mfhn.h -- C ABI of the B200-native matrix-free hanging-node operator engine.
 *
 * This is the drop-in boundary for the hot path of
 * peterrum/dealii-matrixfree-hanging-nodes: the 3D Laplace vmult on FE_Q(k),
 * k = 1..8, on octree meshes with hanging-node constraints.  Plain pointers and
 * sizes only; no C++ or torch types cross it.  Every entry point returns an
 * int status (0 = ok) and never throws; mfhn_last_error() gives the message.
 *
 * Each entry point cites the reference interface (file:line in the reference
 * repository) it stands in for.  The arithmetic behind those interfaces lives
 * in deal.II (not vendored by the reference), see DESIGN.md.
 */
#ifndef MFHN_H
#define MFHN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFHN_OK 0
#define MFHN_ERR_INVALID 1     /* bad argument (e.g. unsupported degree, unknown geometry) */
#define MFHN_ERR_CUDA 2        /* CUDA runtime failure */
#define MFHN_ERR_NOT_IMPL 3    /* like the reference's ExcNotImplemented */

typedef struct mfhn_mesh_s *mfhn_mesh;
typedef struct mfhn_dofs_s *mfhn_dofs;
typedef struct mfhn_op_s *mfhn_op;

/* number types (reference: `using Number = double`, benchmark_03.h:390) */
#define MFHN_F64 0
#define MFHN_F32 1

/* mesh flavours: which 2:1 balance the refinement enforces */
#define MFHN_SERIAL 0 /* dealii::Triangulation: faces + edges  (benchmark_01.h:183)           */
#define MFHN_P4EST 1  /* parallel::distributed::Triangulation: + corners (benchmark_03.h:397) */

/* geometry descriptions accepted by mfhn_op_create */
#define MFHN_GEOM_CARTESIAN 0 /* one double per cell: edge length h                       */
#define MFHN_GEOM_AFFINE 1    /* nine doubles per cell: Jacobian J[r][c] = dx_r/dxi_c      */
#define MFHN_GEOM_GENERAL 2   /* [cell][6][(k+1)^3] doubles: symmetric JxW J^-1 J^-T (xx,xy,xz,yy,yz,zz)
                                 per quadrature point, lexicographic -- curved cells / high-order
                                 mappings (TestHighOrderMapping, benchmark_01.h:225-242)   */

/* cell kernels */
#define MFHN_KERNEL_AUTO 0
#define MFHN_KERNEL_QPOINT 1    /* collocation sum factorisation + quadrature-point operation */
#define MFHN_KERNEL_SEPARABLE 2 /* Cartesian cells only: 1D mass/stiffness tensor form        */
#define MFHN_KERNEL_BASELINE 3  /* restatement of the deal.II CUDA design (one thread per DoF) */
#define MFHN_KERNEL_PLANE 4     /* Cartesian cells, register-tiled separable kernel, per-cell gather */
#define MFHN_KERNEL_PATCH 5     /* removed (round-1 experiment: sorted-unique patch gather, slower than PLANE); returns MFHN_ERR_NOT_IMPL */
#define MFHN_KERNEL_BULK 6      /* plane kernel; cell-interior and face blocks moved by the bulk-copy engine
                                   (cp.async.bulk / cp.reduce.async.bulk), degrees 3..5, 16-byte aligned vectors */
#define MFHN_KERNEL_RUNS 7      /* plane kernel; every cell's vector entries cut into contiguous runs at setup (any
                                   numbering), long runs moved by the bulk-copy engine, the rest entry by entry;
                                   degrees 1..5, 16-byte aligned vectors */
#define MFHN_KERNEL_QPOINT_ROWS 8 /* QPOINT on Cartesian cells with the GENERAL-PURPOSE constraint algorithm: hanging-node
                                   constraints resolved entry by entry through weighted rows in the gather / scatter, no
                                   interpolation passes (use_fast_hanging_node_algorithm = false, benchmark_01.h:286-293;
                                   t6 / t7 of benchmark_01.cc:222-234) */

const char *mfhn_last_error(void);
const char *mfhn_version(void);

/* ---------------------------------------------------------------------------
 * Mesh layer.  Replaces GridGenerator::create_{quadrant,annulus,step,
 * quadrant_flexible} on hyper_cube(-1,1)^3   (reference benchmark.h:7-144,
 * benchmark_03.h:26-104).
 * ------------------------------------------------------------------------ */
int mfhn_mesh_create(const char *geometry, int n_refinements, int flavour, mfhn_mesh *out);
void mfhn_mesh_destroy(mfhn_mesh m);
int64_t mfhn_mesh_n_cells(mfhn_mesh m);  /* tria.n_active_cells()   benchmark_01.h:302 */
int mfhn_mesh_n_levels(mfhn_mesh m);     /* tria.n_global_levels()  benchmark_03.h:406 */
/* Active cells in storage order: out[4*c + {0,1,2,3}] = level, ix, iy, iz. */
int mfhn_mesh_cells(mfhn_mesh m, int32_t *out);
/* Number of cells with hanging nodes: Helper::is_constrained
 * (constraint_helper.h:89-125; counted in benchmark_03.h:415-432). */
int64_t mfhn_mesh_n_cells_hn(mfhn_mesh m);
/* Morton-curve partition of the active cells into n_ranks contiguous chunks of
 * equal weight; weight = 1 + 10*w for hanging-node cells, 11 otherwise
 * (benchmark_02.cc:15-37; w == 1 gives the plain p4est partition).
 * rank_of_cell is in storage order. */
int mfhn_mesh_partition(mfhn_mesh m, int n_ranks, double hn_weight, int32_t *rank_of_cell);
/* Position of every active cell (storage order) along the Morton curve. */
int mfhn_mesh_morton_position(mfhn_mesh m, int64_t *position_of_cell);

/* ---------------------------------------------------------------------------
 * DoF layer.  Replaces DoFHandler::distribute_dofs(FE_Q(degree)) and the
 * index / mask part of MatrixFree::reinit  (benchmark_01.h:247,255-256;
 * benchmark_03.h:227-228,338-339,438-439).
 * With rank_of_cell == NULL (or n_ranks == 1) the numbering is the serial one.
 * ------------------------------------------------------------------------ */
int mfhn_dofs_create(mfhn_mesh m, int degree, int n_ranks, const int32_t *rank_of_cell,
                     mfhn_dofs *out);
void mfhn_dofs_destroy(mfhn_dofs d);
int64_t mfhn_dofs_n_dofs(mfhn_dofs d); /* dof_handler.n_dofs()  benchmark_03.h:441 */
/* Global range [begin,end) owned by rank. */
int mfhn_dofs_owned_range(mfhn_dofs d, int rank, int64_t *begin, int64_t *end);
int64_t mfhn_dofs_n_cells_of_rank(mfhn_dofs d, int rank);
/* Storage-order indices of the cells of `rank`, ascending. */
int mfhn_dofs_cells_of_rank(mfhn_dofs d, int rank, int64_t *cell_ids);
/* Fill the arrays for the n cells listed in cell_ids (storage-order indices,
 * any order -- MatrixFree is free to reorder its cell batches).  Any output
 * pointer may be NULL.
 *   raw_indices  uint64[n*(k+1)^3]   lexicographic, global, before substitution
 *   dof_indices  uint64[n*(k+1)^3]   lexicographic, global, coarse-substituted
 *   masks        uint8[n]            compressed_constraint_kind
 *                                    (dof_info.hanging_node_constraint_masks, benchmark_01.h:335)
 *   h            double[n]           cell edge length */
int mfhn_dofs_fill(mfhn_dofs d, int64_t n, const int64_t *cell_ids, uint64_t *raw_indices,
                   uint64_t *dof_indices, uint8_t *masks, double *h);
/* Support points (x,y,z) of the DoFs in [begin,end) -- what
 * VectorTools::interpolate needs (benchmark_03.h:459-461). */
int mfhn_dofs_support_points(mfhn_dofs d, int64_t begin, int64_t end, double *xyz);

/* ---------------------------------------------------------------------------
 * MatrixFree::reinit (benchmark_03.h:326-340; benchmark_01.h:251-284 with the
 * cell_vectorization_category of the Categorize option): the rank-local setup
 * product -- loop order of the cells (Morton curve, interior cells first,
 * grouped by constraint mask inside windows), rank-local DoF numbering (owned
 * range, then ghosts sorted by global index), compressed masks, cell sizes and
 * the Utilities::MPI::Partitioner data.  Host memory; the arrays returned by
 * mfhn_mf_arrays / mfhn_mf_partitioner belong to the handle.
 * ------------------------------------------------------------------------ */
typedef struct mfhn_mf_s *mfhn_mf;
typedef struct
{
  int rank;            /* rank whose cells are set up (0 for a serial DoF handler)                     */
  int categorize;      /* 0: Morton order, 1: group by constraint mask, 2: by constrained/unconstrained  */
  int window;          /* cells per categorisation window, 0 = default (3840)                          */
  int batch_alignment; /* partition boundaries on multiples of this many cells, 0 = default (240)      */
} mfhn_mf_options;
typedef struct
{
  int degree, rank, n_ranks;
  int64_t n_cells, n_cells_hn;
  int64_t n_owned, n_ghost, owned_begin;
  int64_t n_interior_a, n_interior; /* cells [0,a) | [a,interior) | [interior,n_cells) touch ghost entries   */
  int n_ghost_peers, n_import_peers;
  int64_t n_import;
} mfhn_mf_sizes;
int mfhn_mf_create(mfhn_dofs d, const mfhn_mf_options *options /* NULL = defaults */, mfhn_mf *out);
void mfhn_mf_destroy(mfhn_mf m);
int mfhn_mf_info(mfhn_mf m, mfhn_mf_sizes *out);
/* cell_ids int64[n_cells] (storage indices in loop order), dof_indices uint32[n_cells*(k+1)^3], masks uint8[n_cells],
 * h double[n_cells], ghost_global int64[n_ghost] (ascending), ghost_owner int32[n_ghost].  Any pointer may be NULL. */
int mfhn_mf_arrays(mfhn_mf m, const int64_t **cell_ids, const uint32_t **dof_indices, const uint8_t **masks, const double **h,
                   const int64_t **ghost_global, const int32_t **ghost_owner);
/* Partitioner: per ghost peer its contiguous range inside the ghost section; per import peer the local owned indices it
 * reads (import_offsets has n_import_peers + 1 entries); rank_begin[n_ranks + 1] = owned ranges of all ranks. */
int mfhn_mf_partitioner(mfhn_mf m, const int32_t **ghost_peers, const int64_t **ghost_begin, const int64_t **ghost_end,
                        const int32_t **import_peers, const int64_t **import_offsets, const int32_t **import_indices,
                        const int64_t **rank_begin);
/* The import lists come from the peers: rank `peer` ghosts the given global indices (all owned here).  Between processes the
 * caller moves the requests (ghost_global ranges) with whatever transport it has; inside one process
 * mfhn_mf_exchange_local does it for the handles of all ranks (ordered by rank). */
int mfhn_mf_set_imports(mfhn_mf m, int peer, const int64_t *global_indices, int64_t n);
int mfhn_mf_exchange_local(mfhn_mf *all_ranks, int n_ranks);

/* ConstraintKinds <-> compressed byte (deal.II hanging_nodes_internal.h,
 * used at benchmark_00_likwid.cc:45-48). */
uint8_t mfhn_compress(uint16_t kind);
uint16_t mfhn_decompress(uint8_t compressed);
int mfhn_check_kind(uint16_t kind);

/* ---------------------------------------------------------------------------
 * Operator.  Replaces LaplaceOperator<3,degree,Number,MemorySpace::CUDA>
 * (benchmark_03.h:319-357): constructor = CUDAWrappers::MatrixFree::reinit,
 * vmult = cell_loop(LaplaceOperatorLocal, src, dst).
 * ------------------------------------------------------------------------ */
typedef struct
{
  int degree;                  /* 1..8 (reference dispatches 1..6, benchmark_03.h:565-614)   */
  int number;                  /* MFHN_F64 / MFHN_F32                                         */
  int64_t n_cells;             /* cells of this rank                                          */
  int64_t n_owned;             /* locally owned DoFs: vector entries [0,n_owned)              */
  int64_t n_ghost;             /* ghost DoFs: vector entries [n_owned,n_owned+n_ghost)        */
  const uint32_t *dof_indices; /* [n_cells*(k+1)^3] rank-local, lexicographic, substituted     */
  const uint8_t *masks;        /* [n_cells] compressed_constraint_kind                         */
  int geometry_type;           /* MFHN_GEOM_*                                                  */
  const double *geometry;      /* per-cell geometry, see MFHN_GEOM_*                           */
  int apply_constraints;       /* benchmark_03.h:255-268: false => plain gather/scatter        */
  int kernel;                  /* MFHN_KERNEL_*                                                */
  int device;                  /* CUDA device ordinal, -1 = current                            */
  const int64_t *segments;     /* optional: ascending first cells of cell segments (segments[0]==0);
                                  mfhn_op_vmult_range may only be called on unions of segments
                                  (deal.II's cell_loop partitions for communication overlap)   */
  int n_segments;
  int vector_padding;          /* the caller promises that every src / dst vector has this many valid (zero, never
                                  inspected) entries BEHIND its n_owned + n_ghost entries.  With >= 4 the bulk-copy
                                  kernel may treat the block that ends the vector like any other (its 16-byte
                                  widened range reaches past the last entry); 0 = such a cell runs separately */
} mfhn_op_desc;

int mfhn_op_create(const mfhn_op_desc *desc, mfhn_op *out);
/* The operator of a MatrixFree handle (Cartesian cells, segments = its three cell partitions): what
 * LaplaceOperator's constructor does with matrix_free.reinit (benchmark_03.h:326-340). */
int mfhn_op_create_mf(mfhn_mf m, int number, int kernel, int apply_constraints, int device, mfhn_op *out);
/* Same with the vector_padding promise of mfhn_op_desc (vectors allocated with MFHN_VECTOR_PADDING spare entries). */
#define MFHN_VECTOR_PADDING 4
int mfhn_op_create_mf_padded(mfhn_mf m, int number, int kernel, int apply_constraints, int device, int vector_padding, mfhn_op *out);
void mfhn_op_destroy(mfhn_op op);

/* dst (+)= A src on device vectors of n_owned+n_ghost entries of the operator's
 * Number type.  Stream-ordered and asynchronous like the reference's cell_loop
 * (the benchmark synchronises explicitly, benchmark_03.h:485).  zero_dst == 0
 * accumulates exactly like the reference (benchmark_03.h:240,352); zero_dst != 0
 * clears dst first inside the same stream.
 * cell_begin/cell_end select a sub-range of the operator's cells (used to
 * overlap the ghost exchange with interior cells); pass 0, -1 for all. */
int mfhn_op_vmult(mfhn_op op, void *dst, const void *src, void *cuda_stream, int zero_dst);
int mfhn_op_vmult_range(mfhn_op op, void *dst, const void *src, void *cuda_stream,
                        int64_t cell_begin, int64_t cell_end);

/* Same operation on HOST vectors (pinned memory recommended): copies src (and
 * dst unless zero_dst) to the device, applies the operator, copies dst back.
 * All copies are issued asynchronously on cuda_stream; synchronise the stream
 * before reading dst.  This is the call a host-vector caller such as
 * LaplaceOperator<...,MemorySpace::Host>::vmult (benchmark_03.h:237-241) binds. */
int mfhn_op_vmult_host(mfhn_op op, void *dst_host, const void *src_host, void *cuda_stream,
                       int zero_dst);
/* Same with an explicit device staging slot (0 or 1): two calls on different
 * streams and different slots may be in flight at once, so the upload of one
 * application overlaps the download of the previous one (full-duplex PCIe). */
int mfhn_op_vmult_host_slot(mfhn_op op, void *dst_host, const void *src_host, void *cuda_stream,
                            int zero_dst, int slot);

/* diag += diagonal of the operator (device vector of n_owned + n_ghost entries; ghost entries
 * hold the contributions to peers' DoFs and must be compressed like a vmult result).  Extension
 * for a point-Jacobi preconditioner (BASELINE.json config 5); the reference has no counterpart. */
int mfhn_op_diagonal(mfhn_op op, void *diag, void *cuda_stream);

/* Change the apply_constraints switch / kernel of an existing operator. */
int mfhn_op_set_apply_constraints(mfhn_op op, int apply_constraints);
int mfhn_op_set_kernel(mfhn_op op, int kernel);

/* Hanging-node interpolation alone on cell-local values (device pointer,
 * [n_cells][(k+1)^3]): FEEvaluationHangingNodesFactory::apply
 * (benchmark_00_likwid.cc:56-59). */
int mfhn_op_apply_hn(mfhn_op op, void *cell_values, int transpose, void *cuda_stream);
/* "DG (C)" stage of the reference's decomposition (benchmark_01.cc:189-199; benchmark_01.h:617-677 with cell-local
 * vectors): dst_cells += [interpolation^T interpolation] src_cells on [n_cells][(k+1)^3] device arrays, no
 * quadrature-point work; constraints as set by mfhn_op_set_apply_constraints. */
int mfhn_op_dg_copy(mfhn_op op, void *dst_cells, const void *src_cells, void *cuda_stream);
/* Hanging-node strategy of the fused cell kernels (analogues of the reference's index / sorted / mask vectorisation
 * types, benchmark_01.cc:70-116): MFHN_HN_BRANCH (default) = a warp takes the interpolation passes only if one of its
 * cells is constrained (with the categorised cell order of mfhn_mf_create this is the "sorted" strategy, with plain
 * Morton order the "index" strategy); MFHN_HN_MASK = every warp takes the passes, constrained lines are selected by
 * per-lane predicates (no data-dependent branch, no dependence on the cell order). */
#define MFHN_HN_BRANCH 0
#define MFHN_HN_MASK 1
int mfhn_op_set_hn_strategy(mfhn_op op, int strategy);

/* Queries: algorithmic bytes / flops of one vmult, cell counts. */
int mfhn_op_query(mfhn_op op, const char *what, double *value);
/* Number of kernels launched by this operator since creation. */
int64_t mfhn_op_launch_count(mfhn_op op);

/* ---------------------------------------------------------------------------
 * Distributed-vector helpers.  Replace the device side of
 * LinearAlgebra::distributed::Vector<Number,MemorySpace::CUDA>::
 * update_ghost_values / compress(add)  (benchmark_03.h:323-324, used inside
 * CUDAWrappers::MatrixFree::cell_loop): pack owned entries into a send buffer,
 * add received contributions.
 * ------------------------------------------------------------------------ */
int mfhn_pack(int number, void *buffer, const void *vec, const int32_t *indices_dev, int64_t n,
              void *cuda_stream);
int mfhn_unpack_add(int number, void *vec, const void *buffer, const int32_t *indices_dev,
                    int64_t n, void *cuda_stream);

/* ---------------------------------------------------------------------------
 * Partitioned operator: the complete vmult of one rank -- ghost import, the three
 * cell partitions, ghost compress -- issued by ONE call (pack / unpack kernels,
 * NCCL send/recv groups on an internal communication stream, events).  This is
 * CUDAWrappers::MatrixFree::cell_loop with its update_ghost_values /
 * compress(add) (benchmark_03.h:348-353) for one process per GPU.  NCCL is
 * resolved at run time (dlopen of libnccl.so.2).
 * ------------------------------------------------------------------------ */
typedef struct mfhn_dist_s *mfhn_dist;
typedef struct
{
  int rank, world;
  const void *unique_id;          /* 128 bytes from mfhn_dist_unique_id on rank 0, broadcast by the caller */
  int n_import_peers;             /* peers that ghost entries owned here                                */
  const int32_t *import_peers;    /* their ranks                                                        */
  const int64_t *import_offsets;  /* [n_import_peers + 1] into import_indices                            */
  const int32_t *import_indices;  /* local owned indices, grouped by peer                                */
  int n_ghost_peers;              /* peers owning this rank's ghosts                                     */
  const int32_t *ghost_peers;
  const int64_t *ghost_begin;     /* per peer: its contiguous range inside the ghost section             */
  const int64_t *ghost_end;
  int64_t segments[4];            /* 0, end of interior A, end of interior B, n_cells (boundary cells last) */
} mfhn_dist_desc;
int mfhn_dist_unique_id(void *id128);
int mfhn_dist_create(mfhn_op op, const mfhn_dist_desc *desc, mfhn_dist *out);
/* Same from a MatrixFree handle whose import lists are set. */
int mfhn_dist_create_mf(mfhn_op op, mfhn_mf m, const void *unique_id, mfhn_dist *out);
void mfhn_dist_destroy(mfhn_dist d);
/* NOTE: the ghost section of src is scratch: it receives the imported entries and is cleared again
 * (zero_out_ghost_values), so src is written although it is declared const. */
int mfhn_dist_vmult(mfhn_dist d, void *dst, const void *src, void *cuda_stream, int zero_dst);
int64_t mfhn_dist_launch_count(mfhn_dist d);

/* Peer-memory variant of the partitioned vmult (fused compute + exchange): the boundary cells read
 * the ghost entries from the OWNER's src and add their contributions into the OWNER's dst through
 * peer-mapped (CUDA IPC) pointers over NVLink -- no pack / unpack kernels, no data-path collective;
 * two 4-byte all-reduces act as barriers.  The vector pair must come from mfhn_vec_alloc so that it
 * can be exported; every rank passes the opened peer pointers of all ranks.  The boundary cells run
 * through the plane kernels (all degrees), the interior cells through the operator's own kernel.
 * mfhn_vec_alloc returns zeroed memory in whole 2 MiB blocks: a CUDA IPC handle maps the allocation BLOCK, so a
 * vector that shared a block with other small allocations would be opened at the wrong address by its peers. */
int mfhn_vec_alloc(int64_t bytes, void **dev_ptr);
int mfhn_vec_free(void *dev_ptr);
int mfhn_ipc_get_handle(void *dev_ptr, void *handle64);
int mfhn_ipc_open_handle(const void *handle64, void **dev_ptr);
int mfhn_ipc_close_handle(void *dev_ptr);
int mfhn_dist_enable_peer(mfhn_dist d, void *src_local, void *dst_local, void *const *peer_src,
                          void *const *peer_dst, const int32_t *ghost_owner,
                          const int64_t *ghost_remote_index);
/* Optional: barriers of the peer path as flag exchanges in each other's memory instead of NCCL all-reduces.
 * flags_local: at least (world + 1) * 4 zeroed bytes from mfhn_vec_alloc on this rank; peer_flags[r]: rank r's array
 * opened with mfhn_ipc_open_handle (entry of the own rank ignored).  With flags the whole vmult consists of CUDA
 * kernels, memsets and events only: it can be captured in a CUDA graph. */
int mfhn_dist_enable_peer_flags(mfhn_dist d, void *flags_local, void *const *peer_flags);
int mfhn_dist_vmult_peer(mfhn_dist d, void *cuda_stream, int zero_dst);

/* ---------------------------------------------------------------------------
 * Conjugate gradients with a point-Jacobi preconditioner (BASELINE.json config 5).  EXTENSION: the reference contains
 * no solver; this is the step either side of vmult in a real solve -- vmult + fused vector updates + one batched
 * dot-product all-reduce per iteration (Chronopoulos / Gear form), all on the device.
 * ------------------------------------------------------------------------ */
typedef struct
{
  int max_iter;
  double rel_tol;  /* stop when |r| <= rel_tol |r_0| (checked every check_every iterations)                    */
  int check_every; /* the host reads the residual only every so many iterations                                */
  int timings;     /* != 0: split the device time of the iterations into vmult / vector kernels / all-reduce   */
} mfhn_cg_options;
typedef struct
{
  int iterations;
  double initial_residual, final_residual;
  double ms_total, ms_vmult, ms_vector_ops, ms_allreduce; /* device time of the iterations (timings != 0) */
} mfhn_cg_result;
/* inv_diag = 1 / diag(A) on the owned entries (0 where the diagonal vanishes: hanging entries); dist may be NULL. */
int mfhn_op_inverse_diagonal(mfhn_op op, mfhn_dist dist, void *inv_diag, void *cuda_stream);
/* Solves A x = b from the start vector x; b must be consistent (the Laplace operator without Dirichlet data is
 * singular).  x, b, inv_diag: device vectors of the operator (inv_diag NULL = no preconditioner); dist NULL = one
 * rank.  residual_history (host, max_iter + 1 doubles) may be NULL.  Synchronises the stream. */
int mfhn_cg_solve(mfhn_op op, mfhn_dist dist, void *x, const void *b, const void *inv_diag, const mfhn_cg_options *options,
                  mfhn_cg_result *result, double *residual_history, void *cuda_stream);

/* Host-only check of the MFHN_KERNEL_BULK layout for a reference index array (read_dof_values order,
 * benchmark_03.h:255-258): n_irregular = cells that do not show contiguous cell-interior / face blocks
 * (they run through the plane kernel), n_mismatch = entries an emulated gather through the layout gets
 * wrong (must be 0).  Needs no GPU. */
int mfhn_bulk_layout_check(int degree, int number, int64_t n_cells, int64_t n_vec, const uint32_t *dof_indices,
                           int64_t *n_irregular, int64_t *n_mismatch);

/* Host-only check of the MFHN_KERNEL_RUNS layout: runs of the sorted per-cell index lists with at most max_gap unused
 * entries inside and at least min_run cell entries become bulk copies, the rest single entries; place != 0 chooses
 * the staging positions against shared-memory bank conflicts.  n_mismatch = entries an emulated gather gets wrong +
 * copied foreign entries that would not be zero in the scatter (must be 0); staging_wavefronts = bank model of the
 * staging reads, summed over the warp batches (2 per plane slot = conflict-free). */
int mfhn_runs_layout_check(int degree, int number, int64_t n_cells, int64_t n_vec, const uint32_t *dof_indices, int max_gap,
                           int min_run, int place, int64_t *n_bulk_copies, int64_t *n_single_entries, int64_t *n_mismatch,
                           int64_t *staging_wavefronts);

/* Microbenchmarks used for the roofline denominators (bench.py). */
int mfhn_bench_dfma(int number, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MFHN_H */
