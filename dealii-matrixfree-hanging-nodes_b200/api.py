"""Host-side mirror of the reference's operator interface over the C ABI.

Names and argument meaning follow the reference drivers so that tests and
benchmarks read like benchmark_01 / benchmark_03:

  Triangulation      GridGenerator::create_* on hyper_cube(-1,1)^3 (benchmark.h:7-144)
  DoFHandler         DoFHandler::distribute_dofs(FE_Q(degree))      (benchmark_03.h:438-439)
  MatrixFree         MatrixFree::reinit: rank-local index arrays, masks, geometry,
                     ghost partitioner                               (benchmark_03.h:338-339)
  LaplaceOperator    LaplaceOperator<3,degree,Number,MemorySpace::CUDA>:
                     initialize_dof_vector, vmult                    (benchmark_03.h:319-357)

All arithmetic happens inside libmfhn.so (CUDA); numpy is used for host index
bookkeeping only and torch for device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi as capi
from ._capi import check, lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Triangulation:
    """Octree mesh on hyper_cube(-1,1)^3 refined like the reference's generators.
    flavour "serial" = dealii::Triangulation (benchmark_01.h:183), "p4est" =
    parallel::distributed::Triangulation (benchmark_03.h:397)."""

    def __init__(self, geometry_type: str, n_refinements: int, flavour: str = "p4est"):
        if flavour not in ("serial", "p4est"):
            raise capi.MfhnError(1, "Unknown mesh flavour!")
        self.geometry_type, self.n_refinements, self.flavour = geometry_type, n_refinements, flavour
        h = C.c_void_p()
        check(lib.mfhn_mesh_create(geometry_type.encode(), n_refinements, capi.SERIAL if flavour == "serial" else capi.P4EST, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib.mfhn_mesh_destroy(self._h)
            self._h = None

    def n_active_cells(self) -> int:
        return int(lib.mfhn_mesh_n_cells(self._h))

    n_global_active_cells = n_active_cells

    def n_global_levels(self) -> int:
        return int(lib.mfhn_mesh_n_levels(self._h))

    def cells(self) -> np.ndarray:
        out = np.empty((self.n_active_cells(), 4), dtype=np.int32)
        check(lib.mfhn_mesh_cells(self._h, _ptr(out)))
        return out

    def n_cells_with_hanging_nodes(self) -> int:
        """Helper::is_constrained count (constraint_helper.h:89-125, benchmark_03.h:415-432)."""
        return int(lib.mfhn_mesh_n_cells_hn(self._h))

    def partition(self, n_ranks: int, hn_weight: float = 1.0) -> np.ndarray:
        """Morton partition with the weights of benchmark_02.cc:15-37."""
        out = np.empty(self.n_active_cells(), dtype=np.int32)
        check(lib.mfhn_mesh_partition(self._h, n_ranks, float(hn_weight), _ptr(out)))
        return out

    def morton_position(self) -> np.ndarray:
        out = np.empty(self.n_active_cells(), dtype=np.int64)
        check(lib.mfhn_mesh_morton_position(self._h, _ptr(out)))
        return out


class DoFHandler:
    def __init__(self, tria: Triangulation, degree: int, n_ranks: int = 1, rank_of_cell=None):
        self.tria, self.degree, self.n_ranks = tria, degree, n_ranks
        if rank_of_cell is not None:
            rank_of_cell = np.ascontiguousarray(rank_of_cell, dtype=np.int32)
            assert rank_of_cell.shape == (tria.n_active_cells(),)
        self.rank_of_cell = rank_of_cell
        h = C.c_void_p()
        check(lib.mfhn_dofs_create(tria._h, degree, n_ranks, _ptr(rank_of_cell), C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib.mfhn_dofs_destroy(self._h)
            self._h = None

    def n_dofs(self) -> int:
        return int(lib.mfhn_dofs_n_dofs(self._h))

    def owned_range(self, rank: int = 0):
        b, e = C.c_int64(), C.c_int64()
        check(lib.mfhn_dofs_owned_range(self._h, rank, C.byref(b), C.byref(e)))
        return int(b.value), int(e.value)

    def cells_of_rank(self, rank: int = 0) -> np.ndarray:
        n = int(lib.mfhn_dofs_n_cells_of_rank(self._h, rank))
        if n < 0:
            raise capi.MfhnError(1, "rank out of range")
        out = np.empty(n, dtype=np.int64)
        check(lib.mfhn_dofs_cells_of_rank(self._h, rank, _ptr(out)))
        return out

    def fill(self, cell_ids, raw=False, substituted=True, masks=True, h=True):
        cell_ids = np.ascontiguousarray(cell_ids, dtype=np.int64)
        n, n3 = len(cell_ids), (self.degree + 1) ** 3
        r = np.empty((n, n3), dtype=np.uint64) if raw else None
        s = np.empty((n, n3), dtype=np.uint64) if substituted else None
        m = np.empty(n, dtype=np.uint8) if masks else None
        hh = np.empty(n, dtype=np.float64) if h else None
        check(lib.mfhn_dofs_fill(self._h, n, _ptr(cell_ids), _ptr(r), _ptr(s), _ptr(m), _ptr(hh)))
        return r, s, m, hh

    def support_points(self, begin: int = 0, end: int | None = None) -> np.ndarray:
        end = self.n_dofs() if end is None else end
        out = np.full((end - begin, 3), np.nan)
        check(lib.mfhn_dofs_support_points(self._h, begin, end, _ptr(out)))
        return out


class Partitioner:
    """Utilities::MPI::Partitioner analogue: owned range, sorted ghost indices,
    per-peer ghost ranges and import index lists."""

    def __init__(self, rank, n_ranks, owned_range, ghost_global, rank_begin):
        self.rank, self.n_ranks = rank, n_ranks
        self.begin, self.end = owned_range
        self.n_owned = self.end - self.begin
        self.ghost_global = ghost_global
        self.n_ghost = len(ghost_global)
        self.rank_begin = np.asarray(rank_begin, dtype=np.int64)
        owner = np.searchsorted(rank_begin, ghost_global, side="right") - 1
        self.ghost_owner = owner.astype(np.int32)
        # ghosts are sorted by global index, owners' ranges are ascending => contiguous per peer
        self.ghost_ranges = {}
        for p in np.unique(owner):
            w = np.nonzero(owner == p)[0]
            self.ghost_ranges[int(p)] = (int(w[0]), int(w[-1]) + 1)
        self.import_indices = {}  # peer -> local owned indices the peer reads (set by exchange)

    def n_ghost_indices(self):
        return self.n_ghost

    def n_import_indices(self):
        return int(sum(len(v) for v in self.import_indices.values()))

    def requests(self):
        """peer -> global indices this rank needs from it."""
        return {p: self.ghost_global[a:b] for p, (a, b) in self.ghost_ranges.items()}

    def set_imports(self, requests_from_peers):
        """requests_from_peers: peer -> global indices (owned here) that the peer ghosts."""
        self.import_indices = {int(p): (np.asarray(g, dtype=np.int64) - self.begin).astype(np.int32) for p, g in requests_from_peers.items() if len(g)}
        for v in self.import_indices.values():
            assert v.min() >= 0 and v.max() < self.n_owned


class MatrixFree:
    """Rank-local setup product of MatrixFree::reinit (benchmark_03.h:338-339):
    cell order (Morton, interior cells first), rank-local substituted DoF
    indices, compressed masks, Cartesian geometry and the ghost partitioner."""

    def __init__(self, dof_handler: DoFHandler, rank: int = 0, categorize: bool = True):
        self.dof_handler, self.rank = dof_handler, rank
        dh = dof_handler
        self.degree = dh.degree
        cells = dh.cells_of_rank(rank)
        pos = dh.tria.morton_position()
        cells = cells[np.argsort(pos[cells], kind="stable")]
        _, sub, masks, h = dh.fill(cells)
        b, e = dh.owned_range(rank)
        rank_begin = np.array([dh.owned_range(r)[0] for r in range(dh.n_ranks)], dtype=np.int64)
        flat = sub.reshape(-1).astype(np.int64)
        is_ghost = (flat < b) | (flat >= e)
        ghost_global = np.unique(flat[is_ghost])
        local = flat - b
        if len(ghost_global):
            local[is_ghost] = (e - b) + np.searchsorted(ghost_global, flat[is_ghost])
        assert local.max(initial=0) < 2 ** 32
        local = local.reshape(sub.shape)
        # cells touching ghost entries go last: [interior | boundary]
        touches = is_ghost.reshape(sub.shape).any(axis=1)
        order = np.concatenate([np.nonzero(~touches)[0], np.nonzero(touches)[0]])
        self.n_interior_cells = int((~touches).sum())
        if dh.n_ranks > 1:
            # partition boundaries on whole warp batches of the cell kernels (16, 10, 8, 6, 5, 4 cells per warp for
            # k = 1..6; lcm 240): the last few interior cells simply join the boundary partition
            self.n_interior_cells -= self.n_interior_cells % 240
        if categorize:
            # the reference's Categorize option (benchmark_01.h:258-284: cell_vectorization_category =
            # constraint mask): inside windows of the Morton order, cells are grouped by their constraint
            # mask, so that fewer warps pay for the interpolation and a warp holds few different
            # constraint kinds (= few different interpolation passes); locality is kept at window scale.
            # Measured on B200 (k=4 / k=5): window 240 by flag 84.3 / 82.0, by kind 86.1 / 89.4,
            # window 960 by kind 86.9 / 93.0, window 3840 by kind 87.2 / 93.8 GDoF/s (plane kernel); bulk-copy
            # kernel at k=4: 960: 95.1, 3840: 95.7, 15360: 93.5.
            w = int(os.environ.get("MFHN_CATEGORIZE_WINDOW", "3840"))
            key = masks[order].astype(np.int64) if os.environ.get("MFHN_CATEGORIZE_BY_KIND", "1") == "1" else (masks[order] != 0).astype(np.int64)
            seg = (np.arange(len(order)) >= self.n_interior_cells).astype(np.int64)
            pos = np.arange(len(order))
            window = np.where(seg == 0, pos, pos - self.n_interior_cells) // w
            order = order[np.lexsort((pos, key, window, seg))]
        # deal.II's cell_loop overlaps the two ghost exchanges with two interior partitions
        self.n_interior_a = (self.n_interior_cells // 2) // 240 * 240 if dh.n_ranks > 1 else self.n_interior_cells
        self.cell_ids = cells[order]
        self.dof_indices = np.ascontiguousarray(local[order].astype(np.uint32))
        self.masks = np.ascontiguousarray(masks[order])
        self.h = np.ascontiguousarray(h[order])
        self.n_cells = len(cells)
        self.partitioner = Partitioner(rank, dh.n_ranks, (b, e), ghost_global, rank_begin)

    def n_cells_hn(self):
        return int((self.masks != 0).sum())


def exchange_import_indices(partitioner: Partitioner, group=None):
    """Setup-time all-to-all of the ghost requests (torch.distributed)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    req = partitioner.requests()
    send = [req.get(p, np.empty(0, dtype=np.int64)) for p in range(world)]
    recv = [None] * world
    gathered = [None] * world
    dist.all_gather_object(gathered, send, group=group)
    me = dist.get_rank(group)
    partitioner.set_imports({p: gathered[p][me] for p in range(world) if p != me})


_TORCH_DTYPE = {}


def _torch():
    import torch

    if not _TORCH_DTYPE:
        _TORCH_DTYPE[capi.F64] = torch.float64
        _TORCH_DTYPE[capi.F32] = torch.float32
    return torch


class LaplaceOperator:
    """LaplaceOperator<3, degree, Number, MemorySpace::CUDA> (benchmark_03.h:319-357).

    vmult(dst, src) ACCUMULATES into dst exactly like the reference's cell_loop
    call (benchmark_03.h:352 passes no zero flag); pass zero_dst=True to clear
    dst first.  Vectors are torch CUDA tensors of n_owned + n_ghost entries."""

    def __init__(self, matrix_free: MatrixFree, number="double", apply_constraints=True, kernel="auto", device=None, geometry=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("LaplaceOperator needs a CUDA device: there is no CPU fallback")
        self.mf = matrix_free
        self.number = {"double": capi.F64, "float": capi.F32}[number]
        self.dtype = _TORCH_DTYPE[self.number]
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        part = matrix_free.partitioner
        kern = capi.KERNELS[kernel]
        cuts = sorted({0, matrix_free.n_interior_a, matrix_free.n_interior_cells} - {matrix_free.n_cells})
        self._segments = np.array(cuts, dtype=np.int64)
        # the C ABI trusts the sizes it is given: check the host arrays before handing over raw pointers
        n3 = (matrix_free.degree + 1) ** 3
        for name, arr, shape, dtype in (("dof_indices", matrix_free.dof_indices, (matrix_free.n_cells, n3), np.uint32),
                                        ("masks", matrix_free.masks, (matrix_free.n_cells,), np.uint8),
                                        ("h", matrix_free.h, (matrix_free.n_cells,), np.float64)):
            if arr.shape != shape or arr.dtype != dtype or not arr.flags["C_CONTIGUOUS"]:
                raise capi.MfhnError(1, f"{name} must be a C-contiguous {np.dtype(dtype).name} array of shape {shape}")
        if geometry is None:
            gtype, geom = capi.GEOM_CARTESIAN, matrix_free.h
        elif np.ndim(geometry) == 3 and np.shape(geometry)[1:] == (3, 3):  # (n_cells, 3, 3) Jacobians
            gtype, geom = capi.GEOM_AFFINE, np.ascontiguousarray(geometry, dtype=np.float64).reshape(-1, 9)
            assert geom.shape[0] == matrix_free.n_cells
        else:  # (n_cells, 6, (k+1)^3): JxW J^-1 J^-T per quadrature point
            gtype, geom = capi.GEOM_GENERAL, np.ascontiguousarray(geometry, dtype=np.float64)
            if geom.shape != (matrix_free.n_cells, 6, n3):
                raise capi.MfhnError(1, f"general geometry must have shape {(matrix_free.n_cells, 6, n3)}")
        desc = capi.OpDesc(
            degree=matrix_free.degree, number=self.number, n_cells=matrix_free.n_cells, n_owned=part.n_owned,
            n_ghost=part.n_ghost, dof_indices=_ptr(matrix_free.dof_indices), masks=_ptr(matrix_free.masks),
            geometry_type=gtype, geometry=_ptr(geom), apply_constraints=int(apply_constraints), kernel=kern,
            device=self.device.index, segments=_ptr(self._segments), n_segments=len(self._segments))
        h = C.c_void_p()
        check(lib.mfhn_op_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.n_owned, self.n_ghost = part.n_owned, part.n_ghost
        self._comm = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib.mfhn_op_destroy(self._h)
            self._h = None

    # -- reference surface ---------------------------------------------------
    def initialize_dof_vector(self):
        return _torch().zeros(self.n_owned + self.n_ghost, dtype=self.dtype, device=self.device)

    def vmult(self, dst, src, zero_dst=False):
        torch = _torch()
        self._check_vec(dst), self._check_vec(src)
        if dst.data_ptr() == src.data_ptr():
            raise capi.MfhnError(1, "dst and src must not alias")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self._comm is None:
            check(lib.mfhn_op_vmult(self._h, dst.data_ptr(), src.data_ptr(), stream, int(zero_dst)))
        else:
            self._comm.vmult(self, dst, src, zero_dst)

    def vmult_host(self, dst_host, src_host, zero_dst=True, slot=0):
        """vmult on HOST vectors (torch CPU tensors, ideally pinned): H2D, kernel, D2H,
        all stream-ordered on the current stream.  Calls with different `slot` (0/1) on
        different streams may overlap (upload of one application, download of the other)."""
        torch = _torch()
        for v in (dst_host, src_host):
            if v.is_cuda or v.dtype != self.dtype or not v.is_contiguous() or v.numel() != self.n_owned + self.n_ghost:
                raise capi.MfhnError(1, "host vector has the wrong type or size")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.mfhn_op_vmult_host_slot(self._h, dst_host.data_ptr(), src_host.data_ptr(), stream, int(zero_dst), slot))

    # -- switches / queries ----------------------------------------------------
    def set_apply_constraints(self, flag: bool):
        check(lib.mfhn_op_set_apply_constraints(self._h, int(flag)))

    def set_kernel(self, kernel: str):
        check(lib.mfhn_op_set_kernel(self._h, capi.KERNELS[kernel]))

    def query(self, what: str) -> float:
        v = C.c_double()
        check(lib.mfhn_op_query(self._h, what.encode(), C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        return int(lib.mfhn_op_launch_count(self._h))

    def vmult_range(self, dst, src, cell_begin, cell_end):
        torch = _torch()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.mfhn_op_vmult_range(self._h, dst.data_ptr(), src.data_ptr(), stream, cell_begin, cell_end))

    def compute_diagonal(self):
        """Diagonal of the operator as a vector (hanging entries are zero).  Extension for a
        point-Jacobi preconditioner (BASELINE.json config 5); no counterpart in the reference."""
        torch = _torch()
        diag = self.initialize_dof_vector()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.mfhn_op_diagonal(self._h, diag.data_ptr(), stream))
        if self._comm is not None:
            self._comm.compress_add(diag)
        return diag

    def apply_hanging_node_constraints(self, cell_values, transpose: bool):
        """FEEvaluationHangingNodesFactory::apply on [n_cells, (k+1)^3] values
        (benchmark_00_likwid.cc:56-59)."""
        torch = _torch()
        assert cell_values.is_cuda and cell_values.dtype == self.dtype and cell_values.is_contiguous()
        assert cell_values.numel() == self.mf.n_cells * (self.mf.degree + 1) ** 3
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.mfhn_op_apply_hn(self._h, cell_values.data_ptr(), int(transpose), stream))

    def _check_vec(self, v):
        if not (v.is_cuda and v.dtype == self.dtype and v.is_contiguous() and v.numel() == self.n_owned + self.n_ghost):
            raise capi.MfhnError(1, "vector must come from initialize_dof_vector()")

    def attach_communicator(self, comm):
        self._comm = comm


def bench_fma(number="double", iters=20000) -> float:
    v = C.c_double()
    check(lib.mfhn_bench_dfma({"double": 0, "float": 1}[number], iters, C.byref(v)))
    return v.value
