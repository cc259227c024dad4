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


def _view(ptr, n, dtype):
    """numpy view of n entries of a host array owned by a C handle (empty array for n == 0)."""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    ct = np.ctypeslib.as_ctypes_type(dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(int(n),))


class Partitioner:
    """Utilities::MPI::Partitioner analogue: owned range, sorted ghost indices, per-peer ghost ranges and
    import index lists.  A view of the data held by the MatrixFree handle (mfhn_mf_partitioner)."""

    def __init__(self, mf_handle=None, keep=None, rank=0, n_ranks=1, owned_range=(0, 0)):
        self._mf, self._keep = mf_handle, keep  # keep: the object that owns the handle
        if mf_handle is None:  # a serial stand-in (no ghosts): used by synthetic layouts in the benchmarks
            self.rank, self.n_ranks = rank, n_ranks
            self.begin, self.end = owned_range
            self.n_owned, self.n_ghost = self.end - self.begin, 0
            self.ghost_global, self.ghost_owner = np.zeros(0, np.int64), np.zeros(0, np.int32)
            self.rank_begin = np.array([self.begin, self.end], dtype=np.int64)
            self.ghost_ranges, self.import_indices = {}, {}
            return
        self._refresh()

    def _refresh(self):
        sz = capi.MfSizes()
        check(lib.mfhn_mf_info(self._mf, C.byref(sz)))
        self.rank, self.n_ranks = sz.rank, sz.n_ranks
        self.begin, self.end = sz.owned_begin, sz.owned_begin + sz.n_owned
        self.n_owned, self.n_ghost = sz.n_owned, sz.n_ghost
        p = [C.c_void_p() for _ in range(6)]
        check(lib.mfhn_mf_arrays(self._mf, *[C.byref(x) for x in p]))
        self.ghost_global = _view(p[4], sz.n_ghost, np.int64)
        self.ghost_owner = _view(p[5], sz.n_ghost, np.int32)
        q = [C.c_void_p() for _ in range(7)]
        check(lib.mfhn_mf_partitioner(self._mf, *[C.byref(x) for x in q]))
        gp = _view(q[0], sz.n_ghost_peers, np.int32)
        gb, ge = _view(q[1], sz.n_ghost_peers, np.int64), _view(q[2], sz.n_ghost_peers, np.int64)
        self.ghost_ranges = {int(gp[i]): (int(gb[i]), int(ge[i])) for i in range(sz.n_ghost_peers)}
        ip = _view(q[3], sz.n_import_peers, np.int32)
        io = _view(q[4], sz.n_import_peers + 1, np.int64)
        ii = _view(q[5], sz.n_import, np.int32)
        self.import_indices = {int(ip[i]): ii[io[i]:io[i + 1]] for i in range(sz.n_import_peers)}
        self.rank_begin = _view(q[6], sz.n_ranks + 1, np.int64)

    def n_ghost_indices(self):
        return self.n_ghost

    def n_import_indices(self):
        return int(sum(len(v) for v in self.import_indices.values()))

    def requests(self):
        """peer -> global indices this rank needs from it."""
        return {p: self.ghost_global[a:b] for p, (a, b) in self.ghost_ranges.items()}

    def set_imports(self, requests_from_peers):
        """requests_from_peers: peer -> global indices (owned here) that the peer ghosts."""
        for p, g in requests_from_peers.items():
            g = np.ascontiguousarray(g, dtype=np.int64)
            check(lib.mfhn_mf_set_imports(self._mf, int(p), _ptr(g), len(g)))
        self._refresh()


class MatrixFree:
    """Rank-local setup product of MatrixFree::reinit (benchmark_03.h:338-339), computed by the library
    (mfhn_mf_create): cell order (Morton, interior cells first, categorised by constraint mask like the reference's
    Categorize option, benchmark_01.h:258-284), rank-local substituted DoF indices, compressed masks, Cartesian
    geometry and the ghost partitioner.  The numpy attributes are views of the handle's host arrays."""

    def __init__(self, dof_handler: DoFHandler, rank: int = 0, categorize: bool = True):
        self.dof_handler, self.rank = dof_handler, rank
        self.degree = dof_handler.degree
        # development switches: window size and grouping key of the categorisation.  Measured on B200 (k=4 / k=5):
        # window 240 by flag 84.3 / 82.0, by kind 86.1 / 89.4, window 960 by kind 86.9 / 93.0, window 3840 by kind
        # 87.2 / 93.8 GDoF/s (plane kernel); bulk-copy kernel at k=4: 960: 95.1, 3840: 95.7, 15360: 93.5.
        by_kind = os.environ.get("MFHN_CATEGORIZE_BY_KIND", "1") == "1"
        opt = capi.MfOptions(rank=rank, categorize=(1 if by_kind else 2) if categorize else 0,
                             window=int(os.environ.get("MFHN_CATEGORIZE_WINDOW", "0")), batch_alignment=0)
        h = C.c_void_p()
        check(lib.mfhn_mf_create(dof_handler._h, C.byref(opt), C.byref(h)))
        self._h = h
        sz = capi.MfSizes()
        check(lib.mfhn_mf_info(h, C.byref(sz)))
        self.n_cells, self.n_interior_cells, self.n_interior_a = sz.n_cells, sz.n_interior, sz.n_interior_a
        self._n_cells_hn = sz.n_cells_hn
        p = [C.c_void_p() for _ in range(6)]
        check(lib.mfhn_mf_arrays(h, *[C.byref(x) for x in p]))
        n3 = (self.degree + 1) ** 3
        self.cell_ids = _view(p[0], sz.n_cells, np.int64)
        self.dof_indices = _view(p[1], sz.n_cells * n3, np.uint32).reshape(sz.n_cells, n3)
        self.masks = _view(p[2], sz.n_cells, np.uint8)
        self.h = _view(p[3], sz.n_cells, np.float64)
        self.partitioner = Partitioner(h)  # a view: valid as long as this object lives

    def __del__(self):
        if getattr(self, "_h", None):
            lib.mfhn_mf_destroy(self._h)
            self._h = None

    def n_cells_hn(self):
        return int(self._n_cells_hn)


def exchange_local(matrix_frees):
    """Import lists of all ranks when every rank's MatrixFree lives in this process (tests, one process driving
    several devices): mfhn_mf_exchange_local."""
    arr = (C.c_void_p * len(matrix_frees))(*[mf._h for mf in matrix_frees])
    check(lib.mfhn_mf_exchange_local(arr, len(matrix_frees)))
    for mf in matrix_frees:
        mf.partitioner._refresh()


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
            device=self.device.index, segments=_ptr(self._segments), n_segments=len(self._segments),
            vector_padding=capi.VECTOR_PADDING)
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
        """n_owned + n_ghost entries (LinearAlgebra::distributed::Vector layout, benchmark_03.h:342-346) with
        MFHN_VECTOR_PADDING spare entries behind them in the same allocation: the bulk-copy kernel moves 16-byte
        aligned ranges, and the range of the block that ends the vector reaches past the last entry."""
        n = self.n_owned + self.n_ghost
        return _torch().zeros(n + capi.VECTOR_PADDING, dtype=self.dtype, device=self.device)[:n]

    def _padded(self, v):
        return v.untyped_storage().nbytes() >= (v.storage_offset() + v.numel() + capi.VECTOR_PADDING) * v.element_size()

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

    def set_hn_strategy(self, strategy: str):
        """"branch" (default: a warp interpolates only if one of its cells is constrained) or "mask" (every warp takes the
        passes, per-lane predicates) -- the reference's vectorisation types, benchmark_01.cc:70-116."""
        check(lib.mfhn_op_set_hn_strategy(self._h, {"branch": 0, "mask": 1}[strategy]))

    def dg_copy(self, dst_cells, src_cells):
        """"DG (C)" stage (benchmark_01.cc:189-199): dst_cells += [W^T W] src_cells on cell-local [n_cells, (k+1)^3] arrays."""
        torch = _torch()
        n = self.mf.n_cells * (self.mf.degree + 1) ** 3
        for v in (dst_cells, src_cells):
            assert v.is_cuda and v.dtype == self.dtype and v.is_contiguous() and v.numel() == n
        check(lib.mfhn_op_dg_copy(self._h, dst_cells.data_ptr(), src_cells.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))

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
        self._check_vec(dst), self._check_vec(src)
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
        if not (v.is_cuda and v.dtype == self.dtype and v.is_contiguous() and v.numel() == self.n_owned + self.n_ghost and self._padded(v)):
            raise capi.MfhnError(1, "vector must come from initialize_dof_vector() (n_owned + n_ghost entries plus padding)")

    def attach_communicator(self, comm):
        self._comm = comm


def bench_fma(number="double", iters=20000) -> float:
    v = C.c_double()
    check(lib.mfhn_bench_dfma({"double": 0, "float": 1}[number], iters, C.byref(v)))
    return v.value
