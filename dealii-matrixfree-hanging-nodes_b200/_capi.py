"""ctypes binding of the C ABI declared in include/mfhn.h (libmfhn.so).

The shared library is the product; there is no Python or CPU fallback.  If it
is missing, importing this module raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFHN_LIB") or os.path.join(_HERE, "libmfhn.so")  # MFHN_LIB: development builds side by side

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python -c 'import __graft_entry__ as g; g.build()' or make -C <package>/csrc)"
    )

lib = C.CDLL(LIB_PATH)

c_void_p, c_int, c_int64, c_double, c_char_p = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_char_p
P = C.POINTER


class OpDesc(C.Structure):
    _fields_ = [
        ("degree", c_int),
        ("number", c_int),
        ("n_cells", c_int64),
        ("n_owned", c_int64),
        ("n_ghost", c_int64),
        ("dof_indices", c_void_p),
        ("masks", c_void_p),
        ("geometry_type", c_int),
        ("geometry", c_void_p),
        ("apply_constraints", c_int),
        ("kernel", c_int),
        ("device", c_int),
        ("segments", c_void_p),
        ("n_segments", c_int),
        ("vector_padding", c_int),
    ]


class MfOptions(C.Structure):
    _fields_ = [("rank", c_int), ("categorize", c_int), ("window", c_int), ("batch_alignment", c_int)]


class MfSizes(C.Structure):
    _fields_ = [("degree", c_int), ("rank", c_int), ("n_ranks", c_int), ("n_cells", c_int64), ("n_cells_hn", c_int64),
                ("n_owned", c_int64), ("n_ghost", c_int64), ("owned_begin", c_int64), ("n_interior_a", c_int64),
                ("n_interior", c_int64), ("n_ghost_peers", c_int), ("n_import_peers", c_int), ("n_import", c_int64)]


class CgOptions(C.Structure):
    _fields_ = [("max_iter", c_int), ("rel_tol", c_double), ("check_every", c_int), ("timings", c_int)]


class CgResult(C.Structure):
    _fields_ = [("iterations", c_int), ("initial_residual", c_double), ("final_residual", c_double), ("ms_total", c_double),
                ("ms_vmult", c_double), ("ms_vector_ops", c_double), ("ms_allreduce", c_double)]


class DistDesc(C.Structure):
    _fields_ = [
        ("rank", c_int),
        ("world", c_int),
        ("unique_id", c_void_p),
        ("n_import_peers", c_int),
        ("import_peers", c_void_p),
        ("import_offsets", c_void_p),
        ("import_indices", c_void_p),
        ("n_ghost_peers", c_int),
        ("ghost_peers", c_void_p),
        ("ghost_begin", c_void_p),
        ("ghost_end", c_void_p),
        ("segments", c_int64 * 4),
    ]


# every symbol include/mfhn.h declares: (restype, argtypes)
SIGNATURES = {
    "mfhn_last_error": (c_char_p, []),
    "mfhn_version": (c_char_p, []),
    "mfhn_mesh_create": (c_int, [c_char_p, c_int, c_int, P(c_void_p)]),
    "mfhn_mesh_destroy": (None, [c_void_p]),
    "mfhn_mesh_n_cells": (c_int64, [c_void_p]),
    "mfhn_mesh_n_levels": (c_int, [c_void_p]),
    "mfhn_mesh_cells": (c_int, [c_void_p, c_void_p]),
    "mfhn_mesh_n_cells_hn": (c_int64, [c_void_p]),
    "mfhn_mesh_partition": (c_int, [c_void_p, c_int, c_double, c_void_p]),
    "mfhn_mesh_morton_position": (c_int, [c_void_p, c_void_p]),
    "mfhn_dofs_create": (c_int, [c_void_p, c_int, c_int, c_void_p, P(c_void_p)]),
    "mfhn_dofs_destroy": (None, [c_void_p]),
    "mfhn_dofs_n_dofs": (c_int64, [c_void_p]),
    "mfhn_dofs_owned_range": (c_int, [c_void_p, c_int, P(c_int64), P(c_int64)]),
    "mfhn_dofs_n_cells_of_rank": (c_int64, [c_void_p, c_int]),
    "mfhn_dofs_cells_of_rank": (c_int, [c_void_p, c_int, c_void_p]),
    "mfhn_dofs_fill": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mfhn_dofs_support_points": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "mfhn_mf_create": (c_int, [c_void_p, P(MfOptions), P(c_void_p)]),
    "mfhn_mf_destroy": (None, [c_void_p]),
    "mfhn_mf_info": (c_int, [c_void_p, P(MfSizes)]),
    "mfhn_mf_arrays": (c_int, [c_void_p] + [P(c_void_p)] * 6),
    "mfhn_mf_partitioner": (c_int, [c_void_p] + [P(c_void_p)] * 7),
    "mfhn_mf_set_imports": (c_int, [c_void_p, c_int, c_void_p, c_int64]),
    "mfhn_mf_exchange_local": (c_int, [P(c_void_p), c_int]),
    "mfhn_compress": (C.c_uint8, [C.c_uint16]),
    "mfhn_decompress": (C.c_uint16, [C.c_uint8]),
    "mfhn_check_kind": (c_int, [C.c_uint16]),
    "mfhn_op_create": (c_int, [P(OpDesc), P(c_void_p)]),
    "mfhn_op_create_mf": (c_int, [c_void_p, c_int, c_int, c_int, c_int, P(c_void_p)]),
    "mfhn_op_create_mf_padded": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, P(c_void_p)]),
    "mfhn_op_destroy": (None, [c_void_p]),
    "mfhn_op_vmult": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "mfhn_op_vmult_range": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64]),
    "mfhn_op_vmult_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "mfhn_op_vmult_host_slot": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "mfhn_op_diagonal": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mfhn_op_set_apply_constraints": (c_int, [c_void_p, c_int]),
    "mfhn_op_set_kernel": (c_int, [c_void_p, c_int]),
    "mfhn_op_apply_hn": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "mfhn_op_dg_copy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mfhn_op_set_hn_strategy": (c_int, [c_void_p, c_int]),
    "mfhn_op_query": (c_int, [c_void_p, c_char_p, P(c_double)]),
    "mfhn_op_launch_count": (c_int64, [c_void_p]),
    "mfhn_pack": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mfhn_unpack_add": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mfhn_bench_dfma": (c_int, [c_int, c_int, P(c_double)]),
    "mfhn_bulk_layout_check": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, P(c_int64), P(c_int64)]),
    "mfhn_runs_layout_check": (c_int, [c_int, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_int, P(c_int64), P(c_int64), P(c_int64), P(c_int64)]),
    "mfhn_dist_unique_id": (c_int, [c_void_p]),
    "mfhn_dist_create": (c_int, [c_void_p, P(DistDesc), P(c_void_p)]),
    "mfhn_dist_create_mf": (c_int, [c_void_p, c_void_p, c_void_p, P(c_void_p)]),
    "mfhn_dist_destroy": (None, [c_void_p]),
    "mfhn_dist_vmult": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "mfhn_dist_launch_count": (c_int64, [c_void_p]),
    "mfhn_op_inverse_diagonal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mfhn_cg_solve": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(CgOptions), P(CgResult), c_void_p, c_void_p]),
    "mfhn_vec_alloc": (c_int, [c_int64, P(c_void_p)]),
    "mfhn_vec_free": (c_int, [c_void_p]),
    "mfhn_ipc_get_handle": (c_int, [c_void_p, c_void_p]),
    "mfhn_ipc_open_handle": (c_int, [c_void_p, P(c_void_p)]),
    "mfhn_ipc_close_handle": (c_int, [c_void_p]),
    "mfhn_dist_enable_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mfhn_dist_enable_peer_flags": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mfhn_dist_vmult_peer": (c_int, [c_void_p, c_void_p, c_int]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args


class MfhnError(RuntimeError):
    """Non-zero status from the C ABI (the reference's AssertThrow analogue)."""

    def __init__(self, status, message):
        super().__init__(f"mfhn error {status}: {message}")
        self.status = status


class NotImplementedMfhn(MfhnError):
    pass


def check(status: int):
    if status != 0:
        msg = lib.mfhn_last_error().decode()
        raise (NotImplementedMfhn if status == 3 else MfhnError)(status, msg)


F64, F32 = 0, 1
VECTOR_PADDING = 4  # MFHN_VECTOR_PADDING: spare entries behind every vector handed to a padded operator
SERIAL, P4EST = 0, 1
GEOM_CARTESIAN, GEOM_AFFINE, GEOM_GENERAL = 0, 1, 2
KERNEL_AUTO, KERNEL_QPOINT, KERNEL_SEPARABLE, KERNEL_BASELINE, KERNEL_PLANE, KERNEL_PATCH = 0, 1, 2, 3, 4, 5
KERNELS = {"auto": 0, "qpoint": 1, "separable": 2, "baseline": 3, "plane": 4, "patch": 5, "bulk": 6, "runs": 7, "qpoint_rows": 8}
