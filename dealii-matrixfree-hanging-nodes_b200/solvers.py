"""Conjugate gradients with a point-Jacobi preconditioner on top of the operator (BASELINE.json config 5).
EXTENSION: the reference contains no solver; this is the step either side of vmult in a real
solve.  The whole iteration runs inside the library (mfhn_cg_solve): vmult, two fused vector kernels and one
batched three-scalar all-reduce per iteration (Chronopoulos / Gear form); this module only passes pointers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import check, lib


def inverse_diagonal(op):
    """1 / diag(A) on the owned entries (0 on hanging entries): the point-Jacobi preconditioner."""
    import torch

    inv = op.initialize_dof_vector()
    stream = torch.cuda.current_stream(op.device).cuda_stream
    native = getattr(op._comm, "_native", None) if op._comm is not None else None
    check(lib.mfhn_op_inverse_diagonal(op._h, native, inv.data_ptr(), stream))
    return inv


def solve_cg(op, x, b, diag=None, rel_tol=1e-8, max_iter=1000, check_every=1, timings=False, inverse=None):
    """Solves A x = b for the (singular, positive semi-definite) Laplace operator `op` starting from x; b must be
    consistent (orthogonal to the constants).  diag: the operator's diagonal (compute_diagonal) or None; inverse: an
    already inverted diagonal (inverse_diagonal).  Returns (iterations, residual history); with timings=True a third
    entry: dict of the device time split."""
    import torch

    op._check_vec(x), op._check_vec(b)
    inv = inverse
    if inv is None and diag is not None:
        inv = op.initialize_dof_vector()
        inv.copy_(torch.where(diag != 0, 1.0 / torch.where(diag != 0, diag, torch.ones_like(diag)), torch.zeros_like(diag)))
    if inv is not None:
        op._check_vec(inv)
    native = None
    if op._comm is not None:
        native = getattr(op._comm, "_native", None)
        if native is None:
            raise capi.MfhnError(1, "solve_cg on a partitioned operator needs the native (NCCL) exchange")
    opt = capi.CgOptions(max_iter=int(max_iter), rel_tol=float(rel_tol), check_every=int(check_every), timings=int(timings))
    res = capi.CgResult()
    hist = np.zeros(max_iter + 1)
    stream = torch.cuda.current_stream(op.device).cuda_stream
    check(lib.mfhn_cg_solve(op._h, native, x.data_ptr(), b.data_ptr(), inv.data_ptr() if inv is not None else None, C.byref(opt), C.byref(res),
                            hist.ctypes.data_as(C.c_void_p), stream))
    history = [float(v) for v in hist[:res.iterations + 1]]
    if timings:
        return res.iterations, history, {"ms_total": res.ms_total, "ms_vmult": res.ms_vmult, "ms_vector_ops": res.ms_vector_ops,
                                         "ms_allreduce": res.ms_allreduce}
    return res.iterations, history
