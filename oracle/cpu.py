"""ctypes front end of oracle/laplace_cpu.c (oracle / CPU baseline; test
infrastructure only, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import fe1d

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.oracle_vmult.restype = C.c_int
        _LIB.oracle_benchmark.restype = C.c_double
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _shape_args(degree):
    sd = fe1d.shape_data(degree)
    return [np.ascontiguousarray(x) for x in (sd.S, sd.Dc, sd.W[0], sd.qw)]


def vmult(degree, dof_indices, masks, h, src, dst=None, apply_constraints=True):
    """dst += A src (single thread), arrays as in oracle.dofs.DoFLayout."""
    idx = np.ascontiguousarray(dof_indices, dtype=np.uint32)
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    h = np.ascontiguousarray(h, dtype=np.float64)
    src = np.ascontiguousarray(src, dtype=np.float64)
    if dst is None:
        dst = np.zeros_like(src)
    sh = _shape_args(degree)
    rc = _lib().oracle_vmult(C.c_int(degree), C.c_long(idx.shape[0]), _p(idx), _p(masks), _p(h), *[_p(s) for s in sh],
                             _p(src), _p(dst), C.c_int(int(apply_constraints)))
    if rc != 0:
        raise ValueError("unsupported degree")
    return dst


def benchmark(degree, dof_indices, masks, h, n_dofs, apply_constraints=True, n_rep=10, n_threads=None):
    """Mean seconds per vmult with benchmark_01's throughput semantics."""
    idx = np.ascontiguousarray(dof_indices, dtype=np.uint32)
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    h = np.ascontiguousarray(h, dtype=np.float64)
    n_threads = n_threads or os.cpu_count()
    sh = _shape_args(degree)
    return float(_lib().oracle_benchmark(C.c_int(degree), C.c_long(idx.shape[0]), C.c_long(n_dofs), _p(idx), _p(masks), _p(h),
                                         *[_p(s) for s in sh], C.c_int(int(apply_constraints)), C.c_int(n_rep), C.c_int(n_threads)))
