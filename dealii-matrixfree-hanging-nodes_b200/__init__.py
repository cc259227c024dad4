"""B200-native matrix-free hanging-node Laplace operator engine.

The directory name follows the reference repository and is not a valid Python
identifier; import it with

    import importlib
    mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")

The product is libmfhn.so (CUDA, sm_100a) behind the C ABI of include/mfhn.h;
this package is the ctypes binding plus the host-side mirror of the reference's
operator interface.  Importing fails loudly if the library has not been built.
"""
from . import _capi as capi
from ._capi import MfhnError, NotImplementedMfhn
from .api import (DoFHandler, LaplaceOperator, MatrixFree, Partitioner, Triangulation, bench_fma,
                  exchange_import_indices, exchange_local)
from .solvers import inverse_diagonal, solve_cg

__all__ = ["capi", "MfhnError", "NotImplementedMfhn", "Triangulation", "DoFHandler", "MatrixFree",
           "Partitioner", "LaplaceOperator", "bench_fma", "exchange_import_indices", "exchange_local", "solve_cg", "inverse_diagonal"]
