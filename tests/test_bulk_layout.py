"""Host-side layout of MFHN_KERNEL_BULK (kernels_bulk.cuh): the block descriptors, the line / vertex index
lists and the staging positions must reproduce the reference index array (read_dof_values order,
benchmark_03.h:255-258) exactly -- checked by an emulated gather with src[i] = i, no GPU needed."""
import ctypes as C
import importlib

import numpy as np
import pytest

capi = importlib.import_module("dealii-matrixfree-hanging-nodes_b200._capi")


def _check(k, number, idx, n_vec):
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    ni, nm = C.c_int64(-1), C.c_int64(-1)
    rc = capi.lib.mfhn_bulk_layout_check(k, number, idx.shape[0], n_vec, idx.ctypes.data, C.byref(ni), C.byref(nm))
    return rc, ni.value, nm.value


@pytest.mark.parametrize("geo,L", [("quadrant", 3), ("annulus", 5), ("step", 3)])
@pytest.mark.parametrize("k", [3, 4, 5])
@pytest.mark.parametrize("n_ranks", [1, 3])
def test_layout_reproduces_index_array(mfhn, geo, L, k, n_ranks):
    tria = mfhn.Triangulation(geo, L, "p4est")
    dh = mfhn.DoFHandler(tria, k) if n_ranks == 1 else mfhn.DoFHandler(tria, k, n_ranks, tria.partition(n_ranks))
    for rank in range(n_ranks):
        mf = mfhn.MatrixFree(dh, rank=rank)
        n_vec = mf.partitioner.n_owned + mf.partitioner.n_ghost
        for number in (capi.F64, capi.F32):
            rc, n_irregular, n_mismatch = _check(k, number, mf.dof_indices, n_vec)
            assert rc == 0 and n_mismatch == 0
            # only blocks whose widened range would leave the vector are left to the plane kernel
            assert 0 <= n_irregular <= 8


def test_other_numberings_are_flagged_not_mangled(mfhn):
    """A numbering without contiguous blocks (random permutation of the DoFs) must be detected: every cell
    irregular, nothing mis-gathered."""
    tria = mfhn.Triangulation("quadrant", 3, "serial")
    k = 4
    dh = mfhn.DoFHandler(tria, k)
    mf = mfhn.MatrixFree(dh)
    perm = np.random.default_rng(1).permutation(dh.n_dofs()).astype(np.uint32)
    rc, n_irregular, n_mismatch = _check(k, capi.F64, perm[mf.dof_indices], dh.n_dofs())
    assert rc == 0 and n_mismatch == 0 and n_irregular == mf.n_cells
    # degrees outside 3..5 have no bulk layout
    rc, _, _ = _check(2, capi.F64, mf.dof_indices[:, :27], dh.n_dofs())
    assert rc != 0
