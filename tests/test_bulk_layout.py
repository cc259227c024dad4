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


def _check_runs(k, number, idx, n_vec, max_gap, min_run, place=1, wavefronts=False):
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    nb, ns, nm, wf = C.c_int64(-1), C.c_int64(-1), C.c_int64(-1), C.c_int64(-1)
    rc = capi.lib.mfhn_runs_layout_check(k, number, idx.shape[0], n_vec, idx.ctypes.data, max_gap, min_run, place, C.byref(nb), C.byref(ns), C.byref(nm),
                                         C.byref(wf))
    return (rc, nb.value, ns.value, nm.value, wf.value) if wavefronts else (rc, nb.value, ns.value, nm.value)


@pytest.mark.parametrize("geo,L", [("quadrant", 3), ("annulus", 5)])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("n_ranks", [1, 3])
def test_runs_layout_reproduces_index_array(mfhn, geo, L, k, n_ranks):
    """MFHN_KERNEL_RUNS (kernels_runs.cuh): bulk copies + single entries + position tables reproduce the index array
    for every run-detection setting, and every copied entry outside the cell is on the zero list."""
    tria = mfhn.Triangulation(geo, L, "p4est")
    dh = mfhn.DoFHandler(tria, k) if n_ranks == 1 else mfhn.DoFHandler(tria, k, n_ranks, tria.partition(n_ranks))
    for rank in range(n_ranks):
        mf = mfhn.MatrixFree(dh, rank=rank)
        n_vec = mf.partitioner.n_owned + mf.partitioner.n_ghost
        n3 = (k + 1) ** 3
        for number in (capi.F64, capi.F32):
            for max_gap, min_run in ((0, 6), (0, 1), (3, 4), (10, 6), (1000, 2)):
                rc, n_bulk, n_single, n_mismatch = _check_runs(k, number, mf.dof_indices, n_vec, max_gap, min_run)
                assert rc == 0 and n_mismatch == 0, (k, number, max_gap, min_run, n_mismatch)
                assert n_bulk >= 0 and 0 <= n_single <= mf.n_cells * n3
            # bank-aware placement: same copies and single entries, fewer modelled shared-memory wavefronts
            r0 = _check_runs(k, number, mf.dof_indices, n_vec, 0, 6, place=0, wavefronts=True)
            r1 = _check_runs(k, number, mf.dof_indices, n_vec, 0, 6, place=1, wavefronts=True)
            assert r0[3] == 0 and r1[3] == 0 and r0[1:3] == r1[1:3] and r1[4] <= r0[4]
            if k >= 3:  # the cell interiors alone give one bulk copy per cell
                rc, n_bulk, n_single, _ = _check_runs(k, number, mf.dof_indices, n_vec, 0, 6)
                assert n_bulk >= mf.n_cells - 8 and n_single < mf.n_cells * (n3 - (k - 1) ** 3)


def test_runs_layout_takes_any_numbering(mfhn):
    """A random permutation of the DoFs leaves (almost) no runs: everything becomes single entries, nothing is
    mis-gathered; repeated entries inside a cell get slots of their own."""
    tria = mfhn.Triangulation("quadrant", 3, "serial")
    k = 4
    dh = mfhn.DoFHandler(tria, k)
    mf = mfhn.MatrixFree(dh)
    perm = np.random.default_rng(1).permutation(dh.n_dofs()).astype(np.uint32)
    rc, n_bulk, n_single, n_mismatch = _check_runs(k, capi.F64, perm[mf.dof_indices], dh.n_dofs(), 0, 6)
    assert rc == 0 and n_mismatch == 0 and n_single > 0.95 * mf.n_cells * 125
    dup = mf.dof_indices.copy()
    dup[:, 7] = dup[:, 3]
    dup[:, 100] = dup[:, 3]
    rc, _, _, n_mismatch = _check_runs(k, capi.F64, dup, dh.n_dofs(), 2, 4)
    assert rc == 0 and n_mismatch == 0
    rc, _, _, _ = _check_runs(6, capi.F64, mf.dof_indices, dh.n_dofs(), 0, 6)  # (k+1)^3 > 256: no layout
    assert rc != 0
