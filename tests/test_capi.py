"""The C-ABI library loads on a machine without a GPU and exports every symbol
include/mfhn.h declares; host-side error behaviour mirrors the reference's
AssertThrow paths.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mfhn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfhn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(mfhn):
    names = _declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(mfhn.capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libmfhn.so does not export {n}"
    # and the binding covers all of them
    assert set(names) == set(mfhn.capi.SIGNATURES)


def test_version_and_error_reporting(mfhn):
    assert mfhn.capi.lib.mfhn_version().decode().startswith("mfhn")
    with pytest.raises(mfhn.MfhnError, match="Unknown geometry type"):  # benchmark_01.h:217, benchmark_03.h:404
        mfhn.Triangulation("torus", 3)
    with pytest.raises(mfhn.MfhnError):
        mfhn.Triangulation("quadrant", 3, "metis")
    tria = mfhn.Triangulation("quadrant", 2, "serial")
    for bad in (0, 9):  # reference dispatches degrees 1..6 and throws otherwise (benchmark_01.cc:64-66)
        with pytest.raises(mfhn.MfhnError, match="degree"):
            mfhn.DoFHandler(tria, bad)
    dh = mfhn.DoFHandler(tria, 2)
    with pytest.raises(mfhn.MfhnError):
        dh.fill(np.array([tria.n_active_cells()]))
    with pytest.raises(mfhn.MfhnError):
        dh.owned_range(1)
    with pytest.raises(mfhn.MfhnError):
        tria.partition(0)


def test_operator_creation_fails_loudly_without_gpu(mfhn):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    tria = mfhn.Triangulation("quadrant", 2, "serial")
    mf = mfhn.MatrixFree(mfhn.DoFHandler(tria, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mfhn.LaplaceOperator(mf)
    # straight through the C ABI: a CUDA error status, never a silent fallback
    desc = mfhn.capi.OpDesc(degree=2, number=0, n_cells=mf.n_cells, n_owned=mf.partitioner.n_owned, n_ghost=0,
                            dof_indices=mf.dof_indices.ctypes.data, masks=mf.masks.ctypes.data, geometry_type=0,
                            geometry=mf.h.ctypes.data, apply_constraints=1, kernel=0, device=-1, segments=None, n_segments=0)
    h = ctypes.c_void_p()
    assert mfhn.capi.lib.mfhn_op_create(ctypes.byref(desc), ctypes.byref(h)) == 2
    assert b"cuda" in mfhn.capi.lib.mfhn_last_error().lower()


def test_matrix_free_host_layout(mfhn):
    tria = mfhn.Triangulation("annulus", 5, "p4est")
    dh = mfhn.DoFHandler(tria, 2)
    mf = mfhn.MatrixFree(dh, categorize=False)
    assert mf.n_cells == tria.n_active_cells() and mf.n_cells_hn() == tria.n_cells_with_hanging_nodes()
    assert mf.dof_indices.dtype == np.uint32 and mf.dof_indices.shape == (mf.n_cells, 27)
    assert mf.masks.dtype == np.uint8 and mf.partitioner.n_ghost == 0 and mf.partitioner.n_owned == dh.n_dofs()
    # cells are visited along the Morton curve (MatrixFree is free to reorder its batches)
    pos = tria.morton_position()
    assert (np.diff(pos[mf.cell_ids]) > 0).all()
    # Categorize (benchmark_01.h:258-284): same cells, grouped by constraint mask inside Morton windows
    mfc = mfhn.MatrixFree(dh)
    assert sorted(mfc.cell_ids) == sorted(mf.cell_ids) and mfc.n_cells_hn() == mf.n_cells_hn()
    assert np.abs(pos[mfc.cell_ids] - np.arange(mf.n_cells)).max() < 3840  # the window (MFHN_CATEGORIZE_WINDOW)
