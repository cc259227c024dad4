"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from
the independent general-purpose oracle O1) against: the fast-algorithm oracle,
its C restatement, the C++ setup (bit-exact indices / masks) and -- with a GPU --
the CUDA kernels through the C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import cpu, dofs, operators

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
assert FIXTURES, "golden fixtures missing"


def _layout(g):
    lay = dofs.DoFLayout()
    lay.degree, lay.n_cells, lay.n_dofs = int(g["degree"]), len(g["cells"]), int(g["n_dofs"])
    lay.dof_indices, lay.masks, lay.h = g["dof_indices"], g["masks"], g["h"]
    lay.kinds = np.array([dofs.decompress(int(m)) for m in g["masks"]], dtype=np.uint16)
    return lay


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_oracles_reproduce_golden(path):
    g = np.load(path)
    lay = _layout(g)
    for v in ("sin", "rnd"):
        ref = g[f"dst_{v}"]
        y = operators.vmult_fast(lay, g[f"src_{v}"])
        assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12
        yc = cpu.vmult(lay.degree, lay.dof_indices, lay.masks, lay.h, g[f"src_{v}"])
        assert np.abs(yc - ref).max() / np.abs(ref).max() < 1e-12


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_cpp_setup_reproduces_golden_bit_exact(path, mfhn):
    g = np.load(path)
    tria = mfhn.Triangulation(str(g["geometry"]), int(g["n_refinements"]), str(g["flavour"]))
    assert np.array_equal(tria.cells(), g["cells"])
    dh = mfhn.DoFHandler(tria, int(g["degree"]))
    assert dh.n_dofs() == int(g["n_dofs"])
    raw, sub, masks, h = dh.fill(np.arange(tria.n_active_cells()), raw=True)
    assert np.array_equal(raw, g["raw_indices"])
    assert np.array_equal(sub, g["dof_indices"])
    assert np.array_equal(masks, g["masks"])
    assert np.array_equal(h, g["h"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
@pytest.mark.parametrize("kernel", ["auto", "qpoint", "separable", "plane", "bulk", "baseline"])
def test_cuda_reproduces_golden(path, kernel, mfhn):
    import torch

    g = np.load(path)
    if kernel == "bulk" and not 3 <= int(g["degree"]) <= 5:
        pytest.skip("the bulk-copy kernel covers degrees 3..5")
    tria = mfhn.Triangulation(str(g["geometry"]), int(g["n_refinements"]), str(g["flavour"]))
    dh = mfhn.DoFHandler(tria, int(g["degree"]))
    mf = mfhn.MatrixFree(dh)
    op = mfhn.LaplaceOperator(mf, kernel=kernel)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    for v in ("sin", "rnd"):
        src.copy_(torch.from_numpy(g[f"src_{v}"]))
        op.vmult(dst, src, zero_dst=True)
        ref = g[f"dst_{v}"]
        assert np.abs(dst.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-12
        # hanging entries of dst are never written by the fast algorithm
        assert dst.cpu().numpy()[g["is_hanging"]].max(initial=0) == 0
