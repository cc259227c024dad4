"""Host-side helpers around the benchmark (no GPU): the setup diff tool of the deal.II shim, the high-order-mapping
geometry of bench.py --mapping high-order, the numbering map of the partitioned parity check, the CPU sample."""
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_diff_setup_self_test():
    """tools/diff_setup.py: a dump in another cell order compares equal, a flipped mask bit is found."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "diff_setup.py"), "--self-test"], check=True, capture_output=True, text=True)
    assert "self-test passed" in out.stdout


def test_high_order_geometry_reduces_to_cartesian(mfhn):
    """Zero displacement: JxW J^-1 J^-T = w_q h I (the Cartesian operator); a displacement makes the tensor full and
    keeps it symmetric positive definite (benchmark_01.h:225-242)."""
    from bench_dist import high_order_geometry

    k = 3
    tria = mfhn.Triangulation("annulus", 4, "p4est")
    mf = mfhn.MatrixFree(mfhn.DoFHandler(tria, k))
    q, w = np.polynomial.legendre.leggauss(k + 1)
    w = 0.5 * w
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :]).ravel()
    G0 = high_order_geometry(mfhn, tria, mf, k, amplitude=0.0)
    for comp in (0, 3, 5):
        assert np.abs(G0[:, comp] - mf.h[:, None] * w3[None, :]).max() < 1e-15
    for comp in (1, 2, 4):
        assert np.abs(G0[:, comp]).max() == 0.0
    G = high_order_geometry(mfhn, tria, mf, k, amplitude=1e-2)
    assert min(np.abs(G[:, comp]).max() for comp in (1, 2, 4)) > 1e-6
    M = np.zeros(G.shape[:1] + G.shape[2:] + (3, 3))
    for comp, (i, j) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
        M[..., i, j] = M[..., j, i] = G[:, comp]
    assert np.linalg.eigvalsh(M).min() > 0


def test_partitioned_numbering_map(mfhn):
    """The parity check of bench.py maps the rank-major numbering onto the serial one through the cell-wise index
    arrays: it must be a bijection that preserves the support points."""
    tria = mfhn.Triangulation("annulus", 4, "p4est")
    k, world = 3, 3
    dh1 = mfhn.DoFHandler(tria, k)
    dhp = mfhn.DoFHandler(tria, k, world, tria.partition(world))
    cells = np.arange(tria.n_active_cells())
    r1 = dh1.fill(cells, raw=True, substituted=False, masks=False, h=False)[0]
    rp = dhp.fill(cells, raw=True, substituted=False, masks=False, h=False)[0]
    to_serial = np.full(dhp.n_dofs(), -1, dtype=np.int64)
    to_serial[rp.reshape(-1).astype(np.int64)] = r1.reshape(-1).astype(np.int64)
    assert (to_serial >= 0).all() and len(np.unique(to_serial)) == dh1.n_dofs()
    assert np.abs(dhp.support_points() - dh1.support_points()[to_serial]).max() == 0.0


def test_cpu_sample_is_spread_over_the_cell_loop(mfhn):
    """bench.py's CPU arm times windows spread over the whole cell loop: the sample's share of cells with hanging nodes
    is the mesh's, its DoFs are renumbered compactly."""
    bench = importlib.import_module("bench")
    tria = mfhn.Triangulation("annulus", 6, "p4est")
    mf = mfhn.MatrixFree(mfhn.DoFHandler(tria, 2))
    idx, masks, h, nd, ns, ns_hn = bench.cpu_sample(mf, 2000, n_windows=20)
    assert ns == 2000 and idx.shape == (ns, 27) and idx.max() == nd - 1 and len(np.unique(idx)) == nd
    assert abs(ns_hn / ns - mf.n_cells_hn() / mf.n_cells) < 0.08
