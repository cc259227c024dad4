"""Generates the golden fixtures under tests/golden/.

The reference repository ships no golden vectors, fixtures or tests for this
path and cannot be built here (its arithmetic is inside an un-vendored deal.II
fork), so these fixtures are produced by the INDEPENDENT
general-purpose oracle O1 (oracle/operators.py: dense element matrices on raw
DoFs + explicit constraint matrix from geometry, no masks, no index
substitution) and pin everything downstream of it: the fast-algorithm oracle
O2, the C restatement, the C++ setup (indices / masks bit-exact) and the CUDA
kernels.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dofs, mesh, operators  # noqa: E402

CASES = [("quadrant", 3, "serial", 3), ("annulus", 5, "serial", 2), ("quadrant", 3, "p4est", 4), ("annulus", 5, "p4est", 1)]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for geo, L, flavour, k in CASES:
        t = mesh.create(geo, L, flavour)
        lay = dofs.setup(t, k)
        O1 = operators.GeneralOperator(t, lay)
        src_sin = np.sin(lay.support_points).sum(axis=1)  # benchmark_03.h:362-378
        src_rnd = np.random.default_rng(12345).uniform(-1, 1, lay.n_dofs)
        np.savez_compressed(
            os.path.join(here, f"{geo}_L{L}_{flavour}_k{k}.npz"),
            geometry=geo, n_refinements=L, flavour=flavour, degree=k, cells=lay.cells, n_dofs=lay.n_dofs,
            raw_indices=lay.raw_indices, dof_indices=lay.dof_indices, masks=lay.masks, h=lay.h,
            src_sin=src_sin, dst_sin=O1.vmult(src_sin), src_rnd=src_rnd, dst_rnd=O1.vmult(src_rnd),
            is_hanging=O1.is_hanging)
        print(geo, L, flavour, k, lay.n_cells, lay.n_dofs)


if __name__ == "__main__":
    main()
