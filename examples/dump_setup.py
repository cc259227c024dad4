#!/usr/bin/env python
"""Dumps the setup product of one mesh + degree (cells, DoF indices, masks, geometry) to an .npz
fixture in the same format as tests/golden/*.npz, so that a deal.II run performed elsewhere
(MatrixFree::get_dof_info().dof_indices / hanging_node_constraint_masks) can be
diffed bit for bit against this engine's setup.

    python examples/dump_setup.py annulus 5 4 out.npz [serial|p4est]

Arrays: cells int32[n_cells,4] (level, ix, iy, iz; storage order), n_dofs, raw_indices /
dof_indices uint64[n_cells,(k+1)^3] (lexicographic, x fastest; before / after the coarse-index
substitution), masks uint8[n_cells] (compressed_constraint_kind), h float64[n_cells],
support_points float64[n_dofs,3].
"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    geo, L, k, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    flavour = sys.argv[5] if len(sys.argv) > 5 else "p4est"
    mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
    tria = mfhn.Triangulation(geo, L, flavour)
    dh = mfhn.DoFHandler(tria, k)
    raw, sub, masks, h = dh.fill(np.arange(tria.n_active_cells()), raw=True)
    np.savez_compressed(out, geometry=geo, n_refinements=L, flavour=flavour, degree=k, cells=tria.cells(), n_dofs=dh.n_dofs(),
                        raw_indices=raw, dof_indices=sub, masks=masks, h=h, support_points=dh.support_points())
    print(f"{geo} L={L} {flavour} k={k}: {tria.n_active_cells()} cells ({tria.n_cells_with_hanging_nodes()} with hanging nodes), "
          f"{dh.n_dofs()} DoFs -> {out}")


if __name__ == "__main__":
    main()
