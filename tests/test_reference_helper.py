"""The reference's own definition of "cell with hanging nodes" -- Helper<dim>::is_constrained,
/root/reference/constraint_helper.h:8-125, restated in oracle/helper.py -- against the constraint
detection of the oracle (ConstraintKinds != 0) and of the C++ setup (compressed masks, n_cells_hn),
plus the 2:1 balance invariants the detection relies on."""
import itertools

import numpy as np
import pytest

from oracle import dofs, mesh
from oracle.helper import Helper

CASES = [("quadrant", 3, "serial"), ("quadrant", 4, "p4est"), ("quadrant", 5, "serial"), ("annulus", 5, "serial"),
         ("annulus", 5, "p4est"), ("annulus", 6, "serial"), ("step", 4, "serial"), ("quadrant_flexible", 3, "serial")]


@pytest.mark.parametrize("geo,L,flavour", CASES)
def test_helper_is_constrained_equals_nonzero_kind(geo, L, flavour, mfhn):
    t = mesh.create(geo, L, flavour)
    cells = t.active_cells()
    helper = Helper(t)
    ref = np.array([helper.is_constrained(c) for c in cells])
    kinds = dofs.constraint_kinds(t, cells)
    assert np.array_equal(ref, kinds != 0)
    # face / edge split: a non-zero face field <=> is_face_constrained; edge-only kinds are edge constrained
    face = np.array([helper.is_face_constrained(c) for c in cells])
    assert np.array_equal(face, ((kinds >> 3) & 7) != 0)
    edge_only = (((kinds >> 3) & 7) == 0) & (((kinds >> 6) & 7) != 0)
    assert all(helper.is_edge_constrained(c) for c, e in zip(cells, edge_only) if e)
    # the C++ setup counts the same cells (benchmark_03.h:415-432)
    tria = mfhn.Triangulation(geo, L, flavour)
    assert tria.n_cells_with_hanging_nodes() == int(ref.sum())
    dh = mfhn.DoFHandler(tria, 1)
    _, _, masks, _ = dh.fill(np.arange(tria.n_active_cells()), substituted=False, h=False)
    assert np.array_equal(masks != 0, ref)


@pytest.mark.parametrize("geo,L,flavour", CASES)
def test_two_to_one_balance(geo, L, flavour):
    """Faces and edges for the serial flavour, corners too for p4est: a neighbouring region of an
    active cell is never covered by a leaf more than one level coarser."""
    t = mesh.create(geo, L, flavour)
    nodes = t.has_children
    offsets = [o for o in itertools.product((-1, 0, 1), repeat=3) if 0 < sum(map(abs, o)) <= (3 if flavour == "p4est" else 2)]
    for (l, i, j, k) in t.active_cells():
        for dx, dy, dz in offsets:
            p = (i + dx, j + dy, k + dz)
            if not all(0 <= c < (1 << l) for c in p):
                continue
            ll, q = l, p
            while (ll,) + q not in nodes:
                ll -= 1
                q = tuple(c >> 1 for c in q)
            assert l - ll <= 1
