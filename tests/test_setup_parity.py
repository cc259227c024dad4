"""C++ setup (mesh generators, DoF enumeration, hanging-node masks, index
substitution, support points, partition numbering) against the Python oracle:
bit-exact, as BASELINE.json north_star requires for indices and masks."""
import numpy as np
import pytest

from oracle import dofs, mesh

CASES = [("quadrant", 5, "serial"), ("annulus", 6, "serial"), ("annulus", 6, "p4est"), ("quadrant", 4, "p4est"),
         ("step", 3, "serial"), ("quadrant_flexible", 2, "serial"), ("quadrant", 0, "serial"), ("quadrant", 1, "p4est"),
         ("annulus", 3, "serial")]


@pytest.mark.parametrize("geo,L,flavour", CASES)
def test_mesh_and_dofs_match_oracle(mfhn, geo, L, flavour):
    tria = mfhn.Triangulation(geo, L, flavour)
    t = mesh.create(geo, L, flavour)
    assert np.array_equal(tria.cells(), mesh.cells_array(t))
    assert tria.n_global_levels() == t.n_levels
    assert tria.n_cells_with_hanging_nodes() == int((dofs.constraint_kinds(t) != 0).sum())
    for k in (1, 2, 3, 4, 7):
        if tria.n_active_cells() * (k + 1) ** 3 > 3e6:
            continue
        dh = mfhn.DoFHandler(tria, k)
        lay = dofs.setup(t, k)
        raw, sub, masks, h = dh.fill(np.arange(tria.n_active_cells()), raw=True)
        assert dh.n_dofs() == lay.n_dofs
        assert np.array_equal(raw, lay.raw_indices) and np.array_equal(sub, lay.dof_indices)
        assert np.array_equal(masks, lay.masks) and np.array_equal(h, lay.h)
        assert np.array_equal(dh.support_points(), lay.support_points)
        # masks are valid compressed kinds
        for m in np.unique(masks):
            kind = mfhn.capi.lib.mfhn_decompress(int(m))
            assert mfhn.capi.lib.mfhn_check_kind(kind) == 1 and mfhn.capi.lib.mfhn_compress(kind) == int(m)


def test_large_mesh_statistics(mfhn):
    """SURVEY.md 6 (benchmark_03 p4est flavour): annulus L=8."""
    tria = mfhn.Triangulation("annulus", 8, "p4est")
    assert tria.n_active_cells() == 272896 and tria.n_cells_with_hanging_nodes() == 104456
    assert mfhn.DoFHandler(tria, 4).n_dofs() == 18578585
    assert mfhn.DoFHandler(tria, 1).n_dofs() == 317063
    tria = mfhn.Triangulation("quadrant", 6, "serial")  # benchmark_01 default (benchmark_01.cc:24-26)
    assert tria.n_active_cells() == 34896 and tria.n_cells_with_hanging_nodes() == 4257
    assert mfhn.DoFHandler(tria, 4).n_dofs() == 2304973


def test_compress_matches_oracle_for_all_kinds(mfhn):
    lib = mfhn.capi.lib
    for kind in range(512):
        assert bool(lib.mfhn_check_kind(kind)) == dofs.check(kind)
        if dofs.check(kind):
            assert lib.mfhn_compress(kind) == dofs.compress(kind)
            assert lib.mfhn_decompress(dofs.compress(kind)) == kind


def test_partitioned_numbering_is_a_permutation(mfhn):
    """Rank-major numbering (owned ranges contiguous per rank) addresses the same
    DoFs as the serial numbering: same count, and cell by cell the same sharing pattern."""
    tria = mfhn.Triangulation("annulus", 5, "p4est")
    cells = np.arange(tria.n_active_cells())
    for k in (1, 3):
        serial = mfhn.DoFHandler(tria, k)
        part = mfhn.DoFHandler(tria, k, 3, tria.partition(3))
        assert serial.n_dofs() == part.n_dofs()
        _, s0, m0, _ = serial.fill(cells)
        _, s1, m1, _ = part.fill(cells)
        assert np.array_equal(m0, m1)
        perm = np.full(serial.n_dofs(), -1, dtype=np.int64)
        perm[s0.ravel()] = s1.ravel()
        assert np.array_equal(perm[s0.ravel()], s1.ravel())  # consistent map
        live = perm >= 0
        assert len(np.unique(perm[live])) == live.sum()  # injective
        ranges = [part.owned_range(r) for r in range(3)]
        assert ranges[0][0] == 0 and ranges[-1][1] == part.n_dofs()
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(2))
