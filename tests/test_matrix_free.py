"""MatrixFree::reinit behind the C ABI (mfhn_mf_create, benchmark_03.h:326-340 / benchmark_01.h:251-284) against a
numpy restatement of the same rules: cell order (Morton curve, interior cells first, categorised by constraint mask
inside windows), rank-local numbering (owned range, ghosts sorted by global index), partitioner.  Bit-exact."""
import importlib

import numpy as np
import pytest


def reinit_numpy(dh, rank, categorize=True, window=3840, by_kind=True):
    cells = dh.cells_of_rank(rank)
    pos = dh.tria.morton_position()
    cells = cells[np.argsort(pos[cells], kind="stable")]
    _, sub, masks, h = dh.fill(cells)
    b, e = dh.owned_range(rank)
    rank_begin = np.array([dh.owned_range(r)[0] for r in range(dh.n_ranks)], dtype=np.int64)
    flat = sub.reshape(-1).astype(np.int64)
    is_ghost = (flat < b) | (flat >= e)
    ghost_global = np.unique(flat[is_ghost])
    local = flat - b
    if len(ghost_global):
        local[is_ghost] = (e - b) + np.searchsorted(ghost_global, flat[is_ghost])
    local = local.reshape(sub.shape)
    touches = is_ghost.reshape(sub.shape).any(axis=1)
    order = np.concatenate([np.nonzero(~touches)[0], np.nonzero(touches)[0]])
    n_interior = int((~touches).sum())
    if dh.n_ranks > 1:
        n_interior -= n_interior % 240
    if categorize:
        key = masks[order].astype(np.int64) if by_kind else (masks[order] != 0).astype(np.int64)
        seg = (np.arange(len(order)) >= n_interior).astype(np.int64)
        p = np.arange(len(order))
        win = np.where(seg == 0, p, p - n_interior) // window
        order = order[np.lexsort((p, key, win, seg))]
    n_interior_a = (n_interior // 2) // 240 * 240 if dh.n_ranks > 1 else n_interior
    owner = (np.searchsorted(rank_begin, ghost_global, side="right") - 1).astype(np.int32)
    return dict(cell_ids=cells[order], dof_indices=local[order].astype(np.uint32), masks=masks[order], h=h[order], n_interior=n_interior,
                n_interior_a=n_interior_a, ghost_global=ghost_global, ghost_owner=owner, n_owned=e - b)


@pytest.mark.parametrize("geo,L,k,world", [("annulus", 5, 2, 1), ("quadrant", 4, 4, 1), ("annulus", 6, 1, 3), ("quadrant", 5, 3, 2), ("annulus", 5, 4, 4)])
@pytest.mark.parametrize("categorize", [True, False])
def test_reinit_matches_numpy_restatement(mfhn, geo, L, k, world, categorize, monkeypatch):
    monkeypatch.setenv("MFHN_CATEGORIZE_WINDOW", "480")  # several windows on these small meshes
    tria = mfhn.Triangulation(geo, L, "p4est")
    dh = mfhn.DoFHandler(tria, k, world, tria.partition(world) if world > 1 else None)
    mfs = [mfhn.MatrixFree(dh, r, categorize=categorize) for r in range(world)]
    if world > 1:
        mfhn.exchange_local(mfs)
    n_cells = 0
    for r, mf in enumerate(mfs):
        ref = reinit_numpy(dh, r, categorize, window=480)
        assert np.array_equal(mf.cell_ids, ref["cell_ids"])
        assert np.array_equal(mf.dof_indices, ref["dof_indices"])
        assert np.array_equal(mf.masks, ref["masks"]) and np.array_equal(mf.h, ref["h"])
        assert (mf.n_interior_cells, mf.n_interior_a) == (ref["n_interior"], ref["n_interior_a"])
        part = mf.partitioner
        assert part.n_owned == ref["n_owned"] and np.array_equal(part.ghost_global, ref["ghost_global"])
        assert np.array_equal(part.ghost_owner, ref["ghost_owner"])
        assert mf.n_cells_hn() == int((ref["masks"] != 0).sum())
        # ghost ranges tile the ghost section, one contiguous range per owner
        covered = sorted(part.ghost_ranges.values())
        assert [a for a, _ in covered] == [0] * (len(covered) > 0) + [b for _, b in covered[:-1]]
        assert (covered[-1][1] if covered else 0) == part.n_ghost
        n_cells += mf.n_cells
    assert n_cells == tria.n_active_cells()
    # imports: what rank r ghosts from owner o is exactly what o imports for r
    for r, mf in enumerate(mfs):
        for o, (a, b) in mf.partitioner.ghost_ranges.items():
            want = mf.partitioner.ghost_global[a:b] - mfs[o].partitioner.begin
            assert np.array_equal(mfs[o].partitioner.import_indices[r], want.astype(np.int32))
    assert sum(mf.partitioner.n_import_indices() for mf in mfs) == sum(mf.partitioner.n_ghost for mf in mfs)


def test_partition_weight_is_truncated_like_the_reference(mfhn):
    """hanging_nodes_weighting returns unsigned int (benchmark_02.cc:19-33): 1 + 10 w is truncated."""
    tria = mfhn.Triangulation("annulus", 6, "p4est")
    assert np.array_equal(tria.partition(64, 2.59), tria.partition(64, 2.5))   # 26.9 -> 26
    assert not np.array_equal(tria.partition(64, 2.6), tria.partition(64, 2.5))  # 27
