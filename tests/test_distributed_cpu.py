"""world_size-2 (and 3) gloo tests of the partitioned path on CPU: Morton
partition, rank-local numbering, ghost import / compress.  The rank-local cell
loop is played by the CPU oracle; the exchange logic is the product's
(dealii-matrixfree-hanging-nodes_b200/distributed.py)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, geo, L, k, hn_weight, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
        distributed = importlib.import_module("dealii-matrixfree-hanging-nodes_b200.distributed")
        from oracle import cpu

        tria = mfhn.Triangulation(geo, L, "p4est")
        rank_of_cell = tria.partition(world, hn_weight)
        dh = mfhn.DoFHandler(tria, k, world, rank_of_cell)
        mf = mfhn.MatrixFree(dh, rank)
        mfhn.exchange_import_indices(mf.partitioner)
        part = mf.partitioner
        b, e = dh.owned_range(rank)
        pts = dh.support_points(b, e)
        src = torch.zeros(part.n_owned + part.n_ghost, dtype=torch.float64)
        src[:part.n_owned] = torch.from_numpy(np.sin(pts).sum(axis=1))  # benchmark_03.h:362-378
        dst = torch.zeros_like(src)

        def local_apply(d, s, cb, ce):
            cpu.vmult(k, mf.dof_indices[cb:ce], mf.masks[cb:ce], mf.h[cb:ce], s.numpy(), d.numpy())

        ex = distributed.GhostExchange(partitioner=part, segments=(0, mf.n_interior_a, mf.n_interior_cells, mf.n_cells),
                                       local_apply=local_apply, device="cpu", dtype=torch.float64)
        ex.vmult(None, dst, src)
        assert float(dst[part.n_owned:].abs().max()) == 0.0 if part.n_ghost else True
        out[rank] = (b, e, dst[:part.n_owned].numpy().copy(), part.n_ghost_indices(), part.n_import_indices(), mf.n_cells)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,geo,L,k,w", [(2, "quadrant", 4, 2, 1.0), (2, "annulus", 5, 3, 1.0), (3, "annulus", 5, 1, 2.5)])
def test_partitioned_vmult_matches_serial(world, geo, L, k, w):
    from oracle import dofs, mesh, operators

    mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
    port = 29500 + (os.getpid() + world * 7 + k) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, geo, L, k, w, out), nprocs=world, join=True)
        res = dict(out)
    # serial answer in the partitioned (rank-major) numbering: same operator, permuted DoFs
    tria = mfhn.Triangulation(geo, L, "p4est")
    dh = mfhn.DoFHandler(tria, k, world, tria.partition(world, w))
    cells = np.arange(tria.n_active_cells())
    _, sub, masks, h = dh.fill(cells)
    lay = dofs.DoFLayout()
    lay.degree, lay.n_cells, lay.n_dofs = k, len(cells), dh.n_dofs()
    lay.dof_indices, lay.masks, lay.h = sub.astype(np.uint32), masks, h
    lay.kinds = np.array([dofs.decompress(int(m)) for m in masks], dtype=np.uint16)
    x = np.sin(dh.support_points()).sum(axis=1)
    ref = operators.vmult_fast(lay, x)
    got = np.zeros_like(ref)
    n_cells = 0
    for r in range(world):
        b, e, y, ng, ni, nc = res[r]
        got[b:e] = y
        n_cells += nc
    assert n_cells == tria.n_active_cells()
    # every ghost entry of one rank is an import entry of its owner (benchmark_02.cc:164-165 logs both)
    assert sum(res[r][3] for r in range(world)) == sum(res[r][4] for r in range(world)) > 0
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
    # the numbering itself: same DoF count as the serial enumeration, ranges tile [0, n_dofs)
    assert dh.n_dofs() == mfhn.DoFHandler(tria, k).n_dofs()
    assert [res[r][0] for r in range(world)] == [0] + [res[r][1] for r in range(world - 1)]


def test_weighted_partition_balances_weight():
    """benchmark_02.cc:15-37: weight 1+10w for hanging-node cells, 11 otherwise."""
    mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
    tria = mfhn.Triangulation("annulus", 6, "p4est")
    dh = mfhn.DoFHandler(tria, 1)
    _, _, masks, _ = dh.fill(np.arange(tria.n_active_cells()), substituted=False, h=False)
    hn = masks != 0
    for w in (1.0, 4.0, 10.0):
        rank = tria.partition(4, w)
        weights = np.where(hn, 1 + 10 * w, 11.0)
        per_rank = np.array([weights[rank == r].sum() for r in range(4)])
        assert per_rank.max() / per_rank.mean() < 1.02
        # contiguous along the Morton curve
        pos = tria.morton_position()
        assert (np.diff(rank[np.argsort(pos)]) >= 0).all()
    counts1 = np.bincount(tria.partition(4, 1.0), minlength=4)
    assert counts1.max() - counts1.min() <= 1
