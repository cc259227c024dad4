"""The partitioned operator on ONE device: every rank of a 2- / 3- / 4-rank Morton partition gets its own operator on
cuda:0, and one vmult is driven through the product's own pieces -- mfhn_pack (update_ghost_values), the cell
partitions with mfhn_op_vmult_range, mfhn_unpack_add (compress(add)) -- with the ghost sections moved by plain device
copies instead of NCCL.  A wrong import list, ghost range or segment boundary shows up here because the input is not
constant (reference: cell_loop with distributed vectors, benchmark_03.h:323-324, 348-353)."""
import ctypes as C
import importlib

import numpy as np
import pytest

from oracle import dofs, operators

pytestmark = pytest.mark.gpu


def _global_reference(mfhn, tria, dh, k, kind):
    cells = np.arange(tria.n_active_cells())
    _, sub, masks, h = dh.fill(cells)
    lay = dofs.DoFLayout()
    lay.degree, lay.n_cells, lay.n_dofs = k, len(cells), dh.n_dofs()
    lay.dof_indices, lay.masks, lay.h = sub.astype(np.uint32), masks, h
    lay.kinds = np.array([dofs.decompress(int(m)) for m in masks], dtype=np.uint16)
    if kind == "sin":  # benchmark_03.h:362-378
        x = np.sin(dh.support_points()).sum(axis=1)
    else:
        x = np.random.default_rng(12345).uniform(-1, 1, dh.n_dofs())
    return x, operators.vmult_fast(lay, x)


@pytest.mark.parametrize("world,geo,L,k,kernel,number", [
    (2, "annulus", 5, 2, "auto", "double"), (3, "annulus", 5, 4, "auto", "double"), (3, "quadrant", 4, 4, "bulk", "double"),
    (2, "quadrant", 4, 2, "plane", "double"), (4, "annulus", 5, 4, "plane", "double"), (2, "quadrant", 4, 6, "auto", "double"),
    (3, "annulus", 5, 3, "qpoint", "double"), (2, "annulus", 5, 4, "auto", "float")])
def test_partitioned_operator_on_one_device(mfhn, world, geo, L, k, kernel, number):
    import torch

    lib, capi = mfhn.capi.lib, mfhn.capi
    tria = mfhn.Triangulation(geo, L, "p4est")
    dh = mfhn.DoFHandler(tria, k, world, tria.partition(world))
    mfs = [mfhn.MatrixFree(dh, r) for r in range(world)]
    mfhn.exchange_local(mfs)
    ops = [mfhn.LaplaceOperator(mf, number=number, kernel=kernel) for mf in mfs]
    num = capi.F64 if number == "double" else capi.F32
    stream = torch.cuda.current_stream().cuda_stream
    imp = [{r: torch.from_numpy(np.ascontiguousarray(idx)).cuda() for r, idx in mf.partitioner.import_indices.items()} for mf in mfs]
    assert sum(mf.partitioner.n_ghost for mf in mfs) > 0 and sum(mf.n_cells for mf in mfs) == tria.n_active_cells()
    for kind in ("sin", "random"):
        x, ref = _global_reference(mfhn, tria, dh, k, kind)
        src, dst = [op.initialize_dof_vector() for op in ops], [op.initialize_dof_vector() for op in ops]
        for r, mf in enumerate(mfs):
            b, e = mf.partitioner.begin, mf.partitioner.end
            src[r][:e - b] = torch.from_numpy(x[b:e]).to(src[r].dtype)
        # update_ghost_values: owners pack, the ghost ranges receive
        for o, mf in enumerate(mfs):
            for r, idx in imp[o].items():
                buf = torch.empty(idx.numel(), dtype=src[o].dtype, device="cuda")
                capi.check(lib.mfhn_pack(num, buf.data_ptr(), src[o].data_ptr(), idx.data_ptr(), idx.numel(), stream))
                a, b_ = mfs[r].partitioner.ghost_ranges[o]
                src[r][mfs[r].partitioner.n_owned + a:mfs[r].partitioner.n_owned + b_] = buf
        # the three cell partitions of every rank
        for r, mf in enumerate(mfs):
            for cb, ce in ((0, mf.n_interior_a), (mf.n_interior_a, mf.n_interior_cells), (mf.n_interior_cells, mf.n_cells)):
                if ce > cb:
                    ops[r].vmult_range(dst[r], src[r], cb, ce)
        # compress(add): ghost sections go back to the owners
        for r, mf in enumerate(mfs):
            for o, (a, b_) in mf.partitioner.ghost_ranges.items():
                buf = dst[r][mf.partitioner.n_owned + a:mf.partitioner.n_owned + b_].clone()
                idx = imp[o][r]
                assert idx.numel() == buf.numel()
                capi.check(lib.mfhn_unpack_add(num, dst[o].data_ptr(), buf.data_ptr(), idx.data_ptr(), idx.numel(), stream))
        torch.cuda.synchronize()
        got = np.zeros_like(ref)
        for r, mf in enumerate(mfs):
            b, e = mf.partitioner.begin, mf.partitioner.end
            got[b:e] = dst[r][:e - b].cpu().numpy()
        err = np.abs(got - ref).max() / np.abs(ref).max()
        if number == "float" and kind == "sin":
            assert err < 1e-4, (world, geo, k, kernel, err)  # cancellation on the smooth input, see test_gpu_parity.py
        else:
            assert err < (1e-12 if number == "double" else 1e-5), (world, geo, k, kernel, kind, err)
