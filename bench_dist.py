"""Problem construction shared by bench.py: one GPU (whole mesh) or one rank
of a Morton-partitioned mesh with ghost exchange (world > 1)."""
from __future__ import annotations

import numpy as np


def build_problem(mfhn, args, L, rank, world):
    import torch

    tria = mfhn.Triangulation(args.geometry, L, "p4est")
    if world == 1:
        dh = mfhn.DoFHandler(tria, args.degree)
        mf = mfhn.MatrixFree(dh)
        op = mfhn.LaplaceOperator(mf, number=args.number, kernel=args.kernel)
        comm = None
        partition = "1 GPU, whole mesh"
        launches = 1
    else:
        from importlib import import_module

        distributed = import_module("dealii-matrixfree-hanging-nodes_b200.distributed")
        rank_of_cell = tria.partition(world, args.hn_weight)
        dh = mfhn.DoFHandler(tria, args.degree, world, rank_of_cell)
        mf = mfhn.MatrixFree(dh, rank)
        mfhn.exchange_import_indices(mf.partitioner)
        op = mfhn.LaplaceOperator(mf, number=args.number, kernel=args.kernel)
        comm = distributed.GhostExchange(op)
        op.attach_communicator(comm)
        how = "one C-ABI call per vmult (NCCL groups issued from C++)" if comm._native is not None else "torch.distributed p2p groups"
        partition = f"Morton (p4est-like) partition into {world} ranks (hanging-node weight {args.hn_weight}), NCCL ghost import/compress overlapped with interior cells, {how}"
        launches = comm.launches_per_vmult()

    def fill_src(src):
        i = torch.arange(src.numel(), device=src.device, dtype=torch.float64)
        src.copy_(torch.sin(1e-3 * i).to(src.dtype))

    kname = {1: "qpoint", 2: "separable", 3: "baseline", 4: "plane", 5: "patch", 6: "bulk"}[int(op.query("kernel"))]
    return {"op": op, "mf": mf, "dh": dh, "tria": tria, "n_dofs": dh.n_dofs(), "n_cells_global": tria.n_active_cells(),
            "n_cells_hn_global": tria.n_cells_with_hanging_nodes(), "fill_src": fill_src, "kernel_name": kname,
            "partition": partition, "launches_per_step": launches, "comm": comm}


def degree_sweep(mfhn, torch, args, time_vmult):
    """degrees 1..8 on the annulus, double and float, with and
    without constraints, all kernels."""
    res = []
    for k in range(1, 9):
        L = 9 if k <= 4 else 8
        tria = mfhn.Triangulation(args.geometry, L, "p4est")
        dh = mfhn.DoFHandler(tria, k)
        mf = mfhn.MatrixFree(dh)
        for number in ("double", "float"):
            row = {"degree": k, "L": L, "number": number, "n_dofs": dh.n_dofs(), "n_cells": mf.n_cells}
            op = mfhn.LaplaceOperator(mf, number=number)
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
            src.fill_(1.0)
            for kern in ("plane", "separable", "qpoint"):
                try:
                    op.set_kernel(kern)
                except mfhn.MfhnError:
                    continue
                for ac in (True, False):
                    op.set_apply_constraints(ac)
                    _, per = time_vmult(torch, op, dst, src, 10, 3)
                    row[f"{kern}{'' if ac else '_noconstr'}_gdofs"] = dh.n_dofs() / (float(np.mean(per)) * 1e-3) / 1e9
            row["algorithmic_bytes"] = op.query("algorithmic_bytes")
            res.append(row)
            del op, src, dst
            torch.cuda.empty_cache()
    return res


def stage_benchmarks(mfhn, torch, args, L, time_vmult):
    """The reference's decomposition (benchmark_01.cc:189-220) on the GPU: "DG (SC)" = every cell owns private
    DoFs (contiguous cell-local gather / scatter, benchmark_01.h:639-677) with and without the interpolation
    (t2, t3, eta3); "CG (SC)" = the real operator (t4, t5, eta5); and the interpolation alone on cell-local
    values (benchmark_00_likwid.cc:56-59)."""
    import numpy as np

    tria = mfhn.Triangulation(args.geometry, L, "p4est")
    dh = mfhn.DoFHandler(tria, args.degree)
    mf = mfhn.MatrixFree(dh)
    n3 = (args.degree + 1) ** 3
    n_hn, n_all = mf.n_cells_hn(), mf.n_cells

    def eta(t_n, t_hn):
        return max((t_hn / (t_n / n_all) - (n_all - n_hn)) / n_hn, 1.0) if n_hn else 1.0

    res = {"n_cells": n_all, "n_cells_hn": n_hn}
    # CG (SC)
    op = mfhn.LaplaceOperator(mf, number=args.number)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.fill_(1.0)  # benchmark_01.h:510-511
    op.set_apply_constraints(False)
    _, t4 = time_vmult(torch, op, dst, src, 10, 3)
    op.set_apply_constraints(True)
    _, t5 = time_vmult(torch, op, dst, src, 10, 3)
    res.update(t4_ms=float(t4.mean()), t5_ms=float(t5.mean()), eta5=eta(float(t4.mean()), float(t5.mean())))
    # the interpolation alone
    vals = torch.ones(mf.n_cells * n3, dtype=src.dtype, device=src.device)
    for _ in range(3):
        op.apply_hanging_node_constraints(vals, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        op.apply_hanging_node_constraints(vals, False)
    e1.record()
    torch.cuda.synchronize()
    res["hn_kernel_alone_ms"] = e0.elapsed_time(e1) / 10
    del op, src, dst, vals
    torch.cuda.empty_cache()
    # DG (SC): private DoFs per cell
    class _DG:
        pass

    dg = _DG()
    dg.degree, dg.n_cells, dg.h, dg.masks = mf.degree, mf.n_cells, mf.h, mf.masks
    dg.dof_indices = np.arange(mf.n_cells * n3, dtype=np.uint32).reshape(mf.n_cells, n3)
    dg.n_interior_cells = dg.n_interior_a = mf.n_cells
    dg.partitioner = mfhn.Partitioner(owned_range=(0, mf.n_cells * n3))
    op = mfhn.LaplaceOperator(dg, number=args.number)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.fill_(1.0)
    op.set_apply_constraints(False)
    _, t2 = time_vmult(torch, op, dst, src, 10, 3)
    op.set_apply_constraints(True)
    _, t3 = time_vmult(torch, op, dst, src, 10, 3)
    res.update(t2_ms=float(t2.mean()), t3_ms=float(t3.mean()), eta3=eta(float(t2.mean()), float(t3.mean())))
    return res
