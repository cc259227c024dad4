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

    kname = {1: "qpoint", 2: "separable", 3: "baseline", 4: "plane", 5: "patch"}[int(op.query("kernel"))]
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
            for kern in ("patch", "plane", "separable", "qpoint"):
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
