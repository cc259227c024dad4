"""Hanging-node weighted repartitioning measured on a GPU (the reference's benchmark_02: cell weight 1 + 10 w for
cells with hanging nodes, w in [1, 10], run with and without communication; benchmark_02.cc:15-37, 63, 196-212).

For every weight the mesh is Morton-partitioned into R ranks; every rank's operator is created on THIS device and its
rank-local cell loop (the "without communication" run of the reference: no ghost exchange) is timed with CUDA events.
One device runs the R loops one after the other, so the load balance of the partition is measured with the real kernel
-- max over ranks = the step time of R GPUs without exchange, max / mean = imbalance -- at 1/R of the GPU time of an
R-GPU job.  The run with communication is `bench.py --gpus R --hn-weight w`.

usage: python examples/partition_study_gpu.py [geometry=quadrant] [L=8] [degree=4] [ranks=8] [weights=1,1.5,2,...]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")

geo = sys.argv[1] if len(sys.argv) > 1 else "quadrant"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
R = int(sys.argv[4]) if len(sys.argv) > 4 else 8
weights = [float(w) for w in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1.0, 1.25, 1.5, 2.0, 3.0, 4.0, 6.0, 10.0]

tria = mfhn.Triangulation(geo, L, "p4est")
out = {"geometry": geo, "n_refinements": L, "degree": k, "n_ranks": R, "n_cells": tria.n_active_cells(),
       "n_cells_hn": tria.n_cells_with_hanging_nodes(), "weights": []}
for w in weights:
    dh = mfhn.DoFHandler(tria, k, R, tria.partition(R, w))
    ms, cells, hn, ghost = [], [], [], []
    for r in range(R):
        mf = mfhn.MatrixFree(dh, r)
        op = mfhn.LaplaceOperator(mf)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.copy_(torch.sin(1e-3 * torch.arange(src.numel(), device=src.device, dtype=torch.float64)))
        for _ in range(3):
            op.vmult_range(dst, src, 0, mf.n_cells)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            op.vmult_range(dst, src, 0, mf.n_cells)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / 10)
        cells.append(mf.n_cells), hn.append(mf.n_cells_hn()), ghost.append(mf.partitioner.n_ghost)
        del op, src, dst, mf
        torch.cuda.empty_cache()
    ms = np.array(ms)
    out["weights"].append({"weight": w, "local_cell_loop_ms": [round(float(x), 4) for x in ms], "step_ms_without_exchange": float(ms.max()),
                           "imbalance": float(ms.max() / ms.mean()), "gdofs_without_exchange": dh.n_dofs() / (float(ms.max()) * 1e-3) / 1e9,
                           "cells_min_max": [int(min(cells)), int(max(cells))], "hn_cells_min_max": [int(min(hn)), int(max(hn))],
                           "ghost_min_max": [int(min(ghost)), int(max(ghost))]})
    print(json.dumps(out["weights"][-1]), file=sys.stderr, flush=True)
best = min(out["weights"], key=lambda e: e["step_ms_without_exchange"])
out["best_weight"] = best["weight"]
print(json.dumps(out))
