"""Hanging-node weighted repartitioning study (the reference's benchmark_02: cell weight 1 + 10 w for cells with
hanging nodes, 1 + 10 otherwise, w in [1, 10]; benchmark_02.cc:15-37, 63; it logs n_ghost_indices / n_import_indices
per rank, :164-165).  Runs on the CPU: for every weight the Morton partition, the per-rank cell / hanging-node-cell
counts, ghost and import sizes, and the load imbalance predicted from the measured cost ratio eta of a cell with
hanging nodes (bench.py: eta5).  Output: one JSON object (profiles/r1_partition_study.json).

usage: python examples/partition_study.py [geometry=annulus] [L=7] [degree=4] [ranks=8] [eta=1.45]"""
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")

geo = sys.argv[1] if len(sys.argv) > 1 else "annulus"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 7
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
R = int(sys.argv[4]) if len(sys.argv) > 4 else 8
eta = float(sys.argv[5]) if len(sys.argv) > 5 else 1.45

tria = mfhn.Triangulation(geo, L, "p4est")
out = {"geometry": geo, "n_refinements": L, "degree": k, "n_ranks": R, "eta": eta, "n_cells": tria.n_active_cells(),
       "n_cells_hn": tria.n_cells_with_hanging_nodes(), "weights": []}
for w in (1.0, 1.25, 1.5, 2.0, 3.0, 4.0, 6.0, 8.0, 10.0):
    rank_of_cell = tria.partition(R, w)
    dh = mfhn.DoFHandler(tria, k, R, rank_of_cell)
    parts = [mfhn.MatrixFree(dh, r, categorize=False) for r in range(R)]
    n_cells = np.array([p.n_cells for p in parts])
    n_hn = np.array([p.n_cells_hn() for p in parts])
    n_ghost = np.array([p.partitioner.n_ghost for p in parts])
    n_import = np.zeros(R, dtype=np.int64)
    for p in parts:  # what rank q ghosts is what its owners import
        owners, counts = np.unique(p.partitioner.ghost_owner, return_counts=True)
        n_import[owners] += counts
    cost = (n_cells - n_hn) + eta * n_hn  # in units of one regular cell
    out["weights"].append({"weight": w, "cells_min_max": [int(n_cells.min()), int(n_cells.max())],
                           "hn_cells_min_max": [int(n_hn.min()), int(n_hn.max())],
                           "ghost_min_max_avg": [int(n_ghost.min()), int(n_ghost.max()), float(n_ghost.mean())],
                           "import_min_max_avg": [int(n_import.min()), int(n_import.max()), float(n_import.mean())],
                           "predicted_imbalance": float(cost.max() / cost.mean())})
best = min(out["weights"], key=lambda e: e["predicted_imbalance"])
out["best_weight"] = best["weight"]
print(json.dumps(out))
