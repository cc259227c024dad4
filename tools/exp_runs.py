#!/usr/bin/env python
"""Development only: run-detection parameter sweep of MFHN_KERNEL_RUNS against the bulk and plane kernels.
usage: exp_runs.py [--degree 4] [--L 9] [--number double] gap:min gap:min ..."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
from tools.exp_kernels import timeit
args = sys.argv[1:]
degree, L, number, occs = 4, None, "double", ["5"]
combos = []
while args:
    a = args.pop(0)
    if a == "--degree": degree = int(args.pop(0))
    elif a == "--L": L = int(args.pop(0))
    elif a == "--number": number = args.pop(0)
    elif a == "--occ": occs = args.pop(0).split(",")
    else: combos.append(tuple(a.split(":")))
L = L or (9 if degree <= 4 else 8)
tria = mfhn.Triangulation("annulus", L, "p4est")
dh = mfhn.DoFHandler(tria, degree)
mf = mfhn.MatrixFree(dh)
nd = dh.n_dofs()
ref = None
for kern, gap, mn in [("plane", "0", "0"), ("bulk", "0", "0")] + [("runs", g, m) for g, m in combos]:
    os.environ["MFHN_RUNS_GAP"], os.environ["MFHN_RUNS_MIN"] = gap, mn if mn != "0" else "6"
    try:
        op = mfhn.LaplaceOperator(mf, number=number, kernel=kern)
    except mfhn.MfhnError as e:
        print(json.dumps({"kernel": kern, "error": str(e)}), flush=True)
        continue
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    i = torch.arange(src.numel(), device=src.device, dtype=torch.float64)
    src.copy_(torch.sin(1e-3 * i).to(src.dtype))
    op.vmult(dst, src, zero_dst=True)
    y = dst.double().clone()
    if ref is None:
        ref = y
    err = float((y - ref).abs().max() / ref.abs().max())
    for occ in occs:
        os.environ["MFHN_OCC"] = occ
        op.set_apply_constraints(True)
        t1 = timeit(op, dst, src, 20, 3)
        op.set_apply_constraints(False)
        t0 = timeit(op, dst, src, 20, 3)
        out = {"degree": degree, "number": number, "kernel": kern, "gap": gap, "min": mn, "occ": occ, "ms": round(t1, 4), "gdofs": round(nd / t1 / 1e6, 2),
               "ms_noconstr": round(t0, 4), "hn_pct": round(100 * (t1 / t0 - 1), 1), "err_vs_plane": err}
        if kern == "runs":
            out.update(copies_per_cell=round(op.query("runs_bulk_copies") / mf.n_cells, 2), singles_per_cell=round(op.query("runs_single_entries") / mf.n_cells, 2),
                       zero_per_cell=round(op.query("runs_zero_entries") / mf.n_cells, 2))
        print(json.dumps(out), flush=True)
    del op, src, dst
    torch.cuda.empty_cache()
