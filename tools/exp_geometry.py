#!/usr/bin/env python
"""Development only: times the operator with per-quadrature-point geometry (MFHN_GEOM_GENERAL, the data class of
TestHighOrderMapping, benchmark_01.h:225-242) against its own byte model.  usage: exp_geometry.py [--degree 4] [--L 8]"""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
from tools.exp_kernels import timeit
from bench_dist import high_order_geometry
degree, L = 4, 8
args = sys.argv[1:]
while args:
    a = args.pop(0)
    if a == "--degree": degree = int(args.pop(0))
    elif a == "--L": L = int(args.pop(0))
tria = mfhn.Triangulation("annulus", L, "p4est")
dh = mfhn.DoFHandler(tria, degree)
mf = mfhn.MatrixFree(dh)
G = high_order_geometry(mfhn, tria, mf, degree)
nd = dh.n_dofs()
for kern in ("qpoint", "auto"):
    op = mfhn.LaplaceOperator(mf, kernel=kern, geometry=G)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.copy_(torch.sin(1e-3 * torch.arange(src.numel(), device=src.device, dtype=torch.float64)))
    t = timeit(op, dst, src, 10, 3)
    b = op.query("algorithmic_bytes_accumulate")
    print(json.dumps({"degree": degree, "L": L, "kernel": kern, "resolved": int(op.query("kernel")), "ms": round(t, 4), "gdofs": round(nd / t / 1e6, 2),
                      "bytes": b, "frac_hbm": round(b / (t * 1e-3) / 1e9 / 6523.7, 4)}), flush=True)
    del op, src, dst
    torch.cuda.empty_cache()
