#!/usr/bin/env python
"""Development only: times one operator under several environment settings.
usage: exp_env.py [--degree 4] [--kernel bulk] NAME=v1,v2 NAME2=w1,w2 ..."""
import importlib, itertools, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
from tools.exp_kernels import timeit
args = sys.argv[1:]
degree, kernel, L = 4, "bulk", None
sweeps = []
while args:
    a = args.pop(0)
    if a == "--degree": degree = int(args.pop(0))
    elif a == "--kernel": kernel = args.pop(0)
    elif a == "--L": L = int(args.pop(0))
    else:
        k, v = a.split("=")
        sweeps.append((k, v.split(",")))
L = L or (9 if degree <= 4 else 8)
tria = mfhn.Triangulation("annulus", L, "p4est")
dh = mfhn.DoFHandler(tria, degree)
mf = mfhn.MatrixFree(dh)
op = mfhn.LaplaceOperator(mf, kernel=kernel)
src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
i = torch.arange(src.numel(), device=src.device, dtype=torch.float64)
src.copy_(torch.sin(1e-3 * i))
nd = dh.n_dofs()
for combo in itertools.product(*[v for _, v in sweeps]):
    for (k, _), v in zip(sweeps, combo):
        os.environ[k] = v
    op.set_apply_constraints(True)
    t1 = timeit(op, dst, src, 20, 3)
    op.set_apply_constraints(False)
    t0 = timeit(op, dst, src, 20, 3)
    print(json.dumps({**{k: v for (k, _), v in zip(sweeps, combo)}, "ms": round(t1, 4), "gdofs": round(nd / t1 / 1e6, 2), "ms_noconstr": round(t0, 4),
                      "hn_pct": round(100 * (t1 / t0 - 1), 1)}), flush=True)
