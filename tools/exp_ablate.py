#!/usr/bin/env python
"""Development only: times the bulk kernel with parts switched off (library built with -DMFHN_ABLATIONS)."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
from tools.exp_kernels import timeit
tria = mfhn.Triangulation("annulus", 9, "p4est")
dh = mfhn.DoFHandler(tria, 4)
mf = mfhn.MatrixFree(dh)
op = mfhn.LaplaceOperator(mf, kernel="bulk")
src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
src.fill_(1.0)
os.environ["MFHN_OCC"] = "5"
names = {0: "full", 1: "no lv gather", 2: "no lv RED", 3: "no lv gather+RED", 4: "no bulk load", 8: "no bulk red", 12: "no bulk load+red", 15: "no global traffic but indices",
         16: "no sweeps", 32: "no staging rd/wr", 48: "no sweeps, no staging", 63: "nothing but index loads"}
for ac in (False, True):
    op.set_apply_constraints(ac)
    for abl, name in names.items():
        if abl:
            os.environ["MFHN_ABL"] = str(abl)
        else:
            os.environ.pop("MFHN_ABL", None)
        t = timeit(op, dst, src, 10, 2)
        print(json.dumps({"constraints": ac, "abl": abl, "what": name, "ms": round(t, 4)}), flush=True)
