#!/usr/bin/env python
"""Experiment driver (not part of the product): times kernel / occupancy variants on one mesh.

    python tools/exp_kernels.py --degrees 4 5 --kernels plane bulk --occ 4 5 6 [--L 9]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")


def timeit(op, dst, src, steps=20, warmup=3):
    for _ in range(warmup):
        op.vmult(dst, src)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        op.vmult(dst, src)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--degrees", type=int, nargs="+", default=[4])
    ap.add_argument("--kernels", nargs="+", default=["plane", "bulk"])
    ap.add_argument("--occ", type=int, nargs="+", default=[4])
    ap.add_argument("--numbers", nargs="+", default=["double"])
    ap.add_argument("--geometry", default="annulus")
    ap.add_argument("--L", type=int, default=None)
    ap.add_argument("--env", nargs="*", default=[], help="NAME=v1,v2 : extra environment sweeps")
    args = ap.parse_args()
    for k in args.degrees:
        L = args.L or (9 if k <= 4 else 8)
        tria = mfhn.Triangulation(args.geometry, L, "p4est")
        dh = mfhn.DoFHandler(tria, k)
        mf = mfhn.MatrixFree(dh)
        nd = dh.n_dofs()
        for number in args.numbers:
            for kern in args.kernels:
                try:
                    op = mfhn.LaplaceOperator(mf, number=number, kernel=kern)
                except mfhn.MfhnError as e:
                    print(json.dumps({"degree": k, "kernel": kern, "number": number, "error": str(e)}), flush=True)
                    continue
                src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
                i = torch.arange(src.numel(), device=src.device, dtype=torch.float64)
                src.copy_(torch.sin(1e-3 * i).to(src.dtype))
                for occ in args.occ:
                    os.environ["MFHN_OCC"] = str(occ)
                    op.set_apply_constraints(True)
                    t1 = timeit(op, dst, src)
                    op.set_apply_constraints(False)
                    t0 = timeit(op, dst, src)
                    b = op.query("algorithmic_bytes_accumulate")
                    print(json.dumps({"degree": k, "L": L, "kernel": kern, "number": number, "occ": occ, "ms": round(t1, 4), "gdofs": round(nd / t1 / 1e6, 2),
                                      "ms_noconstr": round(t0, 4), "hn_overhead_pct": round(100 * (t1 / t0 - 1), 2),
                                      "frac_hbm": round(b / (t1 * 1e-3) / 1e9 / 6523.7, 4)}), flush=True)
                del op, src, dst
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
