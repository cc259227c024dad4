"""Times the Cartesian fast-path kernels against each other: degrees 3..5, double and float
(annulus, p4est balance; L = 9 for k <= 4, 8 for k = 5).  Output: one JSON line per case."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")

kernels = sys.argv[1].split(",") if len(sys.argv) > 1 else ["plane", "bulk"]
degrees = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3, 4, 5]
for k in degrees:
    L = 9 if k <= 4 else 8
    tria = mfhn.Triangulation("annulus", L, "p4est")
    dh = mfhn.DoFHandler(tria, k)
    mf = mfhn.MatrixFree(dh)
    for number in ("double", "float"):
        op = mfhn.LaplaceOperator(mf, number=number, kernel="plane")
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.copy_(torch.sin(torch.arange(src.numel(), device=src.device, dtype=src.dtype)))
        out = {"k": k, "L": L, "number": number, "n_dofs": dh.n_dofs()}
        for kernel in kernels:
            try:
                op.set_kernel(kernel)
            except mfhn.MfhnError as e:
                out[kernel] = str(e)
                continue
            for _ in range(3):
                op.vmult(dst, src)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                op.vmult(dst, src)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            out[kernel] = {"ms": round(ms, 4), "gdofs": round(dh.n_dofs() / ms / 1e6, 2)}
        print(json.dumps(out), flush=True)
        del op, src, dst
