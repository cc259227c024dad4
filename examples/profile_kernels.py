"""Launches the Cartesian fast-path kernels (plane, bulk) twice each on the benchmark problem:
target for one `ncu -k regex:_cell_kernel` capture (profiles/)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")

L = int(sys.argv[1]) if len(sys.argv) > 1 else 9
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kernels = sys.argv[3].split(",") if len(sys.argv) > 3 else ["plane", "bulk"]
tria = mfhn.Triangulation("annulus", L, "p4est")
dh = mfhn.DoFHandler(tria, k)
mf = mfhn.MatrixFree(dh)
op = mfhn.LaplaceOperator(mf, number="double", kernel="plane")
src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
src.copy_(torch.sin(torch.arange(src.numel(), device=src.device, dtype=src.dtype)))
for kernel in kernels:
    op.set_kernel(kernel)
    for _ in range(2):
        op.vmult(dst, src)
    torch.cuda.synchronize()
    print(kernel, float(dst.abs().max()))
