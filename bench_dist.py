"""Problem construction shared by bench.py: one GPU (whole mesh) or one rank
of a Morton-partitioned mesh with ghost exchange (world > 1)."""
from __future__ import annotations

import numpy as np

def _log(msg):
    import os, sys, time

    if os.environ.get("MFHN_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


KERNEL_NAMES = {1: "qpoint", 2: "separable", 3: "baseline", 4: "plane", 5: "patch", 6: "bulk", 7: "runs", 8: "qpoint_rows"}


def build_problem(mfhn, args, L, rank, world):
    import torch

    tria = mfhn.Triangulation(args.geometry, L, "p4est")
    if world == 1:
        dh = mfhn.DoFHandler(tria, args.degree)
        mf = mfhn.MatrixFree(dh)
        geometry = high_order_geometry(mfhn, tria, mf, args.degree) if getattr(args, "mapping", "cartesian") == "high-order" else None
        op = mfhn.LaplaceOperator(mf, number=args.number, kernel=args.kernel, geometry=geometry)
        del geometry
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

    kname = KERNEL_NAMES[int(op.query("kernel"))]
    return {"op": op, "mf": mf, "dh": dh, "tria": tria, "n_dofs": dh.n_dofs(), "n_cells_global": tria.n_active_cells(),
            "n_cells_hn_global": tria.n_cells_with_hanging_nodes(), "fill_src": fill_src, "kernel_name": kname,
            "partition": partition, "launches_per_step": launches, "comm": comm}


def degree_sweep(mfhn, torch, args, time_vmult, hbm_peak, full=False):
    """BASELINE.json metric "per degree 1-8" (configs 2 and 3): every degree on the annulus mesh in double and float,
    with and without constraints, through the kernel AUTO picks; frac = accumulating-vmult bytes / time / HBM peak.
    Every degree runs on the headline mesh (L = 9: 2.3 M DoFs at k = 1 ... 1.12 B at k = 8; the L = 8 mesh SURVEY 8d
    suggests for k >= 5 has twice the share of constrained cells and too few cells to fill the GPU: k = 5 reads 120 GDoF/s /
    20 % overhead there and 135 / 6 % here); full=True adds L = 10 for k = 1, 2 and the other kernels."""
    res = []
    for k in range(1, 9):
        base = 9 if args.refinements is None else max(args.refinements - (0 if k <= 4 else 1), 2)
        for L in ([base, base + 1] if (full and k <= 2) else [base]):
            tria = mfhn.Triangulation(args.geometry, L, "p4est")
            dh = mfhn.DoFHandler(tria, k)
            mf = mfhn.MatrixFree(dh)
            for number in ("double", "float"):
                op = mfhn.LaplaceOperator(mf, number=number)
                row = {"degree": k, "L": L, "number": number, "n_dofs": dh.n_dofs(), "kernel": KERNEL_NAMES[int(op.query("kernel"))]}
                src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
                i = torch.arange(src.numel(), device=src.device, dtype=torch.float64)
                src.copy_(torch.sin(1e-3 * i).to(src.dtype))
                t = {}
                for ac in (True, False):
                    op.set_apply_constraints(ac)
                    _, per = time_vmult(torch, op, dst, src, 8, 3)
                    t[ac] = float(np.mean(per))
                row["gdofs"] = round(dh.n_dofs() / (t[True] * 1e-3) / 1e9, 2)
                row["gdofs_no_constraints"] = round(dh.n_dofs() / (t[False] * 1e-3) / 1e9, 2)
                row["hn_overhead_percent"] = round(100.0 * (t[True] / t[False] - 1.0), 1)
                row["frac_hbm"] = round(op.query("algorithmic_bytes_accumulate") / (t[True] * 1e-3) / 1e9 / hbm_peak, 3)
                if full:
                    op.set_apply_constraints(True)
                    for kern in ("plane", "bulk", "runs", "separable", "qpoint"):
                        try:
                            op.set_kernel(kern)
                            _, per = time_vmult(torch, op, dst, src, 5, 2)
                            row[f"{kern}_gdofs"] = round(dh.n_dofs() / (float(np.mean(per)) * 1e-3) / 1e9, 2)
                        except mfhn.MfhnError:
                            continue
                res.append(row)
                del op, src, dst
                torch.cuda.empty_cache()
    return res


def parity_check(mfhn, torch, dist, args, rank, world, device):
    """Partitioned vmult (both exchanges: NCCL import / compress and peer-memory access) against the one-GPU operator
    of the same small mesh with the reference's non-constant test vector src = sum_d sin(x_d) (benchmark_03.h:362-378);
    the one-GPU operator itself is pinned against the oracle in tests/.  A constant vector would not notice a broken
    ghost -> owner compress.  Returns {exchange: max |y_partitioned - y_one_gpu| / max |y_one_gpu|}."""
    from importlib import import_module

    distributed = import_module("dealii-matrixfree-hanging-nodes_b200.distributed")
    L = 6 if args.geometry == "annulus" else 5
    if args.degree >= 6:
        L -= 1
    tria = mfhn.Triangulation(args.geometry, L, "p4est")
    dtype = torch.float64 if args.number == "double" else torch.float32
    # whole mesh on this GPU (serial numbering of the same mesh: compare through the support points)
    dh1 = mfhn.DoFHandler(tria, args.degree)
    op1 = mfhn.LaplaceOperator(mfhn.MatrixFree(dh1), number=args.number, kernel=args.kernel)
    x1 = torch.from_numpy(np.sin(dh1.support_points()).sum(axis=1)).to(device=device, dtype=dtype)
    s1, d1 = op1.initialize_dof_vector(), op1.initialize_dof_vector()
    s1.copy_(x1)
    op1.vmult(d1, s1)
    ref_scale = float(d1.abs().max())
    y1 = d1.double().cpu().numpy()
    dhp = mfhn.DoFHandler(tria, args.degree, world, tria.partition(world, args.hn_weight))
    # serial index of every DoF of the rank-major partitioned numbering, through the cell-wise (unsubstituted) index arrays
    cells = np.arange(tria.n_active_cells())
    r1 = dh1.fill(cells, raw=True, substituted=False, masks=False, h=False)[0]
    rp = dhp.fill(cells, raw=True, substituted=False, masks=False, h=False)[0]
    to_serial = np.zeros(dhp.n_dofs(), dtype=np.int64)
    to_serial[rp.reshape(-1).astype(np.int64)] = r1.reshape(-1).astype(np.int64)
    mf = mfhn.MatrixFree(dhp, rank)
    mfhn.exchange_import_indices(mf.partitioner)
    op = mfhn.LaplaceOperator(mf, number=args.number, kernel=args.kernel)
    _log("parity: communicator")
    comm = distributed.GhostExchange(op)
    op.attach_communicator(comm)
    b, e = mf.partitioner.begin, mf.partitioner.end
    xo = x1[torch.from_numpy(to_serial[b:e]).to(device)]
    ref = torch.from_numpy(y1[to_serial[b:e]]).to(device)
    out = {}
    for exchange in ("nccl", "peer"):
        _log(f"parity: {exchange}")
        if exchange == "peer":
            dst, src = comm.enable_peer()
        else:
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        _log("parity: vectors")
        src.zero_()
        src[:e - b] = xo
        for _ in range(2):  # twice: the second application starts from a used ghost section
            op.vmult(dst, src, zero_dst=True)
        torch.cuda.synchronize()
        _log("parity: applied")
        err = (dst[:e - b].double() - ref).abs().max().reshape(1) / ref_scale
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        out[exchange] = float(err.item())
    out["mesh"] = f"{args.geometry} L={L}, {tria.n_active_cells()} cells, {dh1.n_dofs()} DoFs, src = sum sin(x_d)"
    dist.barrier()
    _log("parity: done")
    # the communicator (a second NCCL communicator and IPC mappings) stays alive until the process ends: tearing it down
    # in the middle of the run blocked one rank on this pool
    return out, (comm, op, mf, dhp, tria)


def stage_benchmarks(mfhn, torch, args, L, time_vmult):
    """The reference's decomposition (benchmark_01.cc:189-220) on the GPU: "DG (SC)" = every cell owns private
    DoFs (contiguous cell-local gather / scatter, benchmark_01.h:639-677) with and without the interpolation
    (t2, t3, eta3); "CG (SC)" = the real operator (t4, t5, eta5); and the interpolation alone on cell-local
    values (benchmark_00_likwid.cc:56-59)."""
    import numpy as np

    tria = mfhn.Triangulation(args.geometry, L, "p4est")
    dh = mfhn.DoFHandler(tria, args.degree)
    mf = mfhn.MatrixFree(dh)
    n3 = (args.degree + 1) ** 3
    n_hn, n_all = mf.n_cells_hn(), mf.n_cells

    def eta(t_n, t_hn):
        return max((t_hn / (t_n / n_all) - (n_all - n_hn)) / n_hn, 1.0) if n_hn else 1.0

    res = {"n_cells": n_all, "n_cells_hn": n_hn}
    # CG (SC)
    op = mfhn.LaplaceOperator(mf, number=args.number)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.fill_(1.0)  # benchmark_01.h:510-511
    op.set_apply_constraints(False)
    _, t4 = time_vmult(torch, op, dst, src, 10, 3)
    op.set_apply_constraints(True)
    _, t5 = time_vmult(torch, op, dst, src, 10, 3)
    res.update(t4_ms=float(t4.mean()), t5_ms=float(t5.mean()), eta5=eta(float(t4.mean()), float(t5.mean())))
    # CG (SC) with the general-purpose constraint algorithm (benchmark_01.cc:222-234: t6, t7, eta7).  Both algorithms on
    # the SAME cell kernel (the q-point kernel): t6 without constraints, t7 constraint rows in gather / scatter,
    # t5_qpoint the fast interpolation passes
    op.set_kernel("qpoint")
    op.set_apply_constraints(False)
    _, t6 = time_vmult(torch, op, dst, src, 5, 2)
    op.set_apply_constraints(True)
    _, t5q = time_vmult(torch, op, dst, src, 5, 2)
    op.set_kernel("qpoint_rows")
    _, t7 = time_vmult(torch, op, dst, src, 5, 2)
    res.update(t6_ms=float(t6.mean()), t7_ms=float(t7.mean()), t5_qpoint_ms=float(t5q.mean()), eta7=eta(float(t6.mean()), float(t7.mean())),
               eta5_qpoint=eta(float(t6.mean()), float(t5q.mean())), constraint_row_entries=int(op.query("constraint_row_entries")))
    op.set_kernel(args.kernel)
    # the interpolation alone
    vals = torch.ones(mf.n_cells * n3, dtype=src.dtype, device=src.device)
    for _ in range(3):
        op.apply_hanging_node_constraints(vals, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        op.apply_hanging_node_constraints(vals, False)
    e1.record()
    torch.cuda.synchronize()
    res["hn_kernel_alone_ms"] = e0.elapsed_time(e1) / 10
    # DG (C): cell-local vectors, no quadrature-point work (benchmark_01.cc:189-199): t0 plain copy, t1 with interpolation
    dvals = torch.zeros_like(vals)

    def time_dg(ac):
        op.set_apply_constraints(ac)
        for _ in range(3):
            op.dg_copy(dvals, vals)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            op.dg_copy(dvals, vals)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 10

    t0, t1 = time_dg(False), time_dg(True)
    op.set_apply_constraints(True)
    res.update(t0_ms=t0, t1_ms=t1, eta1=eta(t0, t1))
    del op, src, dst, vals, dvals
    torch.cuda.empty_cache()
    # DG (SC): private DoFs per cell
    class _DG:
        pass

    dg = _DG()
    dg.degree, dg.n_cells, dg.h, dg.masks = mf.degree, mf.n_cells, mf.h, mf.masks
    dg.dof_indices = np.arange(mf.n_cells * n3, dtype=np.uint32).reshape(mf.n_cells, n3)
    dg.n_interior_cells = dg.n_interior_a = mf.n_cells
    dg.partitioner = mfhn.Partitioner(owned_range=(0, mf.n_cells * n3))
    op = mfhn.LaplaceOperator(dg, number=args.number)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.fill_(1.0)
    op.set_apply_constraints(False)
    _, t2 = time_vmult(torch, op, dst, src, 10, 3)
    op.set_apply_constraints(True)
    _, t3 = time_vmult(torch, op, dst, src, 10, 3)
    res.update(t2_ms=float(t2.mean()), t3_ms=float(t3.mean()), eta3=eta(float(t2.mean()), float(t3.mean())))
    return res


def cg_benchmark(mfhn, torch, dist, args, L, rank, world, keep=None):
    """BASELINE.json config 5: conjugate gradients with point-Jacobi on the adaptive mesh (degree 6 by default), the whole
    iteration inside the library (mfhn_cg_solve: one vmult, two fused vector kernels, one batched 3-scalar all-reduce per
    iteration).  The right-hand side is A x* for a random x*, so the system is consistent.  Reports the device time per
    iteration split into vmult / vector kernels / all-reduce."""
    prob = build_problem(mfhn, args, L, rank, world)
    op, mf = prob["op"], prob["mf"]
    n = op.n_owned
    xs, b, x = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
    g = torch.Generator(device=xs.device)
    g.manual_seed(1234 + rank)
    xs[:n] = torch.rand(n, generator=g, device=xs.device, dtype=torch.float64).to(xs.dtype) - 0.5
    op.vmult(b, xs, zero_dst=True)
    inv = mfhn.inverse_diagonal(op)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    its, hist, split = mfhn.solve_cg(op, x, b, inverse=inv, rel_tol=args.cg_tol, max_iter=args.cg_iterations, check_every=10, timings=True)
    t = torch.tensor([split["ms_total"], split["ms_vmult"], split["ms_vector_ops"], split["ms_allreduce"]], device=x.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot, mv, vec, ar = (float(v) / max(its, 1) for v in t.tolist())
    if keep is not None:
        keep.append((prob, op, mf, xs, b, x, inv))  # communicators stay alive until the process ends
    s = 8 if args.number == "double" else 4
    n_dofs = prob["n_dofs"]
    return {"metric": "cg_jacobi_time_per_iteration", "value": tot, "unit": "ms", "higher_is_better": False, "n_gpus": world,
            "iterations": its, "residual_reduction": hist[-1] / hist[0] if hist and hist[0] > 0 else None,
            "ms_per_iteration": {"total": tot, "vmult": mv, "vector_kernels": vec, "all_reduce": ar},
            "gdofs_per_iteration": n_dofs / (tot * 1e-3) / 1e9,
            "vector_kernels_bytes_per_iteration": 15 * s * n_dofs,  # update: 6 reads + 5 writes, dots: 3 reads (+ inverse diagonal)
            "vector_kernels_gbs": 15 * s * n_dofs / (vec * 1e-3) / 1e9 if vec > 0 else None,
            "dtype": "f64" if args.number == "double" else "f32", "data": "synthetic (b = A x*, x* random)",
            "config": {"workload": f"{args.geometry} L={L}, FE_Q({args.degree}), {args.number}, CG + point-Jacobi", "n_dofs": int(n_dofs),
                       "n_cells": int(prob["n_cells_global"]), "kernel": prob["kernel_name"], "partition": prob["partition"],
                       "tolerance": args.cg_tol, "max_iterations": args.cg_iterations}}


def high_order_benchmark(mfhn, torch, args, time_vmult, hbm_peak, L=8):
    """Per-quadrature-point geometry (TestHighOrderMapping, benchmark_01.h:225-242) through the q-point kernel on the
    annulus mesh one level below the headline mesh (six coefficients per quadrature point: 1.6 GB at L=8, k=4)."""
    tria = mfhn.Triangulation(args.geometry, L, "p4est")
    dh = mfhn.DoFHandler(tria, args.degree)
    mf = mfhn.MatrixFree(dh)
    G = high_order_geometry(mfhn, tria, mf, args.degree)
    op = mfhn.LaplaceOperator(mf, number=args.number, geometry=G)
    del G
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.copy_(torch.sin(1e-3 * torch.arange(src.numel(), device=src.device, dtype=torch.float64)).to(src.dtype))
    _, per = time_vmult(torch, op, dst, src, 10, 3)
    t = float(np.mean(per))
    b = op.query("algorithmic_bytes_accumulate")
    return {"workload": f"{args.geometry} L={L}, FE_Q({args.degree}), {args.number}, high-order mapping (per-quadrature-point JxW J^-1 J^-T)",
            "n_dofs": dh.n_dofs(), "kernel": KERNEL_NAMES[int(op.query("kernel"))], "ms_per_step": t, "gdofs": dh.n_dofs() / (t * 1e-3) / 1e9,
            "algorithmic_bytes_per_launch": b, "frac_hbm": b / (t * 1e-3) / 1e9 / hbm_peak}


def high_order_geometry(mfhn, tria, mf, degree, amplitude=1e-6):
    """Per-quadrature-point coefficients JxW J^-1 J^-T ([cell][6][(k+1)^3], xx xy xz yy yz zz) of the reference's
    TestHighOrderMapping (benchmark_01.h:225-242): the Cartesian mesh displaced by amplitude * sin(pi x_d) in every
    coordinate direction (a smooth deformation resolved by the high-order mapping), evaluated at the Gauss points."""
    k, n = degree, degree + 1
    q, w = np.polynomial.legendre.leggauss(n)
    q, w = 0.5 * (q + 1.0), 0.5 * w
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :]).ravel()
    cells = tria.cells()[mf.cell_ids]  # (level, i, j, k)
    h = mf.h
    org = -1.0 + cells[:, 1:4] * h[:, None]
    qq = (np.tile(q, n * n), np.tile(np.repeat(q, n), n), np.repeat(q, n * n))
    G = np.zeros((mf.n_cells, 6, n ** 3))
    chunk = 65536
    for a in range(0, mf.n_cells, chunk):
        b = min(a + chunk, mf.n_cells)
        X = [org[a:b, d, None] + h[a:b, None] * qq[d][None, :] for d in range(3)]
        # x_d = X_d + amplitude sin(pi X_d) sin(pi X_{d+1}): a full (non-diagonal) Jacobian
        J = np.zeros((b - a, n ** 3, 3, 3))
        for d in range(3):
            e = (d + 1) % 3
            J[..., d, d] = 1.0 + amplitude * np.pi * np.cos(np.pi * X[d]) * np.sin(np.pi * X[e])
            J[..., d, e] = amplitude * np.pi * np.sin(np.pi * X[d]) * np.cos(np.pi * X[e])
        J *= h[a:b, None, None, None]
        det = np.linalg.det(J)
        Ji = np.linalg.inv(J)
        M = np.einsum("cqik,cqjk->cqij", Ji, Ji) * (det * w3[None, :])[..., None, None]
        for c, (i, j) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
            G[a:b, c, :] = M[..., i, j]
    return G
