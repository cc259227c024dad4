#!/usr/bin/env python
"""Benchmark of the hot path: 3D Laplace vmult, FE_Q(k), adaptively refined mesh
with hanging nodes (reference driver: benchmark_03.h:382-546, `./benchmark_03
cuda annulus 4`, cuda/run.sh).

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference

Prints ONE JSON line (rank 0).  metric = Laplace vmult throughput in GDoF/s,
n_dofs = dof_handler.n_dofs() as tabulated by the reference (benchmark_03.h:441).
A step is one vmult (dst += A src, accumulating like benchmark_03.h:352) over
the whole mesh.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun pins OMP_NUM_THREADS=1; the host-side setup (mesh, DoF enumeration) is OpenMP-parallel
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))
PKG = "dealii-matrixfree-hanging-nodes_b200"

METRIC = "laplace_vmult_throughput"
UNIT = "GDoF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)  # n_repetitions = 100, benchmark_03.h:393
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--geometry", default="annulus")
    ap.add_argument("--refinements", type=int, default=None)
    ap.add_argument("--degree", type=int, default=None)
    ap.add_argument("--number", default="double", choices=["double", "float"])
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="partitioned runs: NCCL ghost import/compress (default) or peer-memory access of the owners' vectors")
    ap.add_argument("--hn-weight", type=float, default=1.0, help="partition weight of cells with hanging nodes (benchmark_02.cc:15-37)")
    ap.add_argument("--sweep", action="store_true", help="degree sweep with L=10 for k=1,2 and all kernel variants (the default line carries the compact sweep)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the compact degree sweep (degrees 1..8, double + float) of the default line")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys cg_jacobi (config 5; N = 1 and 8) and high_order_mapping (N = 1) of the default line")
    ap.add_argument("--no-weak", action="store_true", help="8 GPUs: skip the extra weak-scaling run on the next finer mesh")
    ap.add_argument("--stages", action="store_true", help="the reference's DG (SC) / CG (SC) decomposition and eta (benchmark_01.cc:189-220)")
    ap.add_argument("--mapping", default="cartesian", choices=["cartesian", "high-order"],
                    help="high-order: per-quadrature-point geometry of the reference's TestHighOrderMapping (benchmark_01.h:225-242), q-point kernel, own byte model")
    ap.add_argument("--cg", action="store_true", help="BASELINE.json config 5: CG + point-Jacobi solve (degree 6 unless --degree is given), time per iteration split")
    ap.add_argument("--cg-iterations", type=int, default=100)
    ap.add_argument("--cg-tol", type=float, default=1e-8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--minimal", action="store_true", help="timed loop only (for profiler runs)")
    ap.add_argument("--graph", action="store_true", help="partitioned runs: replay a CUDA graph of one vmult (experimental: the capture of the NCCL p2p groups hung on this pool, so it is off by default)")
    args = ap.parse_args()
    if args.mapping == "high-order" and (args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1):
        ap.error("--mapping high-order is a one-GPU mode")
    if args.degree is None:
        args.degree = 6 if args.cg else 4
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def default_refinements(args):
    """The same mesh at every N (strong scaling) and every degree: annulus (p4est flavour) L=9 (142.8 M DoFs at k=4;
    SURVEY 8d / BASELINE.md C2).  The 8-GPU weak-scaling run on L+1 is an extra key of the N=8 line."""
    if args.refinements is not None:
        return args.refinements
    if args.mapping == "high-order":
        return 8 if args.degree <= 4 else 7  # 6 s (k+1)^3 bytes of coefficients per cell: 1.6 GB at L=8, k=4
    if args.geometry == "annulus":
        return 9  # every degree on the same mesh: 142.8 M DoFs at k=4, 477 M at k=6, 1.12 B at k=8
    return 8 if args.degree <= 4 else 7


def time_vmult(torch, op, dst, src, steps, warmup, barrier=None, graph=None):
    """K steps bracketed by synchronize (+barrier), per-launch CUDA events on the launching stream."""
    if graph is not None:
        class _G:
            @staticmethod
            def vmult(d, s):
                graph.replay()
        op = _G
    for _ in range(warmup):
        op.vmult(dst, src)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(steps):
        op.vmult(dst, src)
        ev[i + 1].record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])  # ms
    return ev[0].elapsed_time(ev[steps]), per


def cpu_sample(mf, n_sample_cells, n_windows=100):
    """Bounded sample of the workload for the CPU arm: n_windows windows of consecutive cells, evenly spaced over the
    whole cell loop (Morton order), so that the sample's share of cells with hanging nodes matches the mesh's; the
    DoFs of the sample are renumbered compactly."""
    ns = min(n_sample_cells, mf.n_cells)
    w = max(ns // n_windows, 1)
    starts = np.linspace(0, mf.n_cells - w, num=max(ns // w, 1)).astype(np.int64)
    cells = np.unique((starts[:, None] + np.arange(w)[None, :]).reshape(-1))
    idx = mf.dof_indices[cells]
    uniq, inv = np.unique(idx, return_inverse=True)
    masks = np.ascontiguousarray(mf.masks[cells])
    return inv.reshape(idx.shape).astype(np.uint32), masks, np.ascontiguousarray(mf.h[cells]), len(uniq), len(cells), int((masks != 0).sum())


def run_cpu(args, mf, degree, n_rep, n_sample_cells=200_000, threads=None, warmup=1):
    from oracle import cpu

    threads = threads or os.cpu_count()
    idx, masks, h, nd, ns, ns_hn = cpu_sample(mf, n_sample_cells)
    for _ in range(max(warmup, 1)):
        cpu.benchmark(degree, idx, masks, h, nd, True, 1, threads)  # spin up the thread pool
    t = cpu.benchmark(degree, idx, masks, h, nd, True, n_rep, threads)
    return {"value": threads * nd / t / 1e9, "unit": UNIT, "cores": threads, "kind": "port", "sample_n_cells": ns, "sample_n_cells_hn": ns_hn,
            "sample_hn_fraction": ns_hn / max(ns, 1), "mesh_hn_fraction": mf.n_cells_hn() / max(mf.n_cells, 1),
            "sample": f"{ns} cells ({ns_hn} with hanging nodes, {nd} DoFs) of the same mesh: 100 windows of consecutive cells evenly spaced over the "
                      f"Morton-ordered cell loop, src=1, {n_rep} reps per thread; every thread applies the operator to its own vectors "
                      f"(benchmark_01.h:536-573); C restatement of the deal.II CPU path (AVX-512 across 8 cells, even-odd sum factorisation); "
                      f"deal.II itself cannot be built here"}, t, nd


def problem_config(args, workload, tria, n_dofs, world):
    """Workload description shared by both arms (the reference arm prints the same object)."""
    return {"workload": workload, "n_cells": int(tria.n_active_cells()), "n_cells_hn": int(tria.n_cells_with_hanging_nodes()), "n_dofs": int(n_dofs),
            "l2": "inputs larger than L2 (vectors + index arrays >> 126 MB), no flush", "dst": "accumulating vmult like benchmark_03.h:352",
            "exchange": args.exchange if world > 1 else None}


def log(msg):
    """Progress on stderr (MFHN_BENCH_VERBOSE=1): where a multi-rank run is when it has to be killed."""
    if os.environ.get("MFHN_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def main():
    # keep stdout clean for the ONE JSON line: library chatter (e.g. "NCCL version ...") goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run()
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    if line is not None:
        print(line, flush=True)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # partitioned runs: leave without tearing down NCCL communicators and IPC mappings rank by rank (the ranks would
        # wait for each other's teardown in interpreter-exit order); the process group was destroyed in run()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    L = default_refinements(args)
    mapping = "Cartesian MappingQ1" if args.mapping == "cartesian" else "high-order mapping (1e-6 sin displacement, per-quadrature-point JxW J^-1 J^-T, benchmark_01.h:225-242)"
    workload = f"{args.geometry} L={L} (p4est-balanced octree on hyper_cube(-1,1)^3), FE_Q({args.degree}), {args.number}, {mapping}"

    mfhn = importlib.import_module(PKG)

    if args.impl == "reference":
        if rank != 0:
            return None
        tria = mfhn.Triangulation(args.geometry, L, "p4est")
        dh = mfhn.DoFHandler(tria, args.degree)
        mf = mfhn.MatrixFree(dh)
        t0 = time.perf_counter()
        cb, t, nd = run_cpu(args, mf, args.degree, args.steps, warmup=args.warmup)
        wall = time.perf_counter() - t0
        v = cb["value"]
        return json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
                          "vs_baseline": None, "dtype": "f64" if args.number == "double" else "f32", "data": "synthetic",
                          "config": problem_config(args, workload, tria, dh.n_dofs(), args.gpus),
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "wall_s": wall})

    import torch

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    barrier = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        barrier = dist.barrier

    from bench_dist import build_problem  # noqa: E402  (shared by the 1-GPU and the partitioned path)

    if args.cg:
        from bench_dist import cg_benchmark

        res = cg_benchmark(mfhn, torch, dist if world > 1 else None, args, L, rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return json.dumps(res) if rank == 0 else None
    t_setup = time.perf_counter()

    log("setup")
    prob = build_problem(mfhn, args, L, rank, world)
    op, mf, n_dofs_global = prob["op"], prob["mf"], prob["n_dofs"]
    log("problem built")
    if world > 1 and args.exchange == "peer":
        dst, src = prob["comm"].enable_peer()  # exportable vector pair, peers' vectors mapped through CUDA IPC
    else:
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    prob["fill_src"](src)
    t_setup = time.perf_counter() - t_setup

    graph = None
    if world > 1 and args.graph:
        graph = prob["comm"].capture(op, dst, src)
        dst.zero_()
    log("timed loop")
    with ClockSampler(local_rank) as clocks:
        total_ms, per = time_vmult(torch, op, dst, src, args.steps, args.warmup, barrier, graph)
    log("timed loop done")
    rank_info = None
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        # per-rank load: time of the rank-local cell loop alone (no exchange), cells, hanging-node cells
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            op.vmult_range(dst, src, 0, mf.n_cells)
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1) / 10, mf.n_cells, mf.n_cells_hn(), mf.n_cells - mf.n_interior_cells,
                             mf.partitioner.n_ghost], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_info = {"local_cell_loop_ms": [round(float(a[0]), 4) for a in allr], "n_cells": [int(a[1]) for a in allr],
                     "n_cells_hn": [int(a[2]) for a in allr], "n_boundary_cells": [int(a[3]) for a in allr],
                     "n_ghost": [int(a[4]) for a in allr]}
        dst.zero_()
    ms_per_step = total_ms / args.steps
    value = n_dofs_global / (ms_per_step * 1e-3) / 1e9
    parity = None
    if world > 1:
        # partitioned vmult against the one-GPU operator on a small mesh, non-constant vector, both exchanges
        from bench_dist import parity_check

        log("parity check")
        parity, parity_keep = parity_check(mfhn, torch, dist, args, rank, world, src.device)
        log("parity check done")
        tol = 1e-12 if args.number == "double" else 1e-5
        assert parity["nccl"] < tol and parity["peer"] < tol, f"partitioned vmult differs from the one-GPU operator: {parity}"

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
           "dtype": "f64" if args.number == "double" else "f32", "data": "synthetic"}
    # kernels per step, counted by the engine itself: cell kernels of one vmult (the bulk-copy kernel adds a
    # plane-kernel launch for every cell it leaves out), plus pack / unpack in partitioned runs
    c0 = op.launch_count()
    op.vmult(dst, src)
    torch.cuda.synchronize()
    cell_launches = op.launch_count() - c0
    launches_per_step = cell_launches if world == 1 else prob["comm"].launches_per_vmult(cell_launches)
    out["gpu_launches"] = int(launches_per_step * args.steps)
    if parity is not None:
        out["parity_max_rel_err"] = parity

    # roofline of the dominant kernel (the fused cell kernel): algorithmic bytes / average launch time
    peak, peak_src = peaks()
    # vmult accumulates (dst += A src like the reference), so dst is read as well: + s n_dofs
    b_alg = op.query("algorithmic_bytes_accumulate")
    b_alg_plain = op.query("algorithmic_bytes")
    flops = op.query("algorithmic_flops")
    if world > 1:
        t = torch.tensor([b_alg, b_alg_plain, flops], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        b_alg, b_alg_plain, flops = (float(x) for x in t.tolist())
    kernel_ms = float(np.mean(per)) if world == 1 else ms_per_step
    achieved = b_alg / (kernel_ms * 1e-3) / 1e9
    peak = peak * world  # aggregate over the GPUs of the job
    # DRAM bytes per launch of this kernel from its ncu capture (profiles/traffic.json names the capture and the
    # commit of the kernel source it was taken from; null when this configuration has no capture)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            tj = json.load(f)
        entry = tj.get(f"{args.geometry}_L{L}_k{args.degree}_{args.number}_{prob['kernel_name']}")
        if isinstance(entry, dict):
            traffic, traffic_src = entry.get("bytes"), entry.get("source")
        else:
            traffic = entry
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                       "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg, "algorithmic_bytes_without_dst_read": b_alg_plain,
                       "bytes_note": "3 s n_dofs + n_cells (4 (k+1)^3 + 1 + 3 s): src read, dst read+write (accumulating vmult), uint32 indices, mask, Cartesian geometry"
                                     if args.mapping == "cartesian" else
                                     "3 s n_dofs + n_cells (4 (k+1)^3 + 1 + 6 s (k+1)^3): vectors, indices, mask, six coefficients per quadrature point",
                       "kernel_ms_min_avg_max": [float(per.min()), float(per.mean()), float(per.max())],
                       "algorithmic_flops_per_launch": flops}
    out["clocks"] = clocks.summary()
    if rank_info is not None:
        out["ranks"] = rank_info
    out["config"] = problem_config(args, workload, prob["tria"], n_dofs_global, world)
    out["engine"] = {"kernel": prob["kernel_name"], "partition": prob["partition"], "setup_s": round(t_setup, 1),
                     "launch": "CUDA graph replay of one partitioned vmult" if graph is not None else "host launches"}

    if rank == 0 and world == 1 and not args.minimal:
        # hanging-node overhead on the same mesh and index arrays (benchmark_03.h:255-268)
        op.set_apply_constraints(False)
        _, per_nc = time_vmult(torch, op, dst, src, max(args.steps // 2, 5), 3)
        op.set_apply_constraints(True)
        out["hn_overhead_percent"] = 100.0 * (float(np.mean(per)) / float(np.mean(per_nc)) - 1.0)
        out["no_constraints_gdofs"] = n_dofs_global / (float(np.mean(per_nc)) * 1e-3) / 1e9
        # cost of a cell with hanging nodes relative to a regular cell: eta of benchmark_01.cc:179-187 (CG (SC): t4, t5)
        n_hn, n_all = int(prob["n_cells_hn_global"]), int(prob["n_cells_global"])
        t_n, t_hn = float(np.mean(per_nc)), float(np.mean(per))
        out["eta5"] = max((t_hn / (t_n / n_all) - (n_all - n_hn)) / n_hn, 1.0) if n_hn else 1.0
    if rank == 0 and world == 1 and not args.minimal and args.mapping == "cartesian":
        # the other kernels on the same problem, for the record
        variants = {}
        for kname in ("plane", "bulk", "runs", "qpoint", "separable", "baseline"):
            try:
                op.set_kernel(kname)
                _, pk = time_vmult(torch, op, dst, src, 10, 3)
                variants[kname] = n_dofs_global / (float(np.mean(pk)) * 1e-3) / 1e9
            except mfhn.MfhnError as e:
                variants[kname] = str(e)
        op.set_kernel(args.kernel)
        out["kernel_variants_gdofs"] = variants
        out["kernel_variants_note"] = ("baseline = restatement of the deal.II CUDAWrappers::MatrixFree design behind cuda/benchmark_03.cu "
                                       "(one thread per DoF, per-q-point 3x3 inverse Jacobian + JxW from global memory, atomics), same box")
        out["fp64_fma_tflops_measured"] = mfhn.bench_fma("double", 20000)
        # hanging-node strategies: cells in plain Morton order ("index" analogue) vs grouped by
        # constraint flag inside Morton windows ("sorted" analogue, the reference's Categorize option; default)
        mf_plain = mfhn.MatrixFree(prob["dh"], categorize=False)
        op_plain = mfhn.LaplaceOperator(mf_plain, number=args.number, kernel=args.kernel)
        _, p1 = time_vmult(torch, op_plain, dst, src, 10, 3)
        op_plain.set_apply_constraints(False)
        _, p0 = time_vmult(torch, op_plain, dst, src, 10, 3)
        # "mask" analogue: every warp takes the interpolation passes, constrained lines picked by per-lane predicates
        op.set_hn_strategy("mask")
        _, pm = time_vmult(torch, op, dst, src, 10, 3)
        op.set_hn_strategy("branch")
        out["hn_strategies"] = {"sorted_overhead_percent": out["hn_overhead_percent"], "sorted_gdofs": value,
                                "index_overhead_percent": 100.0 * (float(np.mean(p1)) / float(np.mean(p0)) - 1.0),
                                "index_gdofs": n_dofs_global / (float(np.mean(p1)) * 1e-3) / 1e9,
                                "mask_overhead_percent": 100.0 * (float(np.mean(pm)) / float(np.mean(per_nc)) - 1.0),
                                "mask_gdofs": n_dofs_global / (float(np.mean(pm)) * 1e-3) / 1e9,
                                "note": "analogues of the reference's vectorisation types (benchmark_01.cc:70-116): index = plain Morton cell order, "
                                        "sorted = cells grouped by constraint kind inside Morton windows (default), mask = branch-free, all warps interpolate"}
        del op_plain, mf_plain

    log("roofline done")
    if not args.no_e2e and not args.minimal:
        # end to end with HOST vectors: H2D of src, vmult, D2H of dst every step
        n_local = src.numel()
        esz = src.element_size()
        e_steps = max(3, min(args.steps, 10))
        hbuf = [(torch.empty(n_local, dtype=src.dtype).pin_memory(), torch.empty(n_local, dtype=src.dtype).pin_memory()) for _ in range(2)]  # (dst, src)
        hbuf[0][1].copy_(src.cpu())
        hbuf[1][1].copy_(hbuf[0][1])
        main = torch.cuda.current_stream()
        if world == 1:
            # C ABI entry point on host vectors; two staging slots on two streams: the upload of step i+1 overlaps
            # the download of step i (full-duplex PCIe); every step still moves its own src up and dst down
            streams = [torch.cuda.Stream(), torch.cuda.Stream()]

            def host_step(i):
                with torch.cuda.stream(streams[i % 2]):
                    op.vmult_host(hbuf[i % 2][0], hbuf[i % 2][1], zero_dst=True, slot=i % 2)

            def drain():
                for s_ in streams:
                    main.wait_stream(s_)

            for s_ in streams:
                s_.wait_stream(main)
        else:
            # partitioned: two device vector pairs; an upload stream and a download stream run beside the vmult
            # stream, so that the copies of steps i+1 / i-1 overlap the vmult (ghost exchange inside) of step i
            up, down = torch.cuda.Stream(), torch.cuda.Stream()
            if args.exchange == "peer":
                dev = [(dst, src), (dst, src)]  # the registered pair is the only one the peer path takes: no overlap across steps
            else:
                dev = [(dst, src), (op.initialize_dof_vector(), op.initialize_dof_vector())]
            ev_up = [torch.cuda.Event() for _ in range(2)]
            ev_mv = [torch.cuda.Event() for _ in range(2)]
            ev_dn = [torch.cuda.Event() for _ in range(2)]
            for e_ in ev_mv + ev_dn:
                e_.record(main)

            def host_step(i):
                d_, s_ = dev[i % 2]
                with torch.cuda.stream(up):
                    up.wait_event(ev_mv[i % 2])  # the vmult that last read this src has finished
                    s_.copy_(hbuf[i % 2][1], non_blocking=True)
                    ev_up[i % 2].record(up)
                main.wait_event(ev_up[i % 2])
                main.wait_event(ev_dn[i % 2])  # the download that last read this dst has finished
                op.vmult(d_, s_, zero_dst=True)  # ghost import / compress inside
                ev_mv[i % 2].record(main)
                with torch.cuda.stream(down):
                    down.wait_event(ev_mv[i % 2])
                    hbuf[i % 2][0].copy_(d_, non_blocking=True)
                    ev_dn[i % 2].record(down)

            def drain():
                main.wait_stream(up)
                main.wait_stream(down)

        log("e2e warm-up")
        for i in range(2):
            host_step(i)
        drain()
        torch.cuda.synchronize()
        log("e2e timed")
        if barrier:
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1:
            for s_ in streams:
                s_.wait_stream(main)
        for i in range(e_steps):
            host_step(i)
        drain()
        e1.record()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        e_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        out["e2e"] = {"value": n_dofs_global / (e_ms / e_steps * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n_local * esz),
                      "d2h_bytes_per_step": int(n_local * esz), "steps": e_steps,
                      "note": "mfhn_op_vmult_host_slot: pinned host src -> device, vmult into a zeroed device dst, dst -> pinned host, every step; "
                              "two staging slots on two streams so that the upload of step i+1 overlaps the download of step i; PCIe-bound"}
        if world > 1:
            out["e2e"]["note"] = ("per rank: pinned host src -> device, partitioned vmult with ghost exchange, dst -> pinned host, every step; two vector pairs, "
                                  "upload / download streams beside the vmult stream; bytes are per rank")
            del hbuf

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.minimal and args.mapping == "cartesian":
        cb, _, _ = run_cpu(args, mf, args.degree, 5)
        out["cpu_baseline"] = cb

    if world == 1 and rank == 0 and not args.minimal and (args.sweep or not args.no_sweep) and args.mapping == "cartesian":
        from bench_dist import degree_sweep

        del src, dst
        torch.cuda.empty_cache()
        src = dst = None
        out["degree_sweep"] = degree_sweep(mfhn, torch, args, time_vmult, peaks()[0], full=args.sweep)
    if args.stages and world == 1 and rank == 0:
        from bench_dist import stage_benchmarks

        del op
        src = dst = None
        torch.cuda.empty_cache()
        out["stages"] = stage_benchmarks(mfhn, torch, args, L, time_vmult)

    force = False
    extras = not args.no_extras and not args.minimal and args.refinements is None and args.mapping == "cartesian" and args.degree == 4 and not args.stages
    force = bool(os.environ.get("MFHN_BENCH_FORCE_EXTRAS"))  # development: the 8-GPU extras at any N > 1
    if extras and (world in (1, 8) or force):
        # BASELINE.json config 5: CG + point-Jacobi at degree 6 on the same mesh (annulus L=9, 477 M DoFs)
        import copy

        from bench_dist import cg_benchmark

        a2 = copy.copy(args)
        a2.degree, a2.cg_iterations = 6, 60
        log("cg extra")
        extras_keep = []
        res = cg_benchmark(mfhn, torch, dist if world > 1 else None, a2, 9, rank, world, keep=extras_keep)
        out["cg_jacobi"] = {k: res[k] for k in ("value", "unit", "iterations", "residual_reduction", "ms_per_iteration", "gdofs_per_iteration",
                                                "vector_kernels_gbs", "config")}
        if world == 1:
            extras_keep.clear()
            torch.cuda.empty_cache()
    if extras and world == 1 and rank == 0:
        from bench_dist import high_order_benchmark

        log("high-order extra")
        out["high_order_mapping"] = high_order_benchmark(mfhn, torch, args, time_vmult, peaks()[0])

    if (world == 8 or (force and world > 1)) and not args.no_weak and not args.minimal and args.refinements is None:
        # weak scaling (BASELINE.json config 4): the next finer mesh on 8 GPUs (annulus L=10, k=4: 1.124 B DoFs, 140.5 M per GPU).
        # Not comparable with the N=1 line DoF for DoF: the finer mesh has half the share of cells with hanging nodes, so
        # its efficiency is quoted against this run's own rank-local cell loops.
        strong_keep = (op, prob, mf, src, dst)  # no communicator teardown in the middle of the run (see parity_check)
        log("weak-scaling problem")
        prob = build_problem(mfhn, args, L + 1, rank, world)
        op, mf = prob["op"], prob["mf"]
        if args.exchange == "peer":
            dst, src = prob["comm"].enable_peer()
        else:
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        prob["fill_src"](src)
        w_ms, _ = time_vmult(torch, op, dst, src, max(args.steps // 2, 5), 3, barrier)
        t = torch.tensor([w_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ms = float(t.item()) / max(args.steps // 2, 5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            op.vmult_range(dst, src, 0, mf.n_cells)
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        loc = [float(a[0]) for a in allr]
        out["weak_scaling"] = {"workload": f"{args.geometry} L={L + 1}", "n_dofs": int(prob["n_dofs"]), "n_cells": int(prob["n_cells_global"]),
                               "n_cells_hn": int(prob["n_cells_hn_global"]), "value": prob["n_dofs"] / (w_ms * 1e-3) / 1e9, "unit": UNIT,
                               "ms_per_step": w_ms, "local_cell_loop_ms": [round(x, 4) for x in loc],
                               "efficiency_vs_local_cell_loop": max(loc) / w_ms, "exchange": args.exchange}

    if world > 1:
        log("leaving")
        dist.barrier()
        dist.destroy_process_group()
    return json.dumps(out) if rank == 0 else None


if __name__ == "__main__":
    main()
