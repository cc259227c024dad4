"""bench.py contract: the reference arm runs on the host (no GPU), prints exactly ONE JSON line on
stdout with the agreed keys; the GPU arm's contract is checked on the B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--refinements", "6", "--degree", "2"], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "laplace_vmult_throughput" and d["unit"] == "GDoF/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["n_cells_hn"] > 0
    cb = d["cpu_baseline"]  # the sample is spread over the whole cell loop: its share of hanging-node cells is the mesh's
    assert cb["sample_n_cells"] > 0 and abs(cb["sample_hn_fraction"] - cb["mesh_hn_fraction"]) < 0.05


@pytest.mark.gpu
def test_gpu_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--refinements", "6"],
                         check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["n_gpus"] == 1 and d["gpu_launches"] == 5 and d["value"] > 0  # 5 steps x one cell kernel
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    sweep = d["degree_sweep"]  # degrees 1..8 in double and float through the kernel AUTO picks
    assert sorted({(r["degree"], r["number"]) for r in sweep}) == sorted((k, nb) for k in range(1, 9) for nb in ("double", "float"))
    assert all(r["gdofs"] > 0 and 0 < r["frac_hbm"] < 1.5 for r in sweep)
    assert set(d["config"]) == {"workload", "n_cells", "n_cells_hn", "n_dofs", "l2", "dst", "exchange"} and "kernel" in d["engine"]
