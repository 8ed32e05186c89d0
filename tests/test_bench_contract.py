"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (``--impl reference``: the C
port of the recurrence on the host cores) prints ONE JSON line with the keys the driver reads, and the last committed
GPU line (``profiles/r2f_bench_line.json``) is arithmetically consistent with itself (value = arcs / time, roofline
fractions = achieved / peak, algorithmic bytes = SURVEY 8(d)'s figures for the workload it names)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "arcs/s" and d["higher_is_better"] is True
    assert d["metric"] == "arcs/sec forward-backward (log semiring)" and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "lattices" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "arcs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "config4" in d["config"]["workload"]


def test_committed_gpu_line_is_self_consistent():
    path = os.path.join(ROOT, "profiles", "r2f_bench_line.json")
    if not os.path.exists(path):
        pytest.skip("no committed GPU line")
    with open(path) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    A, S = d["config"]["arcs_per_gpu"], d["config"]["states_per_gpu"]
    assert abs(d["value"] - A / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    both = r["algorithmic_bytes_per_launch"] + r["other_kernel"]["algorithmic_bytes_per_launch"]
    assert both == 24 * A + 4 * S  # pull 12 A + 4 S, flow 12 A; = 20 A + 20 S at S = A/4
    assert abs(r["step"]["achieved"] - (20 * A + 20 * S) / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["step"]["achieved"]
    assert d["warmup"] >= 3 and d["gpu_launches"] == 2 and d["dtype"] == "f32"
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert abs(e["value"] - A / (e["ms_per_step"] * 1e-3)) < 1e-6 * e["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert [s["arcs_per_lattice"] for s in d["sweep"]] == [10000, 30000, 100000, 300000, 1000000]
