"""bench.py contract on CPU: the reference arm prints one JSON line with the keys the driver reads, other ranks of a
torchrun launch exit quietly, and the B200 arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-columns", "4096"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "column-layer-steps/sec" and d["unit"] == "column-layer-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert abs(d["value"] - 4096 * 30 * 2 / (d["ms_per_step"] * 2e-3)) <= 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_reference_benchmark_design_script_runs_on_the_cpu_arm(tmp_path):
    out = str(tmp_path / "design.csv")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "reference_benchmark_design.py"), "--engine", "oracle",
                        "--max-i", "2", "--samples", "2", "--out", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = open(out).read().strip().splitlines()
    assert rows[0].split(",")[:4] == ["nrings", "npoints", "min_time_ms", "mid_time_ms"] and len(rows) == 3
    assert [int(x.split(",")[1]) for x in rows[1:]] == [32, 128]          # FullGaussianGrid(2), (4): 8 n^2 points
