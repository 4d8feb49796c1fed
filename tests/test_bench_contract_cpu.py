"""CPU-only checks of bench.py's contract: the reference arm prints one JSON line with the required keys, and the B200 arm
refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--nx", "512", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("grid-point RK4 steps/sec") and d["unit"] == "grid-point-steps/s" and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    r = _run("--steps", "1", "--warmup", "3", timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
