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


def test_reference_arm_under_torchrun_reports_the_threads_it_has():
    # the driver launches the reference arm like the B200 arm (torchrun for N > 1, which exports OMP_NUM_THREADS=1):
    # rank 0 alone works, undoes the thread limit and reports what it really used; the other ranks exit 0 silently
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--nx", "256",
                        "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-1500:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    cb = d["cpu_baseline"]
    assert d["n_gpus"] == 2 and cb["launched_with_world_size"] == 2
    assert cb["cores"] == len(os.sched_getaffinity(0)) and cb["os_cpu_count"] == os.cpu_count()
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--nx", "256",
                        "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_partitioned_reference_and_algorithmic_bytes():
    # the committed N = 1 time of the partitioned leg (what efficiency_vs_n1 divides by at N > 1) and SURVEY 8(d)'s B_alg
    sys.path.insert(0, ROOT)
    import bench
    ms, src = bench.partitioned_n1_reference(1024)
    assert src == "profiles/r02_partitioned_n1.json" and 100.0 < ms < 400.0
    assert bench.partitioned_n1_reference(512) == (None, None)
    k = bench.partitioned_n1_kernels(1024)
    assert set(k) == {"zkernel", "yinv", "xkernel", "yfwd"} and all(v["ms"] > 0 for v in k.values())
    assert bench.b_alg(2) == 432 and bench.b_alg(3) == 560 and bench.b_alg(1) == 304
    assert bench.b_alg(2, "FilteredRK4") == 436 and bench.b_alg(2, "ETDRK4") == 416
    assert bench.host_threads() == len(os.sched_getaffinity(0))
