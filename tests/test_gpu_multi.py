"""Multi-GPU parity (-m gpu, needs >= 2 GPUs: skipped on a single-GPU box): slab-decomposed 3-D steps under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port, marker, nproc=2, timeout=900):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs >= {nproc} GPUs (run {script} under torchrun on a multi-GPU box)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert marker in r.stdout


def test_slab2d_parity_two_ranks():
    _torchrun("mgpu_slab2d_check.py", 29518, "SLAB2D PARITY OK")


def test_batch_decomposition_parity_two_ranks():
    # PTF_DECOMP_BATCH (BASELINE configs[4]'s sharding): every member against the oracle
    _torchrun("mgpu_batch_check.py", 29519, "BATCH PARITY OK")


def test_slab_parity_four_ranks():
    _torchrun("mgpu_slab_check.py", 29520, "SLAB PARITY OK", nproc=4)


def test_slab_parity_two_ranks():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run tests/mgpu_slab_check.py under torchrun on a multi-GPU box)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_slab_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SLAB PARITY OK" in r.stdout
