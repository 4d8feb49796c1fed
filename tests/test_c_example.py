"""The drop-in boundary is a real C ABI: include/ptf_b200.h compiles as pedantic C99 and a plain-C program drives the
library (examples/c_api_example.c).  CPU: it must fail loudly with PTF_ENODEVICE; GPU: it steps the cellular flow."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "passivetracerflows.jl_b200")


def _build(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "c_api_example")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_api_example.c"), "-L", LIBDIR, "-lptf_b200", "-lm", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_example_compiles_as_c99_and_fails_loudly_without_a_gpu(tmp_path):
    import ptf_b200  # noqa: F401  (makes sure the library is built)
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("CPU-only part")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "PTF_ENODEVICE" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_device(tmp_path):
    import ptf_b200  # noqa: F401
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.findall(r"mean\(c\) = ([-\d.e+]+)\s+var\(c\) = ([-\d.e+]+)", r.stdout)
    assert len(m) == 2
    (m0, v0), (m1, v1) = [(float(a), float(b)) for a, b in m]
    assert abs(m0 - m1) < 1e-12 and 0 < v1 < v0          # mass conserved, variance decays (advection-diffusion)
    assert "step 100" in r.stdout
