"""GPU parity of the 2-D slab-decomposed engine (csrc/engine_slab2d.cu) against the CPU oracle.

Single-process part (any GPU box): the engine is run with P = 1 (`B200(decomposition="slab")` on one rank), which
exercises every kernel, plan and layout change except the NCCL exchange.  The 2-rank run lives in
tests/mgpu_slab2d_check.py (driven by test_slab2d_parity_two_ranks below when >= 2 GPUs are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.ptf_oracle import OracleProblem, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_STEP = 1e-12


def P():
    import ptf_b200
    return ptf_b200


U = lambda x, y: 0.2 * np.cos(x) * np.sin(2 * np.pi / 4.0 * y)
V = lambda x, y: -0.3 * np.sin(x) * np.cos(2 * np.pi / 4.0 * y)


def _c0(X, Y):
    return 0.5 * np.exp(-((X - 0.4) ** 2 / 0.3 + Y ** 2 / 0.2))


@pytest.mark.parametrize("stepper", ["RK4", "FilteredRK4", "ETDRK4", "LSRK54", "AB3", "ForwardEuler"])
@pytest.mark.parametrize("n", [(96, 64), (128, 80)])
def test_slab2d_single_rank_matches_oracle(stepper, n):
    nx, ny = n
    L = (2 * np.pi, 4.0)
    kw = dict(kappa=0.01, eta=0.02, dt=2e-3)
    prob = P().Problem(P().B200(decomposition="slab"), P().TwoDAdvectingFlow(u=U, v=V), nx=nx, Lx=L[0], ny=ny, Ly=L[1],
                       stepper=stepper, kappa_h=1e-6, n_kappa_h=2, **kw)
    assert prob.engine == "cufft" and prob.ny_phys_local == ny and prob.nkr_local == nx // 2 + 1
    X, Y = P().gridpoints(prob.grid)
    c0 = _c0(X, Y)
    o = OracleProblem(n=n, L=L, kappa=(0.01, 0.02), dt=2e-3, stepper=stepper, velocity=[U(X, Y), V(X, Y)], steady=True,
                      kappa_h=1e-6, n_kappa_h=2)
    o.set_c(c0)
    prob.set_c(c0)
    assert rel_l2(o.sol, prob.sol) < 1e-14
    o.stepforward(1)
    prob.stepforward(1)
    assert rel_l2(o.updatevars(), prob.updatevars()) < TOL_STEP
    o.stepforward(5)
    prob.stepforward(5)
    assert rel_l2(o.updatevars(), prob.updatevars()) < 5 * TOL_STEP
    assert rel_l2(o.sol, prob.sol) < 5 * TOL_STEP
    d = prob.diagnostics()
    assert abs(d["mean_c"] - o.c.mean()) < 1e-13 and abs(d["variance_c"] - o.c.var()) < 1e-13
    prob.close()


def test_slab2d_time_varying_separable_and_dealias():
    nx, ny, L = 64, 96, (2 * np.pi, 2 * np.pi)
    u = lambda x, y, t: (1 + 0.5 * np.sin(3 * t)) * np.cos(x) * np.sin(y)
    v = lambda x, y, t: -(1 + 0.5 * np.sin(3 * t)) * np.sin(x) * np.cos(y)
    flows = {
        "callback": P().TwoDAdvectingFlow(u=u, v=v, steadyflow=False),
        "separable": P().SeparableFlow(terms=[[(np.cos, np.sin)], [(np.sin, np.cos)]],
                                      coeffs=lambda t, a: [(1 + 0.5 * np.sin(3 * t)) * (1 if a == 0 else -1)],
                                      steadyflow=False),
    }
    for name, flow in flows.items():
        prob = P().Problem(P().B200(decomposition="slab"), flow, nx=nx, ny=ny, kappa=0.01, dt=5e-3, stepper="RK4",
                           dealias=True)
        X, Y = P().gridpoints(prob.grid)
        o = OracleProblem(n=(nx, ny), L=L, kappa=(0.01, 0.01), dt=5e-3, stepper="RK4", velocity=[u, v], steady=False,
                          dealias=True)
        o.set_c(_c0(X, Y))
        prob.set_c(_c0(X, Y))
        o.stepforward(6)
        prob.stepforward(6)
        assert rel_l2(o.updatevars(), prob.updatevars()) < 6 * TOL_STEP, name
        prob.close()


def test_slab2d_sol_round_trip_and_agrees_with_single_gpu_engine():
    nx = ny = 256
    flow = P().TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
    a = P().Problem(P().B200(decomposition="slab"), flow, nx=nx, kappa=0.002, dt=0.01, stepper="RK4")
    b = P().Problem(P().B200(), flow, nx=nx, kappa=0.002, dt=0.01, stepper="RK4")       # fused engine
    X, Y = P().gridpoints(a.grid)
    a.set_c(_c0(X, Y))
    b.set_c(_c0(X, Y))
    a.stepforward(10)
    b.stepforward(10)
    assert rel_l2(a.updatevars(), b.updatevars()) < 1e-12
    s = a.sol.copy()
    a.set_sol(s * 0.5)
    a.updatevars()
    assert rel_l2(a.sol, 0.5 * s) == 0.0
    a.close()
    b.close()


def test_slab2d_parity_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run tests/mgpu_slab2d_check.py under torchrun on a multi-GPU box)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "mgpu_slab2d_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SLAB2D PARITY OK" in r.stdout


def test_slab2d_chunked_forward_pipeline_single_rank(monkeypatch):
    """PTF_SLAB2D_CHUNKS pipelines the forward transform in row chunks (default 4 when P > 1); with P = 1 the same code
    path runs with local copies instead of the exchange."""
    monkeypatch.setenv("PTF_SLAB2D_CHUNKS", "4")
    nx, ny, L = 96, 64, (2 * np.pi, 4.0)
    prob = P().Problem(P().B200(decomposition="slab"), P().TwoDAdvectingFlow(u=U, v=V), nx=nx, Lx=L[0], ny=ny, Ly=L[1],
                       stepper="RK4", kappa=0.01, eta=0.02, dt=2e-3)
    X, Y = P().gridpoints(prob.grid)
    o = OracleProblem(n=(nx, ny), L=L, kappa=(0.01, 0.02), dt=2e-3, stepper="RK4", velocity=[U(X, Y), V(X, Y)], steady=True)
    o.set_c(_c0(X, Y))
    prob.set_c(_c0(X, Y))
    assert rel_l2(o.sol, prob.sol) < 1e-14
    o.stepforward(5)
    prob.stepforward(5)
    assert rel_l2(o.updatevars(), prob.updatevars()) < 5 * TOL_STEP
    prob.close()
