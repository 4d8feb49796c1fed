"""Golden vectors (tests/golden/oracle_golden.npz, written by tests/golden/make_golden.py from the CPU oracle):
CPU — the oracle still reproduces them; GPU — the CUDA path, through the C ABI, reproduces them."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases as C                      # noqa: E402
import make_golden as MG               # noqa: E402
from oracle.ptf_oracle import Grid, irfft, make_filter, rel_l2, rfft   # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))


@pytest.mark.parametrize("name", list(C.TRACER) + list(C.MQG))
def test_oracle_reproduces_golden(name):
    got = MG.run_tracer(name) if name in C.TRACER else MG.run_mqg(name)
    for k, v in got.items():
        assert rel_l2(GOLD[f"{name}/{k}"], v) < 1e-13, (name, k)    # other pocketfft builds may differ in the last bits


def P():
    import ptf_b200
    return ptf_b200


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(C.TRACER))
@pytest.mark.parametrize("engine", ["cufft", "auto"])
def test_cuda_tracer_reproduces_golden(name, engine):
    n, L, stepper, nsteps, dt, steady, kw = C.TRACER[name]
    nd = len(n)
    fs = C.velocity_functions(L)
    flow = [P().OneDAdvectingFlow, P().TwoDAdvectingFlow, P().ThreeDAdvectingFlow][nd - 1](*fs, steadyflow=steady)
    gk = {k: v for k, v in zip(("nx", "ny", "nz"), n)}
    gk.update({k: v for k, v in zip(("Lx", "Ly", "Lz"), L)})
    kap = dict(zip(("kappa", "eta", "iota"), C.KAPPA[:nd]))
    prob = P().Problem(P().B200(engine=engine), flow, dt=dt, stepper=stepper, **gk, **kap, **kw)
    pts = P().gridpoints(prob.grid)
    pts = pts if isinstance(pts, tuple) else (pts,)
    prob.set_c(C.initial_c(pts))
    prob.stepforward(nsteps)
    assert rel_l2(GOLD[f"{name}/c"], prob.updatevars()) < nsteps * 1e-12
    assert rel_l2(GOLD[f"{name}/sol"], prob.sol) < nsteps * 1e-12
    prob.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(C.MQG))
def test_cuda_mqg_reproduces_golden(name):
    nl, n, stepper, nsteps, kw = C.MQG[name]
    g = Grid((n, n), (2 * np.pi, 2 * np.pi))
    mq = P().MultiLayerQG.Problem(nl, P().B200(), nx=n, dt=C.MQG_DT, stepper=stepper, **kw)
    mq.set_q(C.mqg_q0(nl, n, make_filter(g), irfft, rfft, g))
    mq.stepforward(nsteps)
    mq.updatevars()
    for k, got in (("sol", mq.sol), ("u", mq.vars.u), ("v", mq.vars.v), ("psi", mq.vars.psi)):
        assert rel_l2(GOLD[f"{name}/{k}"], got) < 1e-10, (name, k)
    mq.close()
