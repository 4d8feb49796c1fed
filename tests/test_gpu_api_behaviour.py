"""GPU tests (-m gpu) of boundary behaviour: error codes, option flags and handle lifetime through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from oracle.ptf_oracle import OracleProblem, rel_l2
from tests.test_gpu_parity import B200Adapter, TOL_STEP, _pts, _cellular

pytestmark = pytest.mark.gpu


def _P():
    import ptf_b200
    return ptf_b200


def test_error_codes_and_messages():
    P = _P()
    capi = P._capi
    lib = capi.load()
    prob = P.Problem(P.B200(), P.TwoDAdvectingFlow(), nx=64, kappa=0.01, dt=0.01, stepper="ETDRK4")
    h = prob._h
    # wrong velocity extent -> PTF_EINVAL (ValueError), message names the problem
    bad = np.zeros(7)
    rc = lib.ptf_set_velocity(h, 0, capi.as_dp(bad), bad.size)
    assert rc == capi.EINVAL and b"velocity count" in lib.ptf_last_error(h)
    rc = lib.ptf_set_velocity(h, 5, capi.as_dp(np.zeros(64 * 64)), 64 * 64)
    assert rc == capi.EINVAL
    # step_until! with ETDRK4 is refused, as FourierFlows does
    with pytest.raises(capi.PtfError) as ei:
        prob.step_until(0.5)
    assert ei.value.status == capi.EUNSUPPORTED
    # NULL arguments
    assert lib.ptf_set_c(h, None, 0) == capi.EINVAL
    assert lib.ptf_step(None, 1) == capi.EINVAL
    assert lib.ptf_step(h, -1) == capi.EINVAL
    # stepping backwards in time is refused
    prob2 = P.Problem(P.B200(), P.OneDAdvectingFlow(), nx=64, kappa=0.01, dt=0.01)
    prob2.stepforward(3)
    with pytest.raises(ValueError):
        prob2.step_until(0.01)
    # a fused engine request for a grid it cannot serve is an error, not a silent fallback
    with pytest.raises(capi.PtfError) as ei:
        P.Problem(P.B200(engine="fused"), P.TwoDAdvectingFlow(), nx=96)
    assert ei.value.status == capi.EUNSUPPORTED
    # a time-varying problem stepped without its callback
    T = P.tracer_advection_diffusion
    g = T.Grid(nx=64, Lx=2 * np.pi, ny=64, Ly=2 * np.pi, ndim=2)
    p3 = T.TracerProblem(P.B200(), g, T.Params(0.1, 0.1, 0.1, 0.0, 0), 0.01, "RK4", capi.FLOW_CALLBACK)
    with pytest.raises(ValueError, match="callback"):
        p3.stepforward(1)


@pytest.mark.parametrize("engine,n", [("cufft", (64, 48)), ("fused", (256, 256))])
def test_without_cuda_graph_matches(engine, n):
    L = (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.01, stepper="RK4", velocity=vel, steady=True)
    a = B200Adapter(engine=engine, use_graph=True, **kw)
    b = B200Adapter(engine=engine, use_graph=False, **kw)
    a.set_c(c0)
    b.set_c(c0)
    a.stepforward(7)
    b.stepforward(7)
    assert rel_l2(a.updatevars(), b.updatevars()) == 0.0      # same kernels, same order: bitwise equal


@pytest.mark.parametrize("engine,n", [("cufft", (64, 48)), ("fused", (256, 512))])
def test_positive_nyquist_sign_option(engine, n):
    L = (2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(3)
    c0 = rng.standard_normal(n[::-1])               # white noise: the ky-Nyquist row matters
    vel = [rng.standard_normal(n[::-1]), rng.standard_normal(n[::-1])]
    kw = dict(n=n, L=L, kappa=(0.0, 0.0), dt=1e-5, stepper="RK4", velocity=vel, steady=True, nyquist_sign=+1)
    o = OracleProblem(**kw)
    g = B200Adapter(engine=engine, **kw)
    o.set_c(c0)
    g.set_c(c0)
    o.stepforward(2)
    g.stepforward(2)
    assert rel_l2(o.sol, g.sol) <= 2 * TOL_STEP
    # and it really differs from the default (negative) convention
    kw["nyquist_sign"] = -1
    o2 = OracleProblem(**kw)
    o2.set_c(c0)
    o2.stepforward(2)
    assert rel_l2(o.sol, o2.sol) > 1e-9


@pytest.mark.parametrize("engine,n", [("cufft", (64, 64)), ("fused", (256, 256))])
def test_changing_dt_recomputes_etd_coefficients(engine, n):
    L = (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.01, 0.01), dt=0.01, stepper="ETDRK4", velocity=vel, steady=True)
    g = B200Adapter(engine=engine, **kw)
    g.set_c(c0)
    g.stepforward(2)
    g.p.clock.dt = 0.004                      # FourierFlows users may change prob.clock.dt between steps
    g.stepforward(3)
    o = OracleProblem(**kw)
    o.set_c(c0)
    o.stepforward(2)
    o.dt = 0.004
    o._init_stepper()
    o.stepforward(3)
    assert rel_l2(o.updatevars(), g.updatevars()) <= 5 * TOL_STEP
    assert abs(g.p.clock.t - (0.02 + 0.012)) < 1e-15


def test_handles_do_not_leak_device_memory():
    import torch
    P = _P()
    flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.1 + 0 * x, v=lambda x, y: 0 * x)
    free0 = None
    for i in range(6):
        prob = P.Problem(P.B200(), flow, nx=1024, kappa=0.01, dt=1e-4)
        assert prob.device_bytes() > 6 * 513 * 1024 * 16
        prob.stepforward(1)
        prob.close()
        free, total = torch.cuda.mem_get_info()
        if i == 1:
            free0 = free
    assert free0 is not None and abs(free - free0) < 64 * 1024 * 1024
