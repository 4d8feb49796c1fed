"""GPU tests (-m gpu) of the fused 3-D engine (engine="fused", ndim = 3: csrc/engine_fused3d.cu) against the CPU oracle
and against the independent cuFFT engine.  calcN! in 3-D: TAD.jl:771-786 (steady), :725-742 (time-varying).
Tolerances: <= 1e-12 relative L2 per step, <= 1e-10 after 1000 steps (north_star)."""
import numpy as np
import pytest

from oracle.ptf_oracle import OracleProblem, rel_l2
from tests.test_gpu_parity import B200Adapter, TOL_STEP, TOL_1000, _pts

pytestmark = pytest.mark.gpu


def _P():
    import ptf_b200
    return ptf_b200


def _fused(kw):
    a = B200Adapter(engine="fused", **kw)
    assert a.p.engine == "fused"
    return a


def _compare(kw, c0, steps, tol=TOL_STEP):
    o = OracleProblem(**kw)
    g = _fused(kw)
    o.set_c(c0)
    g.set_c(c0)
    assert rel_l2(o.sol, g.sol) <= 1e-14, f"set_c mismatch {rel_l2(o.sol, g.sol):.3e}"
    assert rel_l2(o.updatevars(), g.updatevars()) <= 1e-14
    done = 0
    for ns in steps:
        o.stepforward(ns - done)
        g.stepforward(ns - done)
        done = ns
        e_sol, e_c = rel_l2(o.sol, g.sol), rel_l2(o.updatevars(), g.updatevars())
        lim = tol * (1 if ns <= 1 else min(ns, 100))
        assert e_sol <= lim and e_c <= lim, f"after {ns} steps: sol {e_sol:.3e}, c {e_c:.3e} > {lim:.1e}"
    g.assert_native()


def _abc(n, L, scale=1.0):
    x, y, z = _pts(n, L)
    kx, ky, kz = (2 * np.pi / Lv for Lv in L)
    u = scale * (np.sin(kz * z) + 0.6 * np.cos(ky * y))
    v = scale * (0.8 * np.sin(kx * x) + np.cos(kz * z))
    w = scale * (0.6 * np.sin(ky * y) + 0.8 * np.cos(kx * x))
    c0 = np.exp(-(x ** 2 / 0.4 + y ** 2 / 0.3 + z ** 2 / 0.2))
    return [np.ascontiguousarray(a) for a in (u, v, w)], c0


STEPPERS = ["ForwardEuler", "RK4", "ETDRK4", "LSRK54", "AB3", "FilteredRK4", "FilteredETDRK4", "FilteredLSRK54",
            "FilteredAB3"]


@pytest.mark.parametrize("stepper", STEPPERS)
def test_fused3d_all_steppers_64(stepper):
    n, L = (64, 64, 64), (2 * np.pi, 4.0, 3.0)
    vel, c0 = _abc(n, L, 0.5)
    kw = dict(n=n, L=L, kappa=(0.01, 0.02, 0.005), dt=2e-3, stepper=stepper, velocity=vel, steady=True, kappa_h=1e-6,
              n_kappa_h=2)
    _compare(kw, c0, [1, 2, 5])


@pytest.mark.parametrize("n", [(64, 128, 256), (256, 64, 128), (128, 256, 64), (512, 64, 64), (64, 64, 512)])
def test_fused3d_rectangular_sizes(n):
    L = (2 * np.pi, 3.0, 5.0)
    vel, c0 = _abc(n, L, 0.3)
    kw = dict(n=n, L=L, kappa=(0.002, 0.001, 0.003), dt=1e-3, stepper="RK4", velocity=vel, steady=True)
    _compare(kw, c0, [1, 3])


def test_fused3d_128_cubed_one_step_against_oracle():
    # the size of the reference's own 3-D tests (test/test_traceradvectiondiffusion.jl:156)
    n, L = (128, 128, 128), (2 * np.pi,) * 3
    vel, c0 = _abc(n, L)
    kw = dict(n=n, L=L, kappa=(0.01,) * 3, dt=2e-3, stepper="RK4", velocity=vel, steady=True)
    _compare(kw, c0, [1, 2])


def test_fused3d_white_noise_nyquist_semantics():
    n, L = (64, 64, 64), (2 * np.pi,) * 3
    rng = np.random.default_rng(11)
    c0 = rng.standard_normal((64, 64, 64))
    vel = [rng.standard_normal((64, 64, 64)) for _ in range(3)]
    kw = dict(n=n, L=L, kappa=(0.0,) * 3, dt=1e-5, stepper="RK4", velocity=vel, steady=True)
    _compare(kw, c0, [1, 2])


def test_fused3d_non_hermitian_sol_matches_c2r_semantics():
    n, L = (64, 64, 64), (2 * np.pi,) * 3
    rng = np.random.default_rng(12)
    vel = [rng.standard_normal((64, 64, 64)) for _ in range(3)]
    kw = dict(n=n, L=L, kappa=(0.01,) * 3, dt=1e-5, stepper="RK4", velocity=vel, steady=True)
    o = OracleProblem(**kw)
    g = _fused(kw)
    s = rng.standard_normal((64, 64, 33)) + 1j * rng.standard_normal((64, 64, 33))
    o.sol = s.copy()
    g.p.set_sol(s)
    assert rel_l2(s, g.p.sol) == 0.0          # set_sol / get_sol round trip is a pure layout change
    assert rel_l2(o.updatevars(), g.updatevars()) <= 1e-14
    o.stepforward(1)
    g.stepforward(1)
    assert rel_l2(o.sol, g.sol) <= TOL_STEP


@pytest.mark.parametrize("stepper", ["RK4", "ETDRK4", "LSRK54", "FilteredRK4"])
def test_fused3d_dealias_option(stepper):
    n, L = (64, 128, 64), (2 * np.pi,) * 3
    rng = np.random.default_rng(7)
    vel, _ = _abc(n, L, 0.3)
    c0 = rng.standard_normal((n[2], n[1], n[0]))
    kw = dict(n=n, L=L, kappa=(0.01,) * 3, dt=2e-4, stepper=stepper, velocity=vel, steady=True, dealias=True)
    _compare(kw, c0, [1, 2, 4])


def test_fused3d_time_varying_callback():
    # TAD.jl:737: u(x, y, z, clock.t) for every stage of the step
    n, L = (64, 64, 64), (2 * np.pi,) * 3
    x, y, z = _pts(n, L)
    u = lambda x, y, z, t: (np.sin(z) + np.cos(y)) * (1 + 0.5 * np.sin(3 * t)) + 0.2 * t
    v = lambda x, y, z, t: (np.sin(x) + np.cos(z)) * (1 + 0.5 * np.sin(3 * t))
    w = lambda x, y, z, t: (np.sin(y) + np.cos(x)) * (1 - 0.3 * t)
    c0 = np.exp(-((x - 0.5) ** 2 + y ** 2 + z ** 2) / 0.4)
    kw = dict(n=n, L=L, kappa=(0.005,) * 3, dt=0.005, stepper="RK4", velocity=[u, v, w], steady=False)
    _compare(kw, c0, [1, 2, 5])


def test_fused3d_separable_flow_time_dependent():
    # BASELINE configs[3] flow: multi-term separable ABC flow with a time-dependent amplitude, evaluated in registers
    P = _P()
    n, L = (64, 128, 64), (2 * np.pi,) * 3
    one = lambda s: 1.0 + 0 * s
    g = lambda t: 1.0 + 0.5 * np.sin(t)
    A, B, Cc = 1.0, 0.8, 0.6
    flow = P.SeparableFlow(
        terms=[[(one, one, np.sin), (one, np.cos, one)],
               [(np.sin, one, one), (one, one, np.cos)],
               [(one, np.sin, one), (np.cos, one, one)]],
        coeffs=lambda t, a: g(t) * np.array([[A, Cc], [B, A], [Cc, B]][a]), steadyflow=False)
    prob = P.Problem(P.B200(engine="fused"), flow, nx=n[0], ny=n[1], nz=n[2], kappa=0.01, dt=0.004)
    assert prob.engine == "fused"
    u = lambda x, y, z, t: g(t) * (A * np.sin(z) + Cc * np.cos(y))
    v = lambda x, y, z, t: g(t) * (B * np.sin(x) + A * np.cos(z))
    w = lambda x, y, z, t: g(t) * (Cc * np.sin(y) + B * np.cos(x))
    o = OracleProblem(n=n, L=L, kappa=(0.01,) * 3, dt=0.004, stepper="RK4", velocity=[u, v, w], steady=False)
    x, y, z = _pts(n, L)
    c0 = np.exp(-((x - 0.5) ** 2 + y ** 2 + (z + 0.3) ** 2) / 0.3)
    o.set_c(c0)
    prob.set_c(c0)
    o.stepforward(5)
    prob.stepforward(5)
    assert rel_l2(o.updatevars(), prob.updatevars()) <= 5 * TOL_STEP
    d = prob.diagnostics()
    assert abs(d["mean_c"] - o.c.mean()) <= 1e-13 and abs(d["variance_c"] - o.c.var()) <= 1e-13
    prob.close()


def test_fused3d_1000_steps_64():
    n, L = (64, 64, 64), (2 * np.pi,) * 3
    vel, c0 = _abc(n, L, 0.5)
    kw = dict(n=n, L=L, kappa=(0.002,) * 3, dt=0.005, stepper="RK4", velocity=vel, steady=True)
    o = OracleProblem(**kw)
    g = _fused(kw)
    o.set_c(c0)
    g.set_c(c0)
    o.stepforward(1000)
    g.stepforward(1000)
    e = rel_l2(o.updatevars(), g.updatevars())
    assert e <= TOL_1000, f"after 1000 steps: {e:.3e}"


def test_fused3d_interleaved_updatevars_and_steps_are_consistent():
    n, L = (64, 64, 64), (2 * np.pi,) * 3
    vel, c0 = _abc(n, L, 0.5)
    kw = dict(n=n, L=L, kappa=(0.002,) * 3, dt=0.005, stepper="RK4", velocity=vel, steady=True)
    a, b = _fused(kw), _fused(kw)
    a.set_c(c0)
    b.set_c(c0)
    a.stepforward(6)
    for _ in range(6):
        b.stepforward(1)
        b.updatevars()
        _ = b.p.sol
    assert rel_l2(a.updatevars(), b.updatevars()) <= 1e-15


def test_fused3d_agrees_with_cufft_engine_256():
    # two independent implementations at a size the oracle does not finish in seconds: 256^3, 5 RK4 steps;
    # linearity and mean conservation of the step as size-independent properties
    P = _P()
    n = 256
    one = lambda s: 1.0 + 0 * s
    flow = P.SeparableFlow(
        terms=[[(one, one, np.sin), (one, np.cos, one)], [(np.sin, one, one), (one, one, np.cos)],
               [(one, np.sin, one), (np.cos, one, one)]],
        coeffs=lambda t, a: np.array([[1.0, 0.6], [0.8, 1.0], [0.6, 0.8]][a]), steadyflow=True)
    dt = 0.2 * 2.83 / (3 * 1.8 * n / 2)
    pf = P.Problem(P.B200(engine="fused"), flow, nx=n, kappa=1e-3, dt=dt)
    pc = P.Problem(P.B200(engine="cufft"), flow, nx=n, kappa=1e-3, dt=dt)
    assert pf.engine == "fused" and pc.engine == "cufft"
    x = pf.grid.x
    c1 = np.exp(-(x[None, None, :] ** 2 + x[None, :, None] ** 2 + x[:, None, None] ** 2) / (2 * 0.5 ** 2))
    c2 = 0.7 * np.roll(c1, (40, -30, 17), axis=(0, 1, 2))
    outs = []
    for c in (c1, c2, 2.0 * c1 - 3.0 * c2):
        pf.set_c(c)
        m0 = pf.diagnostics()["mean_c"]
        pf.stepforward(5)
        assert abs(pf.diagnostics()["mean_c"] - m0) <= 1e-13 * max(1.0, abs(m0))
        outs.append(pf.updatevars().copy())
    assert rel_l2(2.0 * outs[0] - 3.0 * outs[1], outs[2]) <= 1e-13
    pc.set_c(c1)
    pc.stepforward(5)
    assert rel_l2(pc.updatevars(), outs[0]) <= 5 * TOL_STEP
    pf.close()
    pc.close()


@pytest.mark.parametrize("chunks,zctas", [(2, None), (4, 1), (7, 2)])
def test_fused3d_chunked_pipeline_on_one_gpu(monkeypatch, chunks, zctas):
    # the slab pipeline's chunked launch sequence (kr chunks, comm-stream events) with the exchange a no-op, and the
    # persistent grid-striding z-column kernel (zctas CTAs per SM) the P2P pipeline uses: same results
    monkeypatch.setenv("PTF_F3_CHUNKS", str(chunks))
    if zctas is not None:
        monkeypatch.setenv("PTF_F3_ZCTAS", str(zctas))
    n, L = (64, 128, 64), (2 * np.pi, 4.0, 3.0)
    vel, c0 = _abc(n, L, 0.5)
    for stepper in ("RK4", "FilteredETDRK4", "LSRK54"):
        kw = dict(n=n, L=L, kappa=(0.01, 0.02, 0.005), dt=2e-3, stepper=stepper, velocity=vel, steady=True)
        _compare(kw, c0, [1, 3])
