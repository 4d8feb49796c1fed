"""GPU tests (-m gpu) of the hand-written FFT core and the fused 2-D engine (engine="fused"), against NumPy FFTs and
the CPU oracle.  Tolerances: FFT core 1e-14 relative L2 (fp64 rounding); engine <= 1e-12 per step (north_star)."""
import ctypes as C

import numpy as np
import pytest

from oracle.ptf_oracle import OracleProblem, rel_l2
from tests.test_gpu_parity import B200Adapter, TOL_STEP, TOL_1000, _pts, _cellular

pytestmark = pytest.mark.gpu


def _P():
    import ptf_b200
    return ptf_b200


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("direction", [-1, 1])
def test_fft_core_matches_numpy(n, direction):
    P = _P()
    lib = P._capi.load()
    rng = np.random.default_rng(n + direction)
    count = 37   # not a multiple of the transforms-per-CTA: exercises the inactive-group path
    x = rng.standard_normal((count, n)) + 1j * rng.standard_normal((count, n))
    x[0] = 0
    x[0, 1] = 1.0          # a pure mode: output must be the twiddle sequence itself
    y = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    P._capi.check(lib.ptf_selftest_fft(n, direction, count, x.ctypes.data_as(dp), y.ctypes.data_as(dp)))
    ref = np.fft.fft(x, axis=1) if direction < 0 else np.fft.ifft(x, axis=1) * n
    for i in range(count):
        assert rel_l2(ref[i], y[i]) < 1e-14, f"transform {i}: {rel_l2(ref[i], y[i]):.2e}"


def _fused(kw):
    a = B200Adapter(engine="fused", **kw)
    assert a.p.engine == "fused"
    return a


def _compare(kw, c0, steps, layered_vel=None, tol=TOL_STEP):
    o = OracleProblem(**kw)
    g = _fused(kw)
    if layered_vel is not None:
        o.set_layered_velocity(*layered_vel)
        g.set_layered_velocity(*layered_vel)
    o.set_c(c0)
    g.set_c(c0)
    assert rel_l2(o.sol, g.sol) <= 1e-14
    done = 0
    for ns in steps:
        o.stepforward(ns - done)
        g.stepforward(ns - done)
        done = ns
        e_sol, e_c = rel_l2(o.sol, g.sol), rel_l2(o.updatevars(), g.updatevars())
        lim = tol * (1 if ns <= 1 else min(ns, 100))
        assert e_sol <= lim and e_c <= lim, f"after {ns} steps: sol {e_sol:.3e}, c {e_c:.3e} > {lim:.1e}"
    g.assert_native()


STEPPERS = ["ForwardEuler", "RK4", "ETDRK4", "LSRK54", "AB3", "FilteredRK4", "FilteredETDRK4", "FilteredLSRK54",
            "FilteredAB3"]


@pytest.mark.parametrize("stepper", STEPPERS)
def test_fused_all_steppers_256(stepper):
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.01, stepper=stepper, velocity=vel, steady=True)
    _compare(kw, c0, [1, 2, 6])


@pytest.mark.parametrize("n", [(256, 512), (512, 256), (1024, 256), (256, 2048), (2048, 512), (4096, 256),
                               (256, 4096), (1024, 1024)])
def test_fused_rectangular_sizes(n):
    L = (2 * np.pi, 3.0)
    x, y = _pts(n, L)
    u = 0.3 * np.cos(x) * np.sin(2 * np.pi * y / L[1]) + 0.1
    v = -0.2 * np.sin(2 * x) * np.cos(2 * np.pi * y / L[1])
    c0 = np.exp(-((x - 0.3) ** 2 / 0.5 + (y + 0.2) ** 2 / 0.3))
    kx2, ky2 = (np.pi * n[0] / L[0]) ** 2, (np.pi * n[1] / L[1]) ** 2
    maxL = kx2 * 0.01 + ky2 * 0.003 + 1e-9 * (kx2 + ky2) ** 2        # RK4 is stable for max|L| dt < 2.785
    kw = dict(n=n, L=L, kappa=(0.01, 0.003), dt=min(1e-3, 1.0 / maxL), stepper="RK4",
              velocity=[np.ascontiguousarray(u), np.ascontiguousarray(v)], steady=True, kappa_h=1e-9, n_kappa_h=2)
    _compare(kw, c0, [1, 3])


def test_fused_white_noise_nyquist_semantics():
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(5)
    c0 = rng.standard_normal((256, 256))
    vel = [rng.standard_normal((256, 256)), rng.standard_normal((256, 256))]
    kw = dict(n=n, L=L, kappa=(0.0, 0.0), dt=1e-5, stepper="RK4", velocity=vel, steady=True)
    _compare(kw, c0, [1, 2])


def test_fused_non_hermitian_sol_matches_c2r_semantics():
    # set_sol with arbitrary (non-Hermitian-consistent) coefficients: c2r must ignore Im of the kr = 0 / Nyquist bins
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(9)
    vel = [rng.standard_normal((256, 256)), rng.standard_normal((256, 256))]
    kw = dict(n=n, L=L, kappa=(0.01, 0.01), dt=1e-5, stepper="RK4", velocity=vel, steady=True)
    o = OracleProblem(**kw)
    g = _fused(kw)
    s = rng.standard_normal((256, 129)) + 1j * rng.standard_normal((256, 129))
    o.sol = s.copy()
    g.p.set_sol(s)
    o.stepforward(1)
    g.stepforward(1)
    assert rel_l2(o.sol, g.sol) <= TOL_STEP


@pytest.mark.parametrize("per_member", [False, True])
def test_fused_ensemble_batch(per_member):
    n, L, B = (256, 256), (2 * np.pi, 2 * np.pi), 3
    x, y = _pts(n, L)
    cx = np.linspace(-1, 1, B).reshape(B, 1, 1)
    c0 = np.exp(-((x - cx) ** 2 + (y + 0.5 * cx) ** 2) / (2 * 0.3 ** 2))
    u, v = 0.2 * np.cos(x) * np.sin(y), -0.2 * np.sin(x) * np.cos(y)
    if per_member:
        s = (1 + 0.1 * np.arange(B)).reshape(B, 1, 1)
        vel = [np.ascontiguousarray(s * u), np.ascontiguousarray(s * v)]
    else:
        vel = [np.ascontiguousarray(u), np.ascontiguousarray(v)]
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.01, stepper="RK4", velocity=vel, steady=True, nbatch=B)
    _compare(kw, c0, [1, 3])


def test_fused_config3_layered_512():
    # BASELINE configs[2]: 2 layers x 512^2, FilteredRK4, kappa = 0.002, dt = 2.5e-3, U = [1, 0]
    n, L, B = (512, 512), (2 * np.pi, 2 * np.pi), 2
    rng = np.random.default_rng(1234)
    x, y = _pts(n, L)
    psi_h = np.zeros((B, n[1], n[0] // 2 + 1), dtype=complex)
    psi_h[:, :12, :12] = rng.standard_normal((B, 12, 12)) + 1j * rng.standard_normal((B, 12, 12))
    kx = np.arange(n[0] // 2 + 1)
    ky = np.where(np.arange(n[1]) < n[1] // 2, np.arange(n[1]), np.arange(n[1]) - n[1])
    u = np.fft.irfft2(-1j * ky[None, :, None] * psi_h, s=(n[1], n[0]))
    v = np.fft.irfft2(1j * kx[None, None, :] * psi_h, s=(n[1], n[0]))
    rms = np.sqrt(np.mean(u ** 2 + v ** 2))
    u, v = np.ascontiguousarray(u / rms), np.ascontiguousarray(v / rms)
    c0 = 10 * np.exp(-(x ** 2 + y ** 2) / (2 * 0.15 ** 2))
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=2.5e-3, stepper="FilteredRK4", velocity="layered", steady=True,
              nbatch=B)
    _compare(kw, c0, [1, 4], layered_vel=(u, v, np.array([1.0, 0.0])))


def test_fused_config3_layered_512_1000_steps():
    # the north star's long-run gate (<= 1e-10 after 1000 steps) on BASELINE configs[2] at full size: 2 layers x 512^2,
    # FilteredRK4, kappa = 0.002, dt = 2.5e-3, frozen flow snapshot + U = [1, 0]   (~70 s of CPU oracle time)
    n, L, B = (512, 512), (2 * np.pi, 2 * np.pi), 2
    rng = np.random.default_rng(1234)
    x, y = _pts(n, L)
    psi_h = np.zeros((B, n[1], n[0] // 2 + 1), dtype=complex)
    psi_h[:, :12, :12] = rng.standard_normal((B, 12, 12)) + 1j * rng.standard_normal((B, 12, 12))
    kx = np.arange(n[0] // 2 + 1)
    ky = np.where(np.arange(n[1]) < n[1] // 2, np.arange(n[1]), np.arange(n[1]) - n[1])
    u = np.fft.irfft2(-1j * ky[None, :, None] * psi_h, s=(n[1], n[0]))
    v = np.fft.irfft2(1j * kx[None, None, :] * psi_h, s=(n[1], n[0]))
    rms = np.sqrt(np.mean(u ** 2 + v ** 2))
    u, v = np.ascontiguousarray(u / rms), np.ascontiguousarray(v / rms)
    c0 = 10 * np.exp(-(x ** 2 + y ** 2) / (2 * 0.15 ** 2))
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=2.5e-3, stepper="FilteredRK4", velocity="layered", steady=True,
              nbatch=B)
    o = OracleProblem(**kw)
    g = _fused(kw)
    for p in (o, g):
        p.set_layered_velocity(u, v, np.array([1.0, 0.0]))
        p.set_c(c0)
    o.stepforward(1000)
    g.stepforward(1000)
    e = rel_l2(o.updatevars(), g.updatevars())
    assert e <= TOL_1000, f"configs[2] after 1000 steps: {e:.3e}"


def test_fused_time_varying_callback():
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    x, y = _pts(n, L)
    u = lambda x, y, t: 0.3 * np.cos(x) * np.sin(y) * (1 + 0.5 * np.sin(3 * t)) + 0.2 * t
    v = lambda x, y, t: -0.3 * np.sin(x) * np.cos(y) * (1 + 0.5 * np.sin(3 * t))
    c0 = np.exp(-((x - 0.5) ** 2 + y ** 2) / 0.2)
    kw = dict(n=n, L=L, kappa=(0.005, 0.005), dt=0.005, stepper="RK4", velocity=[u, v], steady=False)
    _compare(kw, c0, [1, 2, 5])


def test_fused_separable_flow():
    P = _P()
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    x, y = _pts(n, L)
    g = lambda t: 1 + 0.5 * np.sin(t)
    flow = P.SeparableFlow(terms=[[(np.cos, np.sin)], [(np.sin, np.cos)]],
                           coeffs=lambda t, a: np.array([0.2 * g(t)]) if a == 0 else np.array([-0.2 * g(t)]),
                           steadyflow=False)
    prob = P.Problem(P.B200(engine="fused"), flow, nx=256, kappa=0.01, dt=0.005)
    assert prob.engine == "fused"
    u = lambda x, y, t: 0.2 * g(t) * np.cos(x) * np.sin(y)
    v = lambda x, y, t: -0.2 * g(t) * np.sin(x) * np.cos(y)
    o = OracleProblem(n=n, L=L, kappa=(0.01, 0.01), dt=0.005, stepper="RK4", velocity=[u, v], steady=False)
    c0 = np.exp(-((x - 0.5) ** 2 + y ** 2) / 0.2)
    o.set_c(c0)
    prob.set_c(c0)
    o.stepforward(5)
    prob.stepforward(5)
    assert rel_l2(o.updatevars(), prob.updatevars()) <= 5 * TOL_STEP


def test_fused_1000_steps_256():
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.01, stepper="RK4", velocity=vel, steady=True)
    o = OracleProblem(**kw)
    g = _fused(kw)
    o.set_c(c0)
    g.set_c(c0)
    o.stepforward(1000)
    g.stepforward(1000)
    e = rel_l2(o.updatevars(), g.updatevars())
    assert e <= TOL_1000, f"after 1000 steps: {e:.3e}"


def test_fused_interleaved_updatevars_and_steps_are_consistent():
    # updatevars!/get_sol clobber the A,B scratch: stepping afterwards must re-run the prologue and give the same
    # trajectory as uninterrupted stepping
    n, L = (256, 256), (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.01, stepper="RK4", velocity=vel, steady=True)
    a, b = _fused(kw), _fused(kw)
    a.set_c(c0)
    b.set_c(c0)
    a.stepforward(6)
    for _ in range(6):
        b.stepforward(1)
        b.updatevars()
    assert rel_l2(a.updatevars(), b.updatevars()) <= 1e-15


def test_fused_4096_one_step_against_oracle():
    # BASELINE configs[1] at full size: one RK4 step, 4096^2, against the oracle (takes ~10 s of CPU)
    n, L = (4096, 4096), (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    dt = 0.5 * 2.785 / (0.1 * 2 * (4096 / 2) ** 2)
    kw = dict(n=n, L=L, kappa=(0.1, 0.1), dt=dt, stepper="RK4", velocity=vel, steady=True)
    o = OracleProblem(**kw)
    g = _fused(kw)
    o.set_c(c0)
    g.set_c(c0)
    o.stepforward(2)
    g.stepforward(2)
    e = rel_l2(o.updatevars(), g.updatevars())
    assert e <= 2 * TOL_STEP, f"4096^2 after 2 steps: {e:.3e}"


# ------------------------------------------------------------------------------------------
# size-independent properties at BASELINE's full size (4096^2), where the oracle is too slow for long runs
# ------------------------------------------------------------------------------------------
def _full_size_problem(stepper="RK4", engine="auto"):
    P = _P()
    nx = 4096
    dt = 0.5 * 2.785 / (0.1 * 2 * (nx / 2) ** 2)
    flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
    prob = P.Problem(P.B200(engine=engine), flow, nx=nx, kappa=0.1, dt=dt, stepper=stepper)
    x = prob.grid.x
    c0 = 0.5 * np.exp(-((x[None, :] - 0.4 * np.pi) ** 2 + x[:, None] ** 2) / (2 * 0.15 ** 2))
    return prob, c0


def test_full_size_linearity_and_mean_conservation():
    # the step is linear in c: step(a*c1 + b*c2) == a*step(c1) + b*step(c2); an incompressible flow conserves the mean
    prob, c1 = _full_size_problem()
    assert prob.engine == "fused"
    c2 = np.roll(c1, (700, -900), axis=(0, 1)) * 0.7
    outs = []
    for c in (c1, c2, 2.0 * c1 - 3.0 * c2):
        prob.set_c(c)
        m0 = prob.diagnostics()["mean_c"]
        prob.stepforward(5)
        d = prob.diagnostics()
        assert abs(d["mean_c"] - m0) <= 1e-13 * max(1.0, abs(m0)), "mean not conserved"
        outs.append(prob.updatevars().copy())
    assert rel_l2(2.0 * outs[0] - 3.0 * outs[1], outs[2]) <= 1e-13


def test_full_size_engines_agree_and_variance_decays():
    # fused and cuFFT engines are independent implementations: they must agree at 4096^2 over 20 steps, and the
    # tracer variance must decay monotonically (advection by an incompressible flow + diffusion)
    pf, c0 = _full_size_problem(engine="fused")
    pc, _ = _full_size_problem(engine="cufft")
    pf.set_c(c0)
    pc.set_c(c0)
    last = pf.diagnostics()["variance_c"]
    for _ in range(4):
        pf.stepforward(5)
        pc.stepforward(5)
        v = pf.diagnostics()["variance_c"]
        assert v < last
        last = v
    assert rel_l2(pc.updatevars(), pf.updatevars()) <= 20 * TOL_STEP
    assert abs(pc.diagnostics()["variance_c"] - last) <= 1e-12 * last


# ------------------------------------------------------------------------------------------
# fused 1-D engine: whole steps (or many steps) in one launch by one CTA
# ------------------------------------------------------------------------------------------
STEPPERS_1D = ["ForwardEuler", "RK4", "ETDRK4", "LSRK54", "AB3", "FilteredRK4", "FilteredETDRK4", "FilteredAB3"]


@pytest.mark.parametrize("stepper", STEPPERS_1D)
@pytest.mark.parametrize("nx", [16, 128, 2048])
def test_fused_1d_all_steppers(stepper, nx):
    n, L = (nx,), (2 * np.pi,)
    (x,) = _pts(n, L)
    u = 0.05 + 0.02 * np.sin(x)
    kw = dict(n=n, L=L, kappa=(0.01,), dt=0.02 if nx <= 128 else 2e-5, stepper=stepper, velocity=[np.ascontiguousarray(u)],
              steady=True, kappa_h=1e-9 if nx <= 128 else 0.0, n_kappa_h=2)
    _compare(kw, np.exp(-x ** 2 / (2 * 0.15 ** 2)), [1, 2, 5, 40])


def test_fused_1d_time_varying_and_batch_and_dealias():
    n, L = (128,), (2 * np.pi,)
    (x,) = _pts(n, L)
    u = lambda x, t: 0.05 * t + 0.01 * np.cos(x) * (1 + t)
    kw = dict(n=n, L=L, kappa=(0.0,), dt=0.002, stepper="RK4", velocity=[u], steady=False)
    _compare(kw, 0.1 * np.exp(-x ** 2 / (2 * 0.2 ** 2)), [1, 3, 10])
    B = 3
    cx = np.linspace(-1, 1, B).reshape(B, 1)
    kw = dict(n=n, L=L, kappa=(0.01,), dt=0.01, stepper="RK4", velocity=[np.full(n, 0.3)], steady=True, nbatch=B,
              dealias=True)
    rng = np.random.default_rng(2)
    _compare(kw, rng.standard_normal((B, 128)) + np.exp(-(x - cx) ** 2), [1, 4])


def test_fused_1d_is_one_launch_per_call():
    P = _P()
    prob = P.Problem(P.B200(engine="fused"), P.OneDAdvectingFlow(u=lambda x: 0.05 + 0 * x), nx=128, kappa=0.01, dt=0.02)
    x = P.gridpoints(prob.grid)
    prob.set_c(np.exp(-x ** 2 / (2 * 0.15 ** 2)))
    own0, lib0 = prob.launch_count()
    prob.stepforward(500)
    own1, lib1 = prob.launch_count()
    assert own1 - own0 == 1 and lib1 == lib0            # 500 RK4 steps = 2000 stages = one kernel launch, no cuFFT
    assert prob.clock.step == 500 and abs(prob.clock.t - 10.0) < 1e-12


@pytest.mark.parametrize("stepper", ["RK4", "ETDRK4", "LSRK54", "AB3", "FilteredRK4"])
def test_fused_dealias_option(stepper):
    # opt-in dealias!(sol) at the top of every calcN (FourierFlows' box mask), fused into the column kernel
    n, L = (256, 512), (2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(7)
    x, y = _pts(n, L)
    c0 = rng.standard_normal((n[1], n[0]))        # white noise: every mode populated, incl. the masked ones
    vel = [np.ascontiguousarray(0.3 * np.cos(x) * np.sin(y)), np.ascontiguousarray(-0.3 * np.sin(x) * np.cos(y))]
    kw = dict(n=n, L=L, kappa=(0.01, 0.01), dt=2e-4, stepper=stepper, velocity=vel, steady=True, dealias=True)
    _compare(kw, c0, [1, 2, 5])
