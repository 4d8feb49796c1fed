"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  (a) the reference's own 14 known-answer tests with the reference's tolerances, and
  (b) the CPU oracle on identical inputs: <= 1e-12 relative L2 per step, <= 1e-10 after 1000 steps
      (BASELINE.json north_star tolerance).
Every test asserts that libptf_b200.so launched kernels (no silent fallback)."""
import numpy as np
import pytest

from oracle.ptf_oracle import OracleProblem, rel_l2
from tests.kat_cases import REFERENCE_KATS

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12     # relative L2 per step (north_star)
TOL_1000 = 1e-10     # after 1000 steps (north_star)
ENGINES = ["cufft", "auto"]


def _P():
    import ptf_b200
    return ptf_b200


class B200Adapter:
    """Gives the B200 problem the keyword vocabulary of tests/kat_cases.py / OracleProblem."""

    def __init__(self, n, L, kappa, dt, stepper="RK4", velocity=None, steady=True, nbatch=1, kappa_h=0.0,
                 n_kappa_h=0, dealias=False, engine="auto", nyquist_sign=-1, use_graph=True):
        P = _P()
        T = P.tracer_advection_diffusion
        nd = len(n)
        kappa = tuple(np.atleast_1d(kappa).astype(float))
        kappa = kappa + (kappa[0],) * (3 - len(kappa))
        pad = lambda v, fill: tuple(v) + (fill,) * (3 - len(v))
        nn, LL = pad(n, 1), pad(L, 1.0)
        grid = T.Grid(nx=nn[0], Lx=LL[0], ny=nn[1], Ly=LL[1], nz=nn[2], Lz=LL[2], ndim=nd)
        params = T.Params(kappa=kappa[0], eta=kappa[1], iota=kappa[2], kappa_h=kappa_h, n_kappa_h=n_kappa_h)
        dev = P.B200(engine=engine, use_graph=use_graph)
        layered = isinstance(velocity, str) and velocity == "layered"
        if layered:
            kind = P._capi.FLOW_LAYERED
        elif velocity is None or steady:
            kind = P._capi.FLOW_STEADY
        else:
            kind = P._capi.FLOW_CALLBACK
        per_batch = layered or (steady and velocity is not None and not layered and
                                np.asarray(velocity[0]).ndim == nd + 1)
        self.p = T.TracerProblem(dev, grid, params, dt, stepper, kind, nbatch=nbatch, velocity_per_batch=per_batch,
                                 dealias=dealias, nyquist_sign=nyquist_sign)
        if kind == P._capi.FLOW_STEADY:
            arrays = velocity if velocity is not None else [np.zeros(grid.pshape) for _ in range(nd)]
            self.p._set_velocity_arrays(arrays)
        elif kind == P._capi.FLOW_CALLBACK:
            self.p._install_velocity_callback(list(velocity))

    def set_layered_velocity(self, u, v, U=None):
        self.p.set_layered_velocity(u, v, U)

    def set_c(self, c):
        self.p.set_c(c)

    def stepforward(self, n=1):
        self.p.stepforward(n)

    def updatevars(self):
        return self.p.updatevars()

    @property
    def sol(self):
        self.p.updatevars()
        return self.p.sol

    def assert_native(self):
        own, lib = self.p.launch_count()
        assert own > 0, "no kernels of libptf_b200.so were launched"


def make_b200(engine):
    def make(**kw):
        return B200Adapter(engine=engine, **kw)
    return make


# ------------------------------------------------------------------------------------------
# (a) the reference's own known-answer tests, on the GPU
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", list(REFERENCE_KATS))
def test_reference_kat_on_b200(name, engine):
    fn, kw = REFERENCE_KATS[name]
    err, rtol = fn(make_b200(engine), stepper="RK4", **kw)
    assert err <= rtol, f"{name}[{engine}]: rel-L2 {err:.3e} > reference rtol {rtol:.3e}"


@pytest.mark.parametrize("stepper", ["ETDRK4", "LSRK54", "AB3", "FilteredRK4"])
@pytest.mark.parametrize("name", [n for n in REFERENCE_KATS if n not in ("constvel3D", "timedependentvel3D")])
def test_reference_kat_other_steppers_on_b200(name, stepper):
    # the reference's harness takes the stepper as a parameter (test/runtests.jl:26): the same analytic answers pin the
    # steppers its own run leaves out; tolerances as in tests/test_kat_all_steppers.py (ETDRK4 / LSRK54: the reference's)
    from tests.test_kat_all_steppers import kat_tolerance
    fn, kw = REFERENCE_KATS[name]
    err, _ = fn(make_b200("auto"), stepper=stepper, **kw)
    tol = kat_tolerance(name, stepper)
    assert err <= tol, f"{name}[{stepper}]: rel-L2 {err:.3e} > {tol:.3e}"


@pytest.mark.parametrize("stepper", ["RK4", "ETDRK4", "LSRK54", "AB3", "ForwardEuler"])
def test_stepper_order_by_dt_halving_on_b200(stepper):
    from tests.test_kat_all_steppers import ORDERS, observed_order
    order, e1, e2 = observed_order(make_b200("auto"), stepper)
    assert abs(order - ORDERS[stepper]) < 0.35, f"{stepper}: observed order {order:.2f} (errors {e1:.2e}, {e2:.2e})"


# ------------------------------------------------------------------------------------------
# (b) step-by-step parity with the oracle
# ------------------------------------------------------------------------------------------
def _pts(n, L):
    nd = len(n)
    coords = [-Lv / 2 + (Lv / nv) * np.arange(nv) for nv, Lv in zip(n, L)]
    out = []
    for a in range(nd):
        shp = [1] * nd
        shp[nd - 1 - a] = n[a]
        out.append(np.broadcast_to(coords[a].reshape(shp), tuple(reversed(n))))
    return out


def _compare(kw, c0, nsteps_list, engine, tol=TOL_STEP, layered_vel=None):
    o = OracleProblem(**kw)
    g = B200Adapter(engine=engine, **kw)
    if layered_vel is not None:
        o.set_layered_velocity(*layered_vel)
        g.set_layered_velocity(*layered_vel)
    o.set_c(c0)
    g.set_c(c0)
    assert rel_l2(o.sol, g.sol) <= 1e-14, "set_c mismatch"
    done = 0
    worst = 0.0
    for ns in nsteps_list:
        o.stepforward(ns - done)
        g.stepforward(ns - done)
        done = ns
        e_sol = rel_l2(o.sol, g.sol)
        e_c = rel_l2(o.updatevars(), g.updatevars())
        worst = max(worst, e_sol, e_c)
        lim = tol * (1 if ns <= 1 else min(ns, 100))
        assert e_sol <= lim and e_c <= lim, f"after {ns} steps: sol {e_sol:.3e} c {e_c:.3e} > {lim:.1e}"
    g.assert_native()
    return worst


STEPPERS = ["ForwardEuler", "RK4", "ETDRK4", "LSRK54", "AB3", "FilteredRK4", "FilteredETDRK4", "FilteredLSRK54",
            "FilteredAB3", "FilteredForwardEuler"]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("stepper", STEPPERS)
def test_config1_onedim_gaussian(stepper, engine):
    # examples/onedim_gaussiandiffusion.jl:27-60
    n, L = (128,), (2 * np.pi,)
    (x,) = _pts(n, L)
    kw = dict(n=n, L=L, kappa=(0.01,), dt=0.02, stepper=stepper, velocity=[np.full(n, 0.05)], steady=True)
    _compare(kw, np.exp(-x ** 2 / (2 * 0.15 ** 2)), [1, 2, 5, 50], engine)


def _cellular(n, L, psi0=0.2):
    x, y = _pts(n, L)
    u = psi0 * np.cos(x) * np.sin(y)
    v = -psi0 * np.sin(x) * np.cos(y)
    c0 = 0.5 * np.exp(-((x - 0.2 * L[0]) ** 2 + y ** 2) / (2 * 0.15 ** 2))
    return [np.ascontiguousarray(u), np.ascontiguousarray(v)], c0


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("stepper", ["RK4", "ETDRK4", "FilteredRK4", "LSRK54", "AB3"])
def test_config2_cellular_flow_128(stepper, engine):
    # examples/cellularflow.jl:30-81 at the example's own parameters
    n, L = (128, 128), (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.02, stepper=stepper, velocity=vel, steady=True)
    _compare(kw, c0, [1, 2, 10], engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_config2_cellular_flow_1000_steps(engine):
    n, L = (128, 128), (2 * np.pi, 2 * np.pi)
    vel, c0 = _cellular(n, L)
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.02, stepper="RK4", velocity=vel, steady=True)
    o = OracleProblem(**kw)
    g = B200Adapter(engine=engine, **kw)
    o.set_c(c0)
    g.set_c(c0)
    o.stepforward(1000)
    g.stepforward(1000)
    e = rel_l2(o.updatevars(), g.updatevars())
    assert e <= TOL_1000, f"after 1000 steps: {e:.3e}"
    g.assert_native()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n", [(256, 256), (512, 256), (96, 80), (1024, 1024)])
def test_2d_sizes_anisotropic(n, engine):
    L = (2 * np.pi, 3.0)
    x, y = _pts(n, L)
    u = 0.3 * np.cos(x) * np.sin(2 * np.pi * y / L[1]) + 0.1
    v = -0.2 * np.sin(2 * x) * np.cos(2 * np.pi * y / L[1])
    c0 = np.exp(-((x - 0.3) ** 2 / 0.5 + (y + 0.2) ** 2 / 0.3))
    kw = dict(n=n, L=L, kappa=(0.01, 0.003), dt=1e-3 if max(n) < 1024 else 1e-5, stepper="RK4",
              velocity=[np.ascontiguousarray(u), np.ascontiguousarray(v)], steady=True, kappa_h=1e-7, n_kappa_h=2)
    _compare(kw, c0, [1, 3], engine)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("stepper", ["RK4", "ETDRK4", "FilteredRK4"])
def test_3d_anisotropic(stepper, engine):
    n, L = (64, 32, 16), (2 * np.pi, 4.0, 3.0)
    x, y, z = _pts(n, L)
    ky, kz = 2 * np.pi / L[1], 2 * np.pi / L[2]
    u = np.sin(kz * z) + np.cos(ky * y)
    v = np.sin(x) + np.cos(kz * z)
    w = np.sin(ky * y) + np.cos(x)
    c0 = np.exp(-(x ** 2 / 0.4 + y ** 2 / 0.3 + z ** 2 / 0.2))
    kw = dict(n=n, L=L, kappa=(0.01, 0.02, 0.005), dt=2e-3, stepper=stepper,
              velocity=[np.ascontiguousarray(a) for a in (u, v, w)], steady=True, kappa_h=1e-6, n_kappa_h=2)
    _compare(kw, c0, [1, 4], engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_3d_cubic_64(engine):
    n, L = (64, 64, 64), (2 * np.pi,) * 3
    x, y, z = _pts(n, L)
    u, v, w = np.sin(z) + np.cos(y), np.sin(x) + np.cos(z), np.sin(y) + np.cos(x)
    c0 = np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * 0.3 ** 2))
    kw = dict(n=n, L=L, kappa=(0.01,) * 3, dt=5e-3, stepper="RK4",
              velocity=[np.ascontiguousarray(a) for a in (u, v, w)], steady=True)
    _compare(kw, c0, [1, 3], engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_time_varying_flow_uses_clock_t(engine):
    # TAD.jl:718: velocities at clock.t for every stage
    n, L = (64, 64), (2 * np.pi, 2 * np.pi)
    x, y = _pts(n, L)
    u = lambda x, y, t: 0.3 * np.cos(x) * np.sin(y) * (1 + 0.5 * np.sin(3 * t)) + 0.2 * t
    v = lambda x, y, t: -0.3 * np.sin(x) * np.cos(y) * (1 + 0.5 * np.sin(3 * t))
    c0 = np.exp(-((x - 0.5) ** 2 + y ** 2) / 0.2)
    kw = dict(n=n, L=L, kappa=(0.005, 0.005), dt=0.01, stepper="RK4", velocity=[u, v], steady=False)
    _compare(kw, c0, [1, 2, 7], engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_dealias_option(engine):
    n, L = (64, 48), (2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(7)
    x, y = _pts(n, L)
    c0 = rng.standard_normal((n[1], n[0]))        # white noise: every mode populated, incl. Nyquist rows
    vel = [np.ascontiguousarray(0.3 * np.cos(x) * np.sin(y)), np.ascontiguousarray(-0.3 * np.sin(x) * np.cos(y))]
    kw = dict(n=n, L=L, kappa=(0.01, 0.01), dt=1e-3, stepper="RK4", velocity=vel, steady=True, dealias=True)
    _compare(kw, c0, [1, 3], engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_white_noise_nyquist_semantics(engine):
    # the y-Nyquist derivative is kept (negative wavenumber), the x-Nyquist one is dropped by c2r
    n, L = (32, 32), (2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(11)
    c0 = rng.standard_normal((32, 32))
    vel = [rng.standard_normal((32, 32)), rng.standard_normal((32, 32))]
    kw = dict(n=n, L=L, kappa=(0.0, 0.0), dt=1e-4, stepper="RK4", velocity=vel, steady=True)
    _compare(kw, c0, [1, 2], engine)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("per_member", [False, True])
def test_ensemble_batch(per_member, engine):
    # config 5 in miniature: independent members; shared or per-member velocity
    n, L, B = (64, 64), (2 * np.pi, 2 * np.pi), 5
    x, y = _pts(n, L)
    cx = np.linspace(-1, 1, B).reshape(B, 1, 1)
    c0 = np.exp(-((x - cx) ** 2 + (y + 0.5 * cx) ** 2) / (2 * 0.3 ** 2))
    u = 0.2 * np.cos(x) * np.sin(y)
    v = -0.2 * np.sin(x) * np.cos(y)
    if per_member:
        s = (1 + 0.1 * np.arange(B)).reshape(B, 1, 1)
        vel = [np.ascontiguousarray(s * u), np.ascontiguousarray(s * v)]
    else:
        vel = [np.ascontiguousarray(u), np.ascontiguousarray(v)]
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=0.01, stepper="RK4", velocity=vel, steady=True, nbatch=B)
    _compare(kw, c0, [1, 3], engine)


@pytest.mark.parametrize("engine", ENGINES)
def test_config3_layered_flow(engine):
    # examples/turbulent_advection-diffusion.jl in miniature: 2 layers, synthetic band-limited snapshot + U = [1, 0]
    n, L, B = (128, 128), (2 * np.pi, 2 * np.pi), 2
    rng = np.random.default_rng(1234)
    x, y = _pts(n, L)
    psi_h = np.zeros((B, n[1], n[0] // 2 + 1), dtype=complex)
    psi_h[:, :8, :8] = rng.standard_normal((B, 8, 8)) + 1j * rng.standard_normal((B, 8, 8))
    psi = np.fft.irfft2(psi_h, s=(n[1], n[0]))
    kx = np.arange(n[0] // 2 + 1)
    ky = np.where(np.arange(n[1]) < n[1] // 2, np.arange(n[1]), np.arange(n[1]) - n[1])
    u = np.fft.irfft2(-1j * ky[None, :, None] * psi_h, s=(n[1], n[0]))
    v = np.fft.irfft2(1j * kx[None, None, :] * psi_h, s=(n[1], n[0]))
    rms = np.sqrt(np.mean(u ** 2 + v ** 2))
    u, v = np.ascontiguousarray(u / rms), np.ascontiguousarray(v / rms)
    U = np.array([1.0, 0.0])
    c0 = 10 * np.exp(-(x ** 2 + y ** 2) / (2 * 0.15 ** 2))
    kw = dict(n=n, L=L, kappa=(0.002, 0.002), dt=2.5e-3, stepper="FilteredRK4", velocity="layered", steady=True,
              nbatch=B)
    _compare(kw, c0, [1, 4], engine, layered_vel=(u, v, U))


def test_separable_flow_matches_array_flow():
    P = _P()
    n, L = (64, 64), (2 * np.pi, 2 * np.pi)
    x, y = _pts(n, L)
    g = lambda t: 1 + 0.5 * np.sin(t)
    flow = P.SeparableFlow(
        terms=[[(np.cos, np.sin)], [(np.sin, np.cos)]],
        coeffs=lambda t, a: np.array([0.2 * g(t)]) if a == 0 else np.array([-0.2 * g(t)]),
        steadyflow=False)
    prob = P.Problem(P.B200(engine="cufft"), flow, nx=64, kappa=0.01, dt=0.01)
    u = lambda x, y, t: 0.2 * g(t) * np.cos(x) * np.sin(y)
    v = lambda x, y, t: -0.2 * g(t) * np.sin(x) * np.cos(y)
    o = OracleProblem(n=n, L=L, kappa=(0.01, 0.01), dt=0.01, stepper="RK4", velocity=[u, v], steady=False)
    c0 = np.exp(-((x - 0.5) ** 2 + y ** 2) / 0.2)
    o.set_c(c0)
    prob.set_c(c0)
    o.stepforward(5)
    prob.stepforward(5)
    assert rel_l2(o.updatevars(), prob.updatevars()) <= 5 * TOL_STEP
    assert abs(prob.clock.t - 0.05) < 1e-15 and prob.clock.step == 5


def test_public_api_mirror_and_roundtrips():
    P = _P()
    TAD = P.TracerAdvectionDiffusion
    flow = TAD.TwoDAdvectingFlow(u=lambda x, y: 0.2 + 0 * x, v=lambda x, y: 0.1 + 0 * x, steadyflow=True)
    prob = TAD.Problem(P.B200(), flow, nx=64, Lx=2 * np.pi, kappa=0.01, dt=0.01, stepper="RK4")
    x, y = P.gridpoints(prob.grid)
    c0 = np.exp(-(x ** 2 + y ** 2) / 0.1)
    TAD.set_c(prob, c0)
    assert rel_l2(prob.vars.c, c0) < 1e-14                     # set_c! -> updatevars! roundtrip
    sol0 = prob.sol.copy()
    assert rel_l2(sol0, np.fft.rfft2(c0)) < 1e-14               # prob.sol is the unnormalised rfft (TAD.jl:847)
    TAD.stepforward(prob, 3)
    TAD.updatevars(prob)
    assert prob.clock.step == 3 and abs(prob.clock.t - 0.03) < 1e-15
    prob.set_sol(sol0)
    TAD.updatevars(prob)
    assert rel_l2(prob.vars.c, c0) < 1e-14
    d = prob.diagnostics()
    assert abs(d["mean_c"] - c0.mean()) < 1e-12 and abs(d["variance_c"] - c0.var()) < 1e-12
    TAD.step_until(prob, 0.0555)
    assert abs(prob.clock.t - 0.0555) < 1e-15
    with pytest.raises(ValueError):
        TAD.set_c(prob, np.zeros((3, 3)))
    own, lib = prob.launch_count()
    assert own > 0


def test_step_until_matches_manual_partial_step():
    P = _P()
    n, L = (64,), (2 * np.pi,)
    (x,) = _pts(n, L)
    kw = dict(n=n, L=L, kappa=(0.01,), dt=0.01, stepper="RK4", velocity=[np.full(n, 0.3)], steady=True)
    g = B200Adapter(**kw)
    o = OracleProblem(**kw)
    c0 = np.exp(-x ** 2 / 0.1)
    g.set_c(c0)
    o.set_c(c0)
    g.p.step_until(0.035)
    o.stepforward(3)
    o.dt = 0.035 - 0.03
    o.stepforward(1)
    assert rel_l2(o.updatevars(), g.updatevars()) <= 4 * TOL_STEP
