"""GPU parity of the device-resident MultiLayerQG flow solver and of the MQG-coupled tracer against the
CPU oracle.  Tolerances: the north star's ≤1e-12 relative L2 per step; the flow is chaotic, so multi-step parity is
checked over short horizons (≤1e-10 after 20 steps), not after 1000 steps."""
import numpy as np
import pytest

from oracle.mqg_oracle import MQGOracle
from oracle.ptf_oracle import OracleProblem, irfft, make_filter, rel_l2, rfft

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12
TOL_20 = 1e-10

EXAMPLE = dict(beta=5.0, f0=1.0, H=[0.2, 0.8], b=[-1.0, -1.2], U=[1.0, 0.0], mu=5e-2)   # examples/turbulent…:38-52


def P():
    import ptf_b200
    return ptf_b200


def _q0(o, amp=1e-2, seed=1234):
    g = o.grid
    q0 = amp * np.random.default_rng(seed).standard_normal((o.nlayers,) + g.pshape)
    return irfft(g, make_filter(g) * rfft(g, q0))          # examples/…:62-65


def _pair(nl=2, n=64, stepper="FilteredRK4", dt=2.5e-3, aliased_fraction=0.0, amp=1e-2, **kw):
    phys = dict(EXAMPLE) if nl == 2 else {}
    phys.update(kw)
    o = MQGOracle(nl, nx=n, dt=dt, stepper=stepper, aliased_fraction=aliased_fraction, **phys)
    g = P().MultiLayerQG.Problem(nl, P().B200(), nx=n, dt=dt, stepper=stepper, aliased_fraction=aliased_fraction, **phys)
    q0 = _q0(o, amp)
    o.set_q(q0)
    g.set_q(q0)
    return o, g


def _check_vars(o, g, tol):
    o.updatevars()
    g.updatevars()
    for name, a, b in (("q", o.q, g.vars.q), ("psi", o.psi, g.vars.psi), ("u", o.u, g.vars.u), ("v", o.v, g.vars.v)):
        assert rel_l2(a, b) < tol, f"{name}: {rel_l2(a, b):.3e}"
    assert rel_l2(o.sol, g.sol) < tol


@pytest.mark.parametrize("stepper", ["RK4", "FilteredRK4", "ETDRK4", "FilteredETDRK4", "LSRK54", "AB3", "ForwardEuler",
                                     "FilteredAB3"])
def test_flow_solver_matches_oracle(stepper):
    o, g = _pair(stepper=stepper, amp=0.5)
    assert rel_l2(o.params.Qy, g.params.Qy) < 1e-15 and np.abs(g.params.Qx).max() == 0.0
    _check_vars(o, g, 1e-13)
    o.stepforward(1)
    g.stepforward(1)
    assert rel_l2(o.sol, g.sol) < TOL_STEP
    o.stepforward(19)
    g.stepforward(19)
    assert g.clock.step == 20 and abs(g.clock.t - o.t) < 1e-15
    _check_vars(o, g, TOL_20)
    own, lib = g.launch_count()
    assert own > 0 and lib > 0
    g.close()


@pytest.mark.parametrize("aliased_fraction", [0.0, 1.0 / 3.0, 0.5])
def test_dealias_fractions(aliased_fraction):
    o, g = _pair(stepper="RK4", aliased_fraction=aliased_fraction, amp=0.5, n=96)
    o.stepforward(5)
    g.stepforward(5)
    _check_vars(o, g, 1e-11)
    g.close()


@pytest.mark.parametrize("nl", [1, 3, 4])
def test_other_layer_counts(nl):
    H = {1: [1.0], 3: [0.2, 0.3, 0.5], 4: [0.1, 0.2, 0.3, 0.4]}[nl]
    b = {1: None, 3: [-1.0, -1.2, -1.5], 4: [-1.0, -1.1, -1.3, -1.6]}[nl]
    U = {1: [0.3], 3: [1.0, 0.5, 0.0], 4: [1.0, 0.6, 0.3, 0.0]}[nl]
    kw = dict(beta=4.0, f0=1.2, H=H, U=U, mu=0.1, nu=1e-5, nnu=2)
    if b is not None:
        kw["b"] = b
    o, g = _pair(nl=nl, n=64, stepper="FilteredRK4", aliased_fraction=1 / 3, amp=0.5, **kw)
    assert rel_l2(o.params.Qy, g.params.Qy) < 1e-15
    o.stepforward(10)
    g.stepforward(10)
    _check_vars(o, g, 1e-11)
    g.close()


def test_sheared_background_flow_and_topography():
    n = 64
    y = -np.pi + 2 * np.pi / n * np.arange(n)
    U = np.stack([1.0 + 0.3 * np.cos(y), 0.2 * np.sin(2 * y)])
    X, Y = np.meshgrid(y, y)
    eta = 0.5 * np.cos(2 * X) * np.sin(3 * Y) + 0.1 * np.sin(X)
    kw = dict(EXAMPLE)
    kw.update(U=U, eta=eta, topographic_pv_gradient=(0.05, -0.1))
    o, g = _pair(n=n, stepper="RK4", aliased_fraction=1 / 3, amp=0.5, **kw)
    assert rel_l2(o.params.Qy, g.params.Qy) < 1e-13 and rel_l2(o.params.Qx, g.params.Qx) < 1e-13
    o.stepforward(10)
    g.stepforward(10)
    _check_vars(o, g, 1e-11)
    g.close()


def test_rossby_wave_known_answer_on_device():
    beta, U0, kx, ky, n = 3.0, 0.7, 2.0, 3.0, 64
    g = P().MultiLayerQG.Problem(2, P().B200(), nx=n, beta=beta, U=[U0, U0], H=[0.2, 0.8], b=[-1.0, -1.2], dt=1e-3,
                                 stepper="RK4", aliased_fraction=0)
    x = -np.pi + 2 * np.pi / n * np.arange(n)
    X, Y = np.meshgrid(x, x)
    psi0 = 1e-3 * np.cos(kx * X + ky * Y)
    g.set_psi(np.stack([psi0, psi0]))
    g.stepforward(200)
    g.updatevars()
    om = U0 * kx - beta * kx / (kx * kx + ky * ky)
    exact = 1e-3 * np.cos(kx * X + ky * Y - om * g.clock.t)
    assert np.abs(g.vars.psi[0] - exact).max() / 1e-3 < 1e-12
    assert np.abs(g.vars.psi[1] - exact).max() / 1e-3 < 1e-12
    g.close()


def test_step_until_and_graph_off():
    o, _g = _pair(stepper="RK4", amp=0.5)
    _g.close()
    g = P().MultiLayerQG.Problem(2, P().B200(use_graph=False), nx=64, dt=2.5e-3, stepper="RK4", aliased_fraction=0.0,
                                 **EXAMPLE)
    g.set_q(_q0(o, 0.5))
    o.step_until(0.0312)
    g.step_until(0.0312)
    assert g.clock.t == 0.0312 and g.clock.step == 13 and g.clock.dt == 2.5e-3
    _check_vars(o, g, TOL_20)
    g.close()


def _tracer_c0(n, L=2 * np.pi):
    x = -L / 2 + L / n * np.arange(n)
    X, Y = np.meshgrid(x, x)
    return 10 * np.exp(-(X ** 2 + Y ** 2) / (2 * 0.15 ** 2))       # examples/…:95-98


@pytest.mark.parametrize("engine,n", [("cufft", 64), ("auto", 256)])
def test_coupled_tracer_matches_oracle(engine, n):
    """Problem(MQGprob; κ, stepper, tracer_release_time) + the example's loop (examples/…:149-151)."""
    kappa, dt, release = 0.002, 2.5e-3, 0.02
    o, g = _pair(n=n, stepper="FilteredRK4", dt=dt, amp=0.5)
    ad = P().Problem(g, kappa=kappa, stepper="FilteredRK4", tracer_release_time=release, dev=P().B200(engine=engine))
    assert ad.engine == ("cufft" if engine == "cufft" else "fused")
    # oracle side of TAD.jl:236-250
    o.step_until(release)
    o.updatevars()
    assert abs(g.clock.t - release) < 1e-15 and g.clock.step == o.step
    ot = OracleProblem(n=(n, n), L=(2 * np.pi,) * 2, kappa=(kappa, kappa), dt=dt, stepper="FilteredRK4",
                       velocity="layered", steady=True, nbatch=2)
    c0 = _tracer_c0(n)
    ot.set_c(c0)
    ad.set_c(c0)
    nsteps = 12
    for i in range(nsteps):
        ot.set_layered_velocity(o.u, o.v, o.params.U)
        ot.stepforward(1)
        o.stepforward(1)
        o.updatevars()
        if i < 6:      # three separate calls, as the example makes them
            ad.stepforward(1)
            g.stepforward(1)
            g.updatevars()
        elif i == 6:   # the same loop inside the library, no host round trips
            ms = P().MultiLayerQG.step_coupled(ad, nsteps - 6)
            assert ms > 0
    assert ad.clock.step == nsteps and g.clock.step == o.step
    c = ad.updatevars()
    assert rel_l2(ot.updatevars(), c) < TOL_20
    assert rel_l2(o.sol, g.sol) < TOL_20
    assert rel_l2(o.u, g.vars.u) < TOL_20
    ad.close()
    g.close()


def test_coupled_loop_per_step_parity_over_a_long_trajectory():
    """the flow is chaotic, so parity of the coupled loop is a PER-STEP statement.  Over 300 iterations of
    the example's loop (examples/turbulent_advection-diffusion.jl:149-151) the device is re-synchronised to the oracle's
    flow state every 10 iterations and must reproduce the next 10 coupled iterations (flow and tracer) to <= 1e-12 per
    step; the TRACER, which is linear in c and never re-synchronised, must still agree after all 300."""
    n, kappa, dt = 128, 0.002, 2.5e-3
    o, g = _pair(n=n, stepper="FilteredRK4", dt=dt, amp=0.5)
    ad = P().Problem(g, kappa=kappa, stepper="FilteredRK4")
    o.updatevars()
    ot = OracleProblem(n=(n, n), L=(2 * np.pi,) * 2, kappa=(kappa, kappa), dt=dt, stepper="FilteredRK4",
                       velocity="layered", steady=True, nbatch=2)
    c0 = _tracer_c0(n)
    ot.set_c(c0)
    ad.set_c(c0)
    worst_flow = 0.0
    for block in range(30):
        g.set_sol(o.sol)                 # same flow state on both sides at the start of the block
        g.updatevars()
        for i in range(10):
            ot.set_layered_velocity(o.u, o.v, o.params.U)
            ot.stepforward(1)
            o.stepforward(1)
            o.updatevars()
        P().MultiLayerQG.step_coupled(ad, 10)
        e = rel_l2(o.sol, g.sol)
        worst_flow = max(worst_flow, e)
        assert e < 10 * TOL_STEP, f"block {block}: flow after 10 coupled iterations {e:.3e}"
    e_c = rel_l2(ot.updatevars(), ad.updatevars())
    assert e_c < TOL_20, f"tracer after 300 coupled iterations: {e_c:.3e} (worst flow block {worst_flow:.2e})"
    ad.close()
    g.close()


def test_two_tracers_coupled_to_one_flow_are_ordered_across_streams():
    """The flow adopts the stream of the tracer that coupled last; stepping the OTHER tracer through step_coupled runs
    its kernels on a different stream than the flow's and must be ordered with events (data race)."""
    n, kappa, dt = 128, 0.002, 2.5e-3
    o, g = _pair(n=n, stepper="FilteredRK4", dt=dt, amp=0.5)
    ad1 = P().Problem(g, kappa=kappa, stepper="FilteredRK4")
    ad2 = P().Problem(g, kappa=2 * kappa, stepper="FilteredRK4")     # couples second: the flow now runs on ad2's stream
    o.updatevars()
    ots = [OracleProblem(n=(n, n), L=(2 * np.pi,) * 2, kappa=(k, k), dt=dt, stepper="FilteredRK4", velocity="layered",
                         steady=True, nbatch=2) for k in (kappa, 2 * kappa)]
    c0 = _tracer_c0(n)
    for ot, ad in zip(ots, (ad1, ad2)):
        ot.set_c(c0)
        ad.set_c(c0)
    nsteps = 10
    for i in range(nsteps):
        for ot in ots:
            ot.set_layered_velocity(o.u, o.v, o.params.U)
            ot.stepforward(1)
        o.stepforward(1)
        o.updatevars()
    for i in range(nsteps // 2):                     # the users' loop with two tracers: tracer 2 alone, then
        ad2.stepforward(1)                           # tracer 1 + flow + updatevars! inside the library (cross-stream)
        P().MultiLayerQG.step_coupled(ad1, 1)
    for i in range(nsteps // 2):
        ad1.stepforward(1)
        P().MultiLayerQG.step_coupled(ad2, 1)        # same-stream path
    assert ad1.clock.step == nsteps and ad2.clock.step == nsteps and g.clock.step == o.step
    assert rel_l2(ots[0].updatevars(), ad1.updatevars()) < TOL_20
    assert rel_l2(ots[1].updatevars(), ad2.updatevars()) < TOL_20
    assert rel_l2(o.sol, g.sol) < TOL_20
    ad1.close()
    ad2.close()
    g.close()


def test_reference_kat_diffusion_multilayerqg_with_device_flow():
    """test/test_traceradvectiondiffusion.jl:302-355 with the flow solver on the device (zero flow, release after 50
    flow steps, both layers must match the analytic diffusion solution to nx·ny·nsteps·1e-12)."""
    n, dt, tfinal, kappa = 128, 0.005, 0.1, 0.01
    mq = P().MultiLayerQG.Problem(2, P().B200(), nx=n, Lx=2 * np.pi, f0=1, H=[0.2, 0.8], b=[-1.0, -1.2], U=[0.0, 0.0],
                                  mu=0, beta=0, dt=dt, stepper="FilteredRK4", aliased_fraction=0)
    mq.set_q(np.zeros((2, n, n)))
    nsteps = round(tfinal / dt)
    ad = P().Problem(mq, kappa=kappa, stepper="RK4", tracer_release_time=dt * 50)
    assert mq.clock.step == 50
    x = -np.pi + 2 * np.pi / n * np.arange(n)
    X, Y = np.meshgrid(x, x)
    amp, sigma = 0.1, 0.1
    st = np.sqrt(2 * kappa * nsteps * dt + sigma ** 2)
    ad.set_c(amp * np.exp(-(X ** 2 + Y ** 2) / (2 * sigma ** 2)))
    ad.stepforward(nsteps)
    c = ad.updatevars()
    cfinal = amp * (sigma ** 2 / st ** 2) * np.exp(-(X ** 2 + Y ** 2) / (2 * st ** 2))
    rtol = n * n * nsteps * 1e-12
    assert rel_l2(cfinal, c[0]) < rtol and rel_l2(cfinal, c[1]) < rtol
    ad.close()
    mq.close()


def test_error_behaviour_and_lifetimes():
    M = P().MultiLayerQG
    with pytest.raises(ValueError):
        M.Problem(5, P().B200(), nx=32)                                     # more layers than the build supports
    with pytest.raises(ValueError):
        M.Problem(2, P().B200(), nx=33)                                     # odd grid
    with pytest.raises(ValueError):
        M.Problem(2, P().B200(), nx=32, b=[-1.0, -1.0])                     # zero reduced gravity
    g = M.Problem(2, P().B200(), nx=64, dt=2.5e-3, stepper="FilteredRK4", aliased_fraction=0.0, **EXAMPLE)
    with pytest.raises(ValueError):
        P().Problem(g, kappa=0.01, tracer_release_time=-1.0)               # ArgumentError, TAD.jl:234
    with pytest.raises(ValueError):
        g.set_q(np.zeros((2, 32, 32)))
    q0 = 0.5 * np.random.default_rng(3).standard_normal((2, 64, 64))
    g.set_q(q0)
    ad = P().Problem(g, kappa=0.002, stepper="FilteredRK4")
    ad.set_c(_tracer_c0(64))
    P().MultiLayerQG.step_coupled(ad, 3)
    # destroying the tracer first leaves a usable flow; destroying the flow first detaches the tracer loudly
    ad.close()
    g.stepforward(2)
    assert g.clock.step == 5
    ad2 = P().Problem(g, kappa=0.002, stepper="FilteredRK4")
    ad2.set_c(_tracer_c0(64))
    g.close()
    with pytest.raises(ValueError):
        ad2.stepforward(1)
    ad2.close()


def test_spectral_background_term_equals_the_transformed_product(monkeypatch):
    """With no topography and uniform U the library applies rfft(v·Qy) = Qy·i kr ψ̂ spectrally (one transform less per
    stage); PTF_MQG_NO_SPECTRAL_BG forces the generic product path — both must agree to rounding."""
    q0 = None
    sols = []
    for force_generic in (False, True):
        if force_generic:
            monkeypatch.setenv("PTF_MQG_NO_SPECTRAL_BG", "1")
        g = P().MultiLayerQG.Problem(2, P().B200(), nx=64, dt=2.5e-3, stepper="FilteredRK4", aliased_fraction=0.0, **EXAMPLE)
        if q0 is None:
            q0 = 0.5 * np.random.default_rng(7).standard_normal((2, 64, 64))
        g.set_q(q0)
        own0, lib0 = g.launch_count()
        g.stepforward(10)
        own1, lib1 = g.launch_count()
        sols.append((g.sol, lib1 - lib0))
        g.close()
    assert rel_l2(sols[0][0], sols[1][0]) < 1e-13
    assert sols[0][1] == sols[1][1]          # same number of cuFFT calls, smaller batches


def test_non_square_grid_and_domain():
    kw = dict(EXAMPLE)
    o = MQGOracle(2, nx=96, ny=64, Lx=2 * np.pi, Ly=4.0, dt=2.5e-3, stepper="FilteredRK4", aliased_fraction=1 / 3, **kw)
    g = P().MultiLayerQG.Problem(2, P().B200(), nx=96, ny=64, Lx=2 * np.pi, Ly=4.0, dt=2.5e-3, stepper="FilteredRK4",
                                 aliased_fraction=1 / 3, **kw)
    q0 = 0.5 * np.random.default_rng(11).standard_normal((2, 64, 96))
    q0 = irfft(o.grid, make_filter(o.grid) * rfft(o.grid, q0))
    o.set_q(q0)
    g.set_q(q0)
    o.stepforward(10)
    g.stepforward(10)
    _check_vars(o, g, 1e-11)
    # coupled tracer on the same non-square grid (cuFFT tracer engine)
    ad = P().Problem(g, kappa=0.002, stepper="FilteredRK4")
    ot = OracleProblem(n=(96, 64), L=(2 * np.pi, 4.0), kappa=(0.002, 0.002), dt=2.5e-3, stepper="FilteredRK4",
                       velocity="layered", steady=True, nbatch=2)
    x = -np.pi + 2 * np.pi / 96 * np.arange(96)
    y = -2.0 + 4.0 / 64 * np.arange(64)
    c0 = 10 * np.exp(-(x[None, :] ** 2 + y[:, None] ** 2) / (2 * 0.15 ** 2))
    ad.set_c(c0)
    ot.set_c(c0)
    P().MultiLayerQG.step_coupled(ad, 5)
    for _ in range(5):
        ot.set_layered_velocity(o.u, o.v, o.params.U)
        ot.stepforward(1)
        o.stepforward(1)
        o.updatevars()
    assert rel_l2(ot.updatevars(), ad.updatevars()) < TOL_20
    assert rel_l2(o.sol, g.sol) < TOL_20
    ad.close()
    g.close()


def test_reference_example_script_runs_on_the_device():
    """examples/turbulent_advection_diffusion.py = the reference's example line for line (short horizon here)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ex_tad", os.path.join(root, "examples", "turbulent_advection_diffusion.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ad, mq, frames = mod.main(n=64, nsteps=100, tracer_release_time=0.25, save_frequency=50, quiet=True)
    assert ad.clock.step == 101 and mq.clock.step == 100 + 101 and len(frames) == 3
    c = ad.updatevars()
    assert np.isfinite(c).all() and c.max() < 10.0 + 1e-9 and c.min() > -0.5      # diffusing, advected Gaussian
    # tracer mass is conserved by advection-diffusion of a periodic field
    assert abs(c[0].mean() - frames[0].mean()) < 1e-12 and abs(c[1].mean() - frames[0].mean()) < 1e-12
    ad.close()
    mq.close()
