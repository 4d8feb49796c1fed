"""The reference's 14 known-answer tests (test/test_traceradvectiondiffusion.jl, parameters from
test/runtests.jl:26-54) expressed once, backend-agnostic.

Each case is a function ``case(make_problem) -> (err, rtol)`` where ``make_problem(**kw)`` builds either the
CPU oracle (tests/test_oracle_reference_kat.py) or the B200 problem (tests/test_gpu_parity.py, -m gpu) with
the same keyword vocabulary:

    make_problem(n=(..), L=(..), kappa=(..), dt=.., stepper=.., velocity=<arrays|callables|None>, steady=bool,
                 nbatch=.., kappa_h=.., n_kappa_h=..)
    -> object with .grid-like ``gridpoints()``, ``set_c(c)``, ``stepforward(n)``, ``updatevars() -> c``

``err`` is Julia's isapprox metric (relative L2), ``rtol`` the reference's tolerance.
"""
import numpy as np

TWO_PI = 2 * np.pi


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    den = max(np.linalg.norm(a), np.linalg.norm(b))
    return float(np.linalg.norm(a - b) / den) if den > 0 else 0.0


def _pts(n, L):
    """gridpoints in physical layout (x fastest): returns (X[,Y[,Z]])"""
    nd = len(n)
    coords = [-Lv / 2 + (Lv / nv) * np.arange(nv) for nv, Lv in zip(n, L)]
    pshape = tuple(reversed(n))
    out = []
    for a in range(nd):
        shp = [1] * nd
        shp[nd - 1 - a] = n[a]
        out.append(np.broadcast_to(coords[a].reshape(shp), pshape))
    return out


def constvel1D(make, stepper="RK4", dt=1e-2, nsteps=40):
    # test/test_traceradvectiondiffusion.jl:8-33
    n, L = (128,), (TWO_PI,)
    uvel = 0.05
    (x,) = _pts(n, L)
    prob = make(n=n, L=L, kappa=(0.0,), dt=dt, stepper=stepper, velocity=[np.full(n[::-1], uvel)], steady=True)
    sigma, amp = 0.1, 0.1
    c0f = lambda x: amp * np.exp(-x ** 2 / (2 * sigma ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x))
    prob.stepforward(nsteps)
    return rel_l2(c0f(x - uvel * tfinal), prob.updatevars()), n[0] * nsteps * 1e-12


def timedependentvel1D(make, stepper="RK4", dt=0.002, tfinal=0.1, uvel=0.05):
    # :42-70 — pins the frozen clock.t semantics (u evaluated at t_n, offset by dt/2 to hit the midpoint)
    n, L = (128,), (TWO_PI,)
    nsteps = round(tfinal / dt)
    (x,) = _pts(n, L)
    u = lambda x, t: uvel * t + uvel * dt / 2 + 0 * x
    prob = make(n=n, L=L, kappa=(0.0,), dt=dt, stepper=stepper, velocity=[u], steady=False)
    sigma = 0.2
    c0f = lambda x: 0.1 * np.exp(-x ** 2 / (2 * sigma ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x))
    prob.stepforward(nsteps)
    return rel_l2(c0f(x - 0.5 * uvel * tfinal ** 2), prob.updatevars()), n[0] * nsteps * 1e-12


def constvel2D(make, stepper="RK4", dt=1e-2, nsteps=40):
    # :79-106
    n, L = (128, 128), (TWO_PI, TWO_PI)
    uvel, vvel = 0.2, 0.1
    x, y = _pts(n, L)
    ps = n[::-1]
    prob = make(n=n, L=L, kappa=(0.0, 0.0), dt=dt, stepper=stepper,
                velocity=[np.full(ps, uvel), np.full(ps, vvel)], steady=True)
    sigma, amp = 0.1, 0.1
    c0f = lambda x, y: amp * np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x, y))
    prob.stepforward(nsteps)
    return rel_l2(c0f(x - uvel * tfinal, y - vvel * tfinal), prob.updatevars()), n[0] * n[1] * nsteps * 1e-12


def timedependentvel2D(make, stepper="RK4", dt=0.002, tfinal=0.1, uvel=0.5, av=0.5):
    # :116-145
    n, L = (128, 128), (TWO_PI, TWO_PI)
    nsteps = round(tfinal / dt)
    x, y = _pts(n, L)
    u = lambda x, y, t: uvel + 0 * x + 0 * y
    v = lambda x, y, t: av * t + av * dt / 2 + 0 * x + 0 * y
    prob = make(n=n, L=L, kappa=(0.0, 0.0), dt=dt, stepper=stepper, velocity=[u, v], steady=False)
    sigma = 0.2
    c0f = lambda x, y: 0.1 * np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x, y))
    prob.stepforward(nsteps)
    return (rel_l2(c0f(x - uvel * tfinal, y - 0.5 * av * tfinal ** 2), prob.updatevars()),
            n[0] * n[1] * nsteps * 1e-12)


def constvel3D(make, stepper="RK4", dt=1e-2, nsteps=40, nx=128):
    # :155-182 (nx=128 in the reference)
    n, L = (nx,) * 3, (TWO_PI,) * 3
    uvel, vvel, wvel = 0.2, 0.1, 0.05
    x, y, z = _pts(n, L)
    ps = n[::-1]
    prob = make(n=n, L=L, kappa=(0.0,) * 3, dt=dt, stepper=stepper,
                velocity=[np.full(ps, uvel), np.full(ps, vvel), np.full(ps, wvel)], steady=True)
    sigma, amp = 0.1, 0.1
    c0f = lambda x, y, z: amp * np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * sigma ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x, y, z))
    prob.stepforward(nsteps)
    return (rel_l2(c0f(x - uvel * tfinal, y - vvel * tfinal, z - wvel * tfinal), prob.updatevars()),
            n[0] * n[1] * n[2] * nsteps * 1e-12)


def timedependentvel3D(make, stepper="RK4", dt=0.002, tfinal=0.1, uvel=0.5, av=0.5, wvel=0.5, nx=128):
    # :192-224
    n, L = (nx,) * 3, (TWO_PI,) * 3
    nsteps = round(tfinal / dt)
    x, y, z = _pts(n, L)
    u = lambda x, y, z, t: uvel + 0 * x + 0 * y + 0 * z
    v = lambda x, y, z, t: av * t + av * dt / 2 + 0 * x + 0 * y + 0 * z
    w = lambda x, y, z, t: wvel + 0 * x + 0 * y + 0 * z
    prob = make(n=n, L=L, kappa=(0.0,) * 3, dt=dt, stepper=stepper, velocity=[u, v, w], steady=False)
    sigma, amp = 0.2, 0.1
    c0f = lambda x, y, z: amp * np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * sigma ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x, y, z))
    prob.stepforward(nsteps)
    return (rel_l2(c0f(x - uvel * tfinal, y - 0.5 * av * tfinal ** 2, z - wvel * tfinal), prob.updatevars()),
            n[0] * n[1] * n[2] * nsteps * 1e-12)


def _noflow_callables(nd):
    return [(lambda *a: 0.0 * a[0]) for _ in range(nd)]


def diffusion1D(make, stepper="RK4", dt=0.005, tfinal=0.1, steadyflow=True):
    # :232-264  (note sigma(t) = sqrt(2 kappa t + sigma0), *not* sigma0^2 — as the reference writes it, :251)
    n, L = (128,), (TWO_PI,)
    kappa = 0.01
    nsteps = round(tfinal / dt)
    (x,) = _pts(n, L)
    prob = make(n=n, L=L, kappa=(kappa,), dt=dt, stepper=stepper,
                velocity=None if steadyflow else _noflow_callables(1), steady=steadyflow)
    amp, s0 = 0.1, 0.1
    sig = lambda t: np.sqrt(2 * kappa * t + s0)
    c0f = lambda x, t: (amp / sig(t)) * np.exp(-x ** 2 / (2 * sig(t) ** 2))
    tfinal = nsteps * dt
    prob.set_c(c0f(x, 0))
    prob.stepforward(nsteps)
    return rel_l2(c0f(x, tfinal), prob.updatevars()), n[0] * nsteps * 1e-12


def diffusion2D(make, stepper="RK4", dt=0.005, tfinal=0.1, steadyflow=True):
    # :272-301
    n, L = (128, 128), (TWO_PI, TWO_PI)
    kappa = 0.01
    nsteps = round(tfinal / dt)
    x, y = _pts(n, L)
    prob = make(n=n, L=L, kappa=(kappa, kappa), dt=dt, stepper=stepper,
                velocity=None if steadyflow else _noflow_callables(2), steady=steadyflow)
    amp, sigma = 0.1, 0.1
    tfinal = nsteps * dt
    st = np.sqrt(2 * kappa * tfinal + sigma ** 2)
    prob.set_c(amp * np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2)))
    prob.stepforward(nsteps)
    cfinal = amp * (sigma ** 2 / st ** 2) * np.exp(-(x ** 2 + y ** 2) / (2 * st ** 2))
    return rel_l2(cfinal, prob.updatevars()), n[0] * n[1] * nsteps * 1e-12


def diffusion3D(make, stepper="RK4", dt=0.005, tfinal=0.1, steadyflow=True):
    # :363-392 (nx = 64, tolerance 1e-10*N*nsteps)
    n, L = (64,) * 3, (TWO_PI,) * 3
    kappa = 0.01
    nsteps = round(tfinal / dt)
    x, y, z = _pts(n, L)
    prob = make(n=n, L=L, kappa=(kappa,) * 3, dt=dt, stepper=stepper,
                velocity=None if steadyflow else _noflow_callables(3), steady=steadyflow)
    amp, sigma = 0.1, 0.1
    tfinal = nsteps * dt
    st = np.sqrt(2 * kappa * tfinal + sigma ** 2)
    prob.set_c(amp * np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * sigma ** 2)))
    prob.stepforward(nsteps)
    cfinal = amp * (sigma / st) ** 3 * np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * st ** 2))
    return rel_l2(cfinal, prob.updatevars()), n[0] * n[1] * n[2] * nsteps * 1e-10


def diffusion_multilayerqg(make, stepper="RK4", dt=0.005, tfinal=0.1):
    # :303-355 — 2 layers, MQG flow identically zero (q0 = 0, U = 0), tracer released after 50 flow steps.
    # Tracer-side restatement: layered problem with u = v = 0, U = 0; both layers must match the analytic solution.
    n, L = (128, 128), (TWO_PI, TWO_PI)
    nlayers = 2
    kappa = 0.01
    nsteps = round(tfinal / dt)
    x, y = _pts(n, L)
    prob = make(n=n, L=L, kappa=(kappa, kappa), dt=dt, stepper=stepper, velocity="layered", steady=True,
                nbatch=nlayers)
    z = np.zeros((nlayers,) + n[::-1])
    prob.set_layered_velocity(z, z, np.zeros(nlayers))
    amp, sigma = 0.1, 0.1
    tfinal = nsteps * dt
    st = np.sqrt(2 * kappa * tfinal + sigma ** 2)
    prob.set_c(amp * np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2)))
    prob.stepforward(nsteps)
    c = prob.updatevars()
    cfinal = amp * (sigma ** 2 / st ** 2) * np.exp(-(x ** 2 + y ** 2) / (2 * st ** 2))
    return max(rel_l2(cfinal, c[0]), rel_l2(cfinal, c[1])), n[0] * n[1] * nsteps * 1e-12


def hyperdiffusion(make, stepper="RK4", dt=0.005, tfinal=0.1):
    # :400-439 — kappa = eta = 0, kappa_h = 0.01, n_kappa_h = 1 (plain diffusion through the hyper term)
    n, L = (128, 128), (TWO_PI, TWO_PI)
    kh = 0.01
    nsteps = round(tfinal / dt)
    x, y = _pts(n, L)
    ps = n[::-1]
    prob = make(n=n, L=L, kappa=(0.0, 0.0), dt=dt, stepper=stepper, velocity=[np.zeros(ps), np.zeros(ps)],
                steady=True, kappa_h=kh, n_kappa_h=1)
    amp, sigma = 0.1, 0.1
    tfinal = nsteps * dt
    st = np.sqrt(2 * kh * tfinal + sigma ** 2)
    prob.set_c(amp * np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2)))
    prob.stepforward(nsteps)
    cfinal = amp * sigma ** 2 / st ** 2 * np.exp(-(x ** 2 + y ** 2) / (2 * st ** 2))
    return rel_l2(cfinal, prob.updatevars()), n[0] * n[1] * nsteps * 1e-12


# name -> (callable, kwargs) exactly as test/runtests.jl:26-54 invokes them
REFERENCE_KATS = {
    "constvel1D": (constvel1D, dict(dt=1e-2, nsteps=40)),
    "timedependentvel1D": (timedependentvel1D, dict(dt=0.002, tfinal=0.1)),
    "constvel2D": (constvel2D, dict(dt=1e-2, nsteps=40)),
    "timedependentvel2D": (timedependentvel2D, dict(dt=0.002, tfinal=0.1)),
    "constvel3D": (constvel3D, dict(dt=1e-2, nsteps=40)),
    "timedependentvel3D": (timedependentvel3D, dict(dt=0.002, tfinal=0.1)),
    "diffusion1D_steady": (diffusion1D, dict(dt=0.005, tfinal=0.1, steadyflow=True)),
    "diffusion1D_varying": (diffusion1D, dict(dt=0.005, tfinal=0.1, steadyflow=False)),
    "diffusion2D_steady": (diffusion2D, dict(dt=0.005, tfinal=0.1, steadyflow=True)),
    "diffusion2D_varying": (diffusion2D, dict(dt=0.005, tfinal=0.1, steadyflow=False)),
    "diffusion3D_steady": (diffusion3D, dict(dt=0.005, tfinal=0.1, steadyflow=True)),
    "diffusion3D_varying": (diffusion3D, dict(dt=0.005, tfinal=0.1, steadyflow=False)),
    "diffusion_multilayerqg": (diffusion_multilayerqg, dict(dt=0.005, tfinal=0.1)),
    "hyperdiffusion": (hyperdiffusion, dict(dt=0.005, tfinal=0.1)),
}
