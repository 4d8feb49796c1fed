"""GPU parity of PTF_FLOW_EXPR (velocities as run-time compiled CUDA expressions evaluated at clock.t in registers)
against the CPU oracle driven by the equivalent NumPy closures (TAD.jl:695-742: frozen clock.t for all stages)."""
import numpy as np
import pytest

from oracle.ptf_oracle import OracleProblem, rel_l2

pytestmark = pytest.mark.gpu
TOL_STEP = 1e-12


def P():
    import ptf_b200
    return ptf_b200


def _run(nd, n, L, exprs, funcs, stepper, dt, nsteps, dev=None, engine="cufft", **kw):
    flow = P().ExpressionFlow(*exprs)
    names = ["nx", "ny", "nz"]
    lens = ["Lx", "Ly", "Lz"]
    gk = {names[a]: n[a] for a in range(nd)}
    gk.update({lens[a]: L[a] for a in range(nd)})
    prob = P().Problem(dev or P().B200(), flow, kappa=0.01, dt=dt, stepper=stepper, **gk, **kw)
    assert prob.engine == engine
    pts = P().gridpoints(prob.grid)
    pts = pts if isinstance(pts, tuple) else (pts,)
    c0 = np.exp(-sum((p - 0.2) ** 2 for p in pts) / 0.3)
    o = OracleProblem(n=n, L=L, kappa=(0.01,) * nd, dt=dt, stepper=stepper, velocity=funcs, steady=False)
    o.set_c(c0)
    prob.set_c(c0)
    o.stepforward(1)
    prob.stepforward(1)
    e1 = rel_l2(o.updatevars(), prob.updatevars())
    o.stepforward(nsteps - 1)
    prob.stepforward(nsteps - 1)
    en = rel_l2(o.updatevars(), prob.updatevars())
    prob.close()
    return e1, en


def test_expr_flow_1d():
    e1, en = _run(1, (128,), (2 * np.pi,), ["0.3 + 0.2*sin(x)*cos(2*t)"],
                  [lambda x, t: 0.3 + 0.2 * np.sin(x) * np.cos(2 * t)], "RK4", 5e-3, 8)
    assert e1 < TOL_STEP and en < 8 * TOL_STEP


@pytest.mark.parametrize("stepper", ["RK4", "FilteredRK4", "ETDRK4"])
def test_expr_flow_2d(stepper):
    L = (2 * np.pi, 4.0)
    ky = 2 * np.pi / L[1]
    ex = [f"(1 + 0.5*sin(3*t)) * cos(x) * sin({ky!r}*y)", f"-(1 + 0.5*sin(3*t)) * sin(x) * cos({ky!r}*y) + 0.1*exp(-t)"]
    fn = [lambda x, y, t: (1 + 0.5 * np.sin(3 * t)) * np.cos(x) * np.sin(ky * y),
          lambda x, y, t: -(1 + 0.5 * np.sin(3 * t)) * np.sin(x) * np.cos(ky * y) + 0.1 * np.exp(-t)]
    e1, en = _run(2, (96, 64), L, ex, fn, stepper, 5e-3, 6)
    assert e1 < TOL_STEP and en < 6 * TOL_STEP


def test_expr_flow_3d_abc():
    g = "(1 + 0.5*sin(t))"
    ex = [f"(sin(z) + 0.6*cos(y))*{g}", f"(0.8*sin(x) + cos(z))*{g}", f"(0.6*sin(y) + 0.8*cos(x))*{g}"]
    G = lambda t: 1 + 0.5 * np.sin(t)
    fn = [lambda x, y, z, t: (np.sin(z) + 0.6 * np.cos(y)) * G(t), lambda x, y, z, t: (0.8 * np.sin(x) + np.cos(z)) * G(t),
          lambda x, y, z, t: (0.6 * np.sin(y) + 0.8 * np.cos(x)) * G(t)]
    e1, en = _run(3, (32, 48, 64), (2 * np.pi,) * 3, ex, fn, "RK4", 5e-3, 4)
    assert e1 < TOL_STEP and en < 4 * TOL_STEP


@pytest.mark.parametrize("stepper", ["RK4", "FilteredETDRK4", "LSRK54"])
def test_expr_flow_2d_fused_engine(stepper):
    # power-of-two grids run the expressions on the fused 2-D engine: written out once per step by the run-time compiled
    # fill kernel (velocities are frozen at clock.t for all stages), read by the row kernel like steady arrays
    L = (2 * np.pi, 4.0)
    ky = 2 * np.pi / L[1]
    ex = [f"(1 + 0.5*sin(3*t)) * cos(x) * sin({ky!r}*y)", f"-(1 + 0.5*sin(3*t)) * sin(x) * cos({ky!r}*y) + 0.1*exp(-t)"]
    fn = [lambda x, y, t: (1 + 0.5 * np.sin(3 * t)) * np.cos(x) * np.sin(ky * y),
          lambda x, y, t: -(1 + 0.5 * np.sin(3 * t)) * np.sin(x) * np.cos(ky * y) + 0.1 * np.exp(-t)]
    e1, en = _run(2, (128, 64), L, ex, fn, stepper, 5e-3, 6, engine="fused")
    assert e1 < TOL_STEP and en < 6 * TOL_STEP


def test_expr_flow_3d_abc_fused_engine():
    g = "(1 + 0.5*sin(t))"
    ex = [f"(sin(z) + 0.6*cos(y))*{g}", f"(0.8*sin(x) + cos(z))*{g}", f"(0.6*sin(y) + 0.8*cos(x))*{g}"]
    G = lambda t: 1 + 0.5 * np.sin(t)
    fn = [lambda x, y, z, t: (np.sin(z) + 0.6 * np.cos(y)) * G(t), lambda x, y, z, t: (0.8 * np.sin(x) + np.cos(z)) * G(t),
          lambda x, y, z, t: (0.6 * np.sin(y) + 0.8 * np.cos(x)) * G(t)]
    e1, en = _run(3, (64, 128, 64), (2 * np.pi,) * 3, ex, fn, "RK4", 5e-3, 4, engine="fused")
    assert e1 < TOL_STEP and en < 4 * TOL_STEP


def test_expr_flow_on_the_2d_slab_engine():
    ex = ["(1 + 0.5*sin(3*t)) * cos(x) * sin(y)", "-(1 + 0.5*sin(3*t)) * sin(x) * cos(y)"]
    fn = [lambda x, y, t: (1 + 0.5 * np.sin(3 * t)) * np.cos(x) * np.sin(y),
          lambda x, y, t: -(1 + 0.5 * np.sin(3 * t)) * np.sin(x) * np.cos(y)]
    e1, en = _run(2, (64, 96), (2 * np.pi,) * 2, ex, fn, "RK4", 5e-3, 5, dev=P().B200(decomposition="slab"))
    assert e1 < TOL_STEP and en < 5 * TOL_STEP


def test_expr_flow_errors():
    with pytest.raises(ValueError, match="does not compile"):
        P().Problem(P().B200(), P().ExpressionFlow("sin(x) + no_such_function(y)", "0.0"), nx=64)
    with pytest.raises(ValueError):
        P().Problem(P().B200(), P().ExpressionFlow("sin(x); while(1){}", "0.0"), nx=64)
    with pytest.raises(P()._capi.PtfError):      # the fused 1-D engine has no expression flows
        P().Problem(P().B200(engine="fused"), P().ExpressionFlow("sin(x)"), nx=256)


@pytest.mark.parametrize("engine", ["cufft", "auto"])
def test_expr_flow_can_be_replaced_between_steps(engine):
    n, L = (64, 64), (2 * np.pi, 2 * np.pi)
    prob = P().Problem(P().B200(engine=engine), P().ExpressionFlow("cos(x)*sin(y)", "-sin(x)*cos(y)"), nx=64, kappa=0.01,
                       dt=5e-3, stepper="RK4")
    X, Y = P().gridpoints(prob.grid)
    c0 = np.exp(-((X - 0.2) ** 2 + (Y - 0.2) ** 2) / 0.3)
    f1 = [lambda x, y, t: np.cos(x) * np.sin(y), lambda x, y, t: -np.sin(x) * np.cos(y)]
    f2 = [lambda x, y, t: 0.5 + 0.1 * t + 0 * x, lambda x, y, t: -np.sin(x) * np.cos(y)]
    o = OracleProblem(n=n, L=L, kappa=(0.01, 0.01), dt=5e-3, stepper="RK4", velocity=f1, steady=False)
    o.set_c(c0)
    prob.set_c(c0)
    o.stepforward(3)
    prob.stepforward(3)
    o.vel_funcs = f2
    prob.set_velocity_expr(0, "0.5 + 0.1*t")          # recompiles; the captured step graph is rebuilt
    o.stepforward(3)
    prob.stepforward(3)
    assert rel_l2(o.updatevars(), prob.updatevars()) < 6 * TOL_STEP
    prob.close()
