"""Small driver for compute-sanitizer, round 2: the fused 3-D engine (all four kernels, chunked pipeline hook, separable +
expression fills), the new 64/128-point transforms and the per-group barriers (warp barriers for T < 32, named barriers for
T = 32 / 64) of the fused 2-D engine.  Graphs off so every launch is checked individually.
  compute-sanitizer --tool memcheck  python tools/sanitize_run3.py
  compute-sanitizer --tool racecheck python tools/sanitize_run3.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

dev = P.B200(engine="fused", use_graph=False)
one = lambda s: 1.0 + 0 * s
# fused 3-D: 64 x 128 x 64 (T = 4 and 8: several transforms per warp), arrays / separable / expressions, three steppers
for stepper, flow in (
        ("RK4", P.ThreeDAdvectingFlow(u=lambda x, y, z: np.sin(z) + 0.6 * np.cos(y) + 0 * x,
                                      v=lambda x, y, z: 0.8 * np.sin(x) + np.cos(z) + 0 * y,
                                      w=lambda x, y, z: 0.6 * np.sin(y) + 0.8 * np.cos(x) + 0 * z)),
        ("FilteredETDRK4", P.SeparableFlow(terms=[[(one, one, np.sin), (one, np.cos, one)], [(np.sin, one, one), (one, one, np.cos)],
                                                  [(one, np.sin, one), (np.cos, one, one)]],
                                           coeffs=lambda t, a: (1 + 0.5 * np.sin(t)) * np.array([[1.0, 0.6], [0.8, 1.0], [0.6, 0.8]][a]),
                                           steadyflow=False)),
        ("LSRK54", P.ExpressionFlow("sin(z) + 0.6*cos(y)", "0.8*sin(x) + cos(z)", "(0.6*sin(y) + 0.8*cos(x))*(1+t)"))):
    prob = P.Problem(dev, flow, nx=64, ny=128, nz=64, kappa=0.01, dt=2e-3, stepper=stepper, dealias=(stepper == "LSRK54"))
    x, y, z = prob.grid.x, prob.grid.y, prob.grid.z
    prob.set_c(np.exp(-(x[None, None, :] ** 2 + y[None, :, None] ** 2 + z[:, None, None] ** 2)))
    prob.stepforward(2)
    c = prob.updatevars()
    _ = prob.sol
    print("3-D", stepper, prob.engine, float(np.abs(c).max()), prob.diagnostics()["variance_c"], flush=True)
    prob.close()
# fused 2-D with several transforms per CTA: 64 (T = 4), 128 (8), 512 (32), 1024 (64)
for n in (64, 128, 512, 1024):
    flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
    prob = P.Problem(dev, flow, nx=n, kappa=0.002, dt=1e-3 if n <= 128 else 1e-5, stepper="FilteredRK4", nbatch=2)
    X, Y = P.gridpoints(prob.grid)
    prob.set_c(0.5 * np.exp(-((X - 0.4) ** 2 + Y ** 2) / 0.1))
    prob.stepforward(2)
    print("2-D", n, prob.engine, float(np.abs(prob.updatevars()).max()), flush=True)
    prob.close()
