"""Small driver for compute-sanitizer, new engines of this round: MultiLayerQG flow solver (2 and 3 layers), the coupled
tracer loop, and the 2-D slab engine with P = 1 (graphs off so every launch is checked individually)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

rng = np.random.default_rng(0)
for nl, stepper, n in ((2, "FilteredRK4", 64), (3, "ETDRK4", 48), (1, "AB3", 32)):
    H = {1: [1.0], 2: [0.2, 0.8], 3: [0.2, 0.3, 0.5]}[nl]
    b = {1: None, 2: [-1.0, -1.2], 3: [-1.0, -1.2, -1.5]}[nl]
    mq = P.MultiLayerQG.Problem(nl, P.B200(use_graph=False), nx=n, ny=n + 16, beta=5.0, H=H, b=b, U=list(np.linspace(1, 0, nl)),
                                mu=5e-2, nu=1e-6, nnu=2, dt=2.5e-3, stepper=stepper, aliased_fraction=1 / 3)
    mq.set_q(0.5 * rng.standard_normal((nl, n + 16, n)))
    mq.stepforward(4)
    mq.updatevars()
    _ = mq.vars.psi, mq.vars.u, mq.sol
    if nl == 2:
        ad = P.Problem(mq, kappa=0.002, stepper="FilteredRK4", dev=P.B200(use_graph=False))
        ad.set_c(np.exp(-rng.standard_normal((n + 16, n)) ** 2))
        P.MultiLayerQG.step_coupled(ad, 2)
        ad.updatevars()
        ad.close()
    mq.close()
flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
for stepper, nx, ny in (("RK4", 96, 64), ("FilteredETDRK4", 64, 80)):
    prob = P.Problem(P.B200(decomposition="slab", use_graph=False), flow, nx=nx, ny=ny, kappa=0.01, dt=1e-3, stepper=stepper,
                     dealias=True)
    X, Y = P.gridpoints(prob.grid)
    prob.set_c(np.exp(-(X ** 2 + Y ** 2)))
    prob.stepforward(2)
    prob.updatevars()
    prob.diagnostics()
    prob.close()
print("sanitize_run2 done")
