"""Small driver for ncu: a few RK4 steps of the 2-D slab engine with P = 1 (graphs off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
flow = P.SeparableFlow(terms=[[(np.cos, np.sin)], [(np.sin, np.cos)]], coeffs=lambda t, a: [0.2 if a == 0 else -0.2], steadyflow=True)
dt = 0.5 * 2.785 / (0.1 * 2 * (n / 2) ** 2)
prob = P.Problem(P.B200(decomposition="slab", use_graph=False), flow, nx=n, kappa=0.1, dt=dt, stepper="RK4")
x = prob.grid.x
prob.set_c(0.5 * np.exp(-((x[None, :] - 0.4 * np.pi) ** 2 + x[:, None] ** 2) / (2 * 0.15 ** 2)))
prob.stepforward(2)
print(prob.engine, prob.launch_count())
