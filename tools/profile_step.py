"""Small driver for ncu: BASELINE configs[1] (4096^2 cellular flow, RK4) — a few steps, nothing else."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
engine = sys.argv[3] if len(sys.argv) > 3 else "auto"
dt = 0.5 * 2.785 / (0.1 * 2 * (nx / 2) ** 2)
flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
prob = P.Problem(P.B200(engine=engine, use_graph=False), flow, nx=nx, kappa=0.1, dt=dt, stepper="RK4")
x, y = P.gridpoints(prob.grid)
P.set_c(prob, 0.5 * np.exp(-((x - 0.4 * np.pi) ** 2 + y ** 2) / (2 * 0.15 ** 2)))
P.stepforward(prob, nsteps)
print(prob.engine, prob.launch_count(), prob.diagnostics())
