"""Driver for ncu / kernel timing of the fused 3-D engine: BASELINE configs[3]-style n^3 problem (separable ABC flow, or
steady array velocities with `arrays`), a few RK4 steps.  usage: python tools/profile3d.py n nsteps [arrays] [time]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
arrays = "arrays" in sys.argv
one = lambda s: 1.0 + 0 * s
if arrays:
    flow = P.ThreeDAdvectingFlow(u=lambda x, y, z: np.sin(z) + 0.6 * np.cos(y) + 0 * x,
                                 v=lambda x, y, z: 0.8 * np.sin(x) + np.cos(z) + 0 * y,
                                 w=lambda x, y, z: 0.6 * np.sin(y) + 0.8 * np.cos(x) + 0 * z, steadyflow=True)
else:
    flow = P.SeparableFlow(terms=[[(one, one, np.sin), (one, np.cos, one)], [(np.sin, one, one), (one, one, np.cos)],
                                  [(one, np.sin, one), (np.cos, one, one)]],
                           coeffs=lambda t, a: np.array([[1.0, 0.6], [0.8, 1.0], [0.6, 0.8]][a]), steadyflow=True)
dt = 0.2 * 2.83 / (3 * 1.8 * n / 2)
prob = P.Problem(P.B200(engine="fused", use_graph="time" in sys.argv), flow, nx=n, kappa=1e-4, dt=dt, stepper="RK4")
x = prob.grid.x
prob.set_c(np.exp(-(x[None, None, :] ** 2 + x[None, :, None] ** 2 + x[:, None, None] ** 2) / (2 * 0.5 ** 2)))
prob.stepforward(nsteps)
out = {"n": n, "engine": prob.engine, "velocity": "arrays" if arrays else "separable"}
if "time" in sys.argv:
    out["ms_per_step"] = prob.step_timed(5) / 5
    spec = (n // 2 + 1) * n * n * 16
    real = n ** 3 * 8
    # compulsory bytes per launch: z kernel averages 7.25 spectral fields over the 4 RK4 stages
    alg = {"zkernel": 7.25 * spec, "yinv": 5 * spec, "xkernel": 4 * spec + (3 * real if arrays else 0), "yfwd": 2 * spec}
    for k in ("zkernel", "yinv", "xkernel", "yfwd"):
        ms = prob.kernel_time_ms(k, 3)
        out[k] = {"ms": ms, "alg_gbs": alg[k] / ms / 1e6}
print(json.dumps(out))
