"""Time one RK4 step of BASELINE configs[1] (4096^2) under different tuning-knob settings (env vars read at
engine creation).  Usage: python tools/tune.py "PTF_PF_STATE=0" "PTF_PF_STATE=1,PTF_PF_AHEAD_X=296" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

nx = int(os.environ.get("TUNE_NX", "4096"))
nbatch = int(os.environ.get("TUNE_BATCH", "1"))
dt = 0.5 * 2.785 / (0.1 * 2 * (nx / 2) ** 2)
flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
x = -np.pi + (2 * np.pi / nx) * np.arange(nx)
c0 = 0.5 * np.exp(-((x[None, :] - 0.4 * np.pi) ** 2 + x[:, None] ** 2) / (2 * 0.15 ** 2))
KEYS = ["PTF_PF_STATE", "PTF_PF_VEL", "PTF_PF_AHEAD_Y", "PTF_PF_AHEAD_X", "PTF_SMEM_PAD", "PTF_ABLATE_X", "PTF_ABLATE_Y", "PTF_STAGGER_X", "PTF_STAGGER_Y", "PTF_NT", "PTF_X_DIRECT"]
for spec in (sys.argv[1:] or [""]):
    for k in KEYS:
        os.environ.pop(k, None)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    prob = P.Problem(P.B200(engine=os.environ.get("TUNE_ENGINE", "fused")), flow, nx=nx, kappa=0.1, dt=dt, stepper="RK4",
                     nbatch=nbatch)
    prob.set_c(c0)
    prob.stepforward(10)
    ms = min(prob.step_timed(25) / 25 for _ in range(3))
    ky = prob.kernel_time_ms("ykernel", 5) if prob.engine == "fused" else float("nan")
    kx = prob.kernel_time_ms("xkernel", 5) if prob.engine == "fused" else float("nan")
    frac = 432 * nx * nx * nbatch / (ms * 1e-3) / 6453.1e9
    print(f"{spec or 'defaults':45s} ms/step {ms:7.3f}  B_alg-frac {frac:5.3f}  ykernel {ky:6.3f} ms  xkernel {kx:6.3f} ms",
          flush=True)
    prob.close()
