"""Device-timed throughput of every BASELINE.json config that fits one GPU (CUDA events on the step stream).
Prints one JSON object per config; results are copied into profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    PEAK = 6451.8e9
out = []


def timed(prob, nsteps, reps=3):
    prob.stepforward(max(3, nsteps // 4))
    return min(prob.step_timed(nsteps) / nsteps for _ in range(reps))


def report(name, prob, npts_total, ms, balg):
    r = {"config": name, "engine": prob.engine, "ms_per_step": ms, "grid_point_steps_per_s": npts_total / (ms * 1e-3),
         "b_alg": balg, "frac_of_hbm_roofline": balg * npts_total / (ms * 1e-3) / PEAK}
    out.append(r)
    print(json.dumps(r), flush=True)


# configs[0]: 1-D Gaussian diffusion, nx = 128, u = 0.05, kappa = 0.01, dt = 0.02, RK4 (latency-bound: report us/step)
prob = P.Problem(P.B200(), P.OneDAdvectingFlow(u=lambda x: 0.05 + 0 * x), nx=128, kappa=0.01, dt=0.02, stepper="RK4")
x = P.gridpoints(prob.grid)
prob.set_c(np.exp(-x ** 2 / (2 * 0.15 ** 2)))
report("cfg0 1-D nx=128 RK4", prob, 128, timed(prob, 2000), 304)
prob.close()

# configs[1] at 128^2 (the example's own size), 1024^2, 2048^2, 4096^2
for nx, kappa, dt in ((64, 0.002, 0.02), (128, 0.002, 0.02), (256, 0.002, 0.01), (1024, 0.1, None), (2048, 0.1, None), (4096, 0.1, None)):
    dt = dt or 0.5 * 2.785 / (kappa * 2 * (nx / 2) ** 2)
    flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
    for stepper in (("RK4", "ETDRK4", "FilteredRK4") if nx == 4096 else ("RK4",)):
        prob = P.Problem(P.B200(), flow, nx=nx, kappa=kappa, dt=dt, stepper=stepper)
        X, Y = P.gridpoints(prob.grid)
        prob.set_c(0.5 * np.exp(-((X - 0.4 * np.pi) ** 2 + Y ** 2) / (2 * 0.15 ** 2)))
        balg = 432 + (4 if stepper.startswith("Filtered") else 0) - (16 if stepper.endswith("ETDRK4") else 0)
        report(f"cfg1 2-D cellular {nx}^2 {stepper}", prob, nx * nx, timed(prob, 200 if nx < 4096 else 50), balg)
        prob.close()

# configs[2]: 2 layers x 512^2, FilteredRK4, kappa = 0.002, dt = 2.5e-3, synthetic layered flow + U = [1, 0]
n, B = 512, 2
rng = np.random.default_rng(1234)
psi_h = np.zeros((B, n, n // 2 + 1), dtype=complex)
psi_h[:, :12, :12] = rng.standard_normal((B, 12, 12)) + 1j * rng.standard_normal((B, 12, 12))
kx = np.arange(n // 2 + 1)
ky = np.where(np.arange(n) < n // 2, np.arange(n), np.arange(n) - n)
u = np.fft.irfft2(-1j * ky[None, :, None] * psi_h, s=(n, n))
v = np.fft.irfft2(1j * kx[None, None, :] * psi_h, s=(n, n))
rms = np.sqrt(np.mean(u ** 2 + v ** 2))
T = P.tracer_advection_diffusion
grid = T.Grid(nx=n, Lx=2 * np.pi, ny=n, Ly=2 * np.pi, ndim=2)
prob = T.TracerProblem(P.B200(), grid, T.Params(0.002, 0.002, 0.002, 0.0, 0), 2.5e-3, "FilteredRK4", P._capi.FLOW_LAYERED,
                       nbatch=B, velocity_per_batch=True)
prob.set_layered_velocity(u / rms, v / rms, np.array([1.0, 0.0]))
X, Y = P.gridpoints(grid)
prob.set_c(10 * np.exp(-(X ** 2 + Y ** 2) / (2 * 0.15 ** 2)))
report("cfg2 2 layers x 512^2 FilteredRK4 (layered flow + U)", prob, B * n * n, timed(prob, 500), 436)
prob.close()

# configs[4] per-GPU share: 32 members x 1024^2, shared cellular velocity, RK4
nx, B = 1024, 32
dt = 0.5 * 2.785 / (0.1 * 2 * (nx / 2) ** 2)
flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
prob = P.Problem(P.B200(), flow, nx=nx, kappa=0.1, dt=dt, stepper="RK4", nbatch=B)
X, Y = P.gridpoints(prob.grid)
cx = np.linspace(-1, 1, B).reshape(B, 1, 1)
prob.set_c(np.exp(-((X - cx) ** 2 + (Y + 0.5 * cx) ** 2) / (2 * 0.3 ** 2)))
report("cfg4 ensemble 32 x 1024^2 RK4 (one GPU's share of 256)", prob, B * nx * nx, timed(prob, 50), 432)
prob.close()

# 3-D single GPU: 256^3 / 512^3 steady ABC arrays, RK4, fused 3-D engine and (256^3) the cuFFT engine beside it
for n, eng in ((256, "cufft"), (256, "auto"), (512, "auto")):
    flow = P.ThreeDAdvectingFlow(u=lambda x, y, z: np.sin(z) + np.cos(y) + 0 * x, v=lambda x, y, z: np.sin(x) + np.cos(z) + 0 * y,
                                 w=lambda x, y, z: np.sin(y) + np.cos(x) + 0 * z)
    prob = P.Problem(P.B200(engine=eng), flow, nx=n, kappa=0.01, dt=1e-3, stepper="RK4")
    x1 = prob.grid.x
    prob.set_c(np.exp(-(x1[None, None, :] ** 2 + x1[None, :, None] ** 2 + x1[:, None, None] ** 2) / (2 * 0.3 ** 2)))
    report(f"3-D {n}^3 steady arrays RK4 [{eng}]", prob, n ** 3, timed(prob, 20 if n == 256 else 6), 560)
    prob.close()
# (kept for the record: the same 256^3 problem through the default engine)
n = 256
flow = P.ThreeDAdvectingFlow(u=lambda x, y, z: np.sin(z) + np.cos(y) + 0 * x, v=lambda x, y, z: np.sin(x) + np.cos(z) + 0 * y,
                             w=lambda x, y, z: np.sin(y) + np.cos(x) + 0 * z)
prob = P.Problem(P.B200(), flow, nx=n, kappa=0.01, dt=1e-3, stepper="RK4")
X, Y, Z = P.gridpoints(prob.grid)
prob.set_c(np.exp(-(X ** 2 + Y ** 2 + Z ** 2) / (2 * 0.3 ** 2)))
report("3-D 256^3 steady arrays RK4", prob, n ** 3, timed(prob, 20), 560)
prob.close()
# 3-D time-varying ABC flow at 256^3, RK4: how the velocity reaches the product kernel
n = 256
G = lambda t: 1 + 0.5 * np.sin(t)
one = lambda s: 1.0 + 0 * s
flows = {
    "host closures + upload per step (PTF_FLOW_CALLBACK)": P.ThreeDAdvectingFlow(
        u=lambda x, y, z, t: (np.sin(z) + 0.6 * np.cos(y)) * G(t) + 0 * x, v=lambda x, y, z, t: (0.8 * np.sin(x) + np.cos(z)) * G(t) + 0 * y,
        w=lambda x, y, z, t: (0.6 * np.sin(y) + 0.8 * np.cos(x)) * G(t) + 0 * z, steadyflow=False),
    "run-time compiled expressions (PTF_FLOW_EXPR)": P.ExpressionFlow(
        "(sin(z) + 0.6*cos(y))*(1 + 0.5*sin(t))", "(0.8*sin(x) + cos(z))*(1 + 0.5*sin(t))", "(0.6*sin(y) + 0.8*cos(x))*(1 + 0.5*sin(t))"),
    "separable tables (PTF_FLOW_SEPARABLE)": P.SeparableFlow(
        terms=[[(one, one, np.sin), (one, np.cos, one)], [(np.sin, one, one), (one, one, np.cos)],
               [(one, np.sin, one), (np.cos, one, one)]],
        coeffs=lambda t, a: G(t) * np.array([[1.0, 0.6], [0.8, 1.0], [0.6, 0.8]][a]), steadyflow=False),
}
import time
for name, flow in flows.items():
    prob = P.Problem(P.B200(), flow, nx=n, kappa=0.01, dt=1e-3, stepper="RK4")
    X, Y, Z = P.gridpoints(prob.grid)
    prob.set_c(np.exp(-(X ** 2 + Y ** 2 + Z ** 2) / (2 * 0.3 ** 2)))
    prob.stepforward(3)
    t0 = time.perf_counter()
    prob.stepforward(10)                      # wall clock: the host-side closure evaluation and upload are the point
    ms = 1e3 * (time.perf_counter() - t0) / 10
    balg = 560 - (0 if "CALLBACK" in name else 96)
    report(f"3-D 256^3 time-varying ABC flow RK4, wall clock per step, {name}", prob, n ** 3, ms, balg)
    prob.close()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_configs_1gpu.json"), "w"), indent=1)
