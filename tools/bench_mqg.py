"""BASELINE configs[2] end to end on the device: the 2-layer MultiLayerQG flow (examples/turbulent_advection-diffusion.jl
parameters) driving a FilteredRK4 tracer, the example's loop body per iteration
(stepforward!(ADprob); stepforward!(MQGprob); MultiLayerQG.updatevars!(MQGprob)).  Device-timed with CUDA events on the
shared stream; the CPU oracle's time for the same iteration is printed beside it (bounded sample)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P
from oracle.mqg_oracle import MQGOracle
from oracle.ptf_oracle import OracleProblem, irfft, make_filter, rfft

EX = dict(beta=5.0, f0=1.0, H=[0.2, 0.8], b=[-1.0, -1.2], U=[1.0, 0.0], mu=5e-2)
out = []
for n in (128, 512, 1024):
    dt = 2.5e-3 * 128 / n if n > 512 else 2.5e-3
    o = MQGOracle(2, nx=n, dt=dt, stepper="FilteredRK4", aliased_fraction=0.0, **EX)
    q0 = 1e-2 * np.random.default_rng(1234).standard_normal((2, n, n))
    q0 = irfft(o.grid, make_filter(o.grid) * rfft(o.grid, q0))
    mq = P.MultiLayerQG.Problem(2, P.B200(), nx=n, dt=dt, stepper="FilteredRK4", aliased_fraction=0.0, **EX)
    mq.set_q(q0)
    mq.stepforward(200)                                   # spin-up (the example runs to t = 25 first)
    flow_ms = min(mq.step_timed(200, with_updatevars=True) for _ in range(3)) / 200
    ad = P.Problem(mq, kappa=0.002, stepper="FilteredRK4")
    x = -np.pi + 2 * np.pi / n * np.arange(n)
    X, Y = np.meshgrid(x, x)
    c0 = 10 * np.exp(-(X ** 2 + Y ** 2) / (2 * 0.15 ** 2))
    ad.set_c(c0)
    P.MultiLayerQG.step_coupled(ad, 50)
    coupled_ms = min(P.MultiLayerQG.step_coupled(ad, 200) for _ in range(3)) / 200
    tracer_ms = min(ad.step_timed(200) for _ in range(3)) / 200
    c = ad.updatevars()
    # e2e: the example's three calls per iteration through the public API (host-synchronous), result read back
    t0 = time.perf_counter()
    for _ in range(100):
        ad.stepforward(1); mq.stepforward(1); mq.updatevars()
    c = ad.updatevars()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / 100
    # CPU oracle: same iteration
    o.set_q(q0)
    ot = OracleProblem(n=(n, n), L=(2 * np.pi,) * 2, kappa=(0.002, 0.002), dt=dt, stepper="FilteredRK4", velocity="layered",
                       steady=True, nbatch=2)
    ot.set_c(c0)
    nit = 20 if n <= 128 else (5 if n <= 512 else 2)
    t0 = time.perf_counter()
    for _ in range(nit):
        ot.set_layered_velocity(o.u, o.v, o.params.U); ot.stepforward(1); o.stepforward(1); o.updatevars()
    cpu_ms = 1e3 * (time.perf_counter() - t0) / nit
    own_m, lib_m = mq.launch_count()
    r = {"config": f"cfg2 coupled MultiLayerQG(2 layers) + tracer, {n}^2 per layer, FilteredRK4", "tracer_engine": ad.engine,
         "ms_per_coupled_iteration": coupled_ms, "flow_step_plus_updatevars_ms": flow_ms, "tracer_step_ms": tracer_ms,
         "e2e_three_calls_ms": e2e_ms, "tracer_grid_point_steps_per_s": 2 * n * n / (coupled_ms * 1e-3),
         "cpu_oracle_ms_per_iteration": cpu_ms, "cpu_threads": os.cpu_count(), "state_finite": bool(np.isfinite(c).all()),
         "h2d_bytes_per_iteration": 0, "reference_h2d_bytes_per_iteration_if_flow_on_host": 2 * 2 * n * n * 8}
    out.append(r)
    print(json.dumps(r), flush=True)
    ad.close(); mq.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r01_mqg_coupled.json", "w"), indent=1)
