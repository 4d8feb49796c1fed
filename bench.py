#!/usr/bin/env python
"""Benchmark of the TracerAdvectionDiffusion hot path (BASELINE.json metric: grid-point RK4 steps/sec, fp64).

Workload at N GPUs: BASELINE.json configs[1] — examples/cellularflow.jl scaled to 4096^2 (kappa = 0.1, RK4,
steady cellular flow, fp64) — one independent tracer problem per GPU (ensemble sharding, no data-path
collective => "scaling": "weak").

One bench "step" = one frame of the example's loop, ``stepforward!(prob, nsubs=25)`` (examples/cellularflow.jl:134):
  * ``value``  : grid-point RK4 steps/s with everything resident in HBM (device-timed, CUDA events, max over ranks)
  * ``e2e``    : the same frame through the C ABI with HOST buffers: ``set_c!`` from pinned host memory (H2D),
                 ``stepforward!(prob, 25)``, ``updatevars!`` into pinned host memory (D2H), all inside the timed region
  * ``roofline``: dominant kernel, algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  * ``cpu_baseline``: the CPU oracle port (NumPy + threaded pocketfft) on a bounded sample of the same workload

``--impl reference`` times the CPU restatement of the reference path (the reference itself is Julia + FFTW, neither
of which exists in this image — see DESIGN.md) on the box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSUBS = 25          # RK4 steps per frame (examples/cellularflow.jl:33)
METRIC = "grid-point RK4 steps/sec (fp64)"
UNIT = "grid-point-steps/s"


def workload(nx):
    """examples/cellularflow.jl:30-81 scaled to nx^2 with an RK4-stable dt."""
    Lx = 2 * np.pi
    kappa = 0.1
    dt = 0.5 * 2.785 / (kappa * 2 * (nx / 2) ** 2)
    x = -Lx / 2 + (Lx / nx) * np.arange(nx)
    X, Y = x[None, :], x[:, None]
    psi0 = 0.2
    u = np.ascontiguousarray(np.broadcast_to(psi0 * np.cos(X) * np.sin(Y), (nx, nx)))
    v = np.ascontiguousarray(np.broadcast_to(-psi0 * np.sin(X) * np.cos(Y), (nx, nx)))
    c0 = np.ascontiguousarray(0.5 * np.exp(-((X - 0.2 * Lx) ** 2 + Y ** 2) / (2 * 0.15 ** 2)))
    return dict(nx=nx, Lx=Lx, kappa=kappa, dt=dt, u=u, v=v, c0=c0)


def b_alg(ndim, stepper="RK4"):
    """Algorithmic bytes per grid point per step (SURVEY section 8d / BASELINE.md section 3)."""
    b = 128 * ndim + 176
    if stepper.startswith("Filtered"):
        b += 4
    if stepper.endswith("ETDRK4"):
        b -= 16
    return b


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_port_rate(w, seconds_budget=20.0, min_steps=2, max_steps=8):
    """Time the CPU oracle (port of the reference path) on a bounded sample: a few RK4 steps of the same workload."""
    from oracle.ptf_oracle import OracleProblem
    cores = os.cpu_count() or 1
    nx = w["nx"]
    o = OracleProblem(n=(nx, nx), L=(w["Lx"], w["Lx"]), kappa=(w["kappa"], w["kappa"]), dt=w["dt"], stepper="RK4",
                      velocity=[w["u"], w["v"]], steady=True, workers=cores)
    o.set_c(w["c0"])
    o.stepforward(1)   # warm-up (thread pool, page faults)
    n = 0
    t0 = time.perf_counter()
    while n < max_steps and (n < min_steps or time.perf_counter() - t0 < seconds_budget):
        o.stepforward(1)
        n += 1
    dtm = time.perf_counter() - t0
    import scipy
    return {"value": nx * nx * n / dtm, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} RK4 steps of the {nx}x{nx} cellular-flow workload, NumPy {np.__version__} elementwise + "
                      f"scipy.fft {scipy.__version__} (pocketfft) workers={cores}; NOT Julia/FFTW (absent from the image)",
            "ms_per_rk4_step": 1e3 * dtm / n}


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = workload(args.nx)
    # each "step" of the reference arm is a bounded sample: one RK4 step of the workload (a frame is 25 of them)
    from oracle.ptf_oracle import OracleProblem
    cores = os.cpu_count() or 1
    nx = w["nx"]
    o = OracleProblem(n=(nx, nx), L=(w["Lx"], w["Lx"]), kappa=(w["kappa"], w["kappa"]), dt=w["dt"], stepper="RK4",
                      velocity=[w["u"], w["v"]], steady=True, workers=cores)
    o.set_c(w["c0"])
    for _ in range(args.warmup):
        o.stepforward(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.stepforward(1)
    dtm = time.perf_counter() - t0
    val = nx * nx * args.steps / dtm
    import scipy
    sample = (f"each step = 1 RK4 step (1/{NSUBS} of a frame) of the {nx}x{nx} cellular-flow workload; CPU restatement of "
              f"the reference path (NumPy {np.__version__} + scipy.fft {scipy.__version__} pocketfft, workers={cores}); "
              "the reference's Julia/FFTW runtime is absent from this image")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dtm / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cellularflow_{nx}x{nx}_RK4_kappa0.1 (BASELINE configs[1])", "nx": nx, "ny": nx,
                       "stepper": "RK4", "dt": w["dt"], "rk4_steps_per_step": 1},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_slab3d(args, rank, world, local_rank, P, dist, barrier, max_over_ranks):
    """BASELINE configs[3]: 3-D prescribed time-dependent (ABC-type, separable) flow, n^3 fp64 with hyperdiffusion,
    slab-decomposed across the GPUs; the NCCL all-to-all transpose is the only collective.  Strong scaling."""
    n = args.n3
    L = 2 * np.pi
    A, B, Cc = 1.0, 0.8, 0.6
    one = lambda s: 1.0 + 0 * s
    g = lambda t: 1.0 + 0.5 * np.sin(t)
    flow = P.SeparableFlow(
        terms=[[(one, one, np.sin), (one, np.cos, one)],      # u = (A sin z + C cos y) g(t)
               [(np.sin, one, one), (one, one, np.cos)],      # v = (B sin x + A cos z) g(t)
               [(one, np.sin, one), (np.cos, one, one)]],     # w = (C sin y + B cos x) g(t)
        coeffs=lambda t, a: g(t) * np.array([[A, Cc], [B, A], [Cc, B]][a]),
        steadyflow=False)
    kmax2 = 3 * (n / 2) ** 2
    kappa_h = 1e-3 / kmax2          # hyperdiffusion (n_kappa_h = 2): max|L| = kappa_h * kmax2^2
    dt = min(0.5 * 2.785 / (kappa_h * kmax2 ** 2), 0.5 * 2.83 / (3 * 1.5 * 1.8 * n / 2))
    dev = P.parallel.init_b200("slab", device=local_rank) if world > 1 else P.B200(device=local_rank)
    prob = P.Problem(dev, flow, nx=n, kappa=0.0, dt=dt, stepper=args.stepper, kappa_h=kappa_h, n_kappa_h=2)
    x = prob.grid.x
    zs = x[prob.z_offset:prob.z_offset + prob.nz_local]
    sig = 0.1 * 8                   # resolved Gaussian
    c0 = np.exp(-(x[None, None, :] ** 2 + x[None, :, None] ** 2 + zs[:, None, None] ** 2) / (2 * sig ** 2))
    prob.set_c(np.ascontiguousarray(c0))
    for _ in range(args.warmup):
        prob.stepforward(1)
    own0, lib0 = prob.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms = prob.step_timed(args.steps)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    own1, lib1 = prob.launch_count()
    dev_ms = max_over_ranks(dev_ms)
    npts = n ** 3
    value = npts * args.steps / (dev_ms * 1e-3)
    d = prob.diagnostics()
    peak, peak_src = peaks()
    balg = b_alg(3, args.stepper) - 32 * 3      # velocities generated in registers: -32*d
    step_ms = dev_ms / args.steps
    # per-GPU NVLink floor: each transposed field moves (P-1)/P of the local spectral slab out of every GPU
    spec_local = (n // 2 + 1) * n * n * 16 / world
    a2a_bytes = 16 * spec_local * (world - 1) / world if world > 1 else 0
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"slab3d_{n}^3_{args.stepper}_hyperdiffusion_separable_ABC_flow (BASELINE configs[3])",
                           "n": n, "stepper": args.stepper, "dt": dt, "engine": prob.engine,
                           "decomposition": "z-slabs (physical) / ky-slabs (spectral), NCCL all-to-all per transform",
                           "state_finite": bool(np.isfinite(d["max_abs_sol"])), "l2": "fields larger than L2"},
                "gpu_launches": own1 - own0, "library_calls": lib1 - lib0,
                "step_roofline": {"b_alg_bytes_per_point_step": balg, "achieved": balg * npts / (step_ms * 1e-3) / 1e9,
                                  "peak": peak * world, "unit": "GB/s",
                                  "frac": balg * npts / (step_ms * 1e-3) / 1e9 / (peak * world)},
                "nvlink": {"alltoall_bytes_out_per_gpu_per_step": a2a_bytes,
                           "floor_ms_per_step_at_770GBs": a2a_bytes / 770e9 * 1e3},
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)


def run_slab2d(args, rank, world, local_rank, P, dist, barrier, max_over_ranks):
    """One large 2-D cellular-flow problem (BASELINE configs[1] scaled up) slab-decomposed over the GPUs: physical rows
    sharded, spectral kr-columns sharded, one NCCL all-to-all per transform.  Strong scaling."""
    n = args.n2
    kappa = 0.1
    dt = 0.5 * 2.785 / (kappa * 2 * (n / 2) ** 2)
    flow = P.SeparableFlow(terms=[[(np.cos, np.sin)], [(np.sin, np.cos)]],
                           coeffs=lambda t, a: [0.2 if a == 0 else -0.2], steadyflow=True)   # cellular flow, in registers
    dev = P.parallel.init_b200("slab", device=local_rank) if world > 1 else P.B200(device=local_rank, decomposition="slab")
    prob = P.Problem(dev, flow, nx=n, kappa=kappa, dt=dt, stepper=args.stepper)
    x = prob.grid.x
    ys = x[prob.y_offset:prob.y_offset + prob.ny_phys_local]
    c0 = 0.5 * np.exp(-((x[None, :] - 0.4 * np.pi) ** 2 + ys[:, None] ** 2) / (2 * 0.15 ** 2))
    prob.set_c(np.ascontiguousarray(c0))
    for _ in range(args.warmup):
        prob.stepforward(1)
    own0, lib0 = prob.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms = prob.step_timed(args.steps)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    own1, lib1 = prob.launch_count()
    dev_ms = max_over_ranks(dev_ms)
    npts = n * n
    value = npts * args.steps / (dev_ms * 1e-3)
    d = prob.diagnostics()
    peak, peak_src = peaks()
    balg = b_alg(2, args.stepper) - 32 * 2      # velocities generated in registers
    step_ms = dev_ms / args.steps
    spec_local = (n // 2 + 1) * n * 16 / world
    a2a_bytes = 12 * spec_local * (world - 1) / world if world > 1 else 0     # 3 exchanged fields per stage
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"slab2d_{n}^2_{args.stepper}_cellular_flow (BASELINE configs[1] beyond one fused-engine grid)",
                           "n": n, "stepper": args.stepper, "dt": dt, "engine": "slab2d (cuFFT batches + own kernels)",
                           "decomposition": "y-rows (physical) / kr-columns (spectral), NCCL all-to-all per transform",
                           "state_finite": bool(np.isfinite(d["max_abs_sol"])), "l2": "fields larger than L2"},
                "gpu_launches": own1 - own0, "library_calls": lib1 - lib0,
                "step_roofline": {"b_alg_bytes_per_point_step": balg, "achieved": balg * npts / (step_ms * 1e-3) / 1e9,
                                  "peak": peak * world, "unit": "GB/s",
                                  "frac": balg * npts / (step_ms * 1e-3) / 1e9 / (peak * world)},
                "nvlink": {"alltoall_bytes_out_per_gpu_per_step": a2a_bytes,
                           "floor_ms_per_step_at_770GBs": a2a_bytes / 770e9 * 1e3},
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--engine", default="auto", choices=["auto", "cufft", "fused"])
    ap.add_argument("--stepper", default="RK4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cellular2d", choices=["cellular2d", "slab3d", "slab2d"],
                    help="cellular2d: BASELINE configs[1], one problem per GPU (default, the graded line); "
                         "slab3d: BASELINE configs[3], ONE n^3 problem slab-decomposed over all GPUs (strong scaling)")
    ap.add_argument("--n3", type=int, default=512, help="grid size of the slab3d workload (n^3)")
    ap.add_argument("--n2", type=int, default=16384, help="grid size of the slab2d workload (n^2)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        args.steps = 4 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args, rank, world)
        return
    args.steps = 8 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else max(3, args.warmup)

    import ctypes as C

    import torch   # plumbing only: pinned host memory + torch.distributed barrier / max-reduce
    import ptf_b200 as P
    capi = P._capi
    lib = capi.load()

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "slab2d":
        run_slab2d(args, rank, world, local_rank, P, dist, barrier, max_over_ranks)
        if dist is not None:
            dist.destroy_process_group()
        return
    if args.workload == "slab3d":
        run_slab3d(args, rank, world, local_rank, P, dist, barrier, max_over_ranks)
        if dist is not None:
            dist.destroy_process_group()
        return

    w = workload(args.nx)
    nx = w["nx"]
    npts = nx * nx
    flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y),
                               steadyflow=True)
    prob = P.Problem(P.B200(device=local_rank, engine=args.engine), flow, nx=nx, Lx=w["Lx"], kappa=w["kappa"],
                     dt=w["dt"], stepper=args.stepper)
    h = prob._h
    # pinned host buffers for the e2e leg
    c_in = torch.from_numpy(w["c0"].copy()).pin_memory()
    c_out = torch.empty_like(c_in).pin_memory()
    p_in = C.cast(c_in.data_ptr(), C.POINTER(C.c_double))
    p_out = C.cast(c_out.data_ptr(), C.POINTER(C.c_double))
    capi.check(lib.ptf_set_c(h, p_in, 0), h)

    # ---------------- device-resident leg (value) ----------------
    for _ in range(args.warmup):
        prob.stepforward(NSUBS)
    own0, libc0 = prob.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_ms += prob.step_timed(NSUBS)      # CUDA events on the step stream, per frame
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.stop_flag.set()
    sampler.join()
    own1, libc1 = prob.launch_count()
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    value = world * npts * NSUBS * args.steps / (dev_ms * 1e-3)

    # sanity: the state must still be finite (a NaN run is not a measurement)
    capi.check(lib.ptf_get_c(h, p_out), h)
    finite = bool(torch.isfinite(c_out).all().item())

    # ---------------- e2e leg: host buffers through the C ABI ----------------
    for _ in range(2):
        capi.check(lib.ptf_set_c(h, p_in, 0), h)
        capi.check(lib.ptf_step(h, NSUBS), h)
        capi.check(lib.ptf_get_c(h, p_out), h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        capi.check(lib.ptf_set_c(h, p_in, 0), h)       # H2D from pinned memory + r2c
        capi.check(lib.ptf_step(h, NSUBS), h)          # stepforward!(prob, 25)
        capi.check(lib.ptf_get_c(h, p_out), h)         # updatevars! + D2H into pinned memory
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * npts * NSUBS * args.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": npts * 8, "d2h_bytes_per_step": npts * 8,
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "call": "ptf_set_c(pinned host) + ptf_step(25) + ptf_get_c(pinned host)"}

    # ---------------- roofline of the dominant kernel (live CUDA-event timing, in place) ----------------
    peak, peak_src = peaks()
    engine = prob.engine
    roof = None
    kernels = {}
    spec_b = (nx // 2 + 1) * nx * 16          # one spectral field
    real_b = npts * 8                           # one physical field
    if engine == "fused":
        # algorithmic (compulsory) bytes per launch, DESIGN.md "Kernels": the column kernel averages 7.25 spectral
        # fields per launch over the 4 RK4 stages (6 + 8 + 8 + 7); the row kernel reads A, B, u, v and writes P^x
        cands = [("ykernel", int(7.25 * spec_b)), ("xkernel", 3 * spec_b + 2 * real_b)]
    else:
        cands = [("deriv", 3 * 16 * (nx // 2 + 1) * nx), ("z2d", 2 * 8 * npts), ("d2z", 2 * 8 * npts)]
    for name, alg_bytes in cands:
        try:
            ms = prob.kernel_time_ms(name, 20)
        except Exception as e:   # a kernel name this engine does not have
            continue
        kernels[name] = {"ms": ms, "alg_bytes": alg_bytes}
    step_bytes = b_alg(2, args.stepper) * npts
    step_ms = dev_ms / (args.steps * NSUBS)
    # measured DRAM traffic per launch (dram__bytes_read + dram__bytes_write) from the committed ncu --set full capture
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_fused_4096.json")))
        for name, tag in (("ykernel", "k_fused_y"), ("xkernel", "k_fused_x")):
            tr = [pk["dram_traffic_bytes"] for pk in prof if tag in pk["kernel"]]
            if tr and name in kernels and nx == 4096:
                kernels[name]["traffic"] = sum(tr) / len(tr)     # mean over the captured launches (4 RK4 stages)
    except Exception:
        pass
    for k in kernels.values():
        if k["alg_bytes"]:
            k["achieved_gbs"] = k["alg_bytes"] / (k["ms"] * 1e-3) / 1e9
            k["frac_of_peak"] = k["achieved_gbs"] / peak
    if kernels:
        top = max(kernels.items(), key=lambda kv: kv[1]["ms"])
        name, k = top
        if k["alg_bytes"]:
            ach = k["alg_bytes"] / (k["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": k.get("traffic"), "peak_source": peak_src, "ms_per_launch": k["ms"],
                    "alg_bytes_per_launch": k["alg_bytes"]}
    step_roof = {"b_alg_bytes_per_point_step": b_alg(2, args.stepper), "achieved": step_bytes / (step_ms * 1e-3) / 1e9,
                 "peak": peak, "unit": "GB/s", "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                 "ms_per_rk4_step": step_ms}
    if roof is None:
        roof = {"bound": "hbm", "kernel": "whole RK4 step", "achieved": step_roof["achieved"], "peak": peak,
                "unit": "GB/s", "frac": step_roof["frac"], "traffic": None, "peak_source": peak_src}

    # ---------------- CPU baseline (rank 0, N == 1) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_rate(w)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"cellularflow_{nx}x{nx}_{args.stepper}_kappa0.1 (BASELINE configs[1]), one "
                                       f"independent problem per GPU", "nx": nx, "ny": nx, "stepper": args.stepper,
                           "dt": w["dt"], "rk4_steps_per_step": NSUBS, "engine": engine,
                           "l2": "inputs larger than L2 (each field 134 MB > 126 MB L2); no explicit flush",
                           "state_finite": finite, "wall_ms_per_step": wall_ms / args.steps},
                "e2e": e2e, "gpu_launches": own1 - own0, "library_calls": libc1 - libc0,
                "roofline": roof, "step_roofline": step_roof, "kernels": kernels, "cpu_baseline": cpu,
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
