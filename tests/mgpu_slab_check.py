"""Run under torchrun on N GPUs (one process per GPU): slab-decomposed 3-D problems vs the CPU oracle.
Exit code 0 = parity.  Used by tests/test_gpu_multi.py, by bench.py (`partitioned.parity_rel_l2`) and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_slab_check.py

Cases: the cuFFT slab pipeline (sizes that are not powers of two) and the fused 3-D engine (powers of two in
[64, 1024]); steady array flows, a multi-term SeparableFlow with time-dependent coefficients (BASELINE configs[3]'s
flow) and an ExpressionFlow; RK4 / FilteredRK4 / ETDRK4; SURVEY 8(d) config 4's 128^3 against the oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptf_b200 as P                                   # noqa: E402
from oracle.ptf_oracle import OracleProblem, rel_l2    # noqa: E402  (checker only)

_one = lambda s: 1.0 + 0 * s


def _abc_callables(L, g=lambda t: 1.0):
    kx, ky, kz = (2 * np.pi / Lv for Lv in L)
    u = lambda x, y, z, t=0.0: g(t) * (np.sin(kz * z) + 0.6 * np.cos(ky * y)) + 0 * x
    v = lambda x, y, z, t=0.0: g(t) * (0.8 * np.sin(kx * x) + np.cos(kz * z)) + 0 * y
    w = lambda x, y, z, t=0.0: g(t) * (0.6 * np.sin(ky * y) + 0.8 * np.cos(kx * x)) + 0 * z
    return u, v, w


def slab_case(dev, n, stepper, flow_kind="arrays", nsteps=3, L=(2 * np.pi, 4.0, 3.0), dt=2e-3, expect_engine=None):
    """One slab-decomposed problem on this rank's slab against the full-grid oracle; returns the worst relative L2
    error over set_c, c, sol and the diagnostics."""
    nx, ny, nz = n
    kx, ky, kz = (2 * np.pi / Lv for Lv in L)
    g = (lambda t: 1.0 + 0.5 * np.sin(3 * t)) if flow_kind in ("separable", "expression") else (lambda t: 1.0)
    u, v, w = _abc_callables(L, g)
    kap = dict(kappa=0.01, eta=0.02, iota=0.005, kappa_h=1e-6, n_kappa_h=2)
    if flow_kind == "arrays":
        flow = P.ThreeDAdvectingFlow(u=lambda x, y, z: u(x, y, z), v=lambda x, y, z: v(x, y, z),
                                     w=lambda x, y, z: w(x, y, z), steadyflow=True)
    elif flow_kind == "separable":   # two terms per component: exercises the z tables of terms m >= 1 on ranks > 0
        sz, cy = (lambda z: np.sin(kz * z)), (lambda y: np.cos(ky * y))
        sx, cz = (lambda x: np.sin(kx * x)), (lambda z: np.cos(kz * z))
        sy, cx = (lambda y: np.sin(ky * y)), (lambda x: np.cos(kx * x))
        flow = P.SeparableFlow(terms=[[(_one, _one, sz), (_one, cy, _one)], [(sx, _one, _one), (_one, _one, cz)],
                                      [(_one, sy, _one), (cx, _one, _one)]],
                               coeffs=lambda t, a: g(t) * np.array([[1.0, 0.6], [0.8, 1.0], [0.6, 0.8]][a]),
                               steadyflow=False)
    else:
        amp = "(1.0 + 0.5*sin(3*t))"
        flow = P.ExpressionFlow(u=f"{amp}*(sin({kz!r}*z) + 0.6*cos({ky!r}*y))", v=f"{amp}*(0.8*sin({kx!r}*x) + cos({kz!r}*z))",
                                w=f"{amp}*(0.6*sin({ky!r}*y) + 0.8*cos({kx!r}*x))")
    prob = P.Problem(dev, flow, nx=nx, Lx=L[0], ny=ny, Ly=L[1], nz=nz, Lz=L[2], dt=dt, stepper=stepper, **kap)
    if expect_engine is not None:
        assert prob.engine == expect_engine, f"{n}: engine {prob.engine}, expected {expect_engine}"
    X, Y, Z = P.gridpoints(prob.grid)
    c0 = np.exp(-(X ** 2 / 0.4 + Y ** 2 / 0.3 + Z ** 2 / 0.2))
    sl = slice(prob.z_offset, prob.z_offset + prob.nz_local)
    prob.set_c(np.ascontiguousarray(c0[sl]))
    if flow_kind == "arrays":
        vel, steady = [np.broadcast_to(f(X, Y, Z), c0.shape) for f in (u, v, w)], True
    else:
        vel, steady = [u, v, w], False
    o = OracleProblem(n=n, L=L, kappa=(kap["kappa"], kap["eta"], kap["iota"]), dt=dt, stepper=stepper, velocity=vel,
                      steady=steady, kappa_h=kap["kappa_h"], n_kappa_h=kap["n_kappa_h"])
    o.set_c(c0)
    ysl = slice(prob.ky_offset, prob.ky_offset + prob.ny_local)
    # a rank's slab may hold almost nothing of the field (the tails of the Gaussian, the high-|l| rows of its spectrum):
    # its error is measured against the norm of the WHOLE field, not of the slab
    gerr = lambda ref, mine, sel: float(np.linalg.norm(ref[sel] - mine) / np.linalg.norm(ref))
    e0 = gerr(o.sol, prob.sol, (slice(None), ysl, slice(None)))
    o.stepforward(nsteps)
    prob.stepforward(nsteps)
    c = prob.updatevars()
    e_c = gerr(o.updatevars(), c, sl)
    e_s = gerr(o.sol, prob.sol, (slice(None), ysl, slice(None)))
    d = prob.diagnostics()
    e_d = abs(d["mean_c"] - o.c.mean()) + abs(d["variance_c"] - o.c.var())
    engine = prob.engine
    prob.close()
    return max(e0, e_c, e_s, e_d), dict(set_c=e0, c=e_c, sol=e_s, diag=e_d, engine=engine)


CASES = [
    # cuFFT slab pipeline (not powers of two)
    ("RK4", (64, 48, 32), "arrays", "cufft"), ("FilteredRK4", (48, 32, 96), "arrays", "cufft"),
    ("ETDRK4", (48, 64, 32), "arrays", "cufft"), ("RK4", (48, 32, 64), "separable", "cufft"),
    ("RK4", (32, 48, 64), "expression", "cufft"),
    # fused 3-D engine, slab-decomposed (exchange of contiguous blocks, pipelined over kr chunks)
    ("RK4", (64, 64, 64), "arrays", "fused"), ("FilteredRK4", (128, 64, 64), "arrays", "fused"),
    ("ETDRK4", (64, 128, 64), "arrays", "fused"), ("LSRK54", (64, 64, 128), "arrays", "fused"),
    ("RK4", (64, 64, 128), "separable", "fused"),
    # 128^3 against the oracle
    ("RK4", (128, 128, 128), "separable", "fused"),
]


def main():
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = P.parallel.init_b200("slab", device=local_rank)
    worst = 0.0
    for stepper, n, kind, engine in CASES:
        if n[1] % world or n[2] % world or (engine == "fused" and n[2] // world < 8):
            continue
        L = (2 * np.pi,) * 3 if n == (128, 128, 128) else (2 * np.pi, 4.0, 3.0)
        e, info = slab_case(dev, n, stepper, kind, nsteps=2 if n == (128, 128, 128) else 3, L=L, expect_engine=engine)
        worst = max(worst, e)
        if rank == 0:
            print(f"[slab x{world}] {stepper} {n} {kind} [{info['engine']}]: set_c {info['set_c']:.2e}  c {info['c']:.2e}  "
                  f"sol {info['sol']:.2e}  diag {info['diag']:.2e}", flush=True)
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t.item() <= 3e-12
    if rank == 0:
        print("SLAB PARITY", "OK" if ok else "FAILED", f"worst {t.item():.2e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
