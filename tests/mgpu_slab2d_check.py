"""Run under torchrun on N GPUs (one process per GPU): slab-decomposed 2-D problem vs the CPU oracle.
Exit code 0 = parity.  Used by tests/test_gpu_slab2d.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/mgpu_slab2d_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptf_b200 as P                                   # noqa: E402
from oracle.ptf_oracle import OracleProblem, rel_l2    # noqa: E402  (checker only)


def slab2d_case(dev, n, stepper, rank, world, verbose=True):
    """One slab-decomposed 2-D problem on this rank's rows / kr columns against the full-grid oracle."""
    L = (2 * np.pi, 4.0)
    ky = 2 * np.pi / L[1]
    u = lambda x, y: 0.2 * np.cos(x) * np.sin(ky * y)
    v = lambda x, y: -0.3 * np.sin(x) * np.cos(ky * y)
    nx, ny = n
    flow = P.TwoDAdvectingFlow(u=u, v=v, steadyflow=True)
    prob = P.Problem(dev, flow, nx=nx, Lx=L[0], ny=ny, Ly=L[1], kappa=0.01, eta=0.02, dt=2e-3, stepper=stepper,
                     kappa_h=1e-6, n_kappa_h=2)
    nkr = nx // 2 + 1
    kc = -(-nkr // world)
    assert prob.ny_phys_local == ny // world and prob.y_offset == rank * (ny // world)
    assert prob.kr_offset == rank * kc and prob.nkr_local == max(0, min(kc, nkr - rank * kc))
    X, Y = P.gridpoints(prob.grid)
    c0 = 0.5 * np.exp(-((X - 0.4) ** 2 / 0.3 + Y ** 2 / 0.2))
    sl = slice(prob.y_offset, prob.y_offset + prob.ny_phys_local)
    ksl = slice(prob.kr_offset, prob.kr_offset + prob.nkr_local)
    prob.set_c(np.ascontiguousarray(c0[sl]))
    o = OracleProblem(n=n, L=L, kappa=(0.01, 0.02), dt=2e-3, stepper=stepper, velocity=[u(X, Y), v(X, Y)], steady=True,
                      kappa_h=1e-6, n_kappa_h=2)
    o.set_c(c0)
    # a rank's block holds only a band of kr: its error is measured against the norm of the WHOLE spectrum (a smooth
    # field leaves ~1e-12 of the energy in the high-kr band, whose own norm is at rounding level)
    blk_err = lambda: float(np.linalg.norm(o.sol[:, ksl] - prob.sol) / np.linalg.norm(o.sol))
    e0 = blk_err()
    o.stepforward(4)
    prob.stepforward(4)
    c = prob.updatevars()
    oc = o.updatevars()
    e_c = float(np.linalg.norm(oc[sl] - c) / np.linalg.norm(oc))   # against the WHOLE field: a rank's rows may hold only tails
    e_s = blk_err()
    d = prob.diagnostics()
    e_d = abs(d["mean_c"] - o.c.mean()) + abs(d["variance_c"] - o.c.var())
    if verbose and (rank == 0 or max(e0, e_c, e_s, e_d) > 5e-12):
        print(f"[slab2d x{world} rank {rank}] {stepper} {n}: set_c {e0:.2e}  c {e_c:.2e}  sol {e_s:.2e}  diag {e_d:.2e}", flush=True)
    prob.close()
    return max(e0, e_c, e_s, e_d)


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = P.parallel.init_b200("slab", device=local_rank)
    worst = 0.0
    for stepper, n in (("RK4", (96, 64)), ("FilteredRK4", (128, 80)), ("ETDRK4", (64, 128)), ("LSRK54", (256, 64))):
        if n[1] % world:
            continue
        worst = max(worst, slab2d_case(dev, n, stepper, rank, world))
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t.item() <= 5e-12
    if rank == 0:
        print("SLAB2D PARITY", "OK" if ok else "FAILED", f"worst {t.item():.2e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
