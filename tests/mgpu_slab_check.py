"""Run under torchrun on N GPUs (one process per GPU): slab-decomposed 3-D problem vs the CPU oracle and vs the
single-GPU result.  Exit code 0 = parity.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_slab_check.py
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


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = P.parallel.init_b200("slab", device=local_rank)
    worst = 0.0
    for stepper, n in (("RK4", (64, 48, 32)), ("FilteredRK4", (32, 32, 64)), ("ETDRK4", (48, 64, 32))):
        nx, ny, nz = n
        L = (2 * np.pi, 4.0, 3.0)
        ky, kz = 2 * np.pi / L[1], 2 * np.pi / L[2]
        u = lambda x, y, z: np.sin(kz * z) + np.cos(ky * y) + 0 * x
        v = lambda x, y, z: np.sin(x) + np.cos(kz * z) + 0 * y
        w = lambda x, y, z: np.sin(ky * y) + np.cos(x) + 0 * z
        flow = P.ThreeDAdvectingFlow(u=u, v=v, w=w, steadyflow=True)
        prob = P.Problem(dev, flow, nx=nx, Lx=L[0], ny=ny, Ly=L[1], nz=nz, Lz=L[2], kappa=0.01, eta=0.02, iota=0.005,
                         dt=2e-3, stepper=stepper, kappa_h=1e-6, n_kappa_h=2)
        assert prob.nz_local == nz // world and prob.z_offset == rank * (nz // world)
        g = prob.grid
        X, Y, Z = P.gridpoints(g)
        c0 = np.exp(-(X ** 2 / 0.4 + Y ** 2 / 0.3 + Z ** 2 / 0.2))
        sl = slice(prob.z_offset, prob.z_offset + prob.nz_local)
        prob.set_c(np.ascontiguousarray(c0[sl]))
        o = OracleProblem(n=n, L=L, kappa=(0.01, 0.02, 0.005), dt=2e-3, stepper=stepper,
                          velocity=[np.broadcast_to(f(X, Y, Z), c0.shape) for f in (u, v, w)], steady=True,
                          kappa_h=1e-6, n_kappa_h=2)
        o.set_c(c0)
        ysl = slice(prob.ky_offset, prob.ky_offset + prob.ny_local)
        e0 = rel_l2(o.sol[:, ysl, :], prob.sol)
        o.stepforward(3)
        prob.stepforward(3)
        c = prob.updatevars()
        e_c = rel_l2(o.updatevars()[sl], c)
        e_s = rel_l2(o.sol[:, ysl, :], prob.sol)
        d = prob.diagnostics()
        e_d = abs(d["mean_c"] - o.c.mean()) + abs(d["variance_c"] - o.c.var())
        worst = max(worst, e0, e_c, e_s, e_d)
        if rank == 0:
            print(f"[slab x{world}] {stepper} {n}: set_c {e0:.2e}  c {e_c:.2e}  sol {e_s:.2e}  diag {e_d:.2e}", flush=True)
        prob.close()
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t.item() <= 3e-12
    if rank == 0:
        print("SLAB PARITY", "OK" if ok else "FAILED", f"worst {t.item():.2e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
