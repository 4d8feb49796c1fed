"""Run under torchrun on N GPUs (one process per GPU): an ENSEMBLE of 2-D tracers sharded over the ranks
(PTF_DECOMP_BATCH: members [rank*B/N, (rank+1)*B/N) per rank, no data-path collective — BASELINE configs[4]'s scheme)
against the CPU oracle, member by member.  Exit code 0 = parity.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_batch_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptf_b200 as P                                   # noqa: E402
from oracle.ptf_oracle import OracleProblem, rel_l2    # noqa: E402  (checker only)


def member_c0(X, Y, b, B):
    """deterministic lattice of Gaussian centres, one per ensemble member"""
    side = int(np.ceil(np.sqrt(B)))
    cx = -1.5 + 3.0 * (b % side) / max(1, side - 1)
    cy = -1.5 + 3.0 * (b // side) / max(1, side - 1)
    return 0.5 * np.exp(-((X - cx) ** 2 + (Y - cy) ** 2) / (2 * 0.3 ** 2))


def batch_case(dev, n, B, stepper, rank, world, nsteps=4):
    u = lambda x, y: 0.2 * np.cos(x) * np.sin(y)
    v = lambda x, y: -0.2 * np.sin(x) * np.cos(y)
    prob = P.Problem(dev, P.TwoDAdvectingFlow(u=u, v=v, steadyflow=True), nx=n, kappa=0.002, dt=0.01, stepper=stepper,
                     nbatch=B)
    assert prob.local_nbatch == B // world and prob.batch_offset == rank * (B // world)
    X, Y = P.gridpoints(prob.grid)
    mine = range(prob.batch_offset, prob.batch_offset + prob.local_nbatch)
    c0 = np.stack([member_c0(X, Y, b, B) for b in mine])
    prob.set_c(c0)
    prob.stepforward(nsteps)
    c = prob.updatevars()
    worst = 0.0
    for i, b in enumerate(mine):
        o = OracleProblem(n=(n, n), L=(2 * np.pi,) * 2, kappa=(0.002, 0.002), dt=0.01, stepper=stepper,
                          velocity=[u(X, Y), v(X, Y)], steady=True)
        o.set_c(member_c0(X, Y, b, B))
        o.stepforward(nsteps)
        worst = max(worst, rel_l2(o.updatevars(), c[i]))
    engine = prob.engine
    prob.close()
    return worst, engine


def main():
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = P.parallel.init_b200("batch", device=local_rank)
    worst = 0.0
    for stepper, n, per_rank in (("RK4", 128, 3), ("FilteredRK4", 256, 2), ("ETDRK4", 96, 2)):
        e, engine = batch_case(dev, n, per_rank * world, stepper, rank, world)
        worst = max(worst, e)
        if rank == 0:
            print(f"[batch x{world}] {stepper} {per_rank * world} x {n}^2 [{engine}]: worst member {e:.2e}", flush=True)
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t.item() <= 5e-12
    if rank == 0:
        print("BATCH PARITY", "OK" if ok else "FAILED", f"worst {t.item():.2e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
