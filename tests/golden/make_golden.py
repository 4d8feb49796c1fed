"""Generate the golden vectors in this directory FROM THE CPU ORACLE (oracle/ptf_oracle.py, oracle/mqg_oracle.py).

The reference (Julia + FFTW) cannot run in this image and ships no golden vectors, so these fixtures do not pin the oracle
to the reference — the reference's 14 analytic known-answer tests do that (tests/test_oracle_reference_kat.py).  They
freeze the oracle's OWN output on small seeded cases so that (a) an accidental change of the oracle is caught on CPU
and (b) the CUDA path is also checked against committed numbers that do not depend on the NumPy / pocketfft build on
the GPU box.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import cases as C                                            # noqa: E402
from oracle.mqg_oracle import MQGOracle                      # noqa: E402
from oracle.ptf_oracle import Grid, OracleProblem, irfft, make_filter, rfft   # noqa: E402


def run_tracer(name):
    n, L, stepper, nsteps, dt, steady, kw = C.TRACER[name]
    nd = len(n)
    g = Grid(n, L)
    pts = g.gridpoints()
    fs = C.velocity_functions(L)
    vel = [np.broadcast_to(f(*pts), g.pshape).copy() for f in fs] if steady else fs
    o = OracleProblem(n=n, L=L, kappa=C.KAPPA[:nd], dt=dt, stepper=stepper, velocity=vel, steady=steady, **kw)
    o.set_c(C.initial_c(pts))
    o.stepforward(nsteps)
    return dict(c=o.updatevars().copy(), sol=o.sol.copy())


def run_mqg(name):
    nl, n, stepper, nsteps, kw = C.MQG[name]
    o = MQGOracle(nl, nx=n, dt=C.MQG_DT, stepper=stepper, **kw)
    o.set_q(C.mqg_q0(nl, n, make_filter(o.grid), irfft, rfft, o.grid))
    o.stepforward(nsteps)
    o.updatevars()
    return dict(sol=o.sol.copy(), u=o.u.copy(), v=o.v.copy(), psi=o.psi.copy())


def main():
    out = {}
    for name in C.TRACER:
        out.update({f"{name}/{k}": v for k, v in run_tracer(name).items()})
    for name in C.MQG:
        out.update({f"{name}/{k}": v for k, v in run_mqg(name).items()})
    path = os.path.join(HERE, "oracle_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", len(C.TRACER) + len(C.MQG), "cases,", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
