"""The reference's examples/turbulent_advection-diffusion.jl, line for line, on the B200 device: a two-layer MultiLayerQG
flow advecting a passive tracer in both layers.  Both problems live on the GPU; nothing crosses PCIe inside the loop.

    python examples/turbulent_advection_diffusion.py [--n 128] [--nsteps 4000] [--release 25.0]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptf_b200 as P                                   # noqa: E402
from ptf_b200 import MultiLayerQG, TracerAdvectionDiffusion   # noqa: E402


def main(n=128, nsteps=4000, tracer_release_time=25.0, save_frequency=50, quiet=False):
    dev = P.B200()                                      # jl:21   dev = CPU()
    stepper, dt = "FilteredRK4", 2.5e-3                 # jl:34-36
    L, mu, beta = 2 * np.pi, 5e-2, 5                    # jl:40-42
    nlayers, f0, H, b = 2, 1, [0.2, 0.8], [-1.0, -1.2]  # jl:44-47
    U = np.zeros(nlayers)                               # jl:49-51
    U[0] = 1.0
    MQGprob = MultiLayerQG.Problem(nlayers, dev, nx=n, Lx=L, f0=f0, H=H, b=b, U=U, mu=mu, beta=beta, dt=dt, stepper=stepper,
                                   aliased_fraction=0)  # jl:56-58
    nx, ny = MQGprob.grid.nx, MQGprob.grid.ny           # jl:59
    rng = np.random.default_rng(1234)                   # jl:62   seed!(1234)
    q0 = 1e-2 * rng.standard_normal((nlayers, ny, nx))  # jl:63
    q0h = MQGprob.timestepper.filter * np.fft.rfft2(q0)                     # jl:64
    q0 = np.fft.irfft2(q0h, s=(ny, nx))                 # jl:65
    MultiLayerQG.set_q(MQGprob, q0)                     # jl:67

    kappa = 0.002                                       # jl:80
    ADprob = TracerAdvectionDiffusion.Problem(MQGprob, kappa=kappa, stepper=stepper,
                                              tracer_release_time=tracer_release_time)   # jl:84
    clock, grid = ADprob.clock, ADprob.grid             # jl:87-88
    x, y = grid.x, grid.y
    amplitude, spread = 10, 0.15                        # jl:97
    c0 = amplitude * np.exp(-(x[None, :] ** 2 + y[:, None] ** 2) / (2 * spread ** 2))   # jl:95-98
    TracerAdvectionDiffusion.set_c(ADprob, c0)          # jl:100

    start = time.time()
    frames = []
    while clock.step <= nsteps:                         # jl:141
        if clock.step % save_frequency == 0:            # jl:142-148
            TracerAdvectionDiffusion.updatevars(ADprob)
            frames.append(ADprob.vars.c[1].copy())      # concentration in the bottom layer
            if not quiet:
                print(f"Output saved, step: {clock.step:04d}, t: {clock.t:.2f}, walltime: {(time.time() - start) / 60:.2f} min, "
                      f"max c (bottom layer): {frames[-1].max():.4f}")
        P.stepforward(ADprob)                           # jl:149
        P.stepforward(MQGprob)                          # jl:150
        MultiLayerQG.updatevars(MQGprob)                # jl:151
    return ADprob, MQGprob, frames


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--nsteps", type=int, default=4000)
    ap.add_argument("--release", type=float, default=25.0)
    a = ap.parse_args()
    main(a.n, a.nsteps, a.release)
