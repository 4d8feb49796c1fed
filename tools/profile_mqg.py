"""Small driver for ncu: a few coupled iterations of BASELINE configs[2] (2-layer MultiLayerQG flow + tracer), graphs off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
EX = dict(beta=5.0, f0=1.0, H=[0.2, 0.8], b=[-1.0, -1.2], U=[1.0, 0.0], mu=5e-2)
mq = P.MultiLayerQG.Problem(2, P.B200(use_graph=False), nx=n, dt=2.5e-3, stepper="FilteredRK4", aliased_fraction=0.0, **EX)
mq.set_q(1e-2 * np.random.default_rng(1234).standard_normal((2, n, n)))
ad = P.Problem(mq, kappa=0.002, stepper="FilteredRK4", dev=P.B200(use_graph=False))
x = -np.pi + 2 * np.pi / n * np.arange(n)
X, Y = np.meshgrid(x, x)
ad.set_c(10 * np.exp(-(X ** 2 + Y ** 2) / (2 * 0.15 ** 2)))
P.MultiLayerQG.step_coupled(ad, nit)
print(ad.engine, ad.launch_count(), mq.launch_count())
