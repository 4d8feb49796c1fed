# needs the experiments build: python passivetracerflows.jl_b200/build.py --variant exp PTF_FFT_EXPERIMENTS=1 ; PTF_LIB_PATH=.../libptf_b200_exp.so
"""Small driver for compute-sanitizer: one fused 2-D step per stepper family + the 1-D engine + FFT self-tests."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ptf_b200 as P

lib = P._capi.load()
dp = C.POINTER(C.c_double)
for n in (256, 1024):
    x = np.random.default_rng(0).standard_normal((6, n)) + 0j
    y = np.empty_like(x)
    P._capi.check(lib.ptf_selftest_fft(n, -1, 6, x.ctypes.data_as(dp), y.ctypes.data_as(dp)))
    xt = np.ascontiguousarray(x.T)
    P._capi.check(lib.ptf_selftest_fft(n, 2, 6, xt.ctypes.data_as(dp), np.empty_like(xt).ctypes.data_as(dp)))
flow = P.TwoDAdvectingFlow(u=lambda x, y: 0.2 * np.cos(x) * np.sin(y), v=lambda x, y: -0.2 * np.sin(x) * np.cos(y))
for stepper, nx, ny, dealias in (("RK4", 256, 512, False), ("FilteredETDRK4", 512, 256, False), ("LSRK54", 256, 256, True)):
    prob = P.Problem(P.B200(engine="fused", use_graph=False), flow, nx=nx, ny=ny, kappa=0.01, dt=1e-3, stepper=stepper,
                     dealias=dealias)
    X, Y = P.gridpoints(prob.grid)
    prob.set_c(np.exp(-(X ** 2 + Y ** 2)))
    prob.stepforward(2)
    prob.updatevars()
    prob.close()
prob = P.Problem(P.B200(engine="fused"), P.OneDAdvectingFlow(u=lambda x: 0.1 + 0 * x), nx=128, kappa=0.01, dt=0.01)
prob.set_c(np.exp(-P.gridpoints(prob.grid) ** 2))
prob.stepforward(3)
prob.updatevars()
print("sanitize_run done")
