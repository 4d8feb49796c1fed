# needs the experiments build: python passivetracerflows.jl_b200/build.py --variant exp PTF_FFT_EXPERIMENTS=1 ; PTF_LIB_PATH=.../libptf_b200_exp.so
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["PTF_SELFTEST_TIME"] = "1"
import ptf_b200 as P
lib = P._capi.load()
dp = C.POINTER(C.c_double)
rng = np.random.default_rng(1)
for n, count in ((256, 70), (1024, 18), (4096, 6), (4096, 8192)):
    x = rng.standard_normal((n, count)) + 1j * rng.standard_normal((n, count))
    y = np.zeros_like(x)
    P._capi.check(lib.ptf_selftest_fft(n, 2, count, x.ctypes.data_as(dp), y.ctypes.data_as(dp)))
    ref = np.fft.fft(x, axis=0)
    err = np.linalg.norm(ref - y) / np.linalg.norm(ref)
    print(f"pair test n={n} count={count}: rel err {err:.2e}", flush=True)
