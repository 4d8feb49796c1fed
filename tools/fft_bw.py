# needs the experiments build: python passivetracerflows.jl_b200/build.py --variant exp PTF_FFT_EXPERIMENTS=1 ; PTF_LIB_PATH=.../libptf_b200_exp.so
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["PTF_SELFTEST_TIME"] = "1"
os.environ["PTF_SELFTEST_REPEAT"] = "16"
import ptf_b200 as P
lib = P._capi.load()
dp = C.POINTER(C.c_double)
for n in (4096, 1024):
    count = (8192 * 4096) // n
    x = np.random.default_rng(0).standard_normal((count, n)) + 0j
    y = np.empty_like(x)
    for pad in (0, 65536):
        os.environ["PTF_SMEM_PAD"] = str(pad)
        print("n", n, "smem pad", pad, flush=True)
        P._capi.check(lib.ptf_selftest_fft(n, -1, count, x.ctypes.data_as(dp), y.ctypes.data_as(dp)))
