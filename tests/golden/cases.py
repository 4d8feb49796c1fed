"""Case definitions shared by make_golden.py (writes oracle_golden.npz), tests/test_golden.py (oracle vs fixtures on CPU,
CUDA path vs fixtures on the GPU)."""
import numpy as np

TRACER = {
    # name: (n, L, stepper, nsteps, dt, steady, extra oracle/problem kwargs)
    "tracer1d_rk4_varying": ((64,), (2 * np.pi,), "RK4", 5, 5e-3, False, {}),
    "tracer2d_rk4_steady": ((48, 32), (2 * np.pi, 4.0), "RK4", 5, 5e-3, True, dict(kappa_h=1e-6, n_kappa_h=2)),
    "tracer2d_filteredrk4_varying": ((32, 48), (2 * np.pi, 4.0), "FilteredRK4", 4, 5e-3, False, {}),
    "tracer2d_etdrk4_dealias": ((32, 32), (2 * np.pi, 2 * np.pi), "ETDRK4", 4, 5e-3, True, dict(dealias=True)),
    "tracer2d_lsrk54": ((32, 32), (2 * np.pi, 2 * np.pi), "LSRK54", 3, 5e-3, True, {}),
    "tracer2d_ab3": ((32, 32), (2 * np.pi, 2 * np.pi), "AB3", 6, 5e-3, True, {}),
    "tracer3d_rk4_varying": ((16, 24, 32), (2 * np.pi, 4.0, 3.0), "RK4", 3, 5e-3, False, {}),
}
KAPPA = (0.01, 0.02, 0.005)

_EX = dict(beta=5.0, f0=1.0, H=[0.2, 0.8], b=[-1.0, -1.2], U=[1.0, 0.0], mu=5e-2)
MQG = {
    # name: (nlayers, n, stepper, nsteps, kwargs)
    "mqg2_filteredrk4_example_params": (2, 32, "FilteredRK4", 8, dict(aliased_fraction=0.0, **_EX)),
    "mqg3_rk4_dealias": (3, 32, "RK4", 5, dict(aliased_fraction=1 / 3, beta=4.0, f0=1.2, H=[0.2, 0.3, 0.5],
                                               b=[-1.0, -1.2, -1.5], U=[1.0, 0.5, 0.0], mu=0.1, nu=1e-5, nnu=2)),
}
MQG_DT = 2.5e-3


def velocity_functions(L):
    """Time-dependent closures u(x[,y[,z]], t); the steady cases sample them at t = 0."""
    k = [2 * np.pi / Lv for Lv in L]
    nd = len(L)
    if nd == 1:
        return [lambda x, t=0.0: 0.3 + 0.2 * np.sin(k[0] * x) * np.cos(2 * t)]
    if nd == 2:
        return [lambda x, y, t=0.0: 0.2 * np.cos(k[0] * x) * np.sin(k[1] * y) * (1 + 0.5 * np.sin(3 * t)),
                lambda x, y, t=0.0: -0.3 * np.sin(k[0] * x) * np.cos(k[1] * y) * (1 + 0.5 * np.sin(3 * t))]
    return [lambda x, y, z, t=0.0: (np.sin(k[2] * z) + 0.6 * np.cos(k[1] * y)) * (1 + 0.5 * np.sin(t)) + 0 * x,
            lambda x, y, z, t=0.0: (0.8 * np.sin(k[0] * x) + np.cos(k[2] * z)) * (1 + 0.5 * np.sin(t)) + 0 * y,
            lambda x, y, z, t=0.0: (0.6 * np.sin(k[1] * y) + 0.8 * np.cos(k[0] * x)) * (1 + 0.5 * np.sin(t)) + 0 * z]


def initial_c(pts):
    return np.exp(-sum((p - 0.2) ** 2 for p in pts) / 0.3)


def mqg_q0(nl, n, filt, irfft, rfft, grid):
    q0 = 0.5 * np.random.default_rng(1234).standard_normal((nl, n, n))
    return irfft(grid, filt * rfft(grid, q0))
