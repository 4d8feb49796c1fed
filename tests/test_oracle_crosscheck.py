"""Independent checks of the ORACLE's transform conventions, so that the checker itself is not only
trusted through pocketfft: (1) a naive O(N²) DFT written from the definition, (2) a second FFT library (torch.fft on
CPU = MKL/pocketfft build of PyTorch) run through one full RK4 step of the oracle."""
import numpy as np
import pytest

import oracle.ptf_oracle as O


def _naive_rfft2(c):
    ny, nx = c.shape
    nkr = nx // 2 + 1
    jx = np.arange(nx)
    jy = np.arange(ny)
    Fx = np.exp(-2j * np.pi * np.outer(np.arange(nkr), jx) / nx)           # e^{-i kr x}, unnormalised
    Fy = np.exp(-2j * np.pi * np.outer(jy, jy) / ny)
    return Fy @ (c @ Fx.T)                                                 # [ny][nkr]


def _naive_irfft2(s, nx):
    """c2c inverse along y first, then a c2r along x that IGNORES Im of the kr = 0 and kr = nx/2 bins; factor 1/(nx ny)."""
    ny, nkr = s.shape
    jy = np.arange(ny)
    t = np.exp(2j * np.pi * np.outer(jy, jy) / ny) @ s                      # inverse along y, unnormalised
    x = np.arange(nx)
    out = np.zeros((ny, nx))
    for k in range(nkr):
        ph = np.exp(2j * np.pi * k * x / nx)
        if k == 0 or k == nx // 2:
            out += np.outer(t[:, k].real, ph.real)
        else:
            out += 2 * (np.outer(t[:, k], ph)).real
    return out / (nx * ny)


def test_forward_and_inverse_conventions_against_a_naive_dft():
    rng = np.random.default_rng(0)
    nx, ny = 12, 10
    g = O.Grid((nx, ny), (2 * np.pi, 3.0))
    c = rng.standard_normal((ny, nx))
    s = O.rfft(g, c)
    assert np.abs(s - _naive_rfft2(c)).max() < 1e-12
    # a generic (non-Hermitian-consistent) spectral array: the imaginary parts of the DC / Nyquist columns must be ignored
    s2 = rng.standard_normal((ny, nx // 2 + 1)) + 1j * rng.standard_normal((ny, nx // 2 + 1))
    # make the y-direction consistent with a real field where the c2r semantics are library-independent
    back = O.irfft(g, O.rfft(g, c) * (1 + 0j))
    assert np.abs(back - c).max() < 1e-13
    assert np.abs(O.irfft(g, s) - _naive_irfft2(s, nx)).max() < 1e-13
    # whatever is purely imaginary in the kr = 0 / kr = nx/2 columns AFTER the inverse y-transform is dropped
    # by the final c2r.  i*fft_y(r) with r real becomes i*r after the inverse along y.
    pert = np.zeros_like(s)
    pert[:, 0] = 1j * np.fft.fft(rng.standard_normal(ny))
    pert[:, nx // 2] = 1j * np.fft.fft(rng.standard_normal(ny))
    assert np.abs(O.irfft(g, s + pert) - O.irfft(g, s)).max() < 1e-13
    assert np.abs(_naive_irfft2(s + pert, nx) - _naive_irfft2(s, nx)).max() < 1e-13
    _ = s2


def test_wavenumbers_and_derivative_against_the_analytic_derivative():
    nx, ny, L = 16, 12, (2 * np.pi, 3.0)
    g = O.Grid((nx, ny), L)
    X, Y = g.gridpoints()
    ky = 2 * np.pi / L[1]
    f = np.sin(3 * X) * np.cos(2 * ky * Y)
    fx = O.irfft(g, 1j * g.kgrid(0) * O.rfft(g, f))
    fy = O.irfft(g, 1j * g.kgrid(1) * O.rfft(g, f))
    assert np.abs(fx - 3 * np.cos(3 * X) * np.cos(2 * ky * Y)).max() < 1e-13
    assert np.abs(fy + 2 * ky * np.sin(3 * X) * np.sin(2 * ky * Y)).max() < 1e-13
    assert g.k[1][ny // 2] == -(ny // 2) * ky                               # Nyquist wavenumber negative (fftfreq)


def test_second_fft_library_gives_the_same_rk4_step(monkeypatch):
    torch = pytest.importorskip("torch")
    n, L = (64, 48), (2 * np.pi, 4.0)
    g = O.Grid(n, L)
    X, Y = g.gridpoints()
    u = 0.2 * np.cos(X) * np.sin(2 * np.pi / 4.0 * Y)
    v = -0.3 * np.sin(X) * np.cos(2 * np.pi / 4.0 * Y)
    c0 = np.exp(-((X - 0.3) ** 2 + Y ** 2) / 0.2)

    def run():
        o = O.OracleProblem(n=n, L=L, kappa=(0.01, 0.02), dt=5e-3, stepper="RK4", velocity=[u, v], steady=True)
        o.set_c(c0)
        o.stepforward(3)
        return o.updatevars().copy()
    a = run()
    monkeypatch.setattr(O, "rfft", lambda grid, c, workers=None: torch.fft.rfftn(torch.from_numpy(np.ascontiguousarray(c)),
                                                                               dim=grid.axes).numpy())
    monkeypatch.setattr(O, "irfft", lambda grid, s, workers=None: torch.fft.irfftn(torch.from_numpy(np.ascontiguousarray(s)),
                                                                                 s=tuple(grid.pshape), dim=grid.axes).numpy())
    b = run()
    assert O.rel_l2(a, b) < 1e-14
