"""CPU ORACLE for the TracerAdvectionDiffusion hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy (fp64) restatement of the reference algorithm.  It is *not*
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``passivetracerflows.jl_b200``) never imports anything from ``oracle/``.

What it restates (all citations relative to /root/reference; TAD.jl =
src/traceradvectiondiffusion.jl):

* grids / wavenumbers      FourierFlows 0.10.5 ``src/domains.jl`` (un-vendored dependency,
                           pinned by Project.toml:24); restated from its published behaviour:
                           ``x = -L/2 + i*L/n``, ``kr = rfftfreq``, ``l,m = fftfreq`` (negative Nyquist).
* linear operator ``L``    TAD.jl:502-566
* nonlinear term calcN!    TAD.jl:695-804 (7 methods: time-varying 1/2/3-D, steady 1/2/3-D, MQG-layered)
* set_c! / updatevars!     TAD.jl:815-872
* RK4 / FilteredRK4 /      FourierFlows 0.10.5 ``src/timesteppers.jl`` (un-vendored): restated from the
  ETDRK4 / ForwardEuler /  published algorithm (classical RK4 with the linear term added per stage,
  AB3 / LSRK54 (+Filtered) Cox-Matthews ETDRK4 with 32-point contour coefficients, Carpenter-Kennedy LSRK54).
* makefilter               FourierFlows 0.10.5 ``src/utils.jl`` (un-vendored)
* dealias!                 FourierFlows 0.10.5 ``src/domains.jl`` (un-vendored; *not* called by TAD.jl -
                           opt-in add-on, default off)

PARITY PINNING.  The reference ships no golden vectors.  The oracle is pinned by the
reference's own 14 known-answer (analytic-solution) tests, test/test_traceradvectiondiffusion.jl
with the parameters of test/runtests.jl:26-54, re-expressed in tests/test_oracle_reference_kat.py
with the reference's tolerances.  Those pin: RK4 + linear term, frozen ``clock.t`` velocities
(TAD.jl:701,718,737), ``L`` including hyperdiffusion, the layered path, the grid origin and FFT
normalisation.  PARITY UNPINNED (no reference test constrains them; restated from the upstream
packages' published algorithms): FilteredRK4 / ETDRK4 / AB3 / LSRK54 on the tracer, the filter
shape, dealias!, the sign of the y/z Nyquist wavenumber, non-zero MQG flow coupling.

Array convention: NumPy C-order arrays whose *last* axis is x, i.e. physical ``(B?, nz?, ny?, nx)``
and spectral ``(B?, nz?, ny?, nx//2+1)``.  This is byte-identical to Julia's column-major
``(nx, ny, nz, B)`` used by the reference (TAD.jl:646-665,678-679).
"""
from __future__ import annotations

import math
import os

import numpy as np

try:  # threaded pocketfft when available (same algorithm family as numpy.fft)
    import scipy.fft as _fft
    _HAVE_SCIPY = True
except Exception:  # pragma: no cover
    import numpy.fft as _fft
    _HAVE_SCIPY = False

STEPPERS = ("ForwardEuler", "RK4", "ETDRK4", "LSRK54", "AB3",
            "FilteredForwardEuler", "FilteredRK4", "FilteredETDRK4", "FilteredLSRK54", "FilteredAB3")


# --------------------------------------------------------------------------------------
# Grid (FourierFlows src/domains.jl: OneDGrid / TwoDGrid / ThreeDGrid)
# --------------------------------------------------------------------------------------
class Grid:
    """Periodic grid. ``n`` and ``L`` are given in (x, y, z) order; arrays are stored x-fastest."""

    def __init__(self, n, L, nyquist_sign=-1):
        self.n = tuple(int(v) for v in n)
        self.L = tuple(float(v) for v in L)
        self.ndim = len(self.n)
        assert 1 <= self.ndim <= 3 and len(self.L) == self.ndim
        for v in self.n:
            assert v % 2 == 0 and v >= 2, "FourierFlows grids need even n"
        self.d = tuple(Lv / nv for nv, Lv in zip(self.n, self.L))
        # x = range(x0, step=dx, length=nx), x0 = -Lx/2
        self.coords = tuple(-Lv / 2 + dv * np.arange(nv) for nv, Lv, dv in zip(self.n, self.L, self.d))
        nx, Lx = self.n[0], self.L[0]
        self.nkr = nx // 2 + 1
        # kr = rfftfreq(nx, 2pi/Lx*nx): 0..nx/2 (Nyquist positive)
        self.kr = np.arange(self.nkr) * (2 * np.pi / Lx * nx) / nx
        self.k = [self.kr]
        for a in range(1, self.ndim):
            na, La = self.n[a], self.L[a]
            j = np.arange(na)
            j = np.where(j < na // 2, j, j - na).astype(np.float64)  # fftfreq: Nyquist negative
            ka = j * (2 * np.pi / La * na) / na
            if nyquist_sign > 0:
                ka[na // 2] = -ka[na // 2]
            self.k.append(ka)
        self.pshape = tuple(reversed(self.n))                    # (nz, ny, nx)
        self.sshape = tuple(reversed(self.n[1:])) + (self.nkr,)  # (nz, ny, nkr)
        self.axes = tuple(range(-self.ndim, 0))
        self.npts = int(np.prod(self.n))

    def kgrid(self, a):
        """Wavenumber of axis ``a`` (0=x) broadcastable against a spectral array."""
        shp = [1] * self.ndim
        shp[self.ndim - 1 - a] = len(self.k[a])
        return self.k[a].reshape(shp)

    def gridpoints(self):
        """Full coordinate arrays (X[, Y[, Z]]) in physical layout (FourierFlows ``gridpoints``)."""
        out = []
        for a in range(self.ndim):
            shp = [1] * self.ndim
            shp[self.ndim - 1 - a] = self.n[a]
            out.append(np.broadcast_to(self.coords[a].reshape(shp), self.pshape))
        return tuple(out)

    @property
    def Krsq(self):
        s = self.kgrid(0) ** 2
        for a in range(1, self.ndim):
            s = s + self.kgrid(a) ** 2
        return s


def rfft(grid, c, workers=None):
    """Unnormalised forward r2c over the grid axes (FourierFlows ``mul!(out, rfftplan, in)``)."""
    if _HAVE_SCIPY:
        return _fft.rfftn(c, axes=grid.axes, workers=workers)
    return _fft.rfftn(c, axes=grid.axes)


def irfft(grid, s, workers=None):
    """1/N-normalised inverse c2r, x-axis last (FourierFlows ``ldiv!(out, rfftplan, in)``)."""
    shp = tuple(grid.pshape)
    if _HAVE_SCIPY:
        return _fft.irfftn(s, s=shp, axes=grid.axes, workers=workers)
    return _fft.irfftn(s, s=shp, axes=grid.axes)


# --------------------------------------------------------------------------------------
# Linear operator (TAD.jl:502-566), filter (FourierFlows utils.jl makefilter), dealias mask
# --------------------------------------------------------------------------------------
def _ipow(x, n):
    """x^n by repeated multiplication, x^0 == 1 (also for 0^0)."""
    out = np.ones_like(x)
    for _ in range(int(n)):
        out = out * x
    return out


def linear_operator(grid, kappa, kappa_h=0.0, n_kappa_h=0):
    """``L = -kappa*kr^2 - eta*l^2 - iota*m^2 - kappa_h*Krsq^n_kappa_h`` (TAD.jl:506,515,524,533,542,551,562).

    ``kappa`` is a length-ndim sequence (kappa, eta, iota).  In 1-D the hyper term uses ``(kr^2)^n``
    which equals ``Krsq^n`` there.
    """
    L = -kappa[0] * grid.kgrid(0) ** 2
    for a in range(1, grid.ndim):
        L = L - kappa[a] * grid.kgrid(a) ** 2
    L = L - kappa_h * _ipow(grid.Krsq, n_kappa_h)
    return np.ascontiguousarray(np.broadcast_to(L, grid.sshape))


def make_filter(grid, order=4, innerK=2.0 / 3.0, outerK=1.0, tol=1e-15):
    """FourierFlows ``makefilter(grid)``: 1 for K<innerK else exp(-decay*(K-innerK)^order)."""
    Ksq = 0.0
    for a in range(grid.ndim):
        Ksq = Ksq + (grid.kgrid(a) * grid.d[a] / np.pi) ** 2
    K = np.sqrt(Ksq)
    decay = -math.log(tol) / (outerK - innerK) ** order
    filt = np.exp(-decay * (K - innerK) ** order)
    filt = np.where(K < innerK, 1.0, filt)
    return np.ascontiguousarray(np.broadcast_to(filt, grid.sshape))


def aliased_ranges(n, nk, aliased_fraction, r2c):
    """0-based [lo, hi) index range FourierFlows ``getaliasedwavenumbers`` zeroes on one axis."""
    iL = int(math.floor((1 - aliased_fraction) / 2 * n)) + 1     # 1-based inclusive
    iR = int(math.ceil((1 + aliased_fraction) / 2 * n))          # 1-based inclusive
    if aliased_fraction <= 0:   # FF dealias!: `grid.aliased_fraction == 0 && return nothing` — no mode is zeroed
        return (0, 0)           # [UPSTREAM-RECALLED, FourierFlows 0.10 src/domains.jl; ADVICE r01]
    if r2c:
        return (iL - 1, nk)
    return (iL - 1, iR)


def dealias_mask(grid, aliased_fraction=1.0 / 3.0):
    """1/0 mask equal to FourierFlows ``dealias!`` (per-axis box truncation)."""
    mask = np.ones(grid.sshape)
    for a in range(grid.ndim):
        nk = grid.nkr if a == 0 else grid.n[a]
        lo, hi = aliased_ranges(grid.n[a], nk, aliased_fraction, a == 0)
        idx = [slice(None)] * grid.ndim
        idx[grid.ndim - 1 - a] = slice(lo, hi)
        mask[tuple(idx)] = 0.0
    return mask


def etd_coefficients(dt, L, ncirc=32, rcirc=1.0):
    """FourierFlows ``getetdcoeffs``: contour-mean ETDRK4 coefficients (zeta, alpha, beta, Gamma), real."""
    circ = rcirc * np.exp(2j * np.pi / ncirc * (np.arange(ncirc) + 0.5))
    zc = dt * L[..., None] + circ
    ez = np.exp(zc)
    zeta = dt * np.mean((np.exp(zc / 2) - 1) / zc, axis=-1).real
    alpha = dt * np.mean((-4 - zc + ez * (4 - 3 * zc + zc ** 2)) / zc ** 3, axis=-1).real
    beta = dt * np.mean((2 + zc + ez * (-2 + zc)) / zc ** 3, axis=-1).real
    gamma = dt * np.mean((-4 - 3 * zc - zc ** 2 + ez * (4 - zc)) / zc ** 3, axis=-1).real
    return zeta, alpha, beta, gamma


# Carpenter & Kennedy (1994) 5-stage 4th-order low-storage RK, as used by FourierFlows LSRK54TimeStepper
LSRK54_A = (0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
            -3550918686646 / 2091501179385, -1275806237668 / 842570457699)
LSRK54_B = (1432997174477 / 9575080441755, 5161836677717 / 13612068292357, 1720146321549 / 2090206949498,
            3134564353537 / 4481467310338, 2277821191437 / 14882151754819)
LSRK54_C = (0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
            2006345519317 / 3224310063776, 2802321613138 / 2924317926251)


# --------------------------------------------------------------------------------------
# The problem: params + vars + equation + timestepper + clock in one object
# --------------------------------------------------------------------------------------
class OracleProblem:
    """Restatement of ``TracerAdvectionDiffusion.Problem`` + ``FourierFlows.Problem`` state.

    velocity:
      * steady:        sequence of ndim arrays (physical layout, optionally with a leading batch axis)
                       - what TAD.jl:426-452 pre-evaluates on ``gridpoints``.
      * time-varying:  sequence of ndim callables ``f(x[,y[,z]], t)`` evaluated on ``gridpoints`` at
                       ``clock.t`` (NOT the stage time: TAD.jl:701,718,737).
      * layered (MQG): call :meth:`set_layered_velocity` with ``u, v`` of shape (B, ny, nx) and ``U`` of
                       shape (B, ny, 1) or (B,) before stepping (TAD.jl:795-796).
    """

    def __init__(self, n, L, kappa, dt, stepper="RK4", velocity=None, steady=True, nbatch=1,
                 kappa_h=0.0, n_kappa_h=0, dealias=False, aliased_fraction=1.0 / 3.0,
                 nyquist_sign=-1, workers=None):
        self.grid = Grid(n, L, nyquist_sign)
        g = self.grid
        kappa = tuple(np.atleast_1d(kappa).astype(float))
        if len(kappa) < g.ndim:  # eta = kappa, iota = kappa defaults (TAD.jl:171,198-199)
            kappa = kappa + (kappa[0],) * (g.ndim - len(kappa))
        self.kappa = kappa
        if stepper not in STEPPERS:
            raise ValueError(f"unknown stepper {stepper!r}")
        self.stepper = stepper
        self.filtered = stepper.startswith("Filtered")
        self.base = stepper[len("Filtered"):] if self.filtered else stepper
        self.dt = float(dt)
        self.t = 0.0
        self.step = 0
        self.nbatch = int(nbatch)
        self.workers = workers if workers is not None else (os.cpu_count() or 1)
        self.batched = self.nbatch > 1
        bs = (self.nbatch,) if self.batched else ()
        self.Lop = linear_operator(g, kappa, kappa_h, n_kappa_h)          # broadcast over batch (TAD.jl:561-563)
        self.filter = make_filter(g) if self.filtered else None
        self.mask = dealias_mask(g, aliased_fraction) if dealias else None
        self.sol = np.zeros(bs + g.sshape, dtype=np.complex128)
        self.c = np.zeros(bs + g.pshape)
        self.steady = steady
        self.vel_arrays = None
        self.vel_funcs = None
        if velocity is None or (isinstance(velocity, str) and velocity == "layered"):
            self.vel_arrays = [np.zeros(g.pshape) for _ in range(g.ndim)]       # noflow (TAD.jl:31)
        elif steady:
            self.vel_arrays = [np.asarray(v, dtype=float) for v in velocity]
        else:
            self.vel_funcs = list(velocity)
        self._ik = [1j * g.kgrid(a) for a in range(g.ndim)]
        self._init_stepper()

    # ---------------- stepper state (FourierFlows timesteppers.jl) ----------------
    def _init_stepper(self):
        if self.base == "ETDRK4":
            self.expLdt = np.exp(self.dt * self.Lop)
            self.expLdt2 = np.exp(self.dt * self.Lop / 2)
            self.zeta, self.alpha, self.beta, self.gamma = etd_coefficients(self.dt, self.Lop)
        if self.base == "AB3":
            self._rhs_m1 = None
            self._rhs_m2 = None

    # ---------------- velocities ----------------
    def set_layered_velocity(self, u, v, U=None):
        """MQG coupling: ``u_total = vars.u + params.U`` (TAD.jl:795), ``v = vars.v`` (TAD.jl:796)."""
        u = np.asarray(u, dtype=float)
        if U is not None:
            U = np.asarray(U, dtype=float)
            if U.ndim == 1:
                U = U.reshape(-1, 1, 1)
            u = u + U
        self.vel_arrays = [u, np.asarray(v, dtype=float)]
        self.steady = True

    def _velocities(self):
        if self.vel_funcs is not None:
            pts = self.grid.gridpoints()
            out = []
            for f in self.vel_funcs:
                out.append(np.broadcast_to(np.asarray(f(*pts, self.t), dtype=float), self.grid.pshape))
            return out
        return self.vel_arrays

    # ---------------- nonlinear term (TAD.jl:695-804) ----------------
    def calcN(self, s, vel):
        g = self.grid
        if self.mask is not None:   # opt-in dealias!(sol) at the top of calcN (GeophysicalFlows convention)
            s *= self.mask
        p = None
        for a in range(g.ndim):
            ga = irfft(g, self._ik[a] * s, self.workers)       # TAD.jl:757-761
            term = vel[a] * ga
            p = -term if p is None else p - term               # TAD.jl:764: -u*cx - v*cy - w*cz
        return rfft(g, p, self.workers)                        # TAD.jl:766

    def _rhs(self, s, vel):
        return self.calcN(s, vel) + self.Lop * s               # addlinearterm!

    # ---------------- steppers ----------------
    def stepforward(self, nsteps=1):
        for _ in range(int(nsteps)):
            vel = self._velocities()        # frozen at clock.t for all stages (TAD.jl:701,718,737)
            getattr(self, "_step_" + self.base)(vel)
            self.t += self.dt
            self.step += 1

    def _finish(self, new):
        self.sol = self.filter * new if self.filtered else new

    def _step_ForwardEuler(self, vel):
        s0 = self.sol
        self._finish(s0 + self.dt * self._rhs(s0, vel))

    def _step_RK4(self, vel):
        dt, s0 = self.dt, self.sol
        k1 = self._rhs(s0, vel)
        k2 = self._rhs(s0 + (dt / 2) * k1, vel)
        k3 = self._rhs(s0 + (dt / 2) * k2, vel)
        k4 = self._rhs(s0 + dt * k3, vel)
        self._finish(s0 + dt * (k1 / 6 + k2 / 3 + k3 / 3 + k4 / 6))

    def _step_ETDRK4(self, vel):
        s0 = self.sol
        E, E2, z = self.expLdt, self.expLdt2, self.zeta
        N1 = self.calcN(s0, vel)
        a = E2 * s0 + z * N1
        N2 = self.calcN(a, vel)
        b = E2 * s0 + z * N2
        N3 = self.calcN(b, vel)
        b = E2 * a + z * (2 * N3 - N1)
        N4 = self.calcN(b, vel)
        self._finish(E * s0 + self.alpha * N1 + 2 * self.beta * (N2 + N3) + self.gamma * N4)

    def _step_LSRK54(self, vel):
        s = self.sol
        S2 = np.zeros_like(s)
        for i in range(5):
            rhs = self._rhs(s, vel)
            S2 = LSRK54_A[i] * S2 + self.dt * rhs
            s = s + LSRK54_B[i] * S2
        self._finish(s)

    def _step_AB3(self, vel):
        s0 = self.sol
        rhs = self._rhs(s0, vel)
        if self.step < 3 or self._rhs_m1 is None or self._rhs_m2 is None:
            new = s0 + self.dt * rhs                       # Euler start-up while clock.step < 3
        else:
            new = s0 + self.dt * (23 / 12 * rhs - 16 / 12 * self._rhs_m1 + 5 / 12 * self._rhs_m2)
        self._rhs_m2, self._rhs_m1 = self._rhs_m1, rhs
        self._finish(new)

    # ---------------- set / get (TAD.jl:815-872) ----------------
    def set_c(self, c):
        c = np.asarray(c, dtype=float)
        if self.batched and c.ndim == self.grid.ndim:
            c = np.broadcast_to(c, (self.nbatch,) + c.shape)   # repeat over layers (TAD.jl:865)
        self.sol = rfft(self.grid, c, self.workers)            # TAD.jl:847,866
        self.updatevars()

    def updatevars(self):
        self.c = irfft(self.grid, self.sol.copy(), self.workers)   # TAD.jl:816-818
        return self.c


def rel_l2(a, b):
    """Julia ``isapprox`` metric on arrays: ||a-b||_2 / max(||a||_2, ||b||_2) (test/...:32)."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    den = max(np.linalg.norm(a), np.linalg.norm(b))
    return float(np.linalg.norm(a - b) / den) if den > 0 else 0.0
