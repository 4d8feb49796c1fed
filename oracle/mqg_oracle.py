"""CPU ORACLE for the MultiLayerQG flow solver that drives the MQG-coupled tracer.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file; the product path
(``passivetracerflows.jl_b200``) never does.

What it restates.  The reference couples the tracer to ``GeophysicalFlows.MultiLayerQG`` (an UN-VENDORED dependency,
pinned ``GeophysicalFlows = "0.16"`` by /root/reference/Project.toml:25).  The reference's own call sites are
  * ``MultiLayerQG.Problem(nlayers, dev; nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction=0)``
    examples/turbulent_advection-diffusion.jl:56-58,
  * ``MultiLayerQG.set_q!``  examples/…:64-69,   ``step_until!(MQGprob, t)``  TAD.jl:238,
  * ``MultiLayerQG.updatevars!(MQGprob)``  TAD.jl:488 and examples/…:151,  ``stepforward!(params.MQGprob)`` examples/…:150,
  * ``MQGprob.vars.u, .v`` and ``MQGprob.params.U`` read by the tracer's calcN!  TAD.jl:795-796,
  * ``streamfunctionfrompv!`` / ``invtransform!``  examples/…:110-118.
GeophysicalFlows' sources are not on this machine, so the algorithm below restates the package's published
equations (its docs "MultiLayerQG module": ``∂ₜq_j + J(ψ_j, q_j + Q_j) + U_j ∂ₓq_j = −ν(−∇²)^{nν} q_j − δ_{jn} μ ∇²ψ_n``
with ``q̂ = S ψ̂``, ``S = −|k|² I + F``, F the tridiagonal stretching matrix built from ``F_{j+½} = f₀²/(g′ H)``) in the
operation order of its ``calcN_advection!``:
    dealias!(sol);  ψ̂ = S⁻¹ q̂;  û = −i l ψ̂,  v̂ = i kr ψ̂;  u = irfft(û) + U;  v = irfft(v̂);  q = irfft(q̂)
    N̂ = −rfft(u·Qx) − rfft(v·Qy) − i kr rfft(u·q) − i l rfft(v·q);   N̂[bottom] += μ |k|² ψ̂[bottom]
and ``L = −ν |k|^{2nν}`` with ``L[0,0] = 0``.

PARITY UNPINNED: the reference has one test through this solver (``test_diffusion_multilayerqg``,
test/test_traceradvectiondiffusion.jl, run with a zero flow) and ships no golden vectors, so nothing in the reference
constrains the flow solver itself.  What pins this restatement instead (tests/test_oracle_mqg.py): analytic linear
Rossby-wave dispersion relations (barotropic with a Doppler shift, baroclinic with the stretching term — they fix the
signs of β, U, S and the rfft conventions), ``S·S⁻¹ = I``, exact conservation of the domain mean, and energy / enstrophy
conservation of the unforced inviscid nonlinear problem to time-stepper accuracy.
"""
from __future__ import annotations

import os

import numpy as np

from .ptf_oracle import Grid, OracleProblem, dealias_mask, irfft, make_filter, rfft, _ipow, STEPPERS


class MQGParams:
    """``MultiLayerQG.Params``: background PV gradients, stretching matrix and its inverse."""

    def __init__(self, grid, nlayers, f0=1.0, beta=0.0, H=None, b=None, U=None, eta=None,
                 topographic_pv_gradient=(0.0, 0.0), mu=0.0, nu=0.0, nnu=1):
        nl = int(nlayers)
        nx, ny = grid.n
        self.nlayers = nl
        self.f0, self.beta, self.mu, self.nu, self.nnu = float(f0), float(beta), float(mu), float(nu), int(nnu)
        self.H = np.full(nl, 1.0 / nl) if H is None else np.asarray(H, dtype=float).reshape(nl)
        self.b = -(1.0 + np.arange(nl) / nl) if b is None else np.asarray(b, dtype=float).reshape(nl)
        U = np.zeros(nl) if U is None else np.asarray(U, dtype=float)
        # U is either one value per layer or a profile U(y) per layer; stored as (nl, ny, 1)
        if U.ndim == 1:
            U = np.broadcast_to(U.reshape(nl, 1, 1), (nl, ny, 1)).copy()
        else:
            U = U.reshape(nl, ny, 1).copy()
        self.U = U
        l = grid.k[1].reshape(1, ny, 1)
        Uyy = np.real(np.fft.ifft(-l ** 2 * np.fft.fft(U, axis=1), axis=1))
        if eta is None:
            etax = np.zeros((ny, nx))
            etay = np.zeros((ny, nx))
        else:
            etah = rfft(grid, np.asarray(eta, dtype=float))
            etax = irfft(grid, 1j * grid.kgrid(0) * etah)
            etay = irfft(grid, 1j * grid.kgrid(1) * etah)
        etax = etax + topographic_pv_gradient[0]
        etay = etay + topographic_pv_gradient[1]
        Qx = np.zeros((nl, ny, nx))
        Qx[nl - 1] += etax
        Qy = np.broadcast_to(self.beta - Uyy, (nl, ny, nx)).copy()
        Qy[nl - 1] += etay
        Krsq = grid.Krsq                      # (ny, nkr)
        if nl >= 2:
            gp = self.b[:-1] - self.b[1:]                     # reduced gravity at each interface
            self.gp = gp
            self.Fm = self.f0 ** 2 / (gp * self.H[1:])
            self.Fp = self.f0 ** 2 / (gp * self.H[:-1])
            F = np.zeros((nl, nl))
            for j in range(nl - 1):
                F[j + 1, j] = self.Fm[j]
                F[j, j + 1] = self.Fp[j]
            diag = -(np.concatenate([self.Fp, [0.0]]) + np.concatenate([[0.0], self.Fm]))
            F[np.arange(nl), np.arange(nl)] = diag
            self.F = F
            eye = np.eye(nl)
            self.S = -Krsq[..., None, None] * eye + F                      # (ny, nkr, nl, nl)
            k2 = np.where(Krsq == 0, 1.0, Krsq)
            self.Sinv = np.linalg.inv(-k2[..., None, None] * eye + F)
            self.Sinv[0, 0] = 0.0
            Qy[0] = Qy[0] - self.Fp[0] * (U[1] - U[0])
            for j in range(1, nl - 1):
                Qy[j] = Qy[j] - self.Fp[j] * (U[j + 1] - U[j]) - self.Fm[j - 1] * (U[j - 1] - U[j])
            Qy[nl - 1] = Qy[nl - 1] - self.Fm[nl - 2] * (U[nl - 2] - U[nl - 1])
        else:
            self.F = np.zeros((1, 1))
            self.S = -Krsq[..., None, None] * np.eye(1)
            inv = np.where(Krsq == 0, 0.0, -1.0 / np.where(Krsq == 0, 1.0, Krsq))
            self.Sinv = inv[..., None, None] * np.eye(1)
        self.Qx, self.Qy = Qx, Qy


class MQGOracle(OracleProblem):
    """``MultiLayerQG.Problem`` + its FourierFlows stepper.  Arrays: physical (nlayers, ny, nx), spectral (nlayers, ny, nkr)."""

    def __init__(self, nlayers, nx=128, ny=None, Lx=2 * np.pi, Ly=None, f0=1.0, beta=0.0, U=None, H=None, b=None,
                 eta=None, topographic_pv_gradient=(0.0, 0.0), mu=0.0, nu=0.0, nnu=1, dt=0.01, stepper="RK4",
                 aliased_fraction=1.0 / 3.0, workers=None):
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        self.grid = Grid((nx, ny), (Lx, Ly))
        g = self.grid
        if stepper not in STEPPERS:
            raise ValueError(f"unknown stepper {stepper!r}")
        self.params = MQGParams(g, nlayers, f0, beta, H, b, U, eta, topographic_pv_gradient, mu, nu, nnu)
        self.nlayers = int(nlayers)
        self.stepper = stepper
        self.filtered = stepper.startswith("Filtered")
        self.base = stepper[len("Filtered"):] if self.filtered else stepper
        self.dt, self.t, self.step = float(dt), 0.0, 0
        self.workers = workers if workers is not None else (os.cpu_count() or 1)
        self.nbatch, self.batched = self.nlayers, True
        Lop = -self.params.nu * _ipow(g.Krsq, self.params.nnu)            # hyperviscosity
        Lop = np.array(np.broadcast_to(Lop, g.sshape))
        Lop[0, 0] = 0.0
        self.Lop = Lop
        self.filter = make_filter(g) if self.filtered else None
        self.mask = dealias_mask(g, aliased_fraction)
        self.sol = np.zeros((self.nlayers,) + g.sshape, dtype=np.complex128)
        self.vel_funcs = None
        self.vel_arrays = None
        self._ikr = 1j * g.kgrid(0)
        self._il = 1j * g.kgrid(1)
        self._init_stepper()
        self.updatevars()

    def _velocities(self):
        return None

    def streamfunctionfrompv(self, qh):
        # psih[j] = sum_m Sinv[j, m] qh[m] per wavenumber
        return np.einsum("yxjm,myx->jyx", self.params.Sinv, qh)

    def pvfromstreamfunction(self, psih):
        return np.einsum("yxjm,myx->jyx", self.params.S, psih)

    def calcN(self, s, vel=None):
        g, p, w = self.grid, self.params, self.workers
        s *= self.mask                                            # dealias!(sol, grid) — in place on the stage state
        psih = self.streamfunctionfrompv(s)
        u = irfft(g, -self._il * psih, w) + p.U
        v = irfft(g, self._ikr * psih, w)
        q = irfft(g, s.copy(), w)
        N = -rfft(g, u * p.Qx, w)
        N = N - rfft(g, v * p.Qy, w)
        N = N - (self._ikr * rfft(g, u * q, w) + self._il * rfft(g, v * q, w))
        N[-1] = N[-1] + p.mu * g.Krsq * psih[-1]                  # bottom linear drag
        return N

    def set_q(self, q):
        qh = rfft(self.grid, np.asarray(q, dtype=float), self.workers)
        qh[:, 0, 0] = 0.0
        self.sol = qh
        self.updatevars()

    def set_psi(self, psi):
        psih = rfft(self.grid, np.asarray(psi, dtype=float), self.workers)
        qh = self.pvfromstreamfunction(psih)
        self.set_q(irfft(self.grid, qh, self.workers))

    def updatevars(self):
        g, w = self.grid, self.workers
        self.sol *= self.mask
        self.psih = self.streamfunctionfrompv(self.sol)
        self.q = irfft(g, self.sol.copy(), w)
        self.psi = irfft(g, self.psih.copy(), w)
        self.u = irfft(g, -self._il * self.psih, w)               # perturbation velocity: U is NOT included
        self.v = irfft(g, self._ikr * self.psih, w)
        return self.q

    def step_until(self, t_stop):
        """FourierFlows ``step_until!``: whole steps, then one shortened step (explicit steppers only)."""
        interval = t_stop - self.t
        nsteps = int(np.floor(interval / self.dt))
        self.stepforward(nsteps)
        rem = interval - nsteps * self.dt
        if rem > 0:
            dt0 = self.dt
            self.dt = rem
            self._init_stepper()
            self.stepforward(1)
            self.dt = dt0
            self._init_stepper()
        self.t = float(t_stop)

    # diagnostics used by the conservation tests (GeophysicalFlows ``energies`` up to normalisation)
    def energy(self):
        p, g = self.params, self.grid
        psih = self.streamfunctionfrompv(self.sol * self.mask)
        w = np.full(g.sshape, 2.0)
        w[:, 0] = 1.0
        if g.n[0] % 2 == 0:
            w[:, -1] = 1.0
        ke = sum(p.H[j] * np.sum(w * g.Krsq * np.abs(psih[j]) ** 2) for j in range(self.nlayers))
        pe = 0.0
        for j in range(self.nlayers - 1):
            pe += p.f0 ** 2 / p.gp[j] * np.sum(w * np.abs(psih[j] - psih[j + 1]) ** 2)
        return 0.5 * (ke + pe) / (g.npts ** 2 * np.sum(p.H))
