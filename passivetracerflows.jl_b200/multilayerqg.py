"""Host-side mirror of ``GeophysicalFlows.MultiLayerQG`` for the ``B200`` device — the flow that advects the tracer of
``TracerAdvectionDiffusion.Problem(MQGprob; ...)`` (TAD.jl:225-250).

Only the surface the reference touches is mirrored (examples/turbulent_advection-diffusion.jl:56-69,110-118,149-151,
TAD.jl:238,488,795-796): ``Problem(nlayers, dev; nx, Lx, f₀, H, b, U, μ, β, dt, stepper, aliased_fraction)``, ``set_q``,
``stepforward``, ``step_until``, ``updatevars``, ``prob.vars.{q, ψ, u, v}``, ``prob.params.{U, nlayers}``,
``prob.clock``, ``prob.sol``.  All numerics run in libptf_b200.so (``ptf_mqg_*`` in include/ptf_b200.h); the state stays
on the device and ``vars`` are host mirrors fetched on demand.

Array convention: NumPy C-order (nlayers, ny, nx) / (nlayers, ny, nkr) — byte-identical to Julia's (nx, ny, nlayers).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _capi
from .tracer_advection_diffusion import B200, Clock, Equation, Grid, TimeStepper, _parse_stepper


@dataclass
class MQGParams:
    nlayers: int
    f0: float
    beta: float
    H: np.ndarray
    b: np.ndarray
    U: np.ndarray       # (nlayers,) or (nlayers, ny)
    mu: float
    nu: float
    nnu: int
    Qx: np.ndarray = None
    Qy: np.ndarray = None


class _Vars:
    """``MQGprob.vars``: q, ψ, u, v as of the last ``updatevars`` — read back from the device when accessed."""

    def __init__(self, prob):
        self._p = prob
        self._cache = {}

    def _invalidate(self):
        self._cache.clear()

    def _get(self, which):
        if which not in self._cache:
            out = np.empty(self._p._pshape)
            _capi.check_mqg(self._p._lib.ptf_mqg_get_var(self._p._h, which, _capi.as_dp(out)), self._p._h)
            self._cache[which] = out
        return self._cache[which]

    u = property(lambda self: self._get(0))     # perturbation velocity; params.U is NOT included (TAD.jl:795 adds it)
    v = property(lambda self: self._get(1))
    q = property(lambda self: self._get(2))
    psi = property(lambda self: self._get(3))


class MultiLayerQGProblem:
    """What ``MultiLayerQG.Problem(nlayers, B200(); ...)`` returns."""

    def __init__(self, nlayers, dev=None, *, nx=128, ny=None, Lx=2 * math.pi, Ly=None, f0=1.0, beta=0.0, U=None, H=None,
                 b=None, eta=None, topographic_pv_gradient=(0.0, 0.0), mu=0.0, nu=0.0, nnu=1, dt=0.01, stepper="RK4",
                 aliased_fraction=1.0 / 3.0, T=np.float64):
        dev = dev or B200()
        if not isinstance(dev, B200):
            raise TypeError("this package implements the B200 device only (no CPU path)")
        if T not in (np.float64, float, "Float64"):
            raise NotImplementedError("the B200 path is fp64")
        self._lib = _capi.load()
        self.dev = dev
        nl = int(nlayers)
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        H = np.full(nl, 1.0 / nl) if H is None else np.ascontiguousarray(H, dtype=np.float64).reshape(nl)
        b = -(1.0 + np.arange(nl) / nl) if b is None else np.ascontiguousarray(b, dtype=np.float64).reshape(nl)
        U = np.zeros(nl) if U is None else np.ascontiguousarray(U, dtype=np.float64)
        if U.shape not in ((nl,), (nl, ny)):
            raise ValueError(f"U must have shape ({nl},) or ({nl}, {ny})")
        self.grid = Grid(nx=nx, Lx=Lx, ny=ny, Ly=Ly, ndim=2, device=dev)
        self.params = MQGParams(nlayers=nl, f0=float(f0), beta=float(beta), H=H, b=b, U=U, mu=float(mu), nu=float(nu),
                                nnu=int(nnu))
        self.stepper = stepper
        self.timestepper = TimeStepper(stepper, self.grid)     # .filter is read at examples/…:64
        self.clock = Clock(dt=float(dt))
        d = _capi.PtfMqgDesc()
        self._lib.ptf_mqg_desc_init(C.byref(d))
        d.nlayers, d.nx, d.ny, d.Lx, d.Ly = nl, nx, ny, Lx, Ly
        d.f0, d.beta, d.mu, d.nu, d.n_nu, d.dt = float(f0), float(beta), float(mu), float(nu), int(nnu), float(dt)
        d.H, d.b, d.U = _capi.as_dp(H), _capi.as_dp(b), _capi.as_dp(U)
        d.U_is_profile = 1 if U.ndim == 2 else 0
        keep = [H, b, U]
        if eta is not None:
            eta = np.ascontiguousarray(eta, dtype=np.float64)
            if eta.shape != (ny, nx):
                raise ValueError(f"eta must have shape ({ny}, {nx})")
            d.eta = _capi.as_dp(eta)
            keep.append(eta)
        d.topographic_pv_gradient[:] = [float(topographic_pv_gradient[0]), float(topographic_pv_gradient[1])]
        d.stepper = _parse_stepper(stepper)
        d.aliased_fraction = float(aliased_fraction)
        d.device = dev.device
        d.use_graph = 1 if dev.use_graph else 0
        h = C.c_void_p()
        _capi.check_mqg(self._lib.ptf_mqg_create(C.byref(d), C.byref(h)))
        self._h = h
        self._desc_dt = float(dt)
        self._pshape = (nl, ny, nx)
        self._sshape = (nl, ny, nx // 2 + 1)
        self.vars = _Vars(self)
        self._tracers = []
        Qx, Qy = np.empty(self._pshape), np.empty(self._pshape)
        _capi.check_mqg(self._lib.ptf_mqg_get_background(h, _capi.as_dp(Qx), _capi.as_dp(Qy)), h)
        self.params.Qx, self.params.Qy = Qx, Qy

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ptf_mqg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def eqn(self) -> Equation:
        """Hyperviscosity ``L = −ν·Krsq^nν`` with ``L[0, 0] = 0``, one copy per layer."""
        g, p = self.grid, self.params
        Ksq = g.kr[None, :] ** 2 + g.l[:, None] ** 2
        hyper = np.ones_like(Ksq)
        for _ in range(p.nnu):
            hyper = hyper * Ksq
        L = -p.nu * hyper
        L[0, 0] = 0.0
        return Equation(L=np.ascontiguousarray(np.broadcast_to(L, self._sshape)), dims=self._sshape)

    # ---- state ----
    @property
    def sol(self):
        out = np.empty(self._sshape, dtype=np.complex128)
        _capi.check_mqg(self._lib.ptf_mqg_get_sol(self._h, out.ctypes.data_as(C.POINTER(C.c_double))), self._h)
        return out

    def set_sol(self, sol):
        sol = np.ascontiguousarray(sol, dtype=np.complex128)
        if sol.shape != self._sshape:
            raise ValueError(f"sol has shape {sol.shape}, expected {self._sshape}")
        _capi.check_mqg(self._lib.ptf_mqg_set_sol(self._h, sol.ctypes.data_as(C.POINTER(C.c_double))), self._h)
        self.vars._invalidate()

    def set_q(self, q):
        """``MultiLayerQG.set_q!(prob, q)``."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        if q.shape != self._pshape:
            raise ValueError(f"q has shape {q.shape}, expected {self._pshape}")
        _capi.check_mqg(self._lib.ptf_mqg_set_q(self._h, _capi.as_dp(q)), self._h)
        self.vars._invalidate()

    def set_psi(self, psi):
        """``MultiLayerQG.set_ψ!(prob, ψ)``."""
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        if psi.shape != self._pshape:
            raise ValueError(f"psi has shape {psi.shape}, expected {self._pshape}")
        _capi.check_mqg(self._lib.ptf_mqg_set_psi(self._h, _capi.as_dp(psi)), self._h)
        self.vars._invalidate()

    def updatevars(self):
        """``MultiLayerQG.updatevars!(prob)`` — on the device; host mirrors are refreshed lazily."""
        _capi.check_mqg(self._lib.ptf_mqg_updatevars(self._h), self._h)
        self.vars._invalidate()

    # ---- stepping ----
    def _sync_clock(self):
        t, s, dt = C.c_double(), C.c_int64(), C.c_double()
        _capi.check_mqg(self._lib.ptf_mqg_get_clock(self._h, C.byref(t), C.byref(s), C.byref(dt)), self._h)
        self.clock.t, self.clock.step, self.clock.dt = t.value, s.value, dt.value

    def _push_dt(self):
        if self.clock.dt != self._desc_dt:
            _capi.check_mqg(self._lib.ptf_mqg_set_dt(self._h, float(self.clock.dt)), self._h)
            self._desc_dt = self.clock.dt

    def stepforward(self, nsteps: int = 1):
        self._push_dt()
        _capi.check_mqg(self._lib.ptf_mqg_step(self._h, int(nsteps)), self._h)
        self._sync_clock()

    def step_until(self, stop_time: float):
        self._push_dt()
        _capi.check_mqg(self._lib.ptf_mqg_step_until(self._h, float(stop_time)), self._h)
        self._sync_clock()

    def step_timed(self, nsteps: int, with_updatevars: bool = False) -> float:
        ms = C.c_float()
        _capi.check_mqg(self._lib.ptf_mqg_step_timed(self._h, int(nsteps), 1 if with_updatevars else 0, C.byref(ms)),
                        self._h)
        self._sync_clock()
        self.vars._invalidate()
        return ms.value

    def launch_count(self):
        a, b = C.c_int64(), C.c_int64()
        _capi.check_mqg(self._lib.ptf_mqg_launch_count(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value


def Problem(nlayers, dev=None, **kwargs):
    """``MultiLayerQG.Problem(nlayers, dev; ...)`` (examples/turbulent_advection-diffusion.jl:56-58)."""
    return MultiLayerQGProblem(nlayers, dev, **kwargs)


def set_q(prob, q):
    prob.set_q(q)


def set_psi(prob, psi):
    prob.set_psi(psi)


def updatevars(prob):
    prob.updatevars()


def couple(tracer, mqg: MultiLayerQGProblem):
    """``ConstDiffTurbulentFlowParams(κ, η, tracer_release_time, MQGprob)`` (TAD.jl:485-491) for a device-resident
    flow: the tracer's calcN! reads ``MQGprob.vars.u .+ MQGprob.params.U`` and ``MQGprob.vars.v`` (TAD.jl:795-796)
    straight from the flow solver's device buffers."""
    _capi.check_mqg(mqg._lib.ptf_mqg_couple(mqg._h, tracer._h), mqg._h)
    mqg._tracers.append(tracer)
    mqg.vars._invalidate()


def step_coupled(tracer, nsteps: int = 1) -> float:
    """The loop of examples/turbulent_advection-diffusion.jl:149-151, ``nsteps`` times, in one library call:
    ``stepforward!(ADprob); stepforward!(params.MQGprob); MultiLayerQG.updatevars!(params.MQGprob)``.
    Returns the device time in ms (CUDA events on the shared stream)."""
    mqg = tracer.params.MQGprob
    if not isinstance(mqg, MultiLayerQGProblem):
        raise TypeError("step_coupled needs a tracer problem built from a B200 MultiLayerQG problem")
    mqg._push_dt()
    ms = C.c_float()
    _capi.check_mqg(mqg._lib.ptf_mqg_step_coupled(mqg._h, tracer._h, int(nsteps), C.byref(ms)), mqg._h)
    mqg._sync_clock()
    tracer._sync_clock()
    mqg.vars._invalidate()
    return ms.value
