"""Host-side mirror of ``PassiveTracerFlows.TracerAdvectionDiffusion`` for the ``B200`` device.

Same names, argument meaning and error behaviour as the reference module (src/traceradvectiondiffusion.jl,
"TAD.jl"), with Julia's ``!`` dropped: ``Problem``, ``set_c``, ``updatevars``, ``stepforward``, ``step_until``,
``OneDAdvectingFlow`` / ``TwoDAdvectingFlow`` / ``ThreeDAdvectingFlow``, ``noflow``.  All numerics run inside
libptf_b200.so through the C ABI (include/ptf_b200.h); this file only builds descriptors, samples the user's
velocity closures on ``gridpoints`` (as TAD.jl:426-452 does on the host) and mirrors ``prob.sol`` /
``prob.vars.c`` / ``prob.clock`` into NumPy arrays.  The Julia glue in INTEGRATION.md does the same with ccall.

Array convention: NumPy C-order with x as the *last* axis — byte-identical to Julia's column-major arrays.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _capi


# ----------------------------------------------------------------------------------------------
# Devices (FourierFlows ``Device`` seam; test/runtests.jl:12).  ``B200`` is the new device type.
# ----------------------------------------------------------------------------------------------
class Device:
    pass


@dataclass
class B200(Device):
    """Selects libptf_b200.so.  ``device`` = CUDA ordinal; ``engine`` in {"auto","cufft","fused"}."""
    device: int = -1
    engine: str = "auto"
    use_graph: bool = True
    # multi-process jobs (one process per GPU): filled by ptf_b200.parallel.init_b200()
    rank: int = 0
    nranks: int = 1
    nccl_id: Optional[bytes] = None
    decomposition: str = "none"      # "none" | "batch" | "slab"


# ----------------------------------------------------------------------------------------------
# Advecting flows (TAD.jl:29-127)
# ----------------------------------------------------------------------------------------------
def noflow(*args):
    """Default u, v, w (TAD.jl:31)."""
    return 0.0


@dataclass
class OneDAdvectingFlow:
    u: Callable = noflow
    steadyflow: bool = True

    def __iter__(self):
        return iter((self.u, self.steadyflow))


@dataclass
class TwoDAdvectingFlow:
    u: Callable = noflow
    v: Callable = noflow
    steadyflow: bool = True

    def __iter__(self):
        return iter((self.u, self.v, self.steadyflow))


@dataclass
class ThreeDAdvectingFlow:
    u: Callable = noflow
    v: Callable = noflow
    w: Callable = noflow
    steadyflow: bool = True

    def __iter__(self):
        return iter((self.u, self.v, self.w, self.steadyflow))


@dataclass
class SeparableFlow:
    """Extension for large grids: ``u_a(x,y,z,t) = sum_m coeff_a(t)[m] * X_a[m](x) * Y_a[m](y) * Z_a[m](z)``.

    ``terms[a]`` is a list of (fx, fy, fz) callables of one coordinate each (fy/fz omitted below 2/3-D);
    ``coeffs(t, a)`` returns the length-``len(terms[a])`` coefficient vector (None = all ones, steady).
    Velocities are evaluated in registers inside the product kernel: zero HBM bytes.
    """
    terms: Sequence[Sequence[Sequence[Callable]]]
    coeffs: Optional[Callable] = None
    steadyflow: bool = True


@dataclass
class ExpressionFlow:
    """Extension for large grids and arbitrary closures: the velocity components written as C/CUDA expressions in
    ``x, y, z, t`` (device math: ``sin``, ``cos``, ``exp``, …, ``pi``), e.g. ``u="(sin(z) + cos(y)) * (1 + 0.5*sin(t))"``.
    They play the role of the reference's ``u(x, y, z, t)`` closures (TAD.jl:268-348): evaluated at ``clock.t`` on the
    grid points, but in registers inside the product kernel (compiled at run time with NVRTC) — zero HBM bytes and no
    per-step upload.  Runs on every engine: in the cuFFT pipelines (any size, 1-D/2-D/3-D, slab-decomposed) inside the product
    kernel, on the fused 2-D / 3-D engines through a fill kernel that writes the fields once per step."""
    u: str = "0.0"
    v: Optional[str] = None
    w: Optional[str] = None
    steadyflow: bool = False

    @property
    def components(self):
        return [c for c in (self.u, self.v, self.w) if c is not None]


# ----------------------------------------------------------------------------------------------
# Small mirrors of the FourierFlows containers the reference's users destructure
# ----------------------------------------------------------------------------------------------
@dataclass
class Clock:
    dt: float
    t: float = 0.0
    step: int = 0


@dataclass
class Grid:
    """Mirror of FourierFlows One/Two/ThreeDGrid fields used by TAD.jl (FF domains.jl)."""
    nx: int
    Lx: float
    ny: int = 1
    Ly: float = 1.0
    nz: int = 1
    Lz: float = 1.0
    ndim: int = 1
    device: Device = None

    def __post_init__(self):
        self.dx, self.dy, self.dz = self.Lx / self.nx, self.Ly / self.ny, self.Lz / self.nz
        self.x = -self.Lx / 2 + self.dx * np.arange(self.nx)
        self.y = -self.Ly / 2 + self.dy * np.arange(self.ny)
        self.z = -self.Lz / 2 + self.dz * np.arange(self.nz)
        self.nkr, self.nl, self.nm = self.nx // 2 + 1, self.ny, self.nz
        self.kr = np.arange(self.nkr) * (2 * np.pi / self.Lx * self.nx) / self.nx

        def fftk(n, L):
            j = np.arange(n)
            j = np.where(j < n // 2, j, j - n).astype(float)
            return j * (2 * np.pi / L * n) / n
        self.l = fftk(self.ny, self.Ly) if self.ndim >= 2 else np.zeros(1)
        self.m = fftk(self.nz, self.Lz) if self.ndim >= 3 else np.zeros(1)

    @property
    def n(self):
        return (self.nx, self.ny, self.nz)[: self.ndim]

    @property
    def pshape(self):
        return tuple(reversed(self.n))

    @property
    def sshape(self):
        return tuple(reversed(self.n[1:])) + (self.nkr,)


def gridpoints(grid: Grid):
    """FourierFlows ``gridpoints(grid)``: full coordinate arrays in physical layout (x fastest)."""
    coords = (grid.x, grid.y, grid.z)[: grid.ndim]
    out = []
    for a in range(grid.ndim):
        shp = [1] * grid.ndim
        shp[grid.ndim - 1 - a] = grid.n[a]
        out.append(np.broadcast_to(coords[a].reshape(shp), grid.pshape))
    return tuple(out) if grid.ndim > 1 else out[0]


def makefilter(grid: "Grid", order: float = 4.0, innerK: float = 2.0 / 3.0, outerK: float = 1.0, tol: float = 1e-15):
    """FourierFlows ``makefilter(grid)`` (FF utils.jl): 1 for K < innerK, exp(−decay·(K − innerK)^order) beyond, with
    K = sqrt(Σ (k_a·d_a/π)²).  Host mirror of what the library evaluates in registers; users read it through
    ``prob.timestepper.filter`` (examples/turbulent_advection-diffusion.jl:64)."""
    shp = grid.sshape
    nd = grid.ndim
    ks = (grid.kr, grid.l, grid.m)[:nd]
    ds = (grid.dx, grid.dy, grid.dz)[:nd]
    Ksq = 0.0
    for a in range(nd):
        v = [1] * nd
        v[nd - 1 - a] = len(ks[a])
        Ksq = Ksq + (ks[a].reshape(v) * ds[a] / np.pi) ** 2
    K = np.sqrt(Ksq)
    decay = -math.log(tol) / (outerK - innerK) ** order
    return np.ascontiguousarray(np.broadcast_to(np.where(K < innerK, 1.0, np.exp(-decay * (K - innerK) ** order)), shp))


class TimeStepper:
    """Mirror of the FourierFlows time-stepper object: behaves like its name (``prob.timestepper == "FilteredRK4"``)
    and exposes ``filter`` for the Filtered steppers."""

    def __init__(self, name: str, grid: "Grid"):
        self.name = name
        self._grid = grid
        self._filter = None

    @property
    def filter(self):
        if not self.name.startswith("Filtered"):
            raise AttributeError(f"{self.name}TimeStepper has no field filter")
        if self._filter is None:
            self._filter = makefilter(self._grid)
        return self._filter

    def __eq__(self, other):
        return self.name == (other.name if isinstance(other, TimeStepper) else other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return f"{self.name}TimeStepper"

    __str__ = lambda self: self.name


@dataclass
class Equation:
    """``prob.eqn``: the diagonal linear operator as a host array (TAD.jl:502-566; MultiLayerQG hyperviscosity)."""
    L: np.ndarray
    dims: tuple


@dataclass
class Vars:
    """``prob.vars``: host mirrors refreshed by ``updatevars`` (TAD.jl:579-636)."""
    c: np.ndarray
    ch: np.ndarray


@dataclass
class Params:
    kappa: float
    eta: float
    iota: float
    kappa_h: float
    n_kappa_h: int
    nlayers: int = 1
    tracer_release_time: float = 0.0
    MQGprob: object = None
    u: object = None
    v: object = None
    w: object = None


def _parse_stepper(stepper: str) -> int:
    name = stepper
    flag = 0
    if name.startswith("Filtered"):
        flag = _capi.STEPPER_FILTERED
        name = name[len("Filtered"):]
    if name not in _capi.STEPPER_IDS:
        raise ValueError(f"unknown stepper {stepper!r}")   # FF: eval(Symbol(stepper, :TimeStepper)) fails
    return _capi.STEPPER_IDS[name] | flag


class TracerProblem:
    """The object ``Problem(...)`` returns: a FourierFlows.Problem look-alike backed by a device handle."""

    def __init__(self, dev: B200, grid: Grid, params: Params, dt: float, stepper: str, flow_kind: int,
                 nbatch: int = 1, velocity_per_batch: bool = False, dealias: bool = False,
                 aliased_fraction: float = 1.0 / 3.0, nyquist_sign: int = -1):
        if not isinstance(dev, B200):
            raise TypeError("this package implements the B200 device only (no CPU path)")
        self._lib = _capi.load()
        self.grid = grid
        self.params = params
        self.stepper = stepper
        self.timestepper = TimeStepper(stepper, grid)
        self.clock = Clock(dt=float(dt))
        self.nbatch = int(nbatch)
        self.dev = dev
        d = _capi.PtfDesc()
        self._lib.ptf_desc_init(C.byref(d))
        d.ndim = grid.ndim
        d.n[:] = [grid.nx, grid.ny, grid.nz]
        d.L[:] = [grid.Lx, grid.Ly, grid.Lz]
        d.nbatch = self.nbatch
        d.stepper = _parse_stepper(stepper)
        d.kappa[:] = [params.kappa, params.eta, params.iota]
        d.kappa_h = params.kappa_h
        d.n_kappa_h = params.n_kappa_h
        d.dealias = 1 if dealias else 0
        d.aliased_fraction = aliased_fraction
        d.dt = float(dt)
        d.nyquist_sign = nyquist_sign
        d.flow_kind = flow_kind
        d.velocity_per_batch = 1 if velocity_per_batch else 0
        d.engine = _capi.ENGINE_NAMES[dev.engine]
        d.device = dev.device
        d.use_graph = 1 if dev.use_graph else 0
        d.nranks, d.rank = dev.nranks, dev.rank
        d.decomposition = {"none": _capi.DECOMP_NONE, "batch": _capi.DECOMP_BATCH,
                           "slab": _capi.DECOMP_SLAB}[dev.decomposition]
        if dev.nccl_id is not None:
            d.nccl_id[:] = list(dev.nccl_id)
        self._desc = d
        h = C.c_void_p()
        _capi.check(self._lib.ptf_create(C.byref(d), C.byref(h)))
        self._h = h
        pn = (C.c_int64 * 4)()
        sn = (C.c_int64 * 4)()
        po = (C.c_int64 * 4)()
        so = (C.c_int64 * 4)()
        _capi.check(self._lib.ptf_local_shape(h, pn, sn, po, so), h)
        self.local_nbatch = int(pn[3])
        self.batch_offset = int(po[3])
        # slab decomposition: this rank owns z-planes [z_offset, z_offset+nz_local) in physical space and
        # ky-rows [ky_offset, ky_offset+ny_local) in spectral space
        self.nz_local, self.z_offset = int(pn[2]), int(po[2])
        self.ny_local, self.ky_offset = int(sn[1]), int(so[1])
        # 2-D slab decomposition: rows [y_offset, y_offset+ny_phys_local) in physical space and kr-columns
        # [kr_offset, kr_offset+nkr_local) (all ky) in spectral space
        self.ny_phys_local, self.y_offset = int(pn[1]), int(po[1])
        self.nkr_local, self.kr_offset = int(sn[0]), int(so[0])
        nd = grid.ndim
        lead = (self.local_nbatch,) if self.nbatch > 1 else ()
        self._pshape = lead + tuple(int(pn[a]) for a in reversed(range(nd)))
        self._sshape = lead + tuple(int(sn[a]) for a in reversed(range(nd)))
        self.sol = np.zeros(self._sshape, dtype=np.complex128)
        self.vars = Vars(c=np.zeros(self._pshape), ch=self.sol)
        self._keep = []          # ctypes callbacks must outlive the handle
        self._vel_funcs = None
        self._mqg = None
        self._mqg_on_device = False

    @property
    def eqn(self) -> Equation:
        """``prob.eqn`` with ``L = −κkr² − ηl² − ιm² − κh·Krsq^nκh`` in the reference's operation order (TAD.jl:502-566)."""
        g, p, nd = self.grid, self.params, self.grid.ndim

        def ax(a, k):
            v = [1] * nd
            v[nd - 1 - a] = len(k)
            return k.reshape(v)
        kr2 = ax(0, g.kr) ** 2
        L, Ksq = -p.kappa * kr2, kr2
        if nd >= 2:
            l2 = ax(1, g.l) ** 2
            L, Ksq = L - p.eta * l2, Ksq + l2
        if nd >= 3:
            m2 = ax(2, g.m) ** 2
            L, Ksq = L - p.iota * m2, Ksq + m2
        hyper = np.ones_like(Ksq)
        for _ in range(int(p.n_kappa_h)):
            hyper = hyper * Ksq
        L = np.broadcast_to(L - p.kappa_h * hyper, g.sshape)
        if self.nbatch > 1:
            L = np.broadcast_to(L, (self.local_nbatch,) + g.sshape)     # copied into each layer (TAD.jl:561-563)
        return Equation(L=np.ascontiguousarray(L), dims=L.shape)

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ptf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ----
    @property
    def engine(self) -> str:
        e = C.c_int32()
        _capi.check(self._lib.ptf_engine(self._h, C.byref(e)), self._h)
        return {1: "cufft", 2: "fused"}[e.value]

    def launch_count(self):
        a, b = C.c_int64(), C.c_int64()
        _capi.check(self._lib.ptf_launch_count(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def device_bytes(self) -> int:
        b = C.c_int64()
        _capi.check(self._lib.ptf_device_bytes(self._h, C.byref(b)), self._h)
        return b.value

    def kernel_time_ms(self, name: str, reps: int = 10) -> float:
        ms = C.c_float()
        _capi.check(self._lib.ptf_kernel_timed(self._h, name.encode(), reps, C.byref(ms)), self._h)
        return ms.value

    def diagnostics(self):
        m, v, s = C.c_double(), C.c_double(), C.c_double()
        _capi.check(self._lib.ptf_diag(self._h, C.byref(m), C.byref(v), C.byref(s)), self._h)
        return {"mean_c": m.value, "variance_c": v.value, "max_abs_sol": s.value}

    # ---- velocities ----
    def _set_velocity_arrays(self, arrays):
        for a, arr in enumerate(arrays):
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            _capi.check(self._lib.ptf_set_velocity(self._h, a, _capi.as_dp(arr), arr.size), self._h)

    def local_gridpoints(self):
        """``gridpoints(grid)`` restricted to the z-planes this rank owns (the whole grid on a single GPU)."""
        pts = gridpoints(self.grid)
        pts = pts if isinstance(pts, tuple) else (pts,)
        if self.grid.ndim == 3 and self.nz_local != self.grid.nz:
            sl = slice(self.z_offset, self.z_offset + self.nz_local)
            pts = tuple(p[sl] for p in pts)
        if self.grid.ndim == 2 and self.ny_phys_local != self.grid.ny:
            sl = slice(self.y_offset, self.y_offset + self.ny_phys_local)
            pts = tuple(p[sl] for p in pts)
        return pts

    def _install_velocity_callback(self, funcs):
        grid = self.grid
        pts = self.local_gridpoints()
        lshape = pts[0].shape
        npts = int(np.prod(lshape))
        nd = grid.ndim

        def cb(user, t, pu, pv, pw):
            # params.u(x, y, clock.t) on gridpoints (TAD.jl:718) — once per step, t = clock.t
            for a, p in enumerate((pu, pv, pw)[:nd]):
                out = np.ctypeslib.as_array(p, shape=(npts,))
                out[:] = np.broadcast_to(np.asarray(funcs[a](*pts, t), dtype=np.float64), lshape).ravel()

        fn = _capi.VELOCITY_FN(cb)
        self._keep.append(fn)
        _capi.check(self._lib.ptf_set_velocity_callback(self._h, fn, None), self._h)

    def _install_separable(self, flow: SeparableFlow):
        grid = self.grid
        coords = (grid.x, grid.y, grid.z)
        nd = grid.ndim
        for a in range(nd):
            terms = flow.terms[a]
            nt = len(terms)
            tabs = []
            for ax in range(3):
                if ax < nd and nt:
                    tabs.append(np.ascontiguousarray(
                        [np.broadcast_to(np.asarray(tm[ax](coords[ax]), dtype=np.float64), coords[ax].shape)
                         for tm in terms]))
                else:
                    tabs.append(None)
            c0 = np.ascontiguousarray(flow.coeffs(0.0, a), dtype=np.float64) if (flow.coeffs and nt) else None
            _capi.check(self._lib.ptf_set_velocity_separable(
                self._h, a, nt, *(None if t is None else _capi.as_dp(t) for t in tabs),
                None if c0 is None else _capi.as_dp(c0)), self._h)
            self._keep.extend(tabs)
        if flow.coeffs is not None and not flow.steadyflow:
            def ccb(user, t, comp, nterms, pa):
                out = np.ctypeslib.as_array(pa, shape=(nterms,))
                out[:] = np.asarray(flow.coeffs(t, comp), dtype=np.float64)
            fn = _capi.COEFF_FN(ccb)
            self._keep.append(fn)
            _capi.check(self._lib.ptf_set_coeff_callback(self._h, fn, None), self._h)

    def set_velocity_expr(self, comp: int, expr: str):
        """Replace one component of an ``ExpressionFlow`` (recompiles the product kernel once all components are set)."""
        _capi.check(self._lib.ptf_set_velocity_expr(self._h, int(comp), str(expr).encode()), self._h)

    def set_layered_velocity(self, u, v, U=None):
        """MQG coupling: u = MQGprob.vars.u (+U broadcast over x), v = MQGprob.vars.v (TAD.jl:795-796)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        Up = None
        if U is not None:
            U = np.asarray(U, dtype=np.float64)
            if U.ndim == 1:
                U = np.repeat(U.reshape(-1, 1), self.grid.ny, axis=1)
            U = np.ascontiguousarray(U.reshape(self.local_nbatch, self.grid.ny))
            Up = _capi.as_dp(U)
        _capi.check(self._lib.ptf_set_layered_velocity(self._h, _capi.as_dp(u), _capi.as_dp(v), Up), self._h)

    def _refresh_mqg_velocity(self):
        q = self._mqg
        if q is None or self._mqg_on_device:   # device-coupled flows are read in place (ptf_mqg_couple)
            return
        self.set_layered_velocity(q.vars.u, q.vars.v, getattr(q.params, "U", None))

    # ---- verbs (exposed as module-level functions too) ----
    def set_c(self, c):
        c = np.ascontiguousarray(c, dtype=np.float64)
        nd = self.grid.ndim
        replicate = self.nbatch > 1 and c.ndim == nd
        expect = self._pshape[-nd:] if replicate or self.nbatch == 1 else self._pshape
        if tuple(c.shape) != tuple(expect):
            raise ValueError(f"c has shape {c.shape}, expected {expect}")
        _capi.check(self._lib.ptf_set_c(self._h, _capi.as_dp(c), 1 if replicate else 0), self._h)
        self.updatevars()

    def updatevars(self):
        _capi.check(self._lib.ptf_get_c(self._h, _capi.as_dp(self.vars.c)), self._h)
        _capi.check(self._lib.ptf_get_sol(self._h, self.sol.ctypes.data_as(C.POINTER(C.c_double))), self._h)
        return self.vars.c

    def set_sol(self, sol):
        sol = np.ascontiguousarray(sol, dtype=np.complex128)
        if sol.shape != self.sol.shape:
            raise ValueError(f"sol has shape {sol.shape}, expected {self.sol.shape}")
        _capi.check(self._lib.ptf_set_sol(self._h, sol.ctypes.data_as(C.POINTER(C.c_double))), self._h)
        self.sol[...] = sol

    def _sync_clock(self):
        t, s, dt = C.c_double(), C.c_int64(), C.c_double()
        _capi.check(self._lib.ptf_get_clock(self._h, C.byref(t), C.byref(s), C.byref(dt)), self._h)
        self.clock.t, self.clock.step, self.clock.dt = t.value, s.value, dt.value

    def stepforward(self, nsteps: int = 1):
        if self.clock.dt != self._desc.dt:       # user changed prob.clock.dt (FF allows it)
            _capi.check(self._lib.ptf_set_dt(self._h, float(self.clock.dt)), self._h)
            self._desc.dt = self.clock.dt
        self._refresh_mqg_velocity()
        _capi.check(self._lib.ptf_step(self._h, int(nsteps)), self._h)
        self._sync_clock()

    def step_until(self, stop_time: float):
        self._refresh_mqg_velocity()
        _capi.check(self._lib.ptf_step_until(self._h, float(stop_time)), self._h)
        self._sync_clock()

    def step_timed(self, nsteps: int) -> float:
        """stepforward with the device time (CUDA events on the step stream) returned in ms."""
        ms = C.c_float()
        _capi.check(self._lib.ptf_step_timed(self._h, int(nsteps), C.byref(ms)), self._h)
        self._sync_clock()
        return ms.value


# ----------------------------------------------------------------------------------------------
# Problem constructors (TAD.jl:143-250)
# ----------------------------------------------------------------------------------------------
def _sample(f, pts, shape):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(f(*pts), dtype=np.float64), shape))


def Problem(dev_or_mqg, advecting_flow=None, *, nx=128, Lx=2 * math.pi, ny=None, Ly=None, nz=None, Lz=None,
            kappa=0.1, eta=None, iota=None, dt=0.01, stepper=None, T=np.float64,
            kappa_h=0.0, n_kappa_h=0, dealias=False, aliased_fraction=1.0 / 3.0, nbatch=1,
            tracer_release_time=0, dev: Optional[B200] = None, nyquist_sign=-1):
    """``Problem(dev, advecting_flow; nx, Lx, ..., κ, η, ι, dt, stepper)`` (TAD.jl:143-216) and
    ``Problem(MQGprob; κ, η, stepper, tracer_release_time)`` (TAD.jl:225-250).

    Extra keywords (not in the reference): ``kappa_h``/``n_kappa_h`` (only reachable through the low-level
    constructors there, test/...:421), ``dealias``, ``nbatch`` (ensemble of independent tracers sharing the flow).
    """
    if T not in (np.float64, float, "Float64"):
        raise NotImplementedError("the B200 path is fp64 (T=Float64); Float32 is listed as future work")
    if not isinstance(dev_or_mqg, Device):
        return _layered_problem(dev_or_mqg, kappa=kappa, eta=eta, stepper=stepper or "FilteredRK4",
                                tracer_release_time=tracer_release_time, dev=dev or B200())
    dev = dev_or_mqg
    stepper = stepper or "RK4"
    flow = advecting_flow
    if isinstance(flow, OneDAdvectingFlow):
        nd = 1
    elif isinstance(flow, TwoDAdvectingFlow):
        nd = 2
    elif isinstance(flow, ThreeDAdvectingFlow):
        nd = 3
    elif isinstance(flow, SeparableFlow):
        nd = len(flow.terms)
    elif isinstance(flow, ExpressionFlow):
        nd = len(flow.components)
    else:
        raise TypeError("advecting_flow must be a One/Two/ThreeDAdvectingFlow (or SeparableFlow / ExpressionFlow)")
    ny = nx if ny is None else ny
    Ly = Lx if Ly is None else Ly
    nz = nx if nz is None else nz
    Lz = Lx if Lz is None else Lz
    eta = kappa if eta is None else eta
    iota = kappa if iota is None else iota
    grid = Grid(nx=nx, Lx=Lx, ny=ny if nd >= 2 else 1, Ly=Ly if nd >= 2 else 1.0,
                nz=nz if nd >= 3 else 1, Lz=Lz if nd >= 3 else 1.0, ndim=nd, device=dev)
    params = Params(kappa=float(kappa), eta=float(eta), iota=float(iota), kappa_h=float(kappa_h),
                    n_kappa_h=int(n_kappa_h))
    if isinstance(flow, SeparableFlow):
        prob = TracerProblem(dev, grid, params, dt, stepper, _capi.FLOW_SEPARABLE, nbatch=nbatch,
                             dealias=dealias, aliased_fraction=aliased_fraction, nyquist_sign=nyquist_sign)
        prob._install_separable(flow)
        return prob
    if isinstance(flow, ExpressionFlow):
        prob = TracerProblem(dev, grid, params, dt, stepper, _capi.FLOW_EXPR, nbatch=nbatch,
                             dealias=dealias, aliased_fraction=aliased_fraction, nyquist_sign=nyquist_sign)
        for a, e in enumerate(flow.components):
            _capi.check(prob._lib.ptf_set_velocity_expr(prob._h, a, str(e).encode()), prob._h)
        params.u, params.v, params.w = (flow.components + [None, None])[:3]
        return prob
    funcs = [flow.u, getattr(flow, "v", None), getattr(flow, "w", None)][:nd]
    if flow.steadyflow:
        # ConstDiffSteadyFlowParams: u.(x, y) evaluated ONCE on gridpoints (TAD.jl:426-452)
        prob = TracerProblem(dev, grid, params, dt, stepper, _capi.FLOW_STEADY, nbatch=nbatch,
                             dealias=dealias, aliased_fraction=aliased_fraction, nyquist_sign=nyquist_sign)
        pts = prob.local_gridpoints()      # this rank's z-slab (everything on a single GPU)
        arrays = [_sample(f, pts, pts[0].shape) for f in funcs]
        params.u, params.v, params.w = (arrays + [None, None])[:3]
        prob._set_velocity_arrays(arrays)
    else:
        # ConstDiffTimeVaryingFlowParams keeps the functions (TAD.jl:341); evaluated at clock.t per step
        prob = TracerProblem(dev, grid, params, dt, stepper, _capi.FLOW_CALLBACK, nbatch=nbatch,
                             dealias=dealias, aliased_fraction=aliased_fraction, nyquist_sign=nyquist_sign)
        params.u, params.v, params.w = (funcs + [None, None])[:3]
        prob._install_velocity_callback(funcs)
    return prob


def _layered_problem(MQGprob, *, kappa, eta, stepper, tracer_release_time, dev):
    """``Problem(MQGprob; κ, η, stepper="FilteredRK4", tracer_release_time=0)`` (TAD.jl:225-250).

    ``MQGprob`` is any object with the fields the reference touches: ``grid`` (nx, Lx, ny, Ly), ``clock.dt``,
    ``vars.u``/``vars.v`` of shape (nlayers, ny, nx), ``params.U`` ((nlayers,) or (nlayers, ny)),
    ``params.nlayers``, and the verbs ``step_until(t)`` / ``updatevars()``.
    """
    g = MQGprob.grid
    if getattr(MQGprob, "dev", None) is not None and isinstance(MQGprob.dev, B200):
        dev = B200(device=MQGprob.dev.device, engine=dev.engine, use_graph=dev.use_graph)   # same device as the flow
    if tracer_release_time < 0:
        raise ValueError("tracer_release_time must be non-negative!")      # ArgumentError, TAD.jl:234
    if tracer_release_time > 0:
        MQGprob.step_until(tracer_release_time)                             # TAD.jl:236-239
    eta = kappa if eta is None else eta
    nlayers = int(MQGprob.params.nlayers)
    MQGprob.updatevars()                                                    # TAD.jl:488
    grid = Grid(nx=g.nx, Lx=g.Lx, ny=g.ny, Ly=g.Ly, ndim=2, device=dev)
    params = Params(kappa=float(kappa), eta=float(eta), iota=float(kappa), kappa_h=0.0, n_kappa_h=0,
                    nlayers=nlayers, tracer_release_time=float(tracer_release_time), MQGprob=MQGprob)
    prob = TracerProblem(dev, grid, params, MQGprob.clock.dt, stepper, _capi.FLOW_LAYERED, nbatch=nlayers,
                         velocity_per_batch=True)
    prob._mqg = MQGprob
    from . import multilayerqg
    if isinstance(MQGprob, multilayerqg.MultiLayerQGProblem):
        # flow solver lives on the same device: alias its u, v buffers instead of uploading every step
        multilayerqg.couple(prob, MQGprob)
        prob._mqg_on_device = True
    else:
        prob._refresh_mqg_velocity()
    return prob


def set_c(prob: TracerProblem, c):
    """``set_c!(prob, c)`` (TAD.jl:844-872)."""
    prob.set_c(c)


def updatevars(prob: TracerProblem):
    """``updatevars!(prob)`` (TAD.jl:815-837)."""
    return prob.updatevars()


def stepforward(prob: TracerProblem, nsteps: int = 1):
    """FourierFlows ``stepforward!(prob[, nsteps])``."""
    prob.stepforward(nsteps)


def step_until(prob: TracerProblem, stop_time: float):
    """FourierFlows ``step_until!(prob, stop_time)`` (used at TAD.jl:238)."""
    prob.step_until(stop_time)
