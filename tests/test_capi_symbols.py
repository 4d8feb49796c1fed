"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/ptf_b200.h
declares, agrees with the ctypes descriptor layout, and fails LOUDLY (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import ptf_b200
    return ptf_b200._capi


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ptf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ptf_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(capi):
    lib = capi.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ptf_b200.h but not exported"
    assert set(syms) == set(capi.SIGNATURES), "ctypes binding and header disagree on the symbol list"


def test_descriptor_layout_matches(capi):
    lib = capi.load()
    d = capi.PtfDesc()
    assert lib.ptf_desc_init(C.byref(d)) == 0
    assert d.struct_size == C.sizeof(capi.PtfDesc)
    # reference defaults (TAD.jl:143-203)
    assert d.ndim == 2 and list(d.n) == [128, 128, 128]
    assert abs(d.L[0] - 6.283185307179586) < 1e-15 and d.kappa[0] == 0.1 and d.dt == 0.01
    assert d.stepper == capi.STEPPER_IDS["RK4"] and d.nyquist_sign == -1 and d.dealias == 0
    assert abs(d.filter_inner_k - 2 / 3) < 1e-16 and d.filter_tol == 1e-15 and d.use_graph == 1


def test_version_and_error_strings(capi):
    lib = capi.load()
    a, b = C.c_int32(), C.c_int32()
    assert lib.ptf_version(C.byref(a), C.byref(b)) == 0 and (a.value, b.value) == (0, 1)
    assert b"PTF_OK" in lib.ptf_error_string(0)
    assert b"no CPU fallback" in lib.ptf_error_string(capi.ENODEVICE)


def test_bad_descriptor_is_rejected(capi):
    lib = capi.load()
    d = capi.PtfDesc()
    lib.ptf_desc_init(C.byref(d))
    d.n[0] = 127   # odd grid
    h = C.c_void_p()
    rc = lib.ptf_create(C.byref(d), C.byref(h))
    assert rc == capi.EINVAL and not h.value
    assert b"even" in lib.ptf_last_error(None)
    d.n[0] = 128
    d.struct_size = 4
    assert lib.ptf_create(C.byref(d), C.byref(h)) == capi.EINVAL


def test_no_gpu_means_loud_failure_not_cpu_fallback(capi):
    from tests.conftest import has_gpu
    if has_gpu():
        pytest.skip("GPU present")
    import numpy as np
    import ptf_b200 as P
    with pytest.raises(P._capi.PtfError) as ei:
        P.Problem(P.B200(), P.TwoDAdvectingFlow(), nx=32)
    assert ei.value.status == capi.ENODEVICE


def test_host_mirror_argument_errors(capi):
    import ptf_b200 as P

    class FakeMQG:
        pass
    with pytest.raises(ValueError, match="non-negative"):
        P.Problem(FakeMQGWithGrid(), tracer_release_time=-1.0)
    with pytest.raises(ValueError, match="unknown stepper"):
        P.tracer_advection_diffusion._parse_stepper("RK5")
    assert P.noflow(3.14) == 0.0          # test/runtests.jl:56


class FakeMQGWithGrid:
    class _G:
        nx, ny, Lx, Ly = 16, 16, 6.28, 6.28
    grid = _G()


def test_mqg_descriptor_layout_matches(capi):
    """ptf_mqg_desc (MultiLayerQG.Problem keyword arguments): ctypes mirror and C struct agree; GeophysicalFlows defaults."""
    lib = capi.load()
    d = capi.PtfMqgDesc()
    assert lib.ptf_mqg_desc_init(C.byref(d)) == 0
    assert d.struct_size == C.sizeof(capi.PtfMqgDesc)
    assert d.nlayers == 2 and d.nx == 128 and d.ny == 128 and abs(d.Lx - 6.283185307179586) < 1e-15
    assert d.f0 == 1.0 and d.beta == 0.0 and d.mu == 0.0 and d.nu == 0.0 and d.n_nu == 1 and d.dt == 0.01
    assert d.stepper == capi.STEPPER_IDS["RK4"] and abs(d.aliased_fraction - 1 / 3) < 1e-16 and d.use_graph == 1
    assert not d.H and not d.b and not d.U and not d.eta


def test_mqg_and_expression_entry_points_fail_loudly_without_a_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    lib = capi.load()
    d = capi.PtfMqgDesc()
    lib.ptf_mqg_desc_init(C.byref(d))
    import numpy as np
    H = np.array([0.5, 0.5])
    b = np.array([-1.0, -1.2])
    d.H, d.b = capi.as_dp(H), capi.as_dp(b)
    h = C.c_void_p()
    assert lib.ptf_mqg_create(C.byref(d), C.byref(h)) == capi.ENODEVICE and not h.value
    assert b"no CPU fallback" in lib.ptf_mqg_last_error(None)
    import ptf_b200 as P
    with pytest.raises(capi.PtfError):
        P.Problem(P.B200(), P.ExpressionFlow("sin(x)", "cos(y)"), nx=64)


def test_host_mirrors_of_filter_and_linear_operator_match_the_oracle():
    """prob.timestepper.filter (examples/turbulent_advection-diffusion.jl:64) and prob.eqn.L are host-side mirrors of what
    the library evaluates in registers; they must equal the oracle's arrays bit for bit."""
    import numpy as np
    import ptf_b200 as P
    from oracle.ptf_oracle import Grid as OGrid, linear_operator, make_filter
    T = P.tracer_advection_diffusion
    for n, L in (((48,), (3.0,)), ((32, 48), (2 * np.pi, 4.0)), ((16, 24, 32), (2 * np.pi, 4.0, 3.0))):
        nd = len(n)
        kw = dict(zip(("nx", "ny", "nz"), n))
        kw.update(dict(zip(("Lx", "Ly", "Lz"), L)))
        g = T.Grid(ndim=nd, **kw)
        og = OGrid(n, L)
        assert np.array_equal(T.makefilter(g), make_filter(og))
        ts = T.TimeStepper("FilteredRK4", g)
        assert ts == "FilteredRK4" and str(ts) == "FilteredRK4" and np.array_equal(ts.filter, make_filter(og))
        with pytest.raises(AttributeError):
            T.TimeStepper("RK4", g).filter
        stub = T.TracerProblem.__new__(T.TracerProblem)      # eqn needs no device
        stub.grid, stub.nbatch, stub.local_nbatch = g, 1, 1
        stub.params = T.Params(kappa=0.01, eta=0.02, iota=0.005, kappa_h=1e-6, n_kappa_h=2)
        assert np.array_equal(stub.eqn.L, linear_operator(og, (0.01, 0.02, 0.005)[:nd], 1e-6, 2))
