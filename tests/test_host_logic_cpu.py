"""CPU-only tests of the host-side mirror (no CUDA device needed): grids, argument validation and error behaviour that is
decided before the library is asked for a device, mirroring the reference's constructors (TAD.jl:143-250)."""
import numpy as np
import pytest

import ptf_b200 as P
from oracle.ptf_oracle import Grid as OGrid

T = P.tracer_advection_diffusion


def test_grid_mirror_matches_the_oracle_grid():
    for n, L in (((48,), (3.0,)), ((32, 48), (2 * np.pi, 4.0)), ((16, 24, 32), (2 * np.pi, 4.0, 3.0))):
        nd = len(n)
        kw = dict(zip(("nx", "ny", "nz"), n))
        kw.update(dict(zip(("Lx", "Ly", "Lz"), L)))
        g = T.Grid(ndim=nd, **kw)
        og = OGrid(n, L)
        for a, (mine, theirs) in enumerate(zip((g.kr, g.l, g.m)[:nd], og.k)):
            assert np.array_equal(mine, theirs), a
        for mine, theirs in zip((g.x, g.y, g.z)[:nd], og.coords):
            assert np.array_equal(mine, theirs)
        assert g.x[0] == -L[0] / 2 and g.nkr == n[0] // 2 + 1          # x0 = -Lx/2, r2c axis
        if nd >= 2:
            assert g.l[n[1] // 2] < 0                                   # fftfreq: Nyquist wavenumber negative
        pts = P.gridpoints(g)
        pts = pts if isinstance(pts, tuple) else (pts,)
        for mine, theirs in zip(pts, og.gridpoints()):
            assert np.array_equal(mine, theirs)
        assert g.pshape == og.pshape and g.sshape == og.sshape


def test_stepper_names():
    assert T._parse_stepper("RK4") == 1 and T._parse_stepper("FilteredRK4") == 1 | 16
    assert T._parse_stepper("FilteredETDRK4") == 2 | 16 and T._parse_stepper("LSRK54") == 3
    with pytest.raises(ValueError):
        T._parse_stepper("RK5")


def test_constructor_argument_validation_happens_before_any_device_work():
    with pytest.raises(TypeError):
        P.Problem(P.B200(), "not a flow")
    with pytest.raises(NotImplementedError):
        P.Problem(P.B200(), P.TwoDAdvectingFlow(), T=np.float32)       # the B200 path is fp64
    with pytest.raises(TypeError):
        T.TracerProblem(object(), T.Grid(nx=8, Lx=1.0), T.Params(0.1, 0.1, 0.1, 0.0, 0), 0.01, "RK4", 0)

    class FakeMQG:                                                      # duck-typed CPU flow, as TAD.jl:225 accepts
        class grid:
            nx, ny, Lx, Ly = 16, 16, 2 * np.pi, 2 * np.pi
        class clock:
            dt = 0.01
        class params:
            nlayers, U = 2, np.zeros(2)
        stepped = []
        def step_until(self, t):
            self.stepped.append(t)
        def updatevars(self):
            pass
    with pytest.raises(ValueError, match="non-negative"):
        P.Problem(FakeMQG(), kappa=0.01, tracer_release_time=-1.0)      # ArgumentError, TAD.jl:234
    with pytest.raises(TypeError):
        P.MultiLayerQG.Problem(2, object())
    with pytest.raises(ValueError):
        P.MultiLayerQG.Problem(2, P.B200(), nx=16, U=np.zeros(3))      # U must be (nlayers,) or (nlayers, ny)
    with pytest.raises(ValueError):
        P.MultiLayerQG.Problem(2, P.B200(), nx=16, eta=np.zeros((8, 8)))


def test_flow_containers():
    f = P.TwoDAdvectingFlow()
    assert f.u is P.noflow and f.v is P.noflow and f.steadyflow is True        # TAD.jl:31,60-63
    u, v, steady = f
    assert u(1.0, 2.0) == 0.0 and steady
    e = P.ExpressionFlow("sin(x)", "cos(y)")
    assert e.components == ["sin(x)", "cos(y)"] and not e.steadyflow
    assert P.ExpressionFlow("1.0").components == ["1.0"]
    assert len(P.ThreeDAdvectingFlow().__iter__.__self__.__dict__) == 4


def test_no_cpu_fallback_anywhere():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(P._capi.PtfError, match="no CPU fallback"):
        P.Problem(P.B200(), P.TwoDAdvectingFlow(), nx=16)
    with pytest.raises(P._capi.PtfError, match="no CPU fallback"):
        P.MultiLayerQG.Problem(2, P.B200(), nx=16)


def test_import_order_with_torch_does_not_matter():
    """ptf_b200 (links libnccl.so.2) before torch (bundles a newer libnccl.so.2) must not break either."""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for code in ("import ptf_b200, torch; print(torch.__version__)", "import torch, ptf_b200; print(torch.__version__)"):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
