"""CPU checks that pin the MultiLayerQG oracle (oracle/mqg_oracle.py).  The reference ships no golden vectors for the
flow solver (it lives in the un-vendored GeophysicalFlows 0.16), so the restatement is pinned by analytic solutions and
invariants of the published equations: linear Rossby-wave dispersion (signs of β, U, the stretching matrix and the
transform conventions), S·S⁻¹ = I, conservation of the mean, 4th-order energy conservation of the inviscid problem."""
import numpy as np
import pytest

from oracle.mqg_oracle import MQGOracle
from oracle.ptf_oracle import irfft, make_filter, rfft


def _wave(o, kx, ky):
    X, Y = o.grid.gridpoints()
    return lambda t, om: 1e-3 * np.cos(kx * X + ky * Y - om * t)


@pytest.mark.parametrize("stepper", ["RK4", "FilteredRK4", "ETDRK4", "LSRK54"])
def test_barotropic_rossby_wave_with_doppler_shift(stepper):
    beta, U0, kx, ky = 3.0, 0.7, 2.0, 3.0
    o = MQGOracle(2, nx=32, beta=beta, U=[U0, U0], H=[0.2, 0.8], b=[-1.0, -1.2], dt=1e-3, stepper=stepper,
                  aliased_fraction=0)
    w = _wave(o, kx, ky)
    o.set_psi(np.stack([w(0, 0), w(0, 0)]))
    o.stepforward(200)
    o.updatevars()
    om = U0 * kx - beta * kx / (kx * kx + ky * ky)
    tol = 1e-12 if stepper != "LSRK54" else 1e-11
    assert np.abs(o.psi[0] - w(o.t, om)).max() / 1e-3 < tol
    assert np.abs(o.psi[1] - w(o.t, om)).max() / 1e-3 < tol


def test_baroclinic_rossby_wave_sees_the_stretching_term():
    beta, f0, kx, ky = 3.0, 1.3, 2.0, 3.0
    o = MQGOracle(2, nx=32, beta=beta, f0=f0, U=[0, 0], H=[0.5, 0.5], b=[-1.0, -1.5], dt=1e-3, stepper="RK4")
    w = _wave(o, kx, ky)
    o.set_psi(np.stack([w(0, 0), -w(0, 0)]))
    o.stepforward(200)
    o.updatevars()
    F = f0 ** 2 / (0.5 * 0.5)
    om = -beta * kx / (kx * kx + ky * ky + 2 * F)
    assert np.abs(o.psi[0] - w(o.t, om)).max() / 1e-3 < 1e-12
    assert np.abs(o.psi[1] + w(o.t, om)).max() / 1e-3 < 1e-12


def test_single_layer_rossby_wave():
    beta, kx, ky = 2.0, 1.0, 2.0
    o = MQGOracle(1, nx=32, beta=beta, U=[0.3], H=[1.0], dt=1e-3, stepper="RK4")
    w = _wave(o, kx, ky)
    o.set_psi(w(0, 0)[None])
    o.stepforward(100)
    o.updatevars()
    om = 0.3 * kx - beta * kx / (kx * kx + ky * ky)
    assert np.abs(o.psi[0] - w(o.t, om)).max() / 1e-3 < 1e-12


@pytest.mark.parametrize("nl", [2, 3, 4])
def test_stretching_matrix_inverse(nl):
    H = np.linspace(0.2, 0.5, nl)
    b = -(1.0 + 0.3 * np.arange(nl))
    o = MQGOracle(nl, nx=16, H=H, b=b, f0=1.1)
    p = o.params
    prod = np.einsum("yxij,yxjk->yxik", p.S, p.Sinv)
    eye = np.broadcast_to(np.eye(nl), prod.shape).copy()
    eye[0, 0] = 0.0       # S⁻¹(0,0) is defined as zero
    assert np.abs(prod - eye).max() < 1e-12
    # rows of F weighted by H sum to zero column-wise: the stretching term conserves the depth-integrated PV
    assert np.abs((p.H[:, None] * p.F).sum(axis=0)).max() < 1e-12


def _turbulent(nl, dt, stepper="RK4", U=None, mu=0.0, n=64):
    H = [0.2, 0.8] if nl == 2 else [0.2, 0.3, 0.5]
    b = [-1.0, -1.2] if nl == 2 else [-1.0, -1.2, -1.5]
    o = MQGOracle(nl, nx=n, beta=5, U=U if U is not None else [0.0] * nl, H=H, b=b, mu=mu, dt=dt, stepper=stepper)
    q0 = 2.0 * np.random.default_rng(1).standard_normal((nl, n, n))
    q0 = irfft(o.grid, make_filter(o.grid) * rfft(o.grid, q0))
    o.set_q(q0)
    return o


def test_inviscid_energy_conservation_converges_at_fourth_order():
    drift = []
    for dt in (5e-3, 2.5e-3):
        o = _turbulent(3, dt)
        e0 = o.energy()
        o.stepforward(int(round(0.5 / dt)))
        drift.append(abs(o.energy() - e0) / e0)
    assert drift[0] < 1e-9 and drift[1] < drift[0] / 12     # RK4: error ∝ dt⁴…dt⁵


def test_mean_pv_stays_zero_and_drag_dissipates():
    o = _turbulent(2, 2.5e-3, stepper="FilteredRK4", U=[1.0, 0.0], mu=5e-2)
    o.stepforward(40)
    assert np.abs(o.sol[:, 0, 0]).max() < 1e-9
    o2 = _turbulent(2, 2.5e-3, mu=0.5)
    e0 = o2.energy()
    o2.stepforward(40)
    assert o2.energy() < e0


def test_step_until_lands_on_the_stop_time():
    o = _turbulent(2, 2.5e-3)
    o.step_until(0.0312)
    assert o.t == 0.0312 and o.step == 13 and o.dt == 2.5e-3
