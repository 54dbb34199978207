"""Pin the CPU oracle to everything the reference ships for this path.

The reference has no tests and cannot run here (no Julia); what pins the scheme are the
checkpoint files written by the real Julia simulation, exact closed-form answers, and the
survey's independently written restatement (tests/golden/survey_crosscheck.json).
"""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

GOLD = Path(__file__).parent / "golden"
DX, DZ = 2 * math.pi / 96, 2 / 64
ACTIONS = {"zero": np.zeros(12), "cos": np.cos(2 * np.pi * np.arange(12) / 12), "ramp": np.linspace(-1, 1, 12)}


def divergence(u, w):
    return (np.roll(u, -1, axis=-1) - u) / DX + (w[..., 1:, :] - w[..., :-1, :]) / DZ


def test_checkpoints_divergence_free_under_staggering(ckpt_ra1e5, ckpt_ra1e4):
    # u on the LEFT face of cell i, w on the BOTTOM face of cell k
    for c in (ckpt_ra1e5, ckpt_ra1e4):
        assert np.abs(divergence(c.u, c.w)).max() < 2e-14
        assert np.abs(c.w[:, 0]).max() == 0 and np.abs(c.w[:, -1]).max() == 0
        wrong = (c.u - np.roll(c.u, 1, axis=-1)) / DX + (c.w[:, 1:] - c.w[:, :-1]) / DZ
        assert np.abs(wrong).max() > 0.1


def test_heater_profile_known_answers():
    P = O.make_params(1e5)
    assert np.all(O.heater_profile(P, np.full(12, 0.37)) == 2.0)               # all-equal -> T = 2 exactly
    a = np.zeros(12); a[0] = 1.0
    T = O.heater_profile(P, a)
    np.testing.assert_allclose(T[:8], 2 + 0.75 * 11 / 12, rtol=0, atol=1e-15)   # one-hot: M = lim/12, K2 = 1
    np.testing.assert_allclose(T[8:], 2 - 0.75 / 12, rtol=0, atol=1e-15)
    T = O.heater_profile(P, np.array([1, -1] * 6, float))
    np.testing.assert_allclose(T[:8], 2.75); np.testing.assert_allclose(T[8:16], 1.25)
    T = O.heater_profile(P, ACTIONS["cos"])                                       # survey table header
    np.testing.assert_allclose(T[::8][:7], [2.75, 2.6495190528, 2.375, 2, 1.625, 1.3504809472, 1.25], atol=1e-9)
    # at Nx = 96 the cubic blends never fire: piecewise constant, 8 cells per segment
    assert all(len(set(T[8 * s:8 * s + 8])) == 1 for s in range(12))
    # at Nx = 192 they do fire on the first/last cell of each segment
    P2 = O.make_params(1e5, nx=192, nz=128)
    T2 = O.heater_profile(P2, ACTIONS["ramp"])
    assert len(set(np.round(T2[16:32], 12))) == 3


def test_conduction_state_is_fixed_point_and_nusselt_quirk():
    # u = w = 0, b linear, zero action: exact fixed point; reference formula gives Nu = dz / 8 dz
    P = O.make_params(1e5)
    z = (np.arange(64) + 0.5) * DZ
    b = np.repeat((2 - z / 2)[:, None], 96, axis=1)
    u, w = np.zeros((64, 96)), np.zeros((65, 96))
    r = O.step(P, b, u, w, np.zeros(12), O.substep_schedule(1.0))
    assert np.abs(r["b"] - b).max() < 1e-13 and np.abs(r["u"]).max() < 1e-13 and np.abs(r["w"]).max() < 1e-13
    ns, no = O.nusselt_state_obs(P, b, u, w)
    assert ns == pytest.approx(0.03125, abs=1e-12) and no == pytest.approx(0.25, abs=1e-12)


def test_survey_crosscheck_table(ckpt_ra1e5):
    tab = json.loads((GOLD / "survey_crosscheck.json").read_text())
    P = O.make_params(1e5)
    c = ckpt_ra1e5
    for row in tab["file"]:
        ep = row["ep"]
        ns, no = O.nusselt_state_obs(P, c.b[ep], c.u[ep], c.w[ep])
        assert ns == pytest.approx(row["nu_state"], abs=2e-10) and no == pytest.approx(row["nu_obs"], abs=2e-10)
        assert O.kinetic_energy(c.u[ep], c.w[ep]) == pytest.approx(row["ke"], abs=2e-12)
        assert c.b[ep][0, 0] == pytest.approx(row["b00"], abs=1e-10) and c.w[ep][32, 48] == pytest.approx(row["w3248"], abs=1e-10)
    for row in tab["rows"]:
        ep = row["ep"]
        r = O.step(P, c.b[ep], c.u[ep], c.w[ep], ACTIONS[row["action"]], O.substep_schedule(row["dt"]))
        ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"])
        assert ns == pytest.approx(row["nu_state"], abs=2e-9), row
        assert no == pytest.approx(row["nu_obs"], abs=2e-9), row
        assert O.kinetic_energy(r["u"], r["w"]) == pytest.approx(row["ke"], abs=1e-11), row
        assert r["b"][0, 0] == pytest.approx(row["b00"], abs=1e-9)
        assert r["b"][32, 48] == pytest.approx(row["b3248"], abs=1e-9)
        assert r["u"][10, 20] == pytest.approx(row["u1020"], abs=1e-9)
        assert r["w"][32, 48] == pytest.approx(row["w3248"], abs=1e-9)
        assert np.abs(divergence(r["u"], r["w"])).max() < 1e-13


def test_poisson_formulations_and_buoyancy_split_agree(ckpt_ra1e5):
    c = ckpt_ra1e5
    a = ACTIONS["ramp"]
    dts = O.substep_schedule(0.09)
    base = O.step(O.make_params(1e5), c.b[3], c.u[3], c.w[3], a, dts, want_pressure=True)
    dense = O.step(O.make_params(1e5, poisson_mode=1), c.b[3], c.u[3], c.w[3], a, dts, want_pressure=True)
    direct = O.step(O.make_params(1e5, split_phy=False), c.b[3], c.u[3], c.w[3], a, dts)
    for k in "buw":
        assert np.abs(base[k] - dense[k]).max() < 1e-13      # FFT/tridiagonal == literal DFT/DCT eigen-division
        assert np.abs(base[k] - direct[k]).max() < 1e-13     # pHY' split == buoyancy in Gw after projection
    assert np.abs(base["pnhs"] - dense["pnhs"]).max() < 1e-12
    assert abs(base["pnhs"].mean()) < 1e-14                  # zero-mean gauge


def test_ra1e4_fixed_point_band(ckpt_ra1e4):
    """The Ra=1e4 reference states sit on a slowly damped oscillation (period ~6 time units, amplitude
    <= 2.5e-4 in Nu) around a fixed point of the full discrete scheme.  The checkpoints span
    Nu_state [3.997490, 3.997789], KE [0.0974366, 0.0974397]; a trajectory of the restated scheme started
    from one of them must stay within that band widened by the oscillation amplitude, and its period mean
    must match the ensemble mean.  (A scheme with a 2nd-order advecting velocity lands at Nu 4.0002,
    KE 0.0973635 — 10x outside these tolerances.)"""
    P = O.make_params(1e4)
    c = ckpt_ra1e4
    nus = []
    for ep in range(c.num_episodes):
        ns, no = O.nusselt_state_obs(P, c.b[ep], c.u[ep], c.w[ep])
        nus.append(ns)
        assert 3.997490 - 1e-6 <= ns <= 3.997789 + 1e-6 and 4.192437 - 1e-6 <= no <= 4.192916 + 1e-6
        assert 0.0974366 - 1e-7 <= O.kinetic_energy(c.u[ep], c.w[ep]) <= 0.0974397 + 1e-7
    b, u, w = c.b[0], c.u[0], c.w[0]
    traj = []
    for _ in range(20):                                   # 6 time units = one oscillation period
        r = O.step(P, b, u, w, np.zeros(12), O.substep_schedule(0.3))
        b, u, w = r["b"], r["u"], r["w"]
        ns, no = O.nusselt_state_obs(P, b, u, w)
        traj.append(ns)
        assert 3.997490 - 2e-4 <= ns <= 3.997789 + 2e-4 and 4.192437 - 2e-4 <= no <= 4.192916 + 2e-4
        assert 0.0974366 - 1e-6 <= O.kinetic_energy(u, w) <= 0.0974397 + 1e-6
    assert abs(np.mean(traj) - 3.997655) < 1e-4           # ensemble mean of the 40 reference states
    assert abs(np.mean(traj) - np.mean(nus)) < 1.5e-4


def test_ra1e4_tracer_residual_pins_wall_order(ckpt_ra1e4):
    """Near-wall tracer tendency of a fixed-point state is small only with the 5/3/1 face rule."""
    P = O.make_params(1e4)
    c = ckpt_ra1e4
    Gb, _, _ = O.tendencies(P, c.b[0], c.u[0], c.w[0], np.zeros(12))
    assert np.abs(Gb[:2]).max() < 1e-3 and np.abs(Gb[-2:]).max() < 1e-3
    assert np.abs(Gb).max() < 2e-3


def test_substep_schedule():
    s = O.substep_schedule(1.0)
    assert len(s) == 34 and np.all(s[:33] == 0.03) and s[-1] == pytest.approx(0.01, abs=1e-12)
    assert len(O.substep_schedule(1.5)) == 50 and len(O.substep_schedule(0.3)) == 10
