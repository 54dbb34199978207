"""Wrappers: the pure-Python restatement against scipy, and the fused CUDA epilogue (through
the host emulator of the kernel source) against the pure-Python restatement."""
import numpy as np
import pytest

from rbc_gym_b200 import wrappers as W
from tests.emu import emu


def test_find_peaks_matches_scipy(ckpt_ra1e5):
    sp = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(0)
    rows = [ckpt_ra1e5.w[e, 31] for e in range(20)] + [rng.standard_normal(96) for _ in range(20)]
    rows.append(np.array([0, 1, 1, 1, 0, 2, 2, 0, -1, 3, 3, 3, 3, 1] + [0.0] * 82))     # plateaus
    rows.append(np.round(rng.standard_normal(96), 1))                                  # many ties
    for r in rows:
        np.testing.assert_array_equal(W.find_peaks_height(r, 0.001), sp.find_peaks(r, height=0.001)[0])


def test_cell_distance_cases():
    x = np.linspace(0, 2 * np.pi, 96, endpoint=False)
    st = np.zeros((3, 64, 96))
    st[2, 31] = np.cos(2 * (x - 0.5))                     # two rolls: peaks half a domain apart, negative flow between
    assert W.cell_distance(st) == pytest.approx(np.pi, abs=2 * np.pi / 96)
    st[2, 31] = np.cos(x - 1.0)                            # one roll pair: a single peak
    assert W.cell_distance(st) == 0
    st[2, 31] = 2 + np.cos(4 * x + 0.3)                    # several peaks but uy > 0 everywhere: same cell
    assert W.cell_distance(st) == 0


def test_normalizers_match_reference_formulas():
    obs = np.random.default_rng(1).uniform(-1, 2.5, (3, 8, 48)).astype(np.float32)
    out = W.normalize_observation(obs.copy(), 0.75)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out[0], 2 * (obs[0] - 1) / 1.75 - 1, rtol=1e-6)
    np.testing.assert_allclose(out[1], obs[1] / 1.3, rtol=1e-6, atol=1e-7)
    s = 0.1 * 1e5 ** 0.4
    assert W.normalize_reward(-5.0, 1e5) == pytest.approx((-5 + s) / (s - 1))
    assert W.shape_reward(0.4, np.pi / 2, 0.1) == pytest.approx(0.9 * 0.4 + 0.1 * 0.5)


@pytest.mark.parametrize("precision", [64, 32])
def test_fused_epilogue_matches_python_wrappers(ckpt_ra1e5, precision):
    c = ckpt_ra1e5
    eps = [0, 7, 16, 3]
    st = emu.pack(c.b[eps], c.u[eps], c.w[eps])
    acts = np.random.default_rng(2).uniform(-1, 1, (4, 12)).astype(np.float32)
    plain = emu.step(st, acts, 1e5, 0.09, precision=precision)
    wr = emu.HostWrappers()
    wr.normalize_obs, wr.obs_clip, wr.obs_maxval = 1, 0, 1.0
    for ch, (lo, hi) in enumerate([(1.0, 2.75), (-1.3, 1.3), (-1.3, 1.3), (-1.3, 1.3)]):
        wr.obs_lo[ch], wr.obs_hi[ch] = lo, hi
    wr.normalize_reward, wr.reward_scale = 1, 0.1 * 1e5 ** 0.4
    wr.shaping, wr.shaping_weight = 1, 0.1
    fused = emu.step(st, acts, 1e5, 0.09, precision=precision, wrappers=wr)
    np.testing.assert_array_equal(fused["state"], plain["state"])
    b, u, w = emu.unpack(plain["state"].astype(np.float64))
    for j in range(4):
        state = np.stack([b[j], u[j], w[j, :-1]]).astype(np.float32)
        cd = W.cell_distance(state)
        assert fused["cell_dist"][j] == pytest.approx(cd, abs=1e-12)
        r = W.shape_reward(W.normalize_reward(float(-plain["nu_obs"][j]), 1e5), cd, 0.1)
        assert fused["reward"][j] == pytest.approx(r, rel=1e-6)
        np.testing.assert_allclose(fused["obs"][j], W.normalize_observation(plain["obs"][j].copy(), 0.75), rtol=2e-6, atol=2e-7)
    assert np.ptp(fused["cell_dist"]) > 0            # the four flows are not all alike
