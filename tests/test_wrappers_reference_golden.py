"""Wrappers against golden vectors produced by EXECUTING the reference's own wrapper modules
(`tools/make_wrapper_golden.py` imports `/root/reference/src/rbc_gym/wrappers/*.py` through stub packages and stores the
outputs in `tests/golden/wrappers_reference.json`).  Checked here: the pure-Python restatement (`rbc_gym_b200.wrappers`)
and the fused CUDA epilogue (kernel source emulated on the host) reproduce them."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

from rbc_gym_b200 import wrappers as W  # noqa: E402
from tests.emu import emu  # noqa: E402

GOLD = json.loads((ROOT / "tests/golden/wrappers_reference.json").read_text())


@pytest.fixture(scope="module")
def inputs():
    from make_wrapper_golden import golden_inputs
    return golden_inputs()


def test_cell_distance_matches_the_reference_code(inputs):
    seen = set()
    for name, st in inputs:
        g = GOLD["cell_distance"][name]
        assert W.cell_distance(st) == pytest.approx(g["cell_dist"], abs=1e-12), name
        assert W.cell_distance(st, use_avg=True) == pytest.approx(g["cell_dist_avg"], abs=1e-12), name
        assert W.shape_reward(-5.0, W.cell_distance(st), 0.1) == pytest.approx(g["shaped_reward_from_-5"], abs=1e-12)
        seen.add(round(g["cell_dist"], 6))
    assert len(seen) >= 4                                  # the cases cover several distinct outcomes (0, pi, in between)


def test_normalisers_match_the_reference_code():
    g = GOLD["normalize_reward"]
    assert 0.1 * 1e5 ** 0.4 == pytest.approx(g["scale_2d_ra1e5"], rel=1e-14)
    assert W.normalize_reward(-5.0, 1e5, dims=2) == pytest.approx(g["reward_2d_of_-5"], rel=1e-13)
    assert W.normalize_reward(-1.8, 2500, dims=3) == pytest.approx(g["reward_3d_of_-1.8"], rel=1e-13)
    g = GOLD["normalize_observation"]
    rng = np.random.default_rng(g["obs_seed"])
    obs = rng.uniform(-1.5, 3.0, (3, 8, 48)).astype(np.float32)
    plain = W.normalize_observation(obs.copy(), 0.75)
    assert float(plain.astype(np.float64).sum()) == pytest.approx(g["plain_sum"], rel=1e-6)
    np.testing.assert_allclose(plain[:, 3, 5], g["plain_sample"], rtol=1e-6)
    clip = W.normalize_observation(obs.copy(), 0.75, maxval=2, clip=True)
    np.testing.assert_allclose(clip[:, 3, 5], g["clip_maxval2_sample"], rtol=1e-6)
    assert clip.min() == g["clip_min"] and clip.max() == g["clip_max"]
    assert W.RBCNormalizeObservation._get_u_limit_3d(2500) == pytest.approx(g["u_limit_3d_ra2500"], rel=1e-13)
    obs3 = rng.uniform(-1.5, 3.0, (4, 4, 4, 4)).astype(np.float32)
    o3 = W.normalize_observation(obs3.copy(), 0.9, u_limit=g["u_limit_3d_ra2500"])
    np.testing.assert_allclose(o3[:, 1, 2, 3], g["obs3_sample"], rtol=1e-6)


def test_fused_cuda_epilogue_matches_the_reference_cell_distance(inputs):
    """The in-kernel peak scan (rbc2d_core.h cell_distance, run through the host emulator with zero RK3 steps is not
    possible — the epilogue runs after a step — so the states are stepped by 0.03 and compared with the reference code
    applied to the stepped state via the Python restatement, which the first test ties to the reference)."""
    states = [st for name, st in inputs if name.startswith("ckpt")][:6]
    st = np.stack(states).astype(np.float64)
    w_full = np.concatenate([st[:, 2], np.zeros((len(states), 1, 96))], axis=1)
    packed = emu.pack(st[:, 0], st[:, 1], w_full)
    wr = emu.HostWrappers()
    wr.shaping, wr.shaping_weight = 1, 0.1
    e = emu.step(packed, np.zeros((len(states), 12), np.float32), 1e5, 0.03, precision=64, wrappers=wr)
    b, u, w = emu.unpack(e["state"])
    for j in range(len(states)):
        ref_cd = W.cell_distance(np.stack([b[j], u[j], w[j, :-1]]))
        assert e["cell_dist"][j] == pytest.approx(ref_cd, abs=1e-12)
        assert e["reward"][j] == pytest.approx(W.shape_reward(float(-e["nu_obs"][j]), ref_cd, 0.1), rel=1e-6)
