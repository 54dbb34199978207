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


def _cell_distance_by_walking(uy, height=0.001):
    """Independent definition for the test below: walk the periodic row from one maximum to the other along the shorter way
    (the reference takes the direct way when it is strictly shorter, otherwise the way over the seam) and see if the flow sinks."""
    n = len(uy)
    tops = [i for i in W.find_peaks_height(uy, height)]
    x = np.linspace(0, 2 * np.pi, n, endpoint=False)
    best = 0.0
    for a_pos, a in enumerate(tops):
        for b in tops[a_pos + 1:]:
            direct = abs(x[b] - x[a]); around = 2 * np.pi - direct
            walk = range(a, b) if direct < around else list(range(b, n)) + list(range(0, a))
            if not all(uy[k] > 0 for k in walk):
                best = max(best, min(direct, around))
    return best


def test_cell_distance_all_pairs_form_equals_walking_definition():
    rng = np.random.default_rng(5)
    x = np.linspace(0, 2 * np.pi, 96, endpoint=False)
    rows = []
    for _ in range(300):
        k = rng.integers(1, 6)
        sig = sum(rng.normal() * np.cos(m * x + rng.uniform(0, 6.3)) for m in range(1, k + 1)) + rng.normal(0, 0.5)
        rows.append(sig + 0.05 * rng.standard_normal(96) * rng.integers(0, 2))
    rows.append(np.round(rng.standard_normal(96), 1))                 # ties and plateaus
    nan_row = rows[0].copy(); nan_row[40] = np.nan; rows.append(nan_row)
    seen = set()
    for r in rows:
        st = np.zeros((3, 64, 96)); st[2, 31] = r; st[2, :31] = 0.5 * r
        got = W.cell_distance(st)
        assert got == _cell_distance_by_walking(r)
        seen.add(got > 0)
        assert W.cell_distance(st, use_avg=True) == _cell_distance_by_walking(st[2].mean(axis=0))
    assert seen == {True, False}


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


class _FakeEnv:
    """Minimal env producing obs = step counter, to check the stacking order without a GPU."""

    def __init__(self):
        from rbc_gym_b200 import spaces
        self.observation_space = spaces.Box(-np.inf, np.inf, shape=(2, 3), dtype=np.float32)
        self.action_space = spaces.Box(-1, 1, shape=(1,), dtype=np.float32)
        self.n = 0

    @property
    def unwrapped(self):
        return self

    def reset(self, seed=None, options=None):
        self.n = 10
        return np.full((2, 3), self.n, np.float32), {}

    def step(self, action):
        self.n += 1
        return np.full((2, 3), self.n, np.float32), -1.0, False, False, {}


def test_flatten_and_frame_stack_follow_gymnasium_semantics():
    env = W.FrameStackObservation(W.FlattenObservation(_FakeEnv()), 4)
    assert env.observation_space.shape == (4, 6)
    obs, _ = env.reset()
    assert obs.shape == (4, 6) and np.all(obs == 10)                     # padding_type="reset"
    for _ in range(2):
        obs, *_ = env.step(None)
    assert [row[0] for row in obs] == [10, 10, 11, 12]                   # oldest first
    envz = W.FrameStackObservation(_FakeEnv(), 3, padding_type="zero")
    obs, _ = envz.reset()
    assert obs.shape == (3, 2, 3) and np.all(obs[:2] == 0) and np.all(obs[2] == 10)


class _FakeVec:
    """CPU stand-in with the vector-env surface and SAME_STEP autoreset: env e truncates when its counter hits 3."""

    def __init__(self, n=3):
        import torch
        self.torch, self.num_envs, self.autoreset_mode = torch, n, "same_step"
        self.c = torch.zeros(n)

    def reset(self, seed=None, options=None):
        self.c = self.torch.arange(self.num_envs, dtype=self.torch.float32)
        return self.c.reshape(-1, 1, 1).expand(-1, 2, 2).clone(), {}

    def step(self, actions):
        self.c = self.c + 1
        trunc = self.c == 3
        self.c = self.torch.where(trunc, self.torch.full_like(self.c, 100.0), self.c)    # reset obs = 100
        # a reset environment reports step == 1 (rbc_sim2D_api.jl:67-68); that is what the ring buffer keys its restart on
        return self.c.reshape(-1, 1, 1).expand(-1, 2, 2).clone(), self.c * 0, trunc & False, trunc, {"step": self.torch.where(trunc, 1, 2)}


def test_vector_frame_stack_ring_buffer_restarts_on_autoreset():
    v = W.VectorFrameStack(_FakeVec(3), stack_size=3, flatten=True)
    obs, _ = v.reset()
    assert obs.shape == (3, 3, 4) and obs[2, :, 0].tolist() == [2, 2, 2]
    obs, _, _, trunc, _ = v.step(None)
    assert trunc.tolist() == [False, False, True]
    assert obs[0, :, 0].tolist() == [0, 0, 1] and obs[2, :, 0].tolist() == [100, 100, 100]   # env 2 restarted its stack
    obs, *_ = v.step(None)
    assert obs[0, :, 0].tolist() == [0, 1, 2] and obs[2, :, 0].tolist() == [100, 100, 101]
