"""Fused vector-env semantics of the step kernel (`VecIO` / `env_epilogue` in rbc2d_core.h), checked on the CPU with the
same-source emulator against a plain restatement of what the vector wrappers do around single environments:

  next_step  gymnasium 1.1.1 (`gym.make_vec`, example/run_vectorized.py:11-31): an env that truncated is reset by the NEXT
             step call — its action is ignored, reward 0, the reset observation is returned;
  same_step  SB3 `SubprocVecEnv` (experiments/run_sarl.py:130-153): reset inside the truncating step, terminal observation
             kept aside.

The restatement below steps every env with the PLAIN emulated kernel and performs the resets in numpy (this is, line for
line, what `rbc_gym_b200/envs/vector.py` did in Python before the fusion); the fused kernel must agree bitwise."""
import numpy as np
import pytest

from tests.emu import emu

RA, DT, EP_LEN = 1e5, 0.09, 0.27          # 3 RK3 steps per action, episodes of 3 actions


def bank_of(ckpt):
    return emu.pack(ckpt.b, ckpt.u, ckpt.w)[:6]


class PlainVec:
    """Vector wrapper semantics over the plain step (no VecIO)."""

    def __init__(self, B, bank, mode, nan_reset, seed, id_offset, precision, cluster=0, pressure=False):
        self.v = emu.VecEmu2D(B, bank, RA, DT, mode, nan_reset, seed, id_offset, precision, episode_length=EP_LEN)   # for draw() only
        self.B, self.bank, self.mode, self.nan_reset, self.precision = B, bank, mode, nan_reset, precision
        self.cluster, self.pressure = cluster, pressure
        self.dtype = np.float64 if precision == 64 else np.float32

    def _reset(self, e):
        self.state[e] = self.bank[self.v.draw(e, int(self.episode[e]))].astype(self.dtype)
        self.episode[e] += 1
        self.ret[e] = 0.0
        self.t[e] = 0.0
        self.stepc[e] = 1

    def _plain(self, state, actions, dt, t0, reset=False):
        if self.cluster:          # reset of a pressure=True env: the set! projection + pressure refresh of rbc2d_reset_from_checkpoints_dev
            return emu.stepx(state, actions, RA, dt, cl=self.cluster, precision=self.precision, nxt_global=self.precision == 64,
                             episode_length=EP_LEN, t0=t0, split=self.pressure, project_first=reset and self.pressure, nsub=0 if reset else -1)
        assert not self.pressure
        return emu.step(state, actions, RA, dt, precision=self.precision, episode_length=EP_LEN, t0=t0)

    def _observe(self, e):
        r = self._plain(self.state[e:e + 1], np.zeros((1, 12), np.float32), DT, np.zeros(1), reset=True) if self.cluster else \
            emu.step(self.state[e:e + 1], np.zeros((1, 12), np.float32), RA, 0.0, precision=self.precision, episode_length=EP_LEN)
        if self.pressure:
            self.state[e] = r["state"][0]                 # the projection is part of the reset
        return r["obs"][0], r["nu_state"][0], r["nu_obs"][0]

    def reset(self):
        B = self.B
        self.state = np.zeros((B, 18528), self.dtype)
        self.episode, self.ret, self.t, self.stepc = np.zeros(B, np.int64), np.zeros(B), np.zeros(B), np.ones(B, np.int32)
        self.pending = np.zeros(B, bool)
        for e in range(B):
            self._reset(e)

    def step(self, actions):
        B = self.B
        r = self._plain(self.state, actions, DT, self.t)
        obs, rew, nus, nuo = r["obs"].copy(), r["reward"].copy(), r["nu_state"].copy(), r["nu_obs"].copy()
        trunc, nan = r["truncated"].astype(bool), r["nan"].astype(bool)
        out = dict(final_obs=np.zeros_like(obs), final_nu_state=np.zeros(B), final_nu_obs=np.zeros(B), final_return=np.zeros(B))
        new_state, new_t, new_step = r["state"], r["t"], self.stepc + 1
        pend = self.pending.copy() if self.mode == 1 else np.zeros(B, bool)
        reset_now = np.zeros(B, bool)
        for e in range(B):
            if pend[e]:
                self._reset(e)
                obs[e], nus[e], nuo[e] = self._observe(e)
                rew[e], trunc[e], nan[e] = 0.0, False, False
                self.pending[e] = False
                continue
            self.state[e], self.t[e], self.stepc[e] = new_state[e], new_t[e], new_step[e]
            bad = self.nan_reset and nan[e]
            if bad:
                rew[e] = 0.0
            self.ret[e] += float(rew[e])
            done = trunc[e] or bad
            if bad or (self.mode == 2 and trunc[e]):
                out["final_obs"][e], out["final_nu_state"][e], out["final_nu_obs"][e], out["final_return"][e] = obs[e], nus[e], nuo[e], self.ret[e]
                self._reset(e)
                obs[e], nus[e], nuo[e] = self._observe(e)
                reset_now[e] = True
            trunc[e] = done
            self.pending[e] = self.mode == 1 and done and not reset_now[e]
        out.update(obs=obs, reward=rew, nu_state=nus, nu_obs=nuo, truncated=trunc.astype(np.int32), nan=nan.astype(np.int32),
                   t=self.t.copy(), step=self.stepc.copy(), episode_return=self.ret.copy(), reset_now=reset_now)
        return out


def test_checkpoint_draw_matches_the_tensor_expression():
    torch = pytest.importorskip("torch")
    v = emu.VecEmu2D(1, np.zeros((20, 18528)), RA, DT, 1, seed=7)
    g, e = torch.arange(5, dtype=torch.int64) + 4090, torch.tensor([0, 1, 2, 3, 40])
    x = (g * 0x9E3779B97F4A7C15 + e * 0xC2B2AE3D27D4EB4F + 7 * 0x165667B19E3779F9) & 0x7FFFFFFFFFFFFFFF
    x = (x ^ (x >> 31)) * 0x7FB5D329728EA185 & 0x7FFFFFFFFFFFFFFF
    x = x ^ (x >> 27)
    assert [v.draw(int(a), int(b)) for a, b in zip(g, e)] == (x % 20).tolist()


@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("precision", [32, 64])
def test_fused_autoreset_equals_the_wrapper_semantics(ckpt_ra1e5, mode, precision):
    bank = bank_of(ckpt_ra1e5)
    B, seed, off = 3, 11, 100
    fused = emu.VecEmu2D(B, bank, RA, DT, mode, False, seed, off, precision, episode_length=EP_LEN)
    plain = PlainVec(B, bank, mode, False, seed, off, precision)
    fused.reset(); plain.reset()
    assert np.array_equal(fused.state, plain.state)
    # stagger the clocks so that the envs truncate on different calls
    for v in (fused, plain):
        v.t[1] = DT
    rng = np.random.default_rng(0)
    n_trunc = 0
    for it in range(9):
        a = rng.uniform(-1, 1, (B, 12)).astype(np.float32)
        f, p = fused.step(a), plain.step(a)
        for k in ("obs", "reward", "nu_state", "nu_obs", "truncated", "nan", "t", "step", "episode_return"):
            assert np.array_equal(f[k], p[k]), (it, k, f[k] if f[k].ndim == 1 else "", p[k] if p[k].ndim == 1 else "")
        m = p["reset_now"]
        for k in ("final_obs", "final_nu_state", "final_nu_obs", "final_return"):
            assert np.array_equal(f[k][m], p[k][m]), (it, k)
        assert np.array_equal(fused.state, plain.state) and np.array_equal(fused.episode, plain.episode)
        assert np.array_equal(fused.pending.astype(bool), plain.pending)
        n_trunc += int(p["truncated"].sum())
    assert n_trunc >= 4                                   # several episodes ended inside the rollout
    if mode == 0:
        assert fused.episode.max() == 1                   # disabled: flags only, nobody was reset


@pytest.mark.parametrize("mode", [1, 2])
def test_nan_reset_policy_reinitialises_only_the_failed_environment(ckpt_ra1e5, mode):
    bank = bank_of(ckpt_ra1e5)
    B = 3
    fused = emu.VecEmu2D(B, bank, RA, DT, mode, True, 5, 0, 32, episode_length=EP_LEN)
    plain = PlainVec(B, bank, mode, True, 5, 0, 32)
    fused.reset(); plain.reset()
    a = np.zeros((B, 12), np.float32)
    fused.step(a); plain.step(a)
    fused.state[1, 3000] = np.nan
    plain.state[1, 3000] = np.nan
    f, p = fused.step(a), plain.step(a)
    assert f["nan"].tolist() == [0, 1, 0] and f["truncated"].tolist() == [0, 1, 0] and f["reward"][1] == 0 and fused.nan_count[0] == 1
    assert np.isfinite(f["obs"]).all() and np.isfinite(fused.state).all() and fused.t[1] == 0 and fused.episode[1] == 2
    for k in ("obs", "reward", "nu_state", "nu_obs", "truncated", "nan", "t", "step", "episode_return"):
        assert np.array_equal(f[k], p[k]), k
    assert not fused.pending.any()


@pytest.mark.parametrize("mode,precision,cl,pressure", [(1, 32, 2, False), (2, 32, 2, False), (2, 64, 2, False), (1, 32, 4, True), (2, 32, 2, True)])
def test_cluster_kernel_fused_autoreset(ckpt_ra1e5, mode, precision, cl, pressure):
    """The same semantics in the cluster kernel (rbc2dx_core.h): the reset gathers slab + halo rows per CTA, the NaN decision is
    a cluster-wide sum; pressure=True adds the set! projection and the refreshed pressure channels of a reset."""
    bank = bank_of(ckpt_ra1e5)
    B, seed, off = 2, 3, 7
    fused = emu.VecEmu2D(B, bank, RA, DT, mode, False, seed, off, precision, episode_length=EP_LEN, cluster=cl, pressure=pressure)
    plain = PlainVec(B, bank, mode, False, seed, off, precision, cluster=cl, pressure=pressure)
    fused.reset(); plain.reset()
    if pressure:                                          # reset() of a pressure env projects (library: rbc2d_vec_reset_dev)
        for e in range(B):
            plain._observe(e)
        fused.state[:] = plain.state
    fused.t[1] = plain.t[1] = DT
    rng = np.random.default_rng(1)
    n_trunc = 0
    for it in range(7):
        a = rng.uniform(-1, 1, (B, 12)).astype(np.float32)
        f, p = fused.step(a), plain.step(a)
        for k in ("obs", "reward", "nu_state", "nu_obs", "truncated", "nan", "t", "step", "episode_return"):
            assert np.array_equal(f[k], p[k]), (it, k)
        m = p["reset_now"]
        for k in ("final_obs", "final_nu_state", "final_nu_obs", "final_return"):
            assert np.array_equal(f[k][m], p[k][m]), (it, k)
        assert np.array_equal(fused.state, plain.state)
        n_trunc += int(p["truncated"].sum())
    assert n_trunc >= 3
