"""On-device vector environment: thousands of independent 2D RBC environments in one batch.

Replaces the reference's process-per-env vectorisation — `gym.make_vec(..., vectorization_mode="async")`
(`example/run_vectorized.py:11-20`) and SB3's `SubprocVecEnv` (`experiments/run_sarl.py:130-153`), where
every env boots its own Julia runtime and ships pickled numpy arrays through pipes — with one batched
handle whose observations, rewards and flags are CUDA tensors that never leave the device.

Auto-reset follows the two conventions the reference's callers rely on:
  "next_step"  gymnasium 1.1.1 default: an env that truncated is reset by the *next* `step` call (its action
               is ignored, reward 0).
  "same_step"  SB3: reset inside the step that truncated; the terminal observation is kept in
               `info["final_obs"]` (rows of truncated envs; other rows are unspecified).
  "disabled"   flags only.
Resets draw a uniformly random episode of the checkpoint bank per env (`rbc_sim2D.jl:176-177`) from a device
generator keyed by (seed, global env id), so results do not depend on how envs are sharded over GPUs.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import backend, spaces


class RBCVectorEnv2D:
    def __init__(self, num_envs: int, rayleigh_number: float = 10_000, episode_length: float = 300, observation_shape=(8, 48),
                 state_shape=(64, 96), heater_segments: int = 12, heater_limit: float = 0.75, heater_duration: float = 1.5,
                 pressure: bool = False, checkpoint: Optional[str] = None, dt_solver: float = 0.03, precision: int = 32,
                 device: int = 0, autoreset_mode: str = "next_step", seed: int = 0, env_id_offset: int = 0, nan_policy: str = "raise"):
        import torch

        if autoreset_mode not in ("next_step", "same_step", "disabled"):
            raise ValueError("autoreset_mode must be next_step, same_step or disabled")
        if nan_policy not in ("raise", "reset"):
            raise ValueError("nan_policy must be raise or reset")
        self.nan_policy = nan_policy      # "raise": the reference's RuntimeError; "reset": re-initialise only the failed envs
        self.torch = torch
        self.num_envs = int(num_envs)
        self.ra = rayleigh_number
        self.episode_length = episode_length
        self.episode_steps = int(episode_length / heater_duration)
        self.observation_shape, self.state_shape = list(observation_shape), list(state_shape)
        self.temperature_difference = [1, 2]
        self.heater_segments, self.heater_limit, self.heater_duration = heater_segments, heater_limit, heater_duration
        self.include_pressure = pressure
        self.checkpoint = checkpoint
        self.autoreset_mode = autoreset_mode
        self.sim = backend.Sim2D(self.num_envs, ra=float(rayleigh_number), dt_action=float(heater_duration),
                                 obs_shape=tuple(observation_shape), state_shape=tuple(state_shape), heaters=heater_segments,
                                 heater_limit=heater_limit, dt_solver=dt_solver, episode_length=float(episode_length),
                                 precision=precision, pressure=pressure, device=device)
        self.device = self.sim.device
        ch = self.sim.channels
        self.single_action_space = spaces.Box(-1, 1, shape=(heater_segments,), dtype=np.float32)
        self.single_observation_space = spaces.Box(-np.inf, np.inf, shape=(ch, *observation_shape), dtype=np.float32)
        self.action_space = spaces.Box(-1, 1, shape=(self.num_envs, heater_segments), dtype=np.float32)
        self.observation_space = spaces.Box(-np.inf, np.inf, shape=(self.num_envs, ch, *observation_shape), dtype=np.float32)
        if checkpoint:
            self.sim.load_checkpoints(checkpoint)
        self.env_ids = torch.arange(env_id_offset, env_id_offset + self.num_envs, device=self.device, dtype=torch.int64)
        self.seed = int(seed)
        self._episode = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        self._pending = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.episode_return = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ helpers
    def _draw_checkpoints(self, ids):
        """Episode index for env `g`, episode counter `e`: a counter-based hash of (seed, g, e) mod n_episodes."""
        t = self.torch
        g, e = self.env_ids[ids], self._episode[ids]
        x = (g * 0x9E3779B97F4A7C15 + e * 0xC2B2AE3D27D4EB4F + self.seed * 0x165667B19E3779F9) & 0x7FFFFFFFFFFFFFFF
        x = (x ^ (x >> 31)) * 0x7FB5D329728EA185 & 0x7FFFFFFFFFFFFFFF
        x = x ^ (x >> 27)
        return (x % self.sim.n_episodes).to(t.int32)

    def _reset_envs(self, ids, ckpt_idx=None):
        t = self.torch
        if ids.numel() == 0:
            return
        if self.sim.n_episodes == 0:
            # noise initialisation drawn and projected on the device (a per-call generator keyed by seed and episode)
            gen = t.Generator(device=self.device)
            gen.manual_seed(self.seed * 1_000_003 + int(self._episode.max().item()))
            self.sim.noise_reset(ids.to(t.int32), kick=0.01, generator=gen)
        else:
            idx = self._draw_checkpoints(ids) if ckpt_idx is None else t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device)
            self.sim.reset_from_checkpoints(idx, env_ids=ids.to(t.int32))
        self._episode[ids] += 1
        self.episode_return[ids] = 0

    # ------------------------------------------------------------------ API
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        t = self.torch
        if seed is not None:
            self.seed = int(seed)
        ids = t.arange(self.num_envs, device=self.device)
        self._episode.zero_()
        self._reset_envs(ids, None if not options else options.get("checkpoint_idx"))
        self._pending.zero_()
        obs, nus, nuo = self.sim.observe()
        return obs, self._info(nus, nuo)

    def _info(self, nus, nuo):
        return {"nusselt_state": nus, "nusselt_obs": nuo}

    def step(self, actions):
        """actions `[num_envs, heater_segments]` (CUDA tensor) -> (obs, reward, terminated, truncated, info), all CUDA tensors."""
        t = self.torch
        pend = self._pending
        have_pending = self.autoreset_mode == "next_step" and bool(pend.any())
        obs, rew, nus, nuo, trunc, nan = self.sim.step(actions)
        bad = nan.to(t.bool) & ~pend if have_pending else nan.to(t.bool)
        truncated = trunc.to(t.bool)
        reward = rew
        info = self._info(nus, nuo)
        already_reset = None
        if bool(bad.any()):
            if self.nan_policy == "raise":
                raise RuntimeError("Error in simulation step, probably NaN values")   # rbc2D.py:170-171
            # a batch of thousands should not die with one environment: re-initialise the failed ones, zero their reward,
            # report them as truncated and in info["nan_reset"]
            self._reset_envs(bad.nonzero().flatten())
            obs, nus, nuo = self.sim.observe()
            info = self._info(nus, nuo)
            reward = rew.clone()
            reward[bad] = 0
            truncated = truncated | bad
            info["nan_reset"] = bad
            already_reset = bad
        if have_pending:
            # gymnasium NEXT_STEP semantics: for an env that truncated on the previous call this call only resets
            # it — the action is ignored, reward 0, and the returned observation is the reset observation.
            ids = pend.nonzero().flatten()
            self._reset_envs(ids)
            obs, nus, nuo = self.sim.observe()
            reward = rew.clone()
            reward[pend] = 0
            if already_reset is not None:
                reward[already_reset] = 0
            truncated = truncated & ~pend
        self.episode_return += reward.to(t.float64)
        to_reset = truncated if already_reset is None else truncated & ~already_reset
        if self.autoreset_mode == "same_step" and bool(to_reset.any()):
            # SB3 semantics: reset inside the truncating step; terminal observation kept in info["final_obs"]
            ids = to_reset.nonzero().flatten()
            info["final_obs"] = obs.clone()
            info["final_info"] = {"nusselt_state": nus.clone(), "nusselt_obs": nuo.clone(),
                                  "episode_return": self.episode_return.clone()}
            self._reset_envs(ids)
            obs, _, _ = self.sim.observe()
        self._pending = truncated.clone() if self.autoreset_mode == "next_step" else t.zeros_like(truncated)
        if already_reset is not None:
            self._pending &= ~already_reset
        terminated = t.zeros_like(truncated)
        return obs, reward, terminated, truncated, info

    def get_state(self):
        return self.sim.get_state()

    def render(self):
        """`rgb_array` frames of all environments as one uint8 CUDA tensor `[num_envs, Nz, Nx, 3]` (e.g. for logging videos
        like `example/run_wandb.py:25-59` without a host round trip per frame)."""
        return self.sim.render_rgb()

    def close(self):
        self.sim.close()


class RBCVectorEnv3D:
    """The 3D twin: B independent `RayleighBenardConvection3D-v0` environments as one on-device batch — what
    `SubprocVecEnv([make_env(i, ...) for i in range(n_envs)])` of `experiments/run_sarl.py:130-153` becomes.

    Observations are the full state `(num_envs, 4, Nz, Ny, Nx)` float32 CUDA tensors (`rbc3D.py:229-232`), actions
    `(num_envs, 8, 8)` in [-1, 1] with `action[:, i, j]` <-> patch i along x, j along y, reward = -Nu, `info["nusselt"]`.
    Resets come from a `3D_ckpt_ra*.h5` bank (device gather) or from the noise initialisation; autoreset as in the 2D class."""

    def __init__(self, num_envs: int, rayleigh_number: float = 2500, prandtl_number: float = 0.7, domain=(2, 4 * np.pi, 4 * np.pi),
                 state_shape=(16, 32, 32), temperature_difference=(1, 2), heater_segments: int = 8, heater_limit: float = 0.9,
                 heater_duration: float = 0.125, episode_length: float = 300, dt_solver: float = 0.01, checkpoint: Optional[str] = None,
                 precision: int = 32, device: int = 0, autoreset_mode: str = "next_step", seed: int = 0, env_id_offset: int = 0):
        import torch

        if autoreset_mode not in ("next_step", "same_step", "disabled"):
            raise ValueError("autoreset_mode must be next_step, same_step or disabled")
        self.torch = torch
        self.num_envs = int(num_envs)
        self.is_3d = True
        self.ra, self.episode_length = rayleigh_number, episode_length
        self.state_shape, self.domain = tuple(state_shape), tuple(domain)
        self.temperature_difference = list(temperature_difference)
        self.heater_segments, self.heater_limit, self.heater_duration = heater_segments, heater_limit, heater_duration
        self.checkpoint, self.autoreset_mode = checkpoint, autoreset_mode
        self.sim = backend.Sim3D(self.num_envs, ra=float(rayleigh_number), pr=float(prandtl_number), domain=self.domain,
                                 state_shape=self.state_shape, temperature_difference=tuple(temperature_difference),
                                 heaters=heater_segments, heater_limit=heater_limit, heater_duration=heater_duration, dt_solver=dt_solver,
                                 episode_length=float(episode_length), precision=precision, device=device)
        self.device = self.sim.device
        self.single_action_space = spaces.Box(-1, 1, shape=(heater_segments, heater_segments), dtype=np.float32)
        self.single_observation_space = spaces.Box(-np.inf, np.inf, shape=(4, *self.state_shape), dtype=np.float32)
        self.action_space = spaces.Box(-1, 1, shape=(self.num_envs, heater_segments, heater_segments), dtype=np.float32)
        self.observation_space = spaces.Box(-np.inf, np.inf, shape=(self.num_envs, 4, *self.state_shape), dtype=np.float32)
        if checkpoint:
            from ..h5lite import load_checkpoint_3d
            self.sim.load_checkpoints(load_checkpoint_3d(checkpoint))
        self.env_ids = torch.arange(env_id_offset, env_id_offset + self.num_envs, device=self.device, dtype=torch.int64)
        self.seed = int(seed)
        self._episode = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        self._pending = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.episode_return = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)

    _draw_checkpoints = RBCVectorEnv2D._draw_checkpoints

    def _reset_envs(self, ids, ckpt_idx=None):
        t = self.torch
        if ids.numel() == 0:
            return
        if self.sim.n_episodes == 0:
            gen = t.Generator(device=self.device)
            gen.manual_seed(self.seed * 1_000_003 + int(self._episode.max().item()))
            self.sim.noise_reset(ids.to(t.int32), kick=0.01, generator=gen)
        else:
            idx = self._draw_checkpoints(ids) if ckpt_idx is None else t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device)
            self.sim.reset_from_checkpoints(idx, env_ids=ids.to(t.int32))
        self._episode[ids] += 1
        self.episode_return[ids] = 0

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        t = self.torch
        if seed is not None:
            self.seed = int(seed)
        self._episode.zero_()
        self._reset_envs(t.arange(self.num_envs, device=self.device), None if not options else options.get("checkpoint_idx"))
        self._pending.zero_()
        obs, nu = self.sim.observe()
        return obs, {"nusselt": nu}

    def step(self, actions):
        t = self.torch
        pend = self._pending
        have_pending = self.autoreset_mode == "next_step" and bool(pend.any())
        obs, rew, nu, trunc, nan = self.sim.step(actions)
        bad = nan.to(t.bool) & ~pend if have_pending else nan.to(t.bool)
        if bool(bad.any()):
            raise RuntimeError("Error in simulation step, probably NaN values")    # rbc3D.py:207-212
        truncated, reward, info = trunc.to(t.bool), rew, {"nusselt": nu}
        if have_pending:
            self._reset_envs(pend.nonzero().flatten())
            obs, nu = self.sim.observe()
            info = {"nusselt": nu}
            reward = rew.clone()
            reward[pend] = 0
            truncated = truncated & ~pend
        self.episode_return += reward.to(t.float64)
        if self.autoreset_mode == "same_step" and bool(truncated.any()):
            info["final_obs"] = obs.clone()
            info["final_info"] = {"nusselt": nu.clone(), "episode_return": self.episode_return.clone()}
            self._reset_envs(truncated.nonzero().flatten())
            obs, _ = self.sim.observe()
        self._pending = truncated.clone() if self.autoreset_mode == "next_step" else t.zeros_like(truncated)
        return obs, reward, t.zeros_like(truncated), truncated, info

    def close(self):
        self.sim.close()
