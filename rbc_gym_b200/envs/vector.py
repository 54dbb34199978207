"""On-device vector environments: thousands of independent RBC environments in one batch.

Replaces the reference's process-per-env vectorisation — `gym.make_vec(..., vectorization_mode="async")`
(`example/run_vectorized.py:11-20`) and SB3's `SubprocVecEnv` (`experiments/run_sarl.py:130-153`), where
every env boots its own Julia runtime and ships pickled numpy arrays through pipes — with one batched
handle whose observations, rewards and flags are CUDA tensors that never leave the device.

`step()` is ONE kernel launch: the march, the observation / Nusselt / reward epilogue, truncation AND the auto-reset
run inside `rbc2d_vec_step_dev` / `rbc3d_vec_step_dev` (`VecIO` in `csrc/rbc2d_core.h`), with no host synchronisation:
the reference's NaN error (`rbc2D.py:170-171`) is raised by the call AFTER the failing one by default
(`nan_policy="raise_deferred"`: the step kernel counts failures on the device, the count is copied to pinned memory behind
the step and read one call later), `"raise"` reads it in the failing call itself at the price of one host synchronisation per
step, `"reset"` re-initialises only the failed environments inside the kernel.  Auto-reset follows the two conventions the
reference's callers rely on:
  "next_step"  gymnasium 1.1.1 default: an env that truncated is reset by the *next* `step` call (its action
               is ignored, it is not marched, reward 0, the reset observation is returned).
  "same_step"  SB3: reset inside the step that truncated; the terminal observation is kept in
               `info["final_obs"]` (rows of truncated envs; other rows are stale).
  "disabled"   flags only.
Resets draw a uniformly random episode of the checkpoint bank per env (`rbc_sim2D.jl:176-177`) from a counter hash of
(seed, global env id, episode), so results do not depend on how envs are sharded over GPUs.  Without a checkpoint
bank resets are the reference's noise initialisation (`rbc_sim2D.jl:163-171`), drawn from the same kind of counter hash
and projected on the device; that path cannot run inside the step kernel and costs one host read per step.

`info` carries the reference's keys (`rbc2D.py:202-212`, `rbc3D.py:241-245`) as batch tensors: `t`, `step`,
`nusselt_state` / `nusselt_obs` (2D) or `nusselt` (3D), plus `episode_return`, `nan`, and — opt-in, `info_state=True`,
it is 74 KB per env and step — `state`.

Returned tensors are copies by default (`copy=True`, like `vector_kwargs={"copy": True}` of the reference example);
with `copy=False` they are the handle's own buffers, which the next `step`/`reset` overwrites in place.
"""
from __future__ import annotations

import collections
import math
from typing import Optional

import numpy as np

from .. import backend, spaces

_M63 = 0x7FFFFFFFFFFFFFFF


def _as_i64(c: int) -> int:
    """A 64-bit constant as the signed value int64 tensor arithmetic wraps to."""
    c &= 0xFFFFFFFFFFFFFFFF
    return c - (1 << 64) if c >= (1 << 63) else c


def _mix63(t, x):
    """One round of the 63-bit multiply-xorshift mixer used for every counter-based draw (same arithmetic as
    `rbc_checkpoint_draw` in the library)."""
    x = ((x ^ (x >> 31)) * _as_i64(0x7FB5D329728EA185)) & _M63
    return x ^ (x >> 27)


def counter_normal(torch, keys, count: int, stream: int):
    """Standard normal draws `[len(keys), count]` (float64) from a counter hash of (key, element index, stream): Box-Muller on
    two 53-bit uniforms.  Stateless, so a reset of any subset of environments draws the same numbers wherever it runs."""
    idx = torch.arange(count, device=keys.device, dtype=torch.int64)[None, :]
    base = (keys[:, None] * _as_i64(0x9E3779B97F4A7C15) + idx * _as_i64(0xC2B2AE3D27D4EB4F) + _as_i64(stream * 0x165667B19E3779F9)) & _M63
    a = _mix63(torch, _mix63(torch, base))
    b = _mix63(torch, _mix63(torch, (base ^ _as_i64(0x2545F4914F6CDD1D)) & _M63))
    u1 = ((a >> 10).to(torch.float64) + 1.0) * (1.0 / (1 << 53))          # (0, 1]
    u2 = (b >> 10).to(torch.float64) * (1.0 / (1 << 53))                  # [0, 1)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


class _VectorBase:
    """Shared machinery of the 2D and 3D vector environments."""

    def _init_common(self, torch, num_envs, autoreset_mode, nan_policy, seed, env_id_offset, copy, fused):
        if autoreset_mode not in ("next_step", "same_step", "disabled"):
            raise ValueError("autoreset_mode must be next_step, same_step or disabled")
        if nan_policy not in ("raise", "raise_deferred", "reset"):
            raise ValueError("nan_policy must be raise, raise_deferred or reset")
        self.torch = torch
        self.num_envs = int(num_envs)
        self.autoreset_mode = autoreset_mode
        # the attributes gymnasium.vector.VectorEnv (1.x) carries besides the spaces: code written against the reference's
        # `gym.make_vec(...)` result (`example/run_vectorized.py:11-31`) reads these
        self.metadata = {"render_modes": ["rgb_array"], "render_fps": 10, "autoreset_mode": autoreset_mode}
        self.render_mode = "rgb_array"
        self.spec = None
        self.closed = False
        # "raise": the reference's RuntimeError on the failing step (one 4-byte host read per step);
        # "raise_deferred": the same error without any synchronisation, raised by the NEXT call; "reset": re-initialise only the
        # failed environments inside the step (reported truncated, reward 0, info["nan"])
        self.nan_policy = nan_policy
        self.seed = int(seed)
        self.env_id_offset = int(env_id_offset)
        self.copy = bool(copy)
        self.fused = bool(fused)
        self.device = self.sim.device
        self.env_ids = torch.arange(env_id_offset, env_id_offset + self.num_envs, device=self.device, dtype=torch.int64)
        self._episode = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        self._pending = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.episode_return = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        self._terminated = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self._nan_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._nan_events = collections.deque()
        self._nan_seen = 0

    # ------------------------------------------------------------------ draws
    def _draw_checkpoints(self, ids):
        """Episode index for env `g`, episode counter `e`: a counter-based hash of (seed, g, e) mod n_episodes
        (`rbc_checkpoint_draw` computes the same number inside the step kernel)."""
        t = self.torch
        g, e = self.env_ids[ids], self._episode[ids]
        x = (g * _as_i64(0x9E3779B97F4A7C15) + e * _as_i64(0xC2B2AE3D27D4EB4F) + _as_i64(self.seed * 0x165667B19E3779F9)) & _M63
        return (_mix63(t, x) % self.sim.n_episodes).to(t.int32)

    def _noise_keys(self, ids):
        """One 63-bit key per (seed, global env id, episode): partial resets at the same episode count never share noise."""
        g, e = self.env_ids[ids], self._episode[ids]
        x = (g * _as_i64(0xD6E8FEB86659FD93) + e * _as_i64(0xA0761D6478BD642F) + _as_i64(self.seed * 0xE7037ED1A0B428DB)) & _M63
        return _mix63(self.torch, x)

    def _use_fused(self) -> bool:
        return self.fused and self.sim.n_episodes > 0 and getattr(self.sim, "fused_autoreset", True)

    def _sync_fused_config(self):
        self.sim.set_autoreset(self.autoreset_mode, nan_reset=self.nan_policy == "reset", seed=self.seed, env_id_offset=self.env_id_offset)

    def _out(self, x):
        return x.clone() if self.copy else x

    # ------------------------------------------------------------------ NaN policy
    def _check_nan_deferred(self, block: bool = False):
        """Read the device-side NaN counter through the copies that have COMPLETED (never waits for the device unless `block`):
        in a loop that runs ahead of the GPU the copy of the previous step is still in flight, and waiting for it would
        serialise host and device — the launch latency of every step would show up as idle GPU time (measured: up to 10 % of the
        step on a slow host).  The error is then raised one or two calls late instead of exactly one."""
        q = self._nan_events
        done = False
        while q and (block or len(q) > 4 or q[0].query()):
            q[0].synchronize()
            q.popleft()
            done = True
        if done:
            n = int(self._nan_host[0])
            if n > self._nan_seen:
                self._nan_seen = n
                raise RuntimeError("Error in simulation step, probably NaN values")      # rbc2D.py:170-171

    def _after_fused_step(self):
        t = self.torch
        if self.nan_policy == "reset":
            return
        self.sim.vec_nan_count_async(self._nan_host)
        ev = t.cuda.Event()
        ev.record(t.cuda.current_stream(self.device))
        self._nan_events.append(ev)
        if self.nan_policy == "raise":
            self._check_nan_deferred(block=True)

    def close(self):
        self.sim.close()
        self.closed = True

    @property
    def unwrapped(self):
        return self


class RBCVectorEnv2D(_VectorBase):
    def __init__(self, num_envs: int, rayleigh_number: float = 10_000, episode_length: float = 300, observation_shape=(8, 48),
                 state_shape=(64, 96), heater_segments: int = 12, heater_limit: float = 0.75, heater_duration: float = 1.5,
                 pressure: bool = False, checkpoint: Optional[str] = None, dt_solver: float = 0.03, precision: int = 32,
                 device: int = 0, autoreset_mode: str = "next_step", seed: int = 0, env_id_offset: int = 0,
                 nan_policy: str = "raise_deferred", info_state: bool = False, copy: bool = True, fused: bool = True):
        import torch

        self.ra = rayleigh_number
        self.episode_length = episode_length
        self.episode_steps = int(episode_length / heater_duration)
        self.observation_shape, self.state_shape = list(observation_shape), list(state_shape)
        self.temperature_difference = [1, 2]
        self.heater_segments, self.heater_limit, self.heater_duration = heater_segments, heater_limit, heater_duration
        self.include_pressure = pressure
        self.checkpoint = checkpoint
        self.info_state = bool(info_state)
        self.sim = backend.Sim2D(int(num_envs), ra=float(rayleigh_number), dt_action=float(heater_duration),
                                 obs_shape=tuple(observation_shape), state_shape=tuple(state_shape), heaters=heater_segments,
                                 heater_limit=heater_limit, dt_solver=dt_solver, episode_length=float(episode_length),
                                 precision=precision, pressure=pressure, device=device)
        self._init_common(torch, num_envs, autoreset_mode, nan_policy, seed, env_id_offset, copy, fused)
        ch = self.sim.channels
        self.single_action_space = spaces.Box(-1, 1, shape=(heater_segments,), dtype=np.float32)
        self.single_observation_space = spaces.Box(-np.inf, np.inf, shape=(ch, *observation_shape), dtype=np.float32)
        self.action_space = spaces.Box(-1, 1, shape=(self.num_envs, heater_segments), dtype=np.float32)
        self.observation_space = spaces.Box(-np.inf, np.inf, shape=(self.num_envs, ch, *observation_shape), dtype=np.float32)
        if checkpoint:
            self.sim.load_checkpoints(checkpoint)

    # ------------------------------------------------------------------ helpers
    def _noise_fields(self, ids):
        """`initialize_model` (`rbc_sim2D.jl:163-171`): u, w = kick*randn, b = clamp(min_b + (Lz - z) db/2 + kick*randn, ...)."""
        t = self.torch
        nz, nx = self.sim.nz, self.sim.nx
        n, kick = int(ids.numel()), 0.01
        keys = self._noise_keys(ids)
        z = (t.arange(nz, device=self.device, dtype=t.float64) + 0.5) * (2.0 / nz)
        b = t.clamp(1.0 + (2.0 - z)[None, :, None] * 0.5 + kick * counter_normal(t, keys, nz * nx, 0).view(n, nz, nx), 1.0, 2.0)
        u = kick * counter_normal(t, keys, nz * nx, 1).view(n, nz, nx)
        w = kick * counter_normal(t, keys, (nz + 1) * nx, 2).view(n, nz + 1, nx)
        w[:, 0] = 0.0
        w[:, -1] = 0.0
        return t.cat([b.reshape(n, -1), u.reshape(n, -1), w.reshape(n, -1)], dim=1)

    def _reset_envs(self, ids, ckpt_idx=None):
        """Python-side reset of the listed environments (the noise path, and the un-fused reference path of the tests)."""
        t = self.torch
        if ids.numel() == 0:
            return
        if self.sim.n_episodes == 0:
            self.sim.reset_from_fields_dev(self._noise_fields(ids), ids.to(t.int32), project=True)
        else:
            idx = self._draw_checkpoints(ids) if ckpt_idx is None else t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device)
            self.sim.reset_from_checkpoints(idx, env_ids=ids.to(t.int32))
        self._episode[ids] += 1
        self.episode_return[ids] = 0

    def _info(self, nus, nuo, extras=None):
        t = self.torch
        if extras is None:
            th, sh = self.sim.info()
            info = {"nusselt_state": self._out(nus), "nusselt_obs": self._out(nuo),
                    "t": t.as_tensor(th, device=self.device), "step": t.as_tensor(sh, device=self.device),
                    "episode_return": self.episode_return.clone()}
        else:
            info = {"nusselt_state": self._out(nus), "nusselt_obs": self._out(nuo), "t": self._out(extras["t"]),
                    "step": self._out(extras["step"]), "episode_return": self._out(extras["episode_return"])}
        if self.info_state:
            info["state"] = self.sim.get_state()                       # rbc2D.py:211 (a fresh tensor per call)
        return info

    # ------------------------------------------------------------------ API
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        t = self.torch
        if seed is not None:
            self.seed = int(seed)
        ckpt_idx = None if not options else options.get("checkpoint_idx")
        self._episode.zero_()
        self._pending.zero_()
        self.episode_return.zero_()
        self._nan_events, self._nan_seen = collections.deque(), 0
        if self._use_fused():
            self._sync_fused_config()
            self.sim.vec_reset(ckpt_idx)
            self._episode += 1
        else:
            self._reset_envs(t.arange(self.num_envs, device=self.device), ckpt_idx)
        obs, nus, nuo = self.sim.observe()
        return self._out(obs), self._info(nus, nuo)

    def step(self, actions):
        """actions `[num_envs, heater_segments]` (CUDA tensor) -> (obs, reward, terminated, truncated, info), all CUDA tensors."""
        if not self._use_fused():
            return self._step_python(actions)
        t = self.torch
        if self.nan_policy == "raise_deferred":
            self._check_nan_deferred()
        obs, rew, nus, nuo, trunc, nan, ex = self.sim.vec_step(actions)
        info = self._info(nus, nuo, ex)
        info["nan"] = nan.to(t.bool)
        if self.nan_policy == "reset":
            info["nan_reset"] = info["nan"]
        if self.autoreset_mode == "same_step" or self.nan_policy == "reset":
            # rows of environments that were re-initialised inside this step (truncated, or NaN under nan_policy="reset")
            info["final_obs"] = self._out(ex["final_obs"])
            info["final_info"] = {"nusselt_state": self._out(ex["final_nu_state"]), "nusselt_obs": self._out(ex["final_nu_obs"]),
                                  "episode_return": self._out(ex["final_return"])}
        self._after_fused_step()
        return self._out(obs), self._out(rew), self._terminated, trunc.to(t.bool), info

    def alloc_host_outputs(self, pinned: bool = True):
        """Page-locked numpy buffers for `step_host` (final_* included in same_step mode / under nan_policy="reset")."""
        return self.sim.alloc_vec_host_outputs(pinned, final=self.autoreset_mode == "same_step" or self.nan_policy == "reset")

    def step_host(self, actions: np.ndarray, out: Optional[dict] = None):
        """`step` for callers that live on the host (numpy actions in, numpy results out — what a policy running elsewhere, or
        the reference's own pipe-based rollout loop, exchanges per step): one call into `rbc2d_vec_step_host`, which overlaps
        the device-to-host copies of a chunk of environments with the march of the next chunk.  Returns
        (obs, reward, terminated, truncated, info) as views of `out` (allocate it once with `alloc_host_outputs`)."""
        if not self._use_fused():
            raise RuntimeError("step_host needs a checkpoint bank (the fused vector step); call step() with device tensors instead")
        if out is None:
            out = self.alloc_host_outputs()
        self.sim.vec_step_host(actions, out)
        if self.nan_policy != "reset" and out["nan"].any():
            raise RuntimeError("Error in simulation step, probably NaN values")      # rbc2D.py:170-171
        info = {"nusselt_state": out["nu_state"], "nusselt_obs": out["nu_obs"], "t": out["t"], "step": out["step"],
                "episode_return": out["episode_return"], "nan": out["nan"].astype(bool)}
        if "final_obs" in out:
            info["final_obs"] = out["final_obs"]
            info["final_info"] = {"nusselt_state": out["final_nu_state"], "nusselt_obs": out["final_nu_obs"],
                                  "episode_return": out["final_return"]}
        return out["obs"], out["reward"], np.zeros(self.num_envs, bool), out["truncated"].astype(bool), info

    def _step_python(self, actions):
        """The same semantics with the resets driven from Python (separate reset + observe launches, host reads of the flags):
        the noise-initialisation path, and the reference the fused kernel is tested against (`fused=False`)."""
        t = self.torch
        pend = self._pending
        have_pending = self.autoreset_mode == "next_step" and bool(pend.any())
        obs, rew, nus, nuo, trunc, nan = self.sim.step(actions)
        bad = nan.to(t.bool) & ~pend if have_pending else nan.to(t.bool)
        truncated = trunc.to(t.bool)
        reward = rew.clone()
        final = None
        already_reset = None
        if bool(bad.any()):
            if self.nan_policy != "reset":
                raise RuntimeError("Error in simulation step, probably NaN values")   # rbc2D.py:170-171
            # a batch of thousands should not die with one environment: re-initialise the failed ones, zero their reward,
            # report them as truncated and in info["nan_reset"]; their terminal outputs go to final_* like a same_step reset
            reward[bad] = 0
            final = {"mask": bad.clone(), "obs": obs.clone(), "nus": nus.clone(), "nuo": nuo.clone()}
            truncated = truncated | bad
            already_reset = bad
        if have_pending:
            # gymnasium NEXT_STEP semantics: for an env that truncated on the previous call this call only resets
            # it — the action is ignored, reward 0, and the returned observation is the reset observation.
            reward[pend] = 0
            truncated = truncated & ~pend
        self.episode_return += reward.to(t.float64)
        to_reset = truncated if already_reset is None else truncated & ~already_reset
        if self.autoreset_mode == "same_step" and bool(to_reset.any()):
            # SB3 semantics: reset inside the truncating step; terminal observation kept in info["final_obs"]
            if final is None:
                final = {"mask": to_reset.clone(), "obs": obs.clone(), "nus": nus.clone(), "nuo": nuo.clone()}
            else:
                final["mask"] |= to_reset
        else:
            to_reset = t.zeros_like(truncated)
        final_return = self.episode_return.clone()
        reset_mask = to_reset.clone()
        if already_reset is not None:
            reset_mask |= already_reset
        if have_pending:
            reset_mask |= pend
        if bool(reset_mask.any()):
            self._reset_envs(reset_mask.nonzero().flatten())
            obs, nus, nuo = self.sim.observe()
        info = self._info(nus, nuo)
        info["nan"] = bad
        if already_reset is not None:
            info["nan_reset"] = bad
        if final is not None:
            info["final_obs"] = final["obs"]
            info["final_info"] = {"nusselt_state": final["nus"], "nusselt_obs": final["nuo"], "episode_return": final_return}
        self._pending = (truncated & ~reset_mask) if self.autoreset_mode == "next_step" else t.zeros_like(truncated)
        return self._out(obs), reward, self._terminated, truncated, info

    def get_state(self):
        return self.sim.get_state()

    def render(self):
        """`rgb_array` frames of all environments as one uint8 CUDA tensor `[num_envs, Nz, Nx, 3]` (e.g. for logging videos
        like `example/run_wandb.py:25-59` without a host round trip per frame)."""
        return self.sim.render_rgb()


class RBCVectorEnv3D(_VectorBase):
    """The 3D twin: B independent `RayleighBenardConvection3D-v0` environments as one on-device batch — what
    `SubprocVecEnv([make_env(i, ...) for i in range(n_envs)])` of `experiments/run_sarl.py:130-153` becomes.

    Observations are the full state `(num_envs, 4, Nz, Ny, Nx)` float32 CUDA tensors (`rbc3D.py:229-232`), actions
    `(num_envs, 8, 8)` in [-1, 1] with `action[:, i, j]` <-> patch i along x, j along y, reward = -Nu, `info` keys `t`, `step`,
    `nusselt` (`rbc3D.py:241-245`).  Resets come from a `3D_ckpt_ra*.h5` bank (gathered inside the step kernel) or from the
    noise initialisation; autoreset as in the 2D class.  `copy` defaults to False here: an observation batch is 262 KB per env."""

    def __init__(self, num_envs: int, rayleigh_number: float = 2500, prandtl_number: float = 0.7, domain=(2, 4 * np.pi, 4 * np.pi),
                 state_shape=(16, 32, 32), temperature_difference=(1, 2), heater_segments: int = 8, heater_limit: float = 0.9,
                 heater_duration: float = 0.125, episode_length: float = 300, dt_solver: float = 0.01, checkpoint: Optional[str] = None,
                 precision: int = 32, device: int = 0, autoreset_mode: str = "next_step", seed: int = 0, env_id_offset: int = 0,
                 nan_policy: str = "raise_deferred", copy: bool = False, fused: bool = True):
        import torch

        self.is_3d = True
        self.ra, self.episode_length = rayleigh_number, episode_length
        self.state_shape, self.domain = tuple(state_shape), tuple(domain)
        self.temperature_difference = list(temperature_difference)
        self.heater_segments, self.heater_limit, self.heater_duration = heater_segments, heater_limit, heater_duration
        self.checkpoint = checkpoint
        self.sim = backend.Sim3D(int(num_envs), ra=float(rayleigh_number), pr=float(prandtl_number), domain=self.domain,
                                 state_shape=self.state_shape, temperature_difference=tuple(temperature_difference),
                                 heaters=heater_segments, heater_limit=heater_limit, heater_duration=heater_duration, dt_solver=dt_solver,
                                 episode_length=float(episode_length), precision=precision, device=device)
        self._init_common(torch, num_envs, autoreset_mode, nan_policy, seed, env_id_offset, copy, fused)
        self.single_action_space = spaces.Box(-1, 1, shape=(heater_segments, heater_segments), dtype=np.float32)
        self.single_observation_space = spaces.Box(-np.inf, np.inf, shape=(4, *self.state_shape), dtype=np.float32)
        self.action_space = spaces.Box(-1, 1, shape=(self.num_envs, heater_segments, heater_segments), dtype=np.float32)
        self.observation_space = spaces.Box(-np.inf, np.inf, shape=(self.num_envs, 4, *self.state_shape), dtype=np.float32)
        if checkpoint:
            from ..h5lite import load_checkpoint_3d
            self.sim.load_checkpoints(load_checkpoint_3d(checkpoint))

    def _noise_fields(self, ids):
        """`initialize_model` (`rbc_sim3D.jl:169-178`) from the counter hash."""
        t = self.torch
        nz, ny, nx = self.state_shape
        n, kick = int(ids.numel()), 0.01
        lz, b0 = float(self.domain[0]), float(self.temperature_difference[0])
        db = float(self.temperature_difference[1]) - b0
        keys = self._noise_keys(ids)
        nc = nz * ny * nx
        z = (t.arange(nz, device=self.device, dtype=t.float64) + 0.5) * (lz / nz)
        b = t.clamp(b0 + (lz - z)[None, :, None, None] * db / 2 + kick * counter_normal(t, keys, nc, 0).view(n, nz, ny, nx), b0, b0 + db)
        u = kick * counter_normal(t, keys, nc, 1)
        v = kick * counter_normal(t, keys, nc, 2)
        w = kick * counter_normal(t, keys, (nz + 1) * ny * nx, 3).view(n, nz + 1, ny, nx)
        w[:, 0] = 0.0
        w[:, -1] = 0.0
        return t.cat([b.reshape(n, -1), u, v, w.reshape(n, -1)], dim=1).contiguous()

    def _reset_envs(self, ids, ckpt_idx=None):
        t = self.torch
        if ids.numel() == 0:
            return
        if self.sim.n_episodes == 0:
            self.sim.reset_from_fields_dev(self._noise_fields(ids), ids.to(t.int32), project=True)
        else:
            idx = self._draw_checkpoints(ids) if ckpt_idx is None else t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device)
            self.sim.reset_from_checkpoints(idx, env_ids=ids.to(t.int32))
        self._episode[ids] += 1
        self.episode_return[ids] = 0

    def render(self, height: int = 608, width: int = 800):
        """`rgb_array` frames of all environments as one uint8 CUDA tensor `[num_envs, height, width, 3]` (volume rendering of the
        temperature, `rbc3D.py:247-318`), e.g. for the videos of `example/run_wandb.py:25-59` without a host round trip per frame."""
        return self.sim.render_rgb(height, width)

    def _info(self, nu, extras=None):
        t = self.torch
        if extras is None:
            th, sh = self.sim.info()
            return {"nusselt": self._out(nu), "t": t.as_tensor(th, device=self.device), "step": t.as_tensor(sh, device=self.device),
                    "episode_return": self.episode_return.clone()}
        return {"nusselt": self._out(nu), "t": self._out(extras["t"]), "step": self._out(extras["step"]),
                "episode_return": self._out(extras["episode_return"])}

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        t = self.torch
        if seed is not None:
            self.seed = int(seed)
        ckpt_idx = None if not options else options.get("checkpoint_idx")
        self._episode.zero_()
        self._pending.zero_()
        self.episode_return.zero_()
        self._nan_events, self._nan_seen = collections.deque(), 0
        if self._use_fused():
            self._sync_fused_config()
            self.sim.vec_reset(ckpt_idx)
            self._episode += 1
        else:
            self._reset_envs(t.arange(self.num_envs, device=self.device), ckpt_idx)
        obs, nu = self.sim.observe()
        return self._out(obs), self._info(nu)

    def step(self, actions):
        if not self._use_fused():
            return self._step_python(actions)
        t = self.torch
        if self.nan_policy == "raise_deferred":
            self._check_nan_deferred()
        same = self.autoreset_mode == "same_step" or self.nan_policy == "reset"
        obs, rew, nu, trunc, nan, ex = self.sim.vec_step(actions, want_final_obs=same)
        info = self._info(nu, ex)
        info["nan"] = nan.to(t.bool)
        if self.nan_policy == "reset":
            info["nan_reset"] = info["nan"]
        if same:
            info["final_obs"] = self._out(ex["final_obs"])
            info["final_info"] = {"nusselt": self._out(ex["final_nusselt"]), "episode_return": self._out(ex["final_return"])}
        self._after_fused_step()
        return self._out(obs), self._out(rew), self._terminated, trunc.to(t.bool), info

    def _step_python(self, actions):
        t = self.torch
        pend = self._pending
        have_pending = self.autoreset_mode == "next_step" and bool(pend.any())
        obs, rew, nu, trunc, nan = self.sim.step(actions)
        bad = nan.to(t.bool) & ~pend if have_pending else nan.to(t.bool)
        if bool(bad.any()):
            raise RuntimeError("Error in simulation step, probably NaN values")    # rbc3D.py:207-212
        truncated, reward = trunc.to(t.bool), rew.clone()
        if have_pending:
            reward[pend] = 0
            truncated = truncated & ~pend
        self.episode_return += reward.to(t.float64)
        final, final_return = None, self.episode_return.clone()
        reset_mask = pend.clone() if have_pending else t.zeros_like(truncated)
        if self.autoreset_mode == "same_step" and bool(truncated.any()):
            final = {"obs": obs.clone(), "nu": nu.clone()}
            reset_mask |= truncated
        if bool(reset_mask.any()):
            self._reset_envs(reset_mask.nonzero().flatten())
            obs, nu = self.sim.observe()
        info = self._info(nu)
        info["nan"] = bad
        if final is not None:
            info["final_obs"] = final["obs"]
            info["final_info"] = {"nusselt": final["nu"], "episode_return": final_return}
        self._pending = (truncated & ~reset_mask) if self.autoreset_mode == "next_step" else t.zeros_like(truncated)
        return self._out(obs), reward, self._terminated, truncated, info
