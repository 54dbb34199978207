"""Stable-Baselines3 `VecEnv` facade over the on-device batches — the drop-in for
`SubprocVecEnv([make_env(i, ...) for i in range(n_envs)])` in `experiments/run_sarl.py:130-153`, including what that
script wraps around every environment: `Monitor` (episode return / length in `infos[i]["episode"]`) and
`RBCNormalizeObservation(maxval=1)` (`run_sarl.py:113-118`).

SB3 semantics: numpy in / numpy out, `step_async` + `step_wait`, reset inside the truncating step with the last
observation in `infos[i]["terminal_observation"]` and `infos[i]["TimeLimit.truncated"] = True`.  When stable_baselines3
is installed the class derives from its `VecEnv`; without it the same methods exist on a plain object (nothing in this
image needs SB3).  The simulation itself never leaves the GPU; only SB3's numpy contract forces the copies here —
policies that live on the device should use `RBCVectorEnv2D/3D` directly.
"""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence

import numpy as np

from .. import spaces

try:  # pragma: no cover - stable_baselines3 is not part of this image
    from stable_baselines3.common.vec_env import VecEnv as _Base
    HAVE_SB3 = True
except Exception:  # noqa: BLE001
    _Base = object
    HAVE_SB3 = False


class RBCSB3VecEnv(_Base):
    def __init__(self, venv, normalize_observation: bool = True, maxval: float = 1.0, u_limit: Optional[float] = None,
                 monitor: bool = True):
        if getattr(venv, "autoreset_mode", None) != "same_step":
            raise ValueError("SB3 resets inside the terminating step: build the vector env with autoreset_mode='same_step'")
        self.venv = venv
        self.torch = venv.torch
        self.is_3d = bool(getattr(venv, "is_3d", False))
        self._monitor = monitor
        obs_space, act_space = venv.single_observation_space, venv.single_action_space
        self._norm = None
        if normalize_observation:
            T = venv.temperature_difference
            if u_limit is None:
                from ..wrappers import RBCNormalizeObservation        # late: wrappers imports the env modules
                u_limit = RBCNormalizeObservation._get_u_limit_3d(venv.ra) if self.is_3d else 1.3
            lo = self.torch.tensor([T[0], -u_limit, -u_limit, -u_limit], device=venv.device, dtype=self.torch.float32)
            hi = self.torch.tensor([T[1] + venv.heater_limit, u_limit, u_limit, u_limit], device=venv.device, dtype=self.torch.float32)
            c = obs_space.shape[0]
            shape = (1, c) + (1,) * (len(obs_space.shape) - 1)
            self._norm = (lo[:c].reshape(shape), hi[:c].reshape(shape), float(maxval))
            limit = maxval * 1.3                                     # rbc_normalize_observation.py:52-61 (eps = 0.3)
            obs_space = spaces.Box(low=-limit, high=limit, shape=obs_space.shape, dtype=np.float32)
        if HAVE_SB3:
            super().__init__(venv.num_envs, obs_space, act_space)
        else:
            self.num_envs, self.observation_space, self.action_space = venv.num_envs, obs_space, act_space
        self._actions = None
        self._ep_len = np.zeros(self.num_envs, np.int64)
        self._ep_ret = np.zeros(self.num_envs, np.float64)
        self._t0 = time.time()

    # ------------------------------------------------------------------ helpers
    def _obs(self, obs):
        if self._norm is not None:
            lo, hi, maxval = self._norm
            obs = maxval * (2 * (obs - lo) / (hi - lo) - 1)
        return obs.cpu().numpy()

    # ------------------------------------------------------------------ VecEnv API
    def reset(self):
        obs, _ = self.venv.reset()
        self._ep_len[:] = 0
        self._ep_ret[:] = 0
        return self._obs(obs)

    def step_async(self, actions) -> None:
        self._actions = np.asarray(actions, dtype=np.float32)

    def step_wait(self):
        t = self.torch
        a = t.from_numpy(self._actions).to(self.venv.device)
        obs, reward, terminated, truncated, info = self.venv.step(a)
        rew = reward.cpu().numpy().astype(np.float32)
        dones = (terminated | truncated).cpu().numpy()
        self._ep_len += 1
        self._ep_ret += rew
        sim_t, sim_step = self.venv.sim.info()
        nu_keys = [k for k in ("nusselt", "nusselt_obs", "nusselt_state") if k in info]
        nus = {k: info[k].cpu().numpy() for k in nu_keys}
        infos: List[dict] = [{**{k: float(v[i]) for k, v in nus.items()}, "t": float(sim_t[i]), "step": int(sim_step[i])}
                             for i in range(self.num_envs)]
        if dones.any():
            final_obs = self._obs(info["final_obs"])
            fin = info["final_info"]
            for i in np.nonzero(dones)[0]:
                infos[i]["terminal_observation"] = final_obs[i]
                infos[i]["TimeLimit.truncated"] = True               # episodes only ever end by the time limit (rbc2D.py:161)
                for k in nu_keys:
                    infos[i][k] = float(fin[k][i].item())
                if self._monitor:                                    # stable_baselines3.common.monitor.Monitor
                    infos[i]["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i]), "t": round(time.time() - self._t0, 6)}
                self._ep_len[i] = 0
                self._ep_ret[i] = 0.0
        return self._obs(obs), rew, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        self.venv.close()

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        if seed is not None:
            self.venv.seed = int(seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.venv, attr_name)] * n

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self.venv, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self.venv, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def get_images(self):
        return list(self.venv.render().cpu().numpy()) if hasattr(self.venv, "render") else []
