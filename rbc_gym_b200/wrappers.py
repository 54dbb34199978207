"""The reference's wrappers (`src/rbc_gym/wrappers/`) for the single-environment API, plus the pure functions
they are made of.  The batched/vector path applies the same arithmetic fused into the CUDA step epilogue
(`Sim2D.set_wrappers`); these classes are the drop-in for code that wraps one `gym.Env`."""
from __future__ import annotations

import logging
from typing import Any, Dict, Tuple

import warnings

import numpy as np

from . import spaces
from .envs.rbc2d import RBCField


# --------------------------------------------------------------------------------------- pure functions
def find_peaks_height(x: np.ndarray, height: float) -> np.ndarray:
    """`scipy.signal.find_peaks(x, height=height)[0]` without scipy: strict local maxima, plateaus give their
    midpoint, the first and last samples are never peaks."""
    x = np.asarray(x)
    peaks, i, imax = [], 1, len(x) - 1
    while i < imax:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < imax and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                mid = (i + ahead - 1) // 2
                if x[mid] >= height:
                    peaks.append(mid)
                i = ahead
        i += 1
    return np.asarray(peaks, dtype=np.int64)


def cell_distance(state: np.ndarray, size_state=(64, 96), use_avg: bool = False) -> float:
    """Distance between the Benard cells of a state; same result as `RBCRewardShaping.compute_cell_distances`
    (`rbc_reward_shaping.py:86-140`), evaluated for all pairs of cells at once.

    Cells are the maxima (height >= 0.001) of the vertical velocity along the row just below mid-height (or its column average).
    The distance of a pair is the shorter way round the periodic domain, and 0 when the velocity stays positive all along that
    way (the two maxima then belong to one up-welling).  The result is the largest pair distance; a single cell gives 0."""
    uy = state[RBCField.UY].mean(axis=0) if use_avg else state[RBCField.UY][int(size_state[0] / 2) - 1]
    peaks = find_peaks_height(uy, 0.001)
    if len(peaks) <= 1:
        return 0
    x = np.linspace(0, 2 * np.pi, size_state[1], endpoint=False)
    lo, hi = np.triu_indices(len(peaks), k=1)                      # every pair once, lo before hi along x
    left, right = peaks[lo], peaks[hi]
    direct = np.abs(x[right] - x[left])
    around = 2 * np.pi - direct
    # down[k] = number of samples among uy[:k] that are not strictly positive (NaN counts as not positive)
    down = np.concatenate(([0], np.cumsum(~(uy > 0))))
    one_upwelling = np.where(direct < around,
                             down[right] == down[left],                              # nothing sinks on the direct way
                             (down[-1] == down[right]) & (down[left] == 0))          # nothing sinks on the way round the seam
    return float(np.where(one_upwelling, 0.0, np.minimum(direct, around)).max())


def shape_reward(reward: float, cell_distances: float, w: float) -> float:
    """`RBCRewardShaping.__apply_reward_shaping` (`rbc_reward_shaping.py:68-84`)."""
    return (1 - w) * reward + w * ((-cell_distances + np.pi) / np.pi)


def normalize_reward(reward: float, ra: float, dims: int = 2) -> float:
    """`RBCNormalizeReward.reward` (`rbc_normalize_reward.py:6-32`): Nu ~ s Ra^a, 2D s=0.1 a=0.4, 3D s=0.22 a=0.27."""
    s, a = (0.1, 0.4) if dims == 2 else (0.22, 0.27)
    scale = s * (ra ** a)
    return (reward + scale) / (scale - 1)


def _affine_to_unit(obs: np.ndarray, min_vals, max_vals, maxval) -> np.ndarray:
    """Map channel c of `obs` from [min_vals[c], max_vals[c]] to [-maxval, maxval], in place, in the array's own precision
    (the reference assigns channel by channel with Python-float limits, i.e. float32 arithmetic on float32 observations)."""
    n = obs.shape[0]
    per_channel = (n,) + (1,) * (obs.ndim - 1)
    lo = np.asarray(min_vals[:n], dtype=obs.dtype).reshape(per_channel)
    span = np.asarray([hi - lo_ for hi, lo_ in zip(max_vals[:n], min_vals[:n])], dtype=obs.dtype).reshape(per_channel)
    obs[...] = maxval * (2 * (obs - lo) / span - 1)
    return obs


def normalize_observation(obs: np.ndarray, heater_limit: float, temperature_difference=(1, 2), maxval=1, u_limit=1.3,
                          clip: bool = False) -> np.ndarray:
    """`RBCNormalizeObservation.observation` (`rbc_normalize_observation.py:64-74`); modifies and returns `obs`."""
    lows = [temperature_difference[0], -u_limit, -u_limit, -u_limit]
    highs = [temperature_difference[1] + heater_limit, u_limit, u_limit, u_limit]
    obs = _affine_to_unit(obs, lows, highs, maxval)
    return np.clip(obs, -maxval, maxval) if clip else obs


# --------------------------------------------------------------------------------------- wrapper classes
class RBCNormalizeObservation(spaces.ObservationWrapper):
    """Normalize the observation to approximately lie in range [-1, 1] (`rbc_normalize_observation.py:10-81`)."""

    def __init__(self, env, heater_limit: float, maxval: int = 1, u_limit: int | None = 1.3, eps: float = 0.3, clip: bool = False):
        spaces.ObservationWrapper.__init__(self, env)
        self.heater_limit, self.clip, self.maxval, self.excursion_eps = heater_limit, clip, maxval, eps
        T = env.unwrapped.temperature_difference
        if u_limit is None:
            if getattr(env.unwrapped, "is_3d", False):
                u_limit = self._get_u_limit_3d(env.unwrapped.ra)
            else:
                raise ValueError("u_limit must be provided for 2D RBC.")
        self.min_vals = [T[0], -u_limit, -u_limit, -u_limit]
        self.max_vals = [T[1] + heater_limit, u_limit, u_limit, u_limit]
        limit = maxval * (1 + eps)
        self.observation_space = spaces.Box(low=-limit, high=limit, shape=env.observation_space.shape, dtype=np.float32)

    def observation(self, obs) -> Any:
        obs = _affine_to_unit(obs, self.min_vals, self.max_vals, self.maxval)
        if self.clip:
            obs = np.clip(obs, -self.maxval, self.maxval)
        peak = float(np.abs(obs).max())
        if peak > (1 + self.excursion_eps) * self.maxval:
            warnings.warn(f"normalised observation reaches {peak:.4g}, outside the declared range of +-{(1 + self.excursion_eps) * self.maxval:.4g}")
        return obs

    @staticmethod
    def _get_u_limit_3d(ra):
        w_inf, Ra_c, n = 0.96549382, 654.37063331, 1.06741877      # rbc_normalize_observation.py:77-81
        return w_inf * ra ** n / (ra ** n + Ra_c ** n)


class RBCNormalizeReward(spaces.RewardWrapper):
    """Normalize the reward to ~[0, 1] (`rbc_normalize_reward.py:6-32`)."""

    def __init__(self, env):
        spaces.RewardWrapper.__init__(self, env)
        ra = env.unwrapped.ra
        s, a = (0.22, 0.27) if getattr(env.unwrapped, "is_3d", False) else (0.1, 0.4)
        self.scale = s * (ra ** a)

    def reward(self, reward):
        return (reward + self.scale) / (self.scale - 1)


class RBCRewardShaping(spaces.Wrapper):
    """Shape the reward with the distance of the Benard cells (`rbc_reward_shaping.py:10-155`; the matplotlib
    debug plotting of the reference is not reproduced)."""

    def __init__(self, env, shaping_weight: float, debug_cell_dist: bool = False):
        spaces.Wrapper.__init__(self, env)
        self.logger = logging.getLogger(__name__)
        self.shaping_weight = shaping_weight
        self.debug_cell_dist = debug_cell_dist
        self.size_state = env.unwrapped.state_shape

    def reset(self, seed: int | None = None, options: Dict[str, Any] | None = None) -> Tuple[Any, Dict[str, Any]]:
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        obs, reward, closed, truncated, info = self.env.step(action)
        cd = self.compute_cell_distances(info["state"])
        reward = shape_reward(reward, cd, self.shaping_weight)
        if np.isnan(reward):
            self.logger.error("Reward is NaN")
        info["cell_dist"] = cd
        return obs, reward, closed, truncated, info

    def compute_cell_distances(self, state, use_avg=False) -> float:
        return cell_distance(state, self.size_state, use_avg)


# --------------------------------------------------------------------------------------- observation plumbing
class FlattenObservation(spaces.ObservationWrapper):
    """`gymnasium.wrappers.FlattenObservation` as used by `example/run_wrapped.py:20`: (C, H, W) -> (C*H*W,)."""

    def __init__(self, env):
        spaces.ObservationWrapper.__init__(self, env)
        sp = env.observation_space
        self.observation_space = spaces.Box(low=np.asarray(sp.low).reshape(-1), high=np.asarray(sp.high).reshape(-1),
                                            shape=(int(np.prod(sp.shape)),), dtype=np.float32)

    def observation(self, obs):
        return np.asarray(obs).reshape(-1)


class FrameStackObservation(spaces.Wrapper):
    """`gymnasium.wrappers.FrameStackObservation(env, stack_size)` (`example/run_wrapped.py:22`): the last
    `stack_size` observations, oldest first; after `reset` the stack is padded with the reset observation
    (`padding_type="reset"`, gymnasium's default) or zeros (`"zero"`)."""

    def __init__(self, env, stack_size: int, padding_type: str = "reset"):
        spaces.Wrapper.__init__(self, env)
        if stack_size < 1:
            raise ValueError("stack_size must be >= 1")
        if padding_type not in ("reset", "zero"):
            raise ValueError("padding_type must be 'reset' or 'zero'")
        self.stack_size, self.padding_type = int(stack_size), padding_type
        sp = env.observation_space
        self.observation_space = spaces.Box(low=np.repeat(np.asarray(sp.low)[None], stack_size, axis=0),
                                            high=np.repeat(np.asarray(sp.high)[None], stack_size, axis=0),
                                            shape=(stack_size, *sp.shape), dtype=np.float32)
        self._frames = None

    def reset(self, seed: int | None = None, options: Dict[str, Any] | None = None):
        obs, info = self.env.reset(seed=seed, options=options)
        pad = obs if self.padding_type == "reset" else np.zeros_like(obs)
        self._frames = np.stack([pad] * (self.stack_size - 1) + [obs])
        return self._frames.copy(), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._frames = np.concatenate([self._frames[1:], np.asarray(obs)[None]])
        return self._frames.copy(), reward, terminated, truncated, info


class VectorFrameStack:
    """Flatten + frame-stack for the on-device vector env (`RBCVectorEnv2D`): a ring buffer `[num_envs, k, ...]` of CUDA
    tensors that never leaves the GPU.  Environments that were (auto-)reset restart their stack from the reset observation.

    Wraps any object with the vector-env surface (`reset`, `step`, `num_envs`); `flatten=True` additionally flattens
    each frame like `FlattenObservation` does in `example/run_wrapped.py:19-22`."""

    def __init__(self, venv, stack_size: int = 4, flatten: bool = True, padding_type: str = "reset"):
        if padding_type not in ("reset", "zero"):
            raise ValueError("padding_type must be 'reset' or 'zero'")
        self.venv, self.k, self.flatten, self.padding_type = venv, int(stack_size), bool(flatten), padding_type
        self.num_envs = venv.num_envs
        self._buf = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.venv, name)

    def _frame(self, obs):
        return obs.reshape(obs.shape[0], -1) if self.flatten else obs

    def _restart(self, frame, rows):
        pad = frame[rows] if self.padding_type == "reset" else frame[rows] * 0
        self._buf[rows] = pad.unsqueeze(1).expand(-1, self.k, *pad.shape[1:])
        self._buf[rows, -1] = frame[rows]

    def reset(self, seed=None, options=None):
        obs, info = self.venv.reset(seed=seed, options=options)
        f = self._frame(obs)
        self._buf = f.new_zeros((f.shape[0], self.k, *f.shape[1:]))
        self._restart(f, slice(None))
        return self._buf.clone(), info

    def step(self, actions):
        obs, reward, terminated, truncated, info = self.venv.step(actions)
        f = self._frame(obs)
        self._buf = self._buf.roll(-1, dims=1)
        self._buf[:, -1] = f
        # an environment that was (auto-)reset by this call — inside the truncating step (same_step) or as the whole of this call
        # (next_step) — reports step == 1 (`rbc_sim2D_api.jl:67-68`); its stack restarts from the reset observation.  A masked
        # select instead of an index list: no host synchronisation.
        was_reset = info["step"] == 1
        pad = f if self.padding_type == "reset" else f * 0
        fresh = pad.unsqueeze(1).expand(-1, self.k, *pad.shape[1:]).clone()
        fresh[:, -1] = f
        mask = was_reset.reshape(-1, *([1] * (self._buf.dim() - 1)))
        self._buf = self.venv.torch.where(mask, fresh, self._buf)
        return self._buf.clone(), reward, terminated, truncated, info
