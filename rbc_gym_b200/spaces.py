"""gymnasium when it is installed, otherwise a minimal stand-in with the same surface.

The reference environments are `gymnasium.Env` subclasses (`src/rbc_gym/envs/rbc2D.py:29`).  gymnasium is
not importable on every box this backend runs on, so the environment classes bind to whichever is
available; the stand-in implements only what the reference uses: `spaces.Box` (shape/dtype/low/high/
sample/contains), `Env` (np_random seeding in `reset`), `Wrapper`/`ObservationWrapper`/`RewardWrapper`.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium exists
    import gymnasium as gym
    from gymnasium import spaces
    HAVE_GYMNASIUM = True
    Env, Wrapper, ObservationWrapper, RewardWrapper = gym.Env, gym.Wrapper, gym.ObservationWrapper, gym.RewardWrapper
    Box = spaces.Box
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Env:
        metadata: dict = {}
        render_mode = None
        _np_random = None
        _np_random_seed = None

        def reset(self, *, seed=None, options=None):
            if seed is not None or self._np_random is None:
                if seed is None:
                    seed = int(np.random.SeedSequence().entropy % (2**31))
                self._np_random = np.random.default_rng(seed)
                self._np_random_seed = seed

        @property
        def np_random(self):
            if self._np_random is None:
                self.reset()
            return self._np_random

        @property
        def np_random_seed(self):
            if self._np_random_seed is None:
                self.reset()
            return self._np_random_seed

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, action):
            return self.env.step(action)

        def render(self):
            return self.env.render()

        def close(self):
            return self.env.close()

    class ObservationWrapper(Wrapper):
        def reset(self, *, seed=None, options=None):
            obs, info = self.env.reset(seed=seed, options=options)
            return self.observation(obs), info

        def step(self, action):
            obs, r, term, trunc, info = self.env.step(action)
            return self.observation(obs), r, term, trunc, info

    class RewardWrapper(Wrapper):
        def step(self, action):
            obs, r, term, trunc, info = self.env.step(action)
            return obs, self.reward(r), term, trunc, info
