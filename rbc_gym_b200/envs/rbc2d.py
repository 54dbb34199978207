"""`RayleighBenardConvection2DEnv` — the reference's single-environment API on the CUDA backend.

Same constructor keywords, spaces, attributes, `info` keys and error behaviour as
`src/rbc_gym/envs/rbc2D.py:29-266`; the Julia module behind `self.sim` (juliacall,
`rbc2D.py:111-115`) is replaced by a batch-of-one `Sim2D` handle of the C-ABI library.
"""
from __future__ import annotations

import logging
import warnings
from enum import IntEnum
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import backend, spaces
from ..colormap import turbo


class RBCField(IntEnum):           # rbc2D.py:16-20
    T = 0
    UX = 1
    UY = 2
    P = 3


def noise_initial_fields(rng: np.random.Generator, shape=(64, 96), kick: float = 0.01, min_b: float = 1.0,
                         delta_b: float = 1.0, lz: float = 2.0) -> np.ndarray:
    """`initialize_model` (`rbc_sim2D.jl:163-171`): u,w = kick*randn, b = clamp(min_b + (Lz-z)db/2 + kick*randn).

    Returns `[1, 18528]` float64 in checkpoint layout.  The projection that Oceananigans' `set!` applies is
    done on the device (`reset_from_fields(project=True)`).  Julia's RNG stream cannot be reproduced; the
    fields are drawn from numpy's generator seeded with the environment seed."""
    nz, nx = shape
    z = (np.arange(nz) + 0.5) * (lz / nz)
    b = np.clip(min_b + (lz - z)[:, None] * delta_b / 2 + kick * rng.standard_normal((nz, nx)), min_b, min_b + delta_b)
    u = kick * rng.standard_normal((nz, nx))
    w = kick * rng.standard_normal((nz + 1, nx))
    w[0] = 0.0
    w[-1] = 0.0                      # impenetrable walls
    return backend.pack_fields(b[None], u[None], w[None])


class RayleighBenardConvection2DEnv(spaces.Env):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 10}

    def __init__(
        self,
        rayleigh_number: Optional[int] = 10_000,
        episode_length: Optional[int] = 300,
        observation_shape: Optional[list] = [8, 48],
        state_shape: Optional[list] = [64, 96],
        heater_segments: Optional[int] = 12,
        heater_limit: Optional[float] = 0.75,
        heater_duration: Optional[float] = 1.5,
        pressure: Optional[bool] = False,
        use_gpu: Optional[bool] = True,
        checkpoint: Optional[str] = None,
        render_mode: Optional[str] = None,
        # extensions (SURVEY §5 "Config / flags"): solver step, explicit episode index, arithmetic, device
        dt_solver: float = 0.03,
        checkpoint_idx: Optional[int] = None,
        precision: int = 64,
        device: int = 0,
    ) -> None:
        super().__init__()
        self.closed = False
        self.use_gpu = use_gpu          # kept for API compatibility; this backend always runs on the GPU
        self.checkpoint = checkpoint
        self.checkpoint_idx = checkpoint_idx

        self.ra = rayleigh_number
        self.episode_length = episode_length
        self.observation_shape = list(observation_shape)
        self.state_shape = list(state_shape)
        self.temperature_difference = [1, 2]
        self.heater_segments = heater_segments
        self.heater_limit = heater_limit
        self.heater_duration = heater_duration
        self.include_pressure = pressure
        self.episode_steps = int(episode_length / heater_duration)

        self.logger = logging.getLogger(__name__)
        self.logger.info("2D environment: Ra=%g, episode of %g time units, %d heaters", self.ra, self.episode_length, self.heater_segments)

        # rbc2D.py:75-108
        self.action_space = spaces.Box(-1, 1, shape=(self.heater_segments,), dtype=np.float32)
        channels = 3
        shp = tuple(self.observation_shape)
        lows = [np.ones(shp) * 1, np.ones(shp) * (-np.inf), np.ones(shp) * (-np.inf)]
        highs = [np.ones(shp) * 2 + self.heater_limit, np.ones(shp) * np.inf, np.ones(shp) * np.inf]
        if self.include_pressure:
            channels += 2
            lows += [np.ones(shp) * (-np.inf)] * 2
            highs += [np.ones(shp) * np.inf] * 2
        self.observation_space = spaces.Box(np.stack(lows, axis=0).astype(np.float32), np.stack(highs, axis=0).astype(np.float32),
                                            shape=(channels, shp[0], shp[1]), dtype=np.float32)

        # the simulation handle replaces juliacall.newmodule("RBCGymAPI") + include(rbc_sim2D_api.jl)
        self.sim = backend.Sim2D(1, ra=float(self.ra), dt_action=float(heater_duration), obs_shape=shp,
                                 state_shape=tuple(self.state_shape), heaters=heater_segments, heater_limit=heater_limit,
                                 dt_solver=dt_solver, episode_length=float(episode_length), precision=precision,
                                 pressure=bool(pressure), device=device)
        self._bank_path = None
        self._initialized = False

        self.render_mode = render_mode
        self.screen_width = 768
        self.screen_height = 512
        self.screen = None
        self.clock = None
        self.last_obs = self.last_reward = self.last_info = self.last_action = None

    # ---------------------------------------------------------------- reset / step
    def reset(self, seed: int | None = None, options: Dict[str, Any] | None = None) -> Tuple[Any, Dict[str, Any]]:
        super().reset(seed=seed)
        import torch

        if self.checkpoint:
            path = Path(self.checkpoint)
            self.logger.info(f"Using checkpoint file {path.absolute()}")
            if not path.exists():
                raise FileNotFoundError(
                    f"Checkpoint file {path} does not exist. Please provide a valid checkpoint directory."
                )
            if self._bank_path != str(path.absolute()):
                self.sim.load_checkpoints(path)
                self._bank_path = str(path.absolute())
            idx = self.checkpoint_idx
            if options and "checkpoint_idx" in options:
                idx = options["checkpoint_idx"]
            if idx is None:                                   # initialize_from_checkpoint: idx = rand(1:n)
                idx = int(self.np_random.integers(self.sim.n_episodes))
            self.logger.info(f"Loading checkpoint with index: {idx} from file: {path}")
            self.sim.reset_from_checkpoints(torch.tensor([idx], dtype=torch.int32))
        else:
            self.sim.reset_from_fields(noise_initial_fields(self.np_random, tuple(self.state_shape)), project=True)
        self._initialized = True
        self.last_action = self.action_space.sample() * 0
        self.sim.observe()
        return self.__get_obs(), self.__get_info()

    def step(self, action: Any = None) -> Tuple[Any, float, bool, bool, Dict[str, Any]]:
        terminated = False       # is always false; no terminal state
        truncated = False
        if action is None:
            action = np.zeros(self.action_space.shape, dtype=np.float32)
            warnings.warn("No action provided, using zero action")
        if not self._initialized:
            raise RuntimeError("Simulation not initialized. Call reset first.")     # rbc_sim2D_api.jl:79-81
        a = np.asarray(action, dtype=np.float32).reshape(1, self.heater_segments)
        import torch

        *_, nan = self.sim.step(torch.from_numpy(a))
        if int(nan.item()):
            raise RuntimeError("Error in simulation step, probably NaN values")      # rbc2D.py:170-171
        self.last_action = a[0]
        self.last_obs = self.__get_obs()
        self.last_reward = self.__get_reward()
        self.last_info = self.__get_info()
        if self.last_info["t"] >= self.episode_length:
            truncated = True
        return self.last_obs, self.last_reward, terminated, truncated, self.last_info

    # ---------------------------------------------------------------- getters (rbc2D.py:184-212)
    def __get_state(self) -> Any:
        return self.sim.get_state(5 if self.include_pressure else 3)[0].cpu().numpy()

    def __get_obs(self) -> Any:
        return self.sim.obs[0].cpu().numpy().copy()

    def __get_reward(self) -> float:
        return -float(self.sim.nu_obs[0].item())

    def __get_info(self) -> dict[str, Any]:
        t, step = self.sim.info()
        return {
            "t": float(t[0]),
            "step": int(step[0]),
            "nusselt_state": float(self.sim.nu_state[0].item()),
            "nusselt_obs": float(self.sim.nu_obs[0].item()),
            "state": self.__get_state(),
        }

    def get_state(self):
        return self.__get_state()

    # ---------------------------------------------------------------- rendering (rbc2D.py:214-261)
    def render(self):
        if self.render_mode is None:
            warnings.warn("You are calling render method without specifying any render mode. "
                          "You can specify the render_mode at initialization, ")
            return None
        data = self.__get_state()[RBCField.T]
        data = np.transpose(data)
        data = np.flip(data, axis=1)            # origin at the top left
        data = turbo(data, vmin=1, vmax=2 + self.heater_limit)
        if self.render_mode == "rgb_array":
            return data.transpose(1, 0, 2)
        if self.render_mode == "human":
            # the window itself is outside the accelerated path: a thin optional viewer (viewer.py) shows the same RGB frame
            if self.screen is None:
                from .viewer import FrameWindow
                self.screen = FrameWindow(self.screen_width, self.screen_height, self.metadata["render_fps"])
            self.screen.show(data)
            return None
        raise ValueError(f"Unknown render mode: {self.render_mode}")

    def close(self):
        if self.screen is not None:
            self.screen.close()
            self.screen = None
        if getattr(self, "sim", None) is not None:
            self.sim.close()
        self.closed = True
