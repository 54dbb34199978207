"""`RayleighBenardConvection3DEnv` — the reference's 3D single-environment API (`src/rbc_gym/envs/rbc3D.py:43-339`)
on the CUDA backend: same kwargs, spaces, `info` keys (`t`, `step`, `nusselt`) and exceptions; the Julia module is
replaced by a batch-of-one `Sim3D` handle.  PyVista rendering (`rbc3D.py:247-318`) is not reproduced."""
from __future__ import annotations

import logging
import warnings
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import backend, spaces


def noise_initial_fields_3d(rng: np.random.Generator, shape=(16, 32, 32), kick: float = 0.01, min_b: float = 1.0, delta_b: float = 1.0,
                            lz: float = 2.0) -> np.ndarray:
    """`initialize_model` (`rbc_sim3D.jl:169-178`); `[1, 66560]` float64, projected on the device afterwards."""
    nz, ny, nx = shape
    z = (np.arange(nz) + 0.5) * (lz / nz)
    b = np.clip(min_b + (lz - z)[:, None, None] * delta_b / 2 + kick * rng.standard_normal((nz, ny, nx)), min_b, min_b + delta_b)
    u = kick * rng.standard_normal((nz, ny, nx))
    v = kick * rng.standard_normal((nz, ny, nx))
    w = kick * rng.standard_normal((nz + 1, ny, nx))
    w[0] = 0.0
    w[-1] = 0.0
    return backend.pack_fields3(b[None], u[None], v[None], w[None])


class RayleighBenardConvection3DEnv(spaces.Env):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 10}
    is_3d = True

    def __init__(
        self,
        rayleigh_number: Optional[int] = 2500,
        prandtl_number: Optional[float] = 0.7,
        domain: Optional[list] = [2, 4 * np.pi, 4 * np.pi],
        state_shape: Optional[list] = (16, 32, 32),
        temperature_difference: Optional[list] = [1, 2],
        heater_segments: Optional[int] = 8,
        heater_limit: Optional[float] = 0.9,
        heater_duration: Optional[float] = 0.125,
        episode_length: Optional[int] = 300,
        dt_solver: Optional[float] = 0.01,
        use_gpu: Optional[bool] = True,
        checkpoint: Optional[str] = None,
        checkpoint_idx: Optional[int] = None,
        render_mode: Optional[str] = None,
        log_dir: str = None,
        env_id: int = 0,
        precision: int = 64,
        device: int = 0,
    ) -> None:
        super().__init__()
        self.closed = False
        self.use_gpu = use_gpu
        self.checkpoint, self.checkpoint_idx = checkpoint, checkpoint_idx
        self.ra, self.pr, self.domain = rayleigh_number, prandtl_number, domain
        self.episode_length, self.dt_solver = episode_length, dt_solver
        self.state_shape = state_shape
        self.temperature_difference = temperature_difference
        self.heater_segments, self.heater_limit, self.heater_duration = heater_segments, heater_limit, heater_duration
        self.logger = logging.getLogger(__name__)
        self.logger.info(f"Using Rayleigh number Ra={self.ra}")
        self.action_space = spaces.Box(-1, 1, shape=(heater_segments, heater_segments), dtype=np.float32)     # rbc3D.py:105-107
        lows = np.stack([np.full(state_shape, temperature_difference[0])] + [np.full(state_shape, -np.inf)] * 3).astype(np.float32)
        highs = np.stack([np.full(state_shape, temperature_difference[1] + heater_limit)] + [np.full(state_shape, np.inf)] * 3).astype(np.float32)
        self.observation_space = spaces.Box(lows, highs, shape=(4, *state_shape), dtype=np.float32)
        self.sim = backend.Sim3D(1, ra=float(rayleigh_number), pr=float(prandtl_number), domain=tuple(domain), state_shape=tuple(state_shape),
                                 temperature_difference=tuple(temperature_difference), heaters=heater_segments, heater_limit=heater_limit,
                                 heater_duration=heater_duration, dt_solver=dt_solver, episode_length=float(episode_length),
                                 precision=precision, device=device)
        self._initialized = False
        self.render_mode = render_mode
        self.last_obs = self.last_reward = self.last_info = self.last_action = None

    def reset(self, seed: int | None = None, options: Dict[str, Any] | None = None) -> Tuple[Any, Dict[str, Any]]:
        super().reset(seed=seed)
        if self.checkpoint:
            path = Path(self.checkpoint)
            if not path.exists():
                raise FileNotFoundError(f"Checkpoint file {path} does not exist. Please provide a valid checkpoint directory.")
            from ..h5lite import load_checkpoint_3d
            import torch
            bank = load_checkpoint_3d(path)
            n = self.sim.load_checkpoints(bank)
            if self.checkpoint_idx is not None:
                # the reference forwards checkpoint_idx unchanged to Julia, where it is 1-BASED: read(h5, "b")[idx, :, :, :]
                # (rbc_sim3D.jl:191; experiments/eval_sarl.py:45 passes 1 for the first episode); 0 or > n is a BoundsError there
                if not 1 <= int(self.checkpoint_idx) <= n:
                    raise IndexError(f"checkpoint_idx {self.checkpoint_idx} out of range 1..{n} (1-based like the reference's Julia side)")
                idx = int(self.checkpoint_idx) - 1
            else:
                idx = int(self.np_random.integers(n))
            self.sim.reset_from_checkpoints(torch.tensor([idx], dtype=torch.int32))
        else:
            lz = float(self.domain[0])
            f = noise_initial_fields_3d(self.np_random, tuple(self.state_shape), min_b=self.temperature_difference[0],
                                        delta_b=self.temperature_difference[1] - self.temperature_difference[0], lz=lz)
            self.sim.reset_from_fields(f, project=True)
        self._initialized = True
        self.last_action = self.action_space.sample() * 0
        self.sim.observe()
        return self.__get_obs(), self.__get_info()

    def step(self, action: Any = None) -> Tuple[Any, float, bool, bool, Dict[str, Any]]:
        terminated, truncated = False, False
        if action is None:
            action = np.zeros(self.action_space.shape, dtype=np.float32)
            warnings.warn("No action provided, using zero action")
        if not self._initialized:
            raise RuntimeError("Simulation not initialized. Call reset first.")
        import torch
        a = np.asarray(action, dtype=np.float32)
        *_, nan = self.sim.step(torch.from_numpy(a.reshape((1,) + a.shape)))
        if int(nan.item()):
            self.logger.error("Simulation step failed, probably NaN values in the simulation.")
            raise RuntimeError("Error in simulation step, probably NaN values")
        self.last_obs, self.last_reward, self.last_info = self.__get_obs(), self.__get_reward(), self.__get_info()
        if self.last_info["t"] >= self.episode_length:
            truncated = True
        return self.last_obs, self.last_reward, terminated, truncated, self.last_info

    def __get_obs(self) -> Any:
        return self.sim.obs[0].cpu().numpy().copy()

    def __get_reward(self) -> float:
        return -float(self.sim.nusselt[0].item())

    def __get_info(self) -> dict[str, Any]:
        t, step = self.sim.info()
        return {"t": float(t[0]), "step": int(step[0]), "nusselt": float(self.sim.nusselt[0].item())}

    def render(self):
        """`rgb_array`: an (608, 800, 3) uint8 volume rendering of the temperature like the reference's PyVista picture
        (`rbc3D.py:247-318`: turbo colormap, clim = temperature_difference, sigmoid opacity, isometric camera), ray-marched on
        the device.  `human` mode (an interactive PyVista window) is not provided."""
        if self.render_mode != "rgb_array":
            return None
        if not self._initialized:
            raise RuntimeError("Simulation not initialized. Call reset first.")
        return self.sim.render_rgb()[0].cpu().numpy()

    def close(self):
        if getattr(self, "sim", None) is not None:
            self.sim.close()            # shutdown_simulation (rbc_sim3D_api.jl:164-171)
        self.closed = True
