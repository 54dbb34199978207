"""ctypes binding of `librbc_b200.so` (C ABI in `include/rbc_b200.h`).

This is the replacement for the reference's juliacall bridge
(`src/rbc_gym/envs/rbc2D.py:111-115`: ``juliacall.newmodule`` + ``include(rbc_sim2D_api.jl)``).
PyTorch is used only for device memory and streams; every numerical operation happens in the
hand-written CUDA kernels behind the C ABI.  There is no CPU fallback: constructing a simulation
without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path
from typing import Optional, Sequence, Union

import numpy as np

from . import build as _build
from .h5lite import Checkpoint2D, load_checkpoint_2d

NX, NZ = 96, 64
NCELL = NX * NZ
NSTATE = 2 * NCELL + NX * (NZ + 1)


class Rbc2dConfig(C.Structure):
    """Mirror of ``rbc2d_config`` (include/rbc_b200.h)."""

    _fields_ = [
        ("num_envs", C.c_int32),
        ("nx", C.c_int32), ("nz", C.c_int32),
        ("obs_nx", C.c_int32), ("obs_nz", C.c_int32),
        ("heaters", C.c_int32),
        ("heater_limit", C.c_double),
        ("ra", C.c_double), ("pr", C.c_double),
        ("dt_action", C.c_double), ("dt_solver", C.c_double),
        ("episode_length", C.c_double),
        ("precision", C.c_int32), ("pressure", C.c_int32), ("device", C.c_int32),
    ]


class Rbc2dWrappers(C.Structure):
    """Mirror of ``rbc2d_wrappers`` (include/rbc_b200.h)."""

    _fields_ = [
        ("normalize_obs", C.c_int32), ("obs_clip", C.c_int32),
        ("obs_lo", C.c_float * 4), ("obs_hi", C.c_float * 4), ("obs_maxval", C.c_float),
        ("normalize_reward", C.c_int32), ("reward_scale", C.c_double),
        ("shaping", C.c_int32), ("shaping_weight", C.c_double),
    ]


class Rbc3dConfig(C.Structure):
    """Mirror of ``rbc3d_config`` (include/rbc_b200.h)."""

    _fields_ = [
        ("num_envs", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("heaters", C.c_int32),
        ("heater_limit", C.c_double), ("ra", C.c_double), ("pr", C.c_double),
        ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double), ("b_min", C.c_double), ("b_max", C.c_double),
        ("heater_duration", C.c_double), ("dt_solver", C.c_double), ("episode_length", C.c_double),
        ("precision", C.c_int32), ("split", C.c_int32), ("device", C.c_int32),
    ]


class RbcAutoreset(C.Structure):
    """Mirror of ``rbc_autoreset`` (include/rbc_b200.h)."""

    _fields_ = [("mode", C.c_int32), ("nan_reset", C.c_int32), ("seed", C.c_int64), ("env_id_offset", C.c_int64)]


class Rbc2dVecOut(C.Structure):
    """Mirror of ``rbc2d_vec_out`` (include/rbc_b200.h): device pointers."""

    _fields_ = [(n, C.c_void_p) for n in ("obs", "reward", "nu_state", "nu_obs", "truncated", "nan", "t", "step", "episode_return",
                                          "final_obs", "final_nu_state", "final_nu_obs", "final_return")]


class Rbc3dVecOut(C.Structure):
    """Mirror of ``rbc3d_vec_out`` (include/rbc_b200.h): device pointers."""

    _fields_ = [(n, C.c_void_p) for n in ("obs", "reward", "nusselt", "truncated", "nan", "t", "step", "episode_return",
                                          "final_obs", "final_nusselt", "final_return")]


AUTORESET_MODES = {"disabled": 0, "next_step": 1, "same_step": 2}

# every symbol include/rbc_b200.h declares; tests check that the built library exports them all
ABI_SYMBOLS = (
    "rbc_abi_version", "rbc_last_error", "rbc2d_create", "rbc2d_destroy", "rbc2d_set_stream", "rbc2d_num_envs",
    "rbc2d_state_values_per_env", "rbc2d_load_checkpoints", "rbc2d_reset_from_checkpoints_dev",
    "rbc2d_reset_from_fields_host", "rbc2d_reset_from_fields_dev", "rbc2d_step_dev", "rbc2d_step_host", "rbc2d_observe_dev", "rbc2d_observe_host",
    "rbc2d_get_state_dev", "rbc2d_get_state_host", "rbc2d_get_fields_host", "rbc2d_get_info_host",
    "rbc2d_launch_count", "rbc2d_last_step_kernel_ms", "rbc2d_step_kernel_ms_history", "rbc2d_set_wrappers", "rbc2d_get_cell_dist_host", "rbc2d_render_rgb_dev",
    "rbc3d_create", "rbc3d_destroy", "rbc3d_set_stream", "rbc3d_state_values_per_env", "rbc3d_load_checkpoints",
    "rbc3d_reset_from_checkpoints_dev", "rbc3d_reset_from_fields_host", "rbc3d_reset_from_fields_dev", "rbc3d_step_dev", "rbc3d_step_host",
    "rbc3d_observe_dev", "rbc3d_get_fields_host", "rbc3d_get_info_host", "rbc3d_launch_count", "rbc3d_last_step_kernel_ms",
    "rbc_checkpoint_draw", "rbc2d_set_autoreset", "rbc2d_vec_reset_dev", "rbc2d_vec_mark_reset_dev", "rbc2d_vec_step_dev", "rbc2d_vec_step_host",
    "rbc2d_vec_nan_count", "rbc2d_vec_nan_count_async", "rbc3d_set_autoreset", "rbc3d_vec_reset_dev", "rbc3d_vec_mark_reset_dev",
    "rbc3d_vec_step_dev", "rbc3d_vec_nan_count", "rbc3d_vec_nan_count_async", "rbc3d_set_rayleigh_per_env", "rbc3d_render_rgb_dev", "rbc2d_set_cfl_guard", "rbc2d_get_cfl_events_host",
)
ABI_VERSION = 2

_lib = None


class BackendUnavailable(RuntimeError):
    pass


def library_path() -> Path:
    return _build.LIB


def load_library(build_if_missing: bool = True):
    """dlopen the C-ABI library; builds it in-tree with nvcc when missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("RBC_B200_LIB")          # developer knob: load another build of the same ABI (the jitter stress build)
    if override:
        if not Path(override).exists():
            raise BackendUnavailable(f"RBC_B200_LIB={override} does not exist")
        L = C.CDLL(override)
        _declare(L)
        _lib = L
        return L
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # a stale library must never be loaded silently: its struct layouts or kernels may differ
            raise BackendUnavailable(
                f"{_build.LIB} is missing or older than its sources and could not be rebuilt ({e}); "
                "run `python -m rbc_gym_b200.build`") from e
    if not _build.LIB.exists():
        raise BackendUnavailable(f"{_build.LIB} is missing; run `python -m rbc_gym_b200.build`")
    L = C.CDLL(str(_build.LIB))
    _declare(L)
    _lib = L
    return L


def _declare(L):
    """argtypes / restypes of every entry point + the ABI version check."""
    vp, ip = C.c_void_p, C.c_int32
    L.rbc_abi_version.restype = C.c_int
    L.rbc_last_error.restype = C.c_char_p
    L.rbc2d_create.argtypes = [C.POINTER(Rbc2dConfig), C.POINTER(vp)]
    L.rbc2d_destroy.argtypes = [vp]
    L.rbc2d_set_stream.argtypes = [vp, vp]
    L.rbc2d_num_envs.argtypes = [vp]
    L.rbc2d_state_values_per_env.argtypes = [vp]
    L.rbc2d_load_checkpoints.argtypes = [vp, vp, vp, vp, ip]
    L.rbc2d_reset_from_checkpoints_dev.argtypes = [vp, vp, vp, ip]
    L.rbc2d_reset_from_fields_host.argtypes = [vp, vp, vp, ip, ip]
    L.rbc2d_reset_from_fields_dev.argtypes = [vp, vp, vp, ip, ip]
    L.rbc2d_step_dev.argtypes = [vp] * 8
    L.rbc2d_step_host.argtypes = [vp] * 8
    L.rbc2d_observe_dev.argtypes = [vp] * 4
    L.rbc2d_observe_host.argtypes = [vp] * 4
    L.rbc2d_get_state_dev.argtypes = [vp, vp, ip]
    L.rbc2d_get_state_host.argtypes = [vp, vp, ip]
    L.rbc2d_get_fields_host.argtypes = [vp, vp]
    L.rbc2d_get_info_host.argtypes = [vp, vp, vp]
    L.rbc2d_set_wrappers.argtypes = [vp, C.POINTER(Rbc2dWrappers)]
    L.rbc2d_get_cell_dist_host.argtypes = [vp, vp]
    L.rbc2d_render_rgb_dev.argtypes = [vp, vp]
    L.rbc2d_set_cfl_guard.argtypes = [vp, C.c_double]
    L.rbc2d_get_cfl_events_host.argtypes = [vp, vp]
    L.rbc2d_launch_count.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(ip), C.POINTER(ip)]
    L.rbc2d_last_step_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.rbc2d_step_kernel_ms_history.argtypes = [vp, C.POINTER(C.c_float), ip]
    L.rbc3d_create.argtypes = [C.POINTER(Rbc3dConfig), C.POINTER(vp)]
    L.rbc3d_destroy.argtypes = [vp]
    L.rbc3d_set_stream.argtypes = [vp, vp]
    L.rbc3d_state_values_per_env.argtypes = [vp]
    L.rbc3d_load_checkpoints.argtypes = [vp, vp, ip]
    L.rbc3d_reset_from_checkpoints_dev.argtypes = [vp, vp, vp, ip]
    L.rbc3d_reset_from_fields_host.argtypes = [vp, vp, vp, ip, ip]
    L.rbc3d_reset_from_fields_dev.argtypes = [vp, vp, vp, ip, ip]
    L.rbc3d_step_dev.argtypes = [vp] * 7
    L.rbc3d_step_host.argtypes = [vp] * 7
    L.rbc3d_observe_dev.argtypes = [vp] * 3
    L.rbc3d_get_fields_host.argtypes = [vp, vp]
    L.rbc3d_get_info_host.argtypes = [vp, vp, vp]
    L.rbc3d_launch_count.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(ip), C.POINTER(ip)]
    L.rbc3d_last_step_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.rbc3d_set_rayleigh_per_env.argtypes = [vp, vp]
    L.rbc3d_render_rgb_dev.argtypes = [vp, vp, ip, ip]
    L.rbc_checkpoint_draw.argtypes = [C.c_int64, C.c_int64, C.c_int64, ip]
    L.rbc_checkpoint_draw.restype = ip
    for dim, out_t in (("2d", Rbc2dVecOut), ("3d", Rbc3dVecOut)):
        getattr(L, f"rbc{dim}_set_autoreset").argtypes = [vp, C.POINTER(RbcAutoreset)]
        getattr(L, f"rbc{dim}_vec_reset_dev").argtypes = [vp, vp]
        getattr(L, f"rbc{dim}_vec_mark_reset_dev").argtypes = [vp, vp, ip]
        getattr(L, f"rbc{dim}_vec_step_dev").argtypes = [vp, vp, C.POINTER(out_t)]
        if dim == "2d":
            L.rbc2d_vec_step_host.argtypes = [vp, vp, C.POINTER(out_t)]
        getattr(L, f"rbc{dim}_vec_nan_count").argtypes = [vp, ip, C.POINTER(C.c_int64)]
        getattr(L, f"rbc{dim}_vec_nan_count_async").argtypes = [vp, vp]
    if L.rbc_abi_version() != ABI_VERSION:
        raise BackendUnavailable(f"librbc_b200.so has ABI version {L.rbc_abi_version()}, this package needs {ABI_VERSION}; "
                                 "rebuild with `python -m rbc_gym_b200.build`")


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def substep_schedule(dt_action: float, dt_solver: float = 0.03):
    """run!'s ``dt' = min(dt_solver, stop_time - t)`` from t = 0 (same rule as the C++ side)."""
    out, t = [], 0.0
    while dt_action - t > 1e-10:
        d = min(dt_solver, dt_action - t)
        out.append(d)
        t += d
    return out


class Sim2D:
    """A batch of B independent 2D RBC environments on one GPU (one `rbc2d_sim` handle).

    Mirrors the Julia API surface of `rbc_sim2D_api.jl` in batched form:
    ``initialize_simulation`` -> constructor + `reset_*`, ``step_simulation`` -> `step`,
    ``get_state`` -> `get_state`, ``get_observation``/``get_nusselt`` -> outputs of `step`/`observe`,
    ``get_info`` -> `info`.
    """

    def __init__(self, num_envs: int, ra: float, dt_action: float = 1.5, *, obs_shape=(8, 48), state_shape=(64, 96),
                 heaters: int = 12, heater_limit: float = 0.75, pr: float = 0.7, dt_solver: float = 0.03,
                 episode_length: float = 300.0, precision: int = 32, pressure: bool = False, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise BackendUnavailable("rbc_gym_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = load_library()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.B = int(num_envs)
        self.obs_shape = (int(obs_shape[0]), int(obs_shape[1]))
        self.state_shape = (int(state_shape[0]), int(state_shape[1]))
        self.nz, self.nx = self.state_shape
        self.ncell = self.nx * self.nz
        self.nstate = 2 * self.ncell + self.nx * (self.nz + 1)
        self.heaters = int(heaters)
        self.channels = 5 if pressure else 3
        self.precision = int(precision)
        self.ra, self.pr = float(ra), float(pr)
        self.kappa = 1.0 / math.sqrt(self.pr * self.ra)
        self.nu = math.sqrt(self.pr / self.ra)
        self.dt_action, self.dt_solver = float(dt_action), float(dt_solver)
        self.nsub = len(substep_schedule(self.dt_action, self.dt_solver))
        self.cfg = Rbc2dConfig(self.B, self.state_shape[1], self.state_shape[0], self.obs_shape[1], self.obs_shape[0],
                               self.heaters, float(heater_limit), self.ra, self.pr, self.dt_action, self.dt_solver,
                               float(episode_length), self.precision, int(bool(pressure)), device)
        h = C.c_void_p()
        self._h = None
        self._check(self._L.rbc2d_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        with torch.cuda.device(self.device):
            f32, f64, i32 = torch.float32, torch.float64, torch.int32
            self.obs = torch.zeros((self.B, self.channels, *self.obs_shape), dtype=f32, device=self.device)
            self.reward = torch.zeros(self.B, dtype=f32, device=self.device)
            self.nu_state = torch.zeros(self.B, dtype=f64, device=self.device)
            self.nu_obs = torch.zeros(self.B, dtype=f64, device=self.device)
            self.truncated = torch.zeros(self.B, dtype=i32, device=self.device)
            self.nan = torch.zeros(self.B, dtype=i32, device=self.device)
        self.n_episodes = 0

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"rbc_b200: {self._L.rbc_last_error().decode()}")

    def _use_current_stream(self):
        self._check(self._L.rbc2d_set_stream(self._h, C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rbc2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reset
    def load_checkpoints(self, ckpt: Union[str, Path, Checkpoint2D]) -> int:
        """Upload a checkpoint bank (`data/checkpoints/*/ckpt_ra*.h5`) to the device once."""
        if not isinstance(ckpt, Checkpoint2D):
            ckpt = load_checkpoint_2d(ckpt)
        if ckpt.shape != self.state_shape:
            raise ValueError(f"checkpoint grid {ckpt.shape} != state_shape {self.state_shape}")
        b, u, w = (np.ascontiguousarray(a, dtype=np.float64) for a in (ckpt.b, ckpt.u, ckpt.w))
        self._use_current_stream()
        self._check(self._L.rbc2d_load_checkpoints(self._h, _np_ptr(b), _np_ptr(u), _np_ptr(w), ckpt.num_episodes))
        self.n_episodes = ckpt.num_episodes
        return self.n_episodes

    def reset_from_checkpoints(self, ckpt_idx, env_ids=None):
        """`initialize_simulation(checkpoint_path=...)` for all or the listed environments.

        ckpt_idx / env_ids: int32 CUDA tensors (or anything convertible)."""
        t = self.torch
        idx = t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device).contiguous()
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        n = self.B if ids is None else int(ids.numel())
        if idx.numel() != n:
            raise ValueError("ckpt_idx must have one entry per environment being reset")
        self._validate_indices(idx, self.n_episodes, "checkpoint_idx")
        if ids is not None:
            self._validate_indices(ids, self.B, "env_ids")
        self._use_current_stream()
        self._check(self._L.rbc2d_reset_from_checkpoints_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                            C.c_void_p(idx.data_ptr()), n))

    def reset_from_fields(self, fields: np.ndarray, env_ids: Optional[Sequence[int]] = None, project: bool = True):
        """Reset from explicit `[n, 2*nx*nz + nx*(nz+1)]` float64 fields (b,u,w); `project` mirrors Oceananigans' `set!`."""
        f = np.ascontiguousarray(fields, dtype=np.float64).reshape(-1, self.nstate)
        ids = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32)
        if ids is not None and (ids.size != f.shape[0] or (ids.size and (ids.min() < 0 or ids.max() >= self.B))):
            raise IndexError(f"env_ids must list {f.shape[0]} environments in [0, {self.B})")
        self._use_current_stream()
        self._check(self._L.rbc2d_reset_from_fields_host(self._h, _np_ptr(ids), _np_ptr(f), f.shape[0], int(project)))

    def reset_from_fields_dev(self, fields, env_ids=None, project: bool = True):
        """`reset_from_fields` with float64 CUDA tensors `[n, nstate]` (and int32 CUDA `env_ids`): nothing touches the host."""
        t = self.torch
        f = t.as_tensor(fields, dtype=t.float64, device=self.device).reshape(-1, self.nstate).contiguous()
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        if ids is not None and ids.numel() != f.shape[0]:
            raise ValueError("env_ids must have one entry per field row")
        if ids is not None:
            self._validate_indices(ids, self.B, "env_ids")
        self._use_current_stream()
        self._check(self._L.rbc2d_reset_from_fields_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                       C.c_void_p(f.data_ptr()), f.shape[0], int(project)))

    def noise_reset(self, env_ids=None, kick: float = 0.01, generator=None):
        """`initialize_model` (`rbc_sim2D.jl:163-171`) for many environments on the device: u, w = kick*randn,
        b = clamp(min_b + (Lz - z) db/2 + kick*randn, min_b, min_b + db), walls impenetrable, then the `set!` projection."""
        t = self.torch
        n = self.B if env_ids is None else int(t.as_tensor(env_ids).numel())
        nz, nx = self.nz, self.nx
        z = (t.arange(nz, device=self.device, dtype=t.float64) + 0.5) * (2.0 / nz)
        rn = lambda *shape: t.randn(shape, device=self.device, dtype=t.float64, generator=generator)
        b = t.clamp(1.0 + (2.0 - z)[None, :, None] * 0.5 + kick * rn(n, nz, nx), 1.0, 2.0)
        u = kick * rn(n, nz, nx)
        w = kick * rn(n, nz + 1, nx)
        w[:, 0] = 0.0
        w[:, -1] = 0.0
        self.reset_from_fields_dev(t.cat([b.reshape(n, -1), u.reshape(n, -1), w.reshape(n, -1)], dim=1), env_ids, project=True)

    # ------------------------------------------------------------------ step / observe
    def step(self, actions):
        """One action step for the whole batch; actions `[B, heaters]` float32 CUDA tensor.

        Returns views of the handle-owned output tensors (obs, reward, nu_state, nu_obs, truncated, nan)."""
        t = self.torch
        a = t.as_tensor(actions, dtype=t.float32, device=self.device).contiguous()
        if a.shape != (self.B, self.heaters):
            raise ValueError(f"actions must have shape {(self.B, self.heaters)}, got {tuple(a.shape)}")
        self._use_current_stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        self._check(self._L.rbc2d_step_dev(self._h, p(a), p(self.obs), p(self.reward), p(self.nu_state), p(self.nu_obs),
                                           p(self.truncated), p(self.nan)))
        return self.obs, self.reward, self.nu_state, self.nu_obs, self.truncated, self.nan

    def step_host(self, actions: np.ndarray, out: Optional[dict] = None) -> dict:
        """Same step through host buffers (H2D of the actions and D2H of all outputs inside the call)."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.B, self.heaters):
            raise ValueError(f"actions must have shape {(self.B, self.heaters)}, got {a.shape}")
        if out is None:
            out = self.alloc_host_outputs()
        self._use_current_stream()
        self._check(self._L.rbc2d_step_host(self._h, _np_ptr(a), _np_ptr(out["obs"]), _np_ptr(out["reward"]),
                                            _np_ptr(out["nu_state"]), _np_ptr(out["nu_obs"]), _np_ptr(out["truncated"]),
                                            _np_ptr(out["nan"])))
        return out

    def alloc_host_outputs(self, pinned: bool = False) -> dict:
        shapes = {"obs": ((self.B, self.channels, *self.obs_shape), np.float32), "reward": ((self.B,), np.float32),
                  "nu_state": ((self.B,), np.float64), "nu_obs": ((self.B,), np.float64),
                  "truncated": ((self.B,), np.int32), "nan": ((self.B,), np.int32)}
        if not pinned:
            return {k: np.zeros(s, d) for k, (s, d) in shapes.items()}
        t = self.torch
        return {k: t.zeros(s, dtype=getattr(t, np.dtype(d).name)).pin_memory().numpy() for k, (s, d) in shapes.items()}

    # ------------------------------------------------------------------ fused vector-env step
    def set_autoreset(self, mode: str = "next_step", nan_reset: bool = False, seed: int = 0, env_id_offset: int = 0):
        """Configure the vector-env semantics fused into the step kernel (`rbc2d_set_autoreset`)."""
        self._check(self._L.rbc2d_set_autoreset(self._h, C.byref(RbcAutoreset(AUTORESET_MODES[mode], int(nan_reset), int(seed), int(env_id_offset)))))
        self.autoreset_mode = mode

    def _vec_buffers(self):
        if getattr(self, "_vec", None) is None:
            t = self.torch
            f32, f64, i32 = t.float32, t.float64, t.int32
            mk = lambda shape, dt: t.zeros(shape, dtype=dt, device=self.device)
            v = {"t": mk(self.B, f64), "step": mk(self.B, i32), "episode_return": mk(self.B, f64),
                 "final_obs": mk((self.B, self.channels, *self.obs_shape), f32), "final_nu_state": mk(self.B, f64),
                 "final_nu_obs": mk(self.B, f64), "final_return": mk(self.B, f64)}
            out = Rbc2dVecOut(*[C.c_void_p(x.data_ptr()) for x in (self.obs, self.reward, self.nu_state, self.nu_obs, self.truncated,
                                                                  self.nan, v["t"], v["step"], v["episode_return"], v["final_obs"],
                                                                  v["final_nu_state"], v["final_nu_obs"], v["final_return"])])
            self._vec, self._vec_out = v, out
        return self._vec

    def vec_reset(self, ckpt_idx=None):
        """`reset()` of the whole batch from the checkpoint bank: episode 0 of every environment (its own draw, or `ckpt_idx`)."""
        t = self.torch
        idx = None if ckpt_idx is None else t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device).contiguous()
        if idx is not None:
            self._validate_indices(idx, self.n_episodes, "checkpoint_idx", self.B)
        self._use_current_stream()
        self._check(self._L.rbc2d_vec_reset_dev(self._h, None if idx is None else C.c_void_p(idx.data_ptr())))

    def vec_mark_reset(self, env_ids=None):
        """Episode bookkeeping for environments the caller re-initialised itself (noise initialisation)."""
        t = self.torch
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        self._use_current_stream()
        self._check(self._L.rbc2d_vec_mark_reset_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                    self.B if ids is None else int(ids.numel())))

    def vec_step(self, actions):
        """One vector-env step in ONE kernel launch: march, observation, reward, truncation AND the auto-reset of the configured
        mode.  Returns (obs, reward, nu_state, nu_obs, truncated, nan, extras) — handle-owned tensors that the next call
        overwrites; `extras` holds t, step, episode_return and the final_* outputs."""
        t = self.torch
        a = t.as_tensor(actions, dtype=t.float32, device=self.device).contiguous()
        if a.shape != (self.B, self.heaters):
            raise ValueError(f"actions must have shape {(self.B, self.heaters)}, got {tuple(a.shape)}")
        v = self._vec_buffers()
        self._use_current_stream()
        self._check(self._L.rbc2d_vec_step_dev(self._h, C.c_void_p(a.data_ptr()), C.byref(self._vec_out)))
        return self.obs, self.reward, self.nu_state, self.nu_obs, self.truncated, self.nan, v

    VEC_HOST_KEYS = ("obs", "reward", "nu_state", "nu_obs", "truncated", "nan", "t", "step", "episode_return",
                     "final_obs", "final_nu_state", "final_nu_obs", "final_return")

    def alloc_vec_host_outputs(self, pinned: bool = True, final: bool = False) -> dict:
        """Host buffers for `vec_step_host` (page-locked by default so that the copies overlap the kernel)."""
        obs = (self.B, self.channels, *self.obs_shape)
        shapes = {"obs": (obs, np.float32), "reward": ((self.B,), np.float32), "nu_state": ((self.B,), np.float64),
                  "nu_obs": ((self.B,), np.float64), "truncated": ((self.B,), np.int32), "nan": ((self.B,), np.int32),
                  "t": ((self.B,), np.float64), "step": ((self.B,), np.int32), "episode_return": ((self.B,), np.float64)}
        if final:
            shapes.update({"final_obs": (obs, np.float32), "final_nu_state": ((self.B,), np.float64),
                           "final_nu_obs": ((self.B,), np.float64), "final_return": ((self.B,), np.float64)})
        t = self.torch
        if not pinned:
            return {k: np.zeros(s, d) for k, (s, d) in shapes.items()}
        return {k: t.zeros(s, dtype=getattr(t, np.dtype(d).name)).pin_memory().numpy() for k, (s, d) in shapes.items()}

    def vec_step_host(self, actions: np.ndarray, out: dict) -> dict:
        """The vector step through host buffers (`rbc2d_vec_step_host`): H2D of the actions, one fused launch per chunk of whole
        waves, D2H of every buffer present in `out` overlapped with the next chunk."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (self.B, self.heaters):
            raise ValueError(f"actions must have shape {(self.B, self.heaters)}, got {a.shape}")
        o = Rbc2dVecOut(*[_np_ptr(out.get(k)) for k in self.VEC_HOST_KEYS])
        self._use_current_stream()
        self._check(self._L.rbc2d_vec_step_host(self._h, _np_ptr(a), C.byref(o)))
        return out

    def vec_nan_count(self, clear: bool = False) -> int:
        """Environments that reported NaNs in vector steps since the counter was cleared (synchronises)."""
        n = C.c_int64()
        self._use_current_stream()
        self._check(self._L.rbc2d_vec_nan_count(self._h, int(clear), C.byref(n)))
        return n.value

    def vec_nan_count_async(self, pinned_int32):
        """Enqueue a copy of the NaN counter into a pinned int32 host tensor (no synchronisation)."""
        self._use_current_stream()
        self._check(self._L.rbc2d_vec_nan_count_async(self._h, C.c_void_p(pinned_int32.data_ptr())))

    def _validate_indices(self, idx, upper: int, what: str, expect_n: Optional[int] = None):
        """Raise like the reference would (Julia BoundsError on a bad checkpoint index) instead of clamping or writing out of
        bounds.  One small reduction + host read; resets are off the hot path."""
        if expect_n is not None and int(idx.numel()) != expect_n:
            raise ValueError(f"{what} must have {expect_n} entries, got {int(idx.numel())}")
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= upper):
            raise IndexError(f"{what} out of range [0, {upper})")

    def set_wrappers(self, *, normalize_obs: bool = False, u_limit: float = 1.3, maxval: float = 1.0, clip: bool = False,
                     normalize_reward: bool = False, shaping_weight: Optional[float] = None):
        """Fuse the reference wrappers into the step epilogue (`src/rbc_gym/wrappers/`), in the order of
        `example/run_wrapped.py`: RBCNormalizeObservation -> RBCNormalizeReward -> RBCRewardShaping."""
        w = Rbc2dWrappers()
        w.normalize_obs, w.obs_clip, w.obs_maxval = int(normalize_obs), int(clip), float(maxval)
        lo = [1.0, -u_limit, -u_limit, -u_limit]                       # rbc_normalize_observation.py:45-57
        hi = [2.0 + self.cfg.heater_limit, u_limit, u_limit, u_limit]
        for c in range(4):
            w.obs_lo[c], w.obs_hi[c] = lo[c], hi[c]
        w.normalize_reward = int(normalize_reward)
        w.reward_scale = 0.1 * self.ra ** 0.4                          # rbc_normalize_reward.py:19-25 (2D)
        w.shaping = int(shaping_weight is not None)
        w.shaping_weight = float(shaping_weight or 0.0)
        self._check(self._L.rbc2d_set_wrappers(self._h, C.byref(w)))
        self.wrappers = w

    def set_cfl_guard(self, limit: float):
        """CFL guard of the grids beyond 96 x 64 (`rbc2d_set_cfl_guard`): RK3 steps above `limit` are split; 0 switches it off."""
        self._check(self._L.rbc2d_set_cfl_guard(self._h, float(limit)))

    def cfl_events(self) -> np.ndarray:
        """Extra RK3 steps the CFL guard inserted per environment since creation, `[B]` int32."""
        out = np.zeros(self.B, np.int32)
        self._use_current_stream()
        self._check(self._L.rbc2d_get_cfl_events_host(self._h, _np_ptr(out)))
        return out

    def cell_dist(self) -> np.ndarray:
        """`info["cell_dist"]` of the last step (`rbc_reward_shaping.py:61-66`), `[B]` float64."""
        out = np.empty(self.B, np.float64)
        self._use_current_stream()
        self._check(self._L.rbc2d_get_cell_dist_host(self._h, _np_ptr(out)))
        return out

    def observe(self):
        """`get_observation` + `get_nusselt` of the current state without stepping."""
        self._use_current_stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        self._check(self._L.rbc2d_observe_dev(self._h, p(self.obs), p(self.nu_state), p(self.nu_obs)))
        return self.obs, self.nu_state, self.nu_obs

    def get_state(self, channels: Optional[int] = None):
        """`get_state` in the Python layout: float32 CUDA tensor `[B, C, Nz, Nx]`."""
        ch = self.channels if channels is None else int(channels)
        out = self.torch.empty((self.B, ch, self.nz, self.nx), dtype=self.torch.float32, device=self.device)
        self._use_current_stream()
        self._check(self._L.rbc2d_get_state_dev(self._h, C.c_void_p(out.data_ptr()), ch))
        return out

    def render_rgb(self):
        """`render("rgb_array")` for every environment, on the device: uint8 CUDA tensor `[B, Nz, Nx, 3]`, origin top left."""
        out = self.torch.empty((self.B, self.nz, self.nx, 3), dtype=self.torch.uint8, device=self.device)
        self._use_current_stream()
        self._check(self._L.rbc2d_render_rgb_dev(self._h, C.c_void_p(out.data_ptr())))
        return out

    def fields(self) -> np.ndarray:
        """Raw (b,u,w) in checkpoint layout `[B, 2*nx*nz + nx*(nz+1)]` float64 on the host."""
        out = np.empty((self.B, self.nstate), np.float64)
        self._use_current_stream()
        self._check(self._L.rbc2d_get_fields_host(self._h, _np_ptr(out)))
        return out

    def info(self):
        """`get_info`: (t[B] float64, step[B] int32) on the host."""
        t = np.empty(self.B, np.float64)
        s = np.empty(self.B, np.int32)
        self._use_current_stream()
        self._check(self._L.rbc2d_get_info_host(self._h, _np_ptr(t), _np_ptr(s)))
        return t, s

    def launch_info(self):
        n, g, s = C.c_int64(), C.c_int32(), C.c_int32()
        self._check(self._L.rbc2d_launch_count(self._h, C.byref(n), C.byref(g), C.byref(s)))
        return {"launches": n.value, "grid": g.value, "smem_bytes": s.value}

    def last_step_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.rbc2d_last_step_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def step_kernel_ms_history(self, n: int):
        """Device times (ms) of the last `n` step kernels (n <= 64), oldest first, from the library's per-launch CUDA events."""
        buf = (C.c_float * n)()
        got = self._L.rbc2d_step_kernel_ms_history(self._h, buf, n)
        if got < 0:
            self._check(got)
        return [buf[i] for i in range(got)]


def split_fields(fields: np.ndarray, shape=(NZ, NX)):
    """`[B, nstate]` -> b[B,nz,nx], u[B,nz,nx], w[B,nz+1,nx] (default grid 64 x 96: `[B, 18528]`)."""
    B = fields.shape[0]
    nz, nx = shape
    nc = nz * nx
    return (fields[:, :nc].reshape(B, nz, nx), fields[:, nc:2 * nc].reshape(B, nz, nx), fields[:, 2 * nc:].reshape(B, nz + 1, nx))


def pack_fields(b, u, w) -> np.ndarray:
    b, u, w = (np.asarray(a, dtype=np.float64) for a in (b, u, w))
    B = b.shape[0]
    return np.concatenate([b.reshape(B, -1), u.reshape(B, -1), w.reshape(B, -1)], axis=1)


# ============================================================================================ 3D
NX3, NY3, NZ3 = 32, 32, 16
NC3 = NX3 * NY3 * NZ3
NSTATE3 = 3 * NC3 + NX3 * NY3 * (NZ3 + 1)


class Sim3D:
    """A batch of B independent 3D RBC environments on one GPU (`rbc3d_sim` handle): the batched form of
    `rbc_sim3D_api.jl` (`initialize_simulation`, `step_simulation`, `get_state`, `get_info`, `get_nusselt`)."""

    def __init__(self, num_envs: int, ra: float = 2500, *, pr: float = 0.7, domain=(2.0, 4 * math.pi, 4 * math.pi),
                 state_shape=(16, 32, 32), temperature_difference=(1.0, 2.0), heaters: int = 8, heater_limit: float = 0.9,
                 heater_duration: float = 0.125, dt_solver: float = 0.01, episode_length: float = 300.0, precision: int = 32,
                 split: bool = False, device: int = 0):
        """`split=True` carries Oceananigans' hydrostatic-pressure anomaly explicitly (buoyancy through pHY'); the default adds
        the buoyancy to G_w instead.  The two differ by a discrete gradient that the projection removes, so velocities and b
        agree to round-off (oracle test), and the 3D state exposes no pressure channel.  fp32 + `split=False` runs the tiled
        kernel (shared-memory tile per half domain)."""
        import torch

        if not torch.cuda.is_available():
            raise BackendUnavailable("rbc_gym_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = load_library()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.B, self.heaters, self.precision = int(num_envs), int(heaters), int(precision)
        self.state_shape = tuple(int(x) for x in state_shape)
        nz, ny, nx = self.state_shape
        self.ncell = nx * ny * nz
        self.nstate = 3 * self.ncell + nx * ny * (nz + 1)
        lz, ly, lx = (float(x) for x in domain)                       # rbc3D.py:173 passes L = domain[::-1]
        self.ra, self.pr = float(ra), float(pr)
        self.kappa = 1.0 / math.sqrt(self.pr * self.ra)
        self.t_ff = lz * lz
        self.nsub = len(substep_schedule(heater_duration * self.t_ff, dt_solver * self.t_ff))
        self.cfg = Rbc3dConfig(self.B, nx, ny, nz, self.heaters, float(heater_limit), self.ra, self.pr, lx, ly, lz,
                               float(temperature_difference[0]), float(temperature_difference[1]), float(heater_duration),
                               float(dt_solver), float(episode_length), self.precision, int(bool(split)), device)
        h = C.c_void_p()
        self._h = None
        self._check(self._L.rbc3d_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        f32, f64, i32 = torch.float32, torch.float64, torch.int32
        self.obs = torch.zeros((self.B, 4, nz, ny, nx), dtype=f32, device=self.device)
        self.reward = torch.zeros(self.B, dtype=f32, device=self.device)
        self.nusselt = torch.zeros(self.B, dtype=f64, device=self.device)
        self.truncated = torch.zeros(self.B, dtype=i32, device=self.device)
        self.nan = torch.zeros(self.B, dtype=i32, device=self.device)
        self.n_episodes = 0

    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"rbc_b200: {self._L.rbc_last_error().decode()}")

    def _use_current_stream(self):
        self._check(self._L.rbc3d_set_stream(self._h, C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rbc3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def fused_autoreset(self) -> bool:
        """The auto-reset is fused into the step on every grid: in the epilogue of the dedicated 32 x 32 x 16 kernel, and as batch
        kernels behind the march of the stage-streaming path (`rbc3dg_lib.cu`)."""
        return True

    def set_rayleigh(self, ra_per_env):
        """One Rayleigh number per environment (`rbc3d_set_rayleigh_per_env`; stage-streaming kernels, i.e. grids other than
        32 x 32 x 16): a whole Rayleigh sweep like `experiments/flowstats/flowstats_ra.py:27-36` runs as one batch."""
        ra = np.ascontiguousarray(ra_per_env, dtype=np.float64)
        if ra.shape != (self.B,):
            raise ValueError(f"need {self.B} Rayleigh numbers")
        self._check(self._L.rbc3d_set_rayleigh_per_env(self._h, _np_ptr(ra)))
        self.ra_per_env = ra
        self.kappa_per_env = 1.0 / np.sqrt(self.pr * ra)

    def load_checkpoints(self, fields: np.ndarray) -> int:
        """Upload a bank `[n_ep, nstate]` float64 (b,u,v,w in checkpoint layout; 66560 values on the registered grid)."""
        f = np.ascontiguousarray(fields, dtype=np.float64).reshape(-1, self.nstate)
        self._use_current_stream()
        self._check(self._L.rbc3d_load_checkpoints(self._h, _np_ptr(f), f.shape[0]))
        self.n_episodes = f.shape[0]
        return self.n_episodes

    def reset_from_checkpoints(self, ckpt_idx, env_ids=None):
        t = self.torch
        idx = t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device).contiguous()
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        n = self.B if ids is None else int(ids.numel())
        self._validate_indices(idx, self.n_episodes, "checkpoint_idx", n)
        if ids is not None:
            self._validate_indices(ids, self.B, "env_ids")
        self._use_current_stream()
        self._check(self._L.rbc3d_reset_from_checkpoints_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                            C.c_void_p(idx.data_ptr()), n))

    def reset_from_fields(self, fields: np.ndarray, env_ids: Optional[Sequence[int]] = None, project: bool = True):
        f = np.ascontiguousarray(fields, dtype=np.float64).reshape(-1, self.nstate)
        ids = None if env_ids is None else np.ascontiguousarray(env_ids, dtype=np.int32)
        self._use_current_stream()
        self._check(self._L.rbc3d_reset_from_fields_host(self._h, _np_ptr(ids), _np_ptr(f), f.shape[0], int(project)))

    def reset_from_fields_dev(self, fields, env_ids=None, project: bool = True):
        """`reset_from_fields` with float64 CUDA tensors `[n, nstate]` (and int32 CUDA `env_ids`): nothing touches the host."""
        t = self.torch
        f = t.as_tensor(fields, dtype=t.float64, device=self.device).reshape(-1, self.nstate).contiguous()
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        if ids is not None:
            self._validate_indices(ids, self.B, "env_ids", f.shape[0])
        self._use_current_stream()
        self._check(self._L.rbc3d_reset_from_fields_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                       C.c_void_p(f.data_ptr()), f.shape[0], int(project)))

    def noise_reset(self, env_ids=None, kick: float = 0.01, generator=None):
        """`initialize_model` (`rbc_sim3D.jl:169-178`) drawn and projected on the device for all or the listed environments."""
        t = self.torch
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        n = self.B if ids is None else int(ids.numel())
        nz, ny, nx = self.state_shape
        lz, b0, db = float(self.cfg.lz), float(self.cfg.b_min), float(self.cfg.b_max - self.cfg.b_min)
        z = (t.arange(nz, device=self.device, dtype=t.float64) + 0.5) * (lz / nz)
        rn = lambda *shape: t.randn(shape, device=self.device, dtype=t.float64, generator=generator)
        b = t.clamp(b0 + (lz - z)[None, :, None, None] * db / 2 + kick * rn(n, nz, ny, nx), b0, b0 + db)
        u, v, w = kick * rn(n, nz, ny, nx), kick * rn(n, nz, ny, nx), kick * rn(n, nz + 1, ny, nx)
        w[:, 0] = 0.0
        w[:, -1] = 0.0
        f = t.cat([x.reshape(n, -1) for x in (b, u, v, w)], dim=1).contiguous()
        self._use_current_stream()
        self._check(self._L.rbc3d_reset_from_fields_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                       C.c_void_p(f.data_ptr()), n, 1))

    # ------------------------------------------------------------------ fused vector-env step
    _validate_indices = Sim2D._validate_indices

    def set_autoreset(self, mode: str = "next_step", nan_reset: bool = False, seed: int = 0, env_id_offset: int = 0):
        self._check(self._L.rbc3d_set_autoreset(self._h, C.byref(RbcAutoreset(AUTORESET_MODES[mode], int(nan_reset), int(seed), int(env_id_offset)))))
        self.autoreset_mode = mode

    def _vec_buffers(self, want_final_obs: bool):
        t = self.torch
        if getattr(self, "_vec", None) is None:
            mk = lambda shape, dt: t.zeros(shape, dtype=dt, device=self.device)
            self._vec = {"t": mk(self.B, t.float64), "step": mk(self.B, t.int32), "episode_return": mk(self.B, t.float64),
                         "final_nusselt": mk(self.B, t.float64), "final_return": mk(self.B, t.float64), "final_obs": None}
        v = self._vec
        if want_final_obs and v["final_obs"] is None:
            v["final_obs"] = t.zeros_like(self.obs)
        return v

    def vec_reset(self, ckpt_idx=None):
        t = self.torch
        idx = None if ckpt_idx is None else t.as_tensor(ckpt_idx, dtype=t.int32, device=self.device).contiguous()
        if idx is not None:
            self._validate_indices(idx, self.n_episodes, "checkpoint_idx", self.B)
        self._use_current_stream()
        self._check(self._L.rbc3d_vec_reset_dev(self._h, None if idx is None else C.c_void_p(idx.data_ptr())))

    def vec_mark_reset(self, env_ids=None):
        t = self.torch
        ids = None if env_ids is None else t.as_tensor(env_ids, dtype=t.int32, device=self.device).contiguous()
        self._use_current_stream()
        self._check(self._L.rbc3d_vec_mark_reset_dev(self._h, None if ids is None else C.c_void_p(ids.data_ptr()),
                                                    self.B if ids is None else int(ids.numel())))

    def vec_step(self, actions, want_obs: bool = True, want_final_obs: bool = True):
        """One vector-env step of the 3D batch in one launch (`rbc3d_vec_step_dev`); see `Sim2D.vec_step`."""
        t = self.torch
        a = t.as_tensor(actions, dtype=t.float32, device=self.device).contiguous()
        if a.shape != (self.B, self.heaters, self.heaters):
            raise RuntimeError(f"Action size does not match the number of actuators. Expected {(self.heaters, self.heaters)}, "
                               f"got {tuple(a.shape[1:])}.")
        v = self._vec_buffers(want_final_obs and want_obs)
        p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
        out = Rbc3dVecOut(p(self.obs) if want_obs else None, p(self.reward), p(self.nusselt), p(self.truncated), p(self.nan), p(v["t"]),
                          p(v["step"]), p(v["episode_return"]), p(v["final_obs"]) if (want_obs and want_final_obs) else None,
                          p(v["final_nusselt"]), p(v["final_return"]))
        self._use_current_stream()
        self._check(self._L.rbc3d_vec_step_dev(self._h, C.c_void_p(a.data_ptr()), C.byref(out)))
        return self.obs, self.reward, self.nusselt, self.truncated, self.nan, v

    def vec_nan_count(self, clear: bool = False) -> int:
        n = C.c_int64()
        self._use_current_stream()
        self._check(self._L.rbc3d_vec_nan_count(self._h, int(clear), C.byref(n)))
        return n.value

    def vec_nan_count_async(self, pinned_int32):
        self._use_current_stream()
        self._check(self._L.rbc3d_vec_nan_count_async(self._h, C.c_void_p(pinned_int32.data_ptr())))

    def step(self, actions, want_obs: bool = True):
        """actions `[B, heaters, heaters]` float32 CUDA tensor -> (obs, reward, nusselt, truncated, nan)."""
        t = self.torch
        a = t.as_tensor(actions, dtype=t.float32, device=self.device).contiguous()
        if a.shape != (self.B, self.heaters, self.heaters):
            # rbc_sim3D.jl:115-117: "Action size does not match the number of actuators"
            raise RuntimeError(f"Action size does not match the number of actuators. Expected {(self.heaters, self.heaters)}, "
                               f"got {tuple(a.shape[1:])}.")
        self._use_current_stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        self._check(self._L.rbc3d_step_dev(self._h, p(a), p(self.obs) if want_obs else None, p(self.reward), p(self.nusselt),
                                           p(self.truncated), p(self.nan)))
        return self.obs, self.reward, self.nusselt, self.truncated, self.nan

    def step_host(self, actions: np.ndarray, out: dict) -> dict:
        a = np.ascontiguousarray(actions, dtype=np.float32)
        self._use_current_stream()
        self._check(self._L.rbc3d_step_host(self._h, _np_ptr(a), _np_ptr(out.get("obs")), _np_ptr(out["reward"]),
                                            _np_ptr(out["nusselt"]), _np_ptr(out["truncated"]), _np_ptr(out["nan"])))
        return out

    def observe(self):
        self._use_current_stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        self._check(self._L.rbc3d_observe_dev(self._h, p(self.obs), p(self.nusselt)))
        return self.obs, self.nusselt

    def render_rgb(self, height: int = 608, width: int = 800):
        """`render("rgb_array")` of every environment on the device (`rbc3D.py:247-318`): uint8 CUDA tensor `[B, height, width, 3]`,
        a ray-marched volume rendering of the temperature (turbo colormap, clim = temperature_difference, sigmoid opacity)."""
        out = self.torch.empty((self.B, height, width, 3), dtype=self.torch.uint8, device=self.device)
        self._use_current_stream()
        self._check(self._L.rbc3d_render_rgb_dev(self._h, C.c_void_p(out.data_ptr()), height, width))
        return out

    def fields(self) -> np.ndarray:
        out = np.empty((self.B, self.nstate), np.float64)
        self._use_current_stream()
        self._check(self._L.rbc3d_get_fields_host(self._h, _np_ptr(out)))
        return out

    def info(self):
        t = np.empty(self.B, np.float64)
        s = np.empty(self.B, np.int32)
        self._use_current_stream()
        self._check(self._L.rbc3d_get_info_host(self._h, _np_ptr(t), _np_ptr(s)))
        return t, s

    def launch_info(self):
        n, g, s = C.c_int64(), C.c_int32(), C.c_int32()
        self._check(self._L.rbc3d_launch_count(self._h, C.byref(n), C.byref(g), C.byref(s)))
        return {"launches": n.value, "grid": g.value, "smem_bytes": s.value}

    def last_step_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self._L.rbc3d_last_step_kernel_ms(self._h, C.byref(ms)))
        return ms.value


def split_fields3(fields: np.ndarray, shape=(NZ3, NY3, NX3)):
    B = fields.shape[0]
    nz, ny, nx = shape
    nc = nz * ny * nx
    shp = (B, nz, ny, nx)
    return (fields[:, :nc].reshape(shp), fields[:, nc:2 * nc].reshape(shp), fields[:, 2 * nc:3 * nc].reshape(shp),
            fields[:, 3 * nc:].reshape(B, nz + 1, ny, nx))


def pack_fields3(b, u, v, w) -> np.ndarray:
    return np.concatenate([np.asarray(x, dtype=np.float64).reshape(np.shape(x)[0], -1) for x in (b, u, v, w)], axis=1)
