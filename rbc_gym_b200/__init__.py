"""rbc_gym_b200 — B200-native backend for RBC-Gym's simulation step.

Keeps the reference's API surface (`src/rbc_gym/__init__.py:4-38`): the gym ids
``rbc_gym/RayleighBenardConvection2D-v0`` (registered here when gymnasium is installed), the single-env class
`RayleighBenardConvection2DEnv`, the wrappers, and the `data/checkpoints` reset layout; adds the batched
on-device `RBCVectorEnv2D` that replaces the process-per-env vector env.
"""
from . import spaces

ENV_ID_2D = "rbc_gym/RayleighBenardConvection2D-v0"
DEFAULT_KWARGS_2D = {                      # src/rbc_gym/__init__.py:7-18
    "rayleigh_number": 10_000,
    "episode_length": 300,
    "observation_shape": (8, 48),
    "state_shape": (64, 96),
    "heater_segments": 12,
    "heater_limit": 0.75,
    "heater_duration": 1.5,
    "checkpoint": None,
    "use_gpu": True,
    "render_mode": None,
}


ENV_ID_3D = "rbc_gym/RayleighBenardConvection3D-v0"
DEFAULT_KWARGS_3D = {                      # src/rbc_gym/__init__.py:24-37
    "rayleigh_number": 500,
    "prandtl_number": 0.7,
    "domain": [2, 4 * 3.141592653589793, 4 * 3.141592653589793],
    "state_shape": (16, 32, 32),
    "temperature_difference": [1, 2],
    "heater_segments": 8,
    "heater_limit": 0.9,
    "heater_duration": 0.125,
    "episode_length": 300,
    "checkpoint": None,
    "use_gpu": True,
    "render_mode": None,
}


def register_envs() -> bool:
    """Register the reference's gym ids against this backend (no-op without gymnasium)."""
    if not spaces.HAVE_GYMNASIUM:
        return False
    from gymnasium.envs.registration import register, registry
    if ENV_ID_2D not in registry:
        register(id=ENV_ID_2D, entry_point="rbc_gym_b200.envs:RayleighBenardConvection2DEnv", kwargs=dict(DEFAULT_KWARGS_2D))
    if ENV_ID_3D not in registry:
        register(id=ENV_ID_3D, entry_point="rbc_gym_b200.envs:RayleighBenardConvection3DEnv", kwargs=dict(DEFAULT_KWARGS_3D))
    return True


def make(env_id: str = ENV_ID_2D, **kwargs):
    """`gym.make` stand-in usable without gymnasium: builds the env with the registered default kwargs."""
    from .envs import RayleighBenardConvection2DEnv, RayleighBenardConvection3DEnv
    if env_id == ENV_ID_2D:
        cls, kw = RayleighBenardConvection2DEnv, dict(DEFAULT_KWARGS_2D)
    elif env_id == ENV_ID_3D:
        cls, kw = RayleighBenardConvection3DEnv, dict(DEFAULT_KWARGS_3D)
    else:
        raise ValueError(f"unknown environment id {env_id!r}")
    kw.update(kwargs)
    return cls(**kw)


register_envs()
