#!/usr/bin/env python
"""Golden vectors for the wrappers, produced by EXECUTING THE REFERENCE'S OWN CODE:
`/root/reference/src/rbc_gym/wrappers/{rbc_reward_shaping,rbc_normalize_reward,rbc_normalize_observation}.py` are
imported from the read-only mount (nothing is copied); the modules they import but that do not exist in this image
(gymnasium, matplotlib, the rbc_gym package itself) are replaced by minimal stubs.  Inputs are reproducible from the
shipped checkpoint file and seeded generators, so only the OUTPUTS are stored.

    python tools/make_wrapper_golden.py        # writes tests/golden/wrappers_reference.json   (needs /root/reference)
"""
import enum
import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/src/rbc_gym/wrappers")


def install_stubs():
    gym = types.ModuleType("gymnasium")

    class Env:
        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        @property
        def unwrapped(self):
            return self.env.unwrapped

    class ObservationWrapper(Wrapper):
        pass

    class RewardWrapper(Wrapper):
        pass

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    gym.Env, gym.Wrapper, gym.ObservationWrapper, gym.RewardWrapper = Env, Wrapper, ObservationWrapper, RewardWrapper
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = Box
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    pkg = types.ModuleType("rbc_gym")
    envs = types.ModuleType("rbc_gym.envs")
    rbc2d = types.ModuleType("rbc_gym.envs.rbc2D")

    class RBCField(enum.IntEnum):                      # src/rbc_gym/envs/rbc2D.py:16-20
        T = 0
        UX = 1
        UY = 2
        P = 3

    class RayleighBenardConvection2DEnv(Env):
        pass

    class RayleighBenardConvection3DEnv(Env):
        pass

    rbc2d.RBCField = RBCField
    envs.RayleighBenardConvection2DEnv, envs.RayleighBenardConvection3DEnv = RayleighBenardConvection2DEnv, RayleighBenardConvection3DEnv
    envs.rbc2D = rbc2d
    pkg.envs = envs
    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", gym.spaces), ("matplotlib", mpl), ("matplotlib.pyplot", mpl.pyplot),
                      ("rbc_gym", pkg), ("rbc_gym.envs", envs), ("rbc_gym.envs.rbc2D", rbc2d)):
        sys.modules[name] = mod
    return gym, envs


def load(name):
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", REF / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_inputs():
    """(name, state[3,64,96] float32) pairs — shared with tests/test_wrappers_reference_golden.py."""
    from rbc_gym_b200.h5lite import load_checkpoint_2d
    c = load_checkpoint_2d(ROOT / "data/checkpoints/train/ckpt_ra100000.h5")
    out = [(f"ckpt_ra1e5_ep{e}", np.stack([c.b[e], c.u[e], c.w[e, :-1]]).astype(np.float32)) for e in range(c.num_episodes)]
    rng = np.random.default_rng(2024)
    x = np.linspace(0, 2 * np.pi, 96, endpoint=False)
    for k in range(12):
        st = rng.standard_normal((3, 64, 96)).astype(np.float32) * 0.3
        st[2, 31] = (np.cos((k % 4 + 1) * x + rng.uniform(0, 6)) * rng.uniform(0.2, 1) + rng.uniform(-0.3, 0.6)
                     + 0.05 * rng.standard_normal(96)).astype(np.float32)
        out.append((f"synthetic_{k}", st))
    plateau = np.zeros((3, 64, 96), np.float32)
    plateau[2, 31, 10:14] = 0.5
    plateau[2, 31, 40:42] = 0.7
    plateau[2, 31, 70] = 0.0009                       # below the height threshold
    out.append(("plateaus", plateau))
    return out


def main():
    inputs = golden_inputs()                    # before the stubs: the real package must not see the fake gymnasium
    gym, envs = install_stubs()
    shaping = load("rbc_reward_shaping")
    nrew = load("rbc_normalize_reward")
    nobs = load("rbc_normalize_observation")

    class Fake2D(envs.RayleighBenardConvection2DEnv):
        state_shape, ra, heater_limit, temperature_difference = [64, 96], 100_000, 0.75, [1, 2]
        observation_space = gym.spaces.Box(-np.inf, np.inf, shape=(3, 8, 48))

    class Fake3D(envs.RayleighBenardConvection3DEnv):
        state_shape, ra, heater_limit, temperature_difference = (16, 32, 32), 2500, 0.9, [1, 2]
        observation_space = gym.spaces.Box(-np.inf, np.inf, shape=(4, 16, 32, 32))

    env2, env3 = Fake2D(), Fake3D()
    sh = shaping.RBCRewardShaping(env2, shaping_weight=0.1)
    apply_shaping = getattr(sh, "_RBCRewardShaping__apply_reward_shaping")
    rows = {}
    for name, st in inputs:
        cd = float(sh.compute_cell_distances(st))
        cd_avg = float(sh.compute_cell_distances(st, use_avg=True))
        rows[name] = {"cell_dist": cd, "cell_dist_avg": cd_avg, "shaped_reward_from_-5": float(apply_shaping(cd, -5.0))}
    r2, r3 = nrew.RBCNormalizeReward(env2), nrew.RBCNormalizeReward(env3)
    rng = np.random.default_rng(7)
    obs = rng.uniform(-1.5, 3.0, (3, 8, 48)).astype(np.float32)
    o_plain = nobs.RBCNormalizeObservation(env2, heater_limit=0.75).observation(obs.copy())
    o_clip = nobs.RBCNormalizeObservation(env2, heater_limit=0.75, maxval=2, clip=True).observation(obs.copy())
    obs3 = rng.uniform(-1.5, 3.0, (4, 4, 4, 4)).astype(np.float32)
    w3 = nobs.RBCNormalizeObservation(env3, heater_limit=0.9, u_limit=None)
    o3 = w3.observation(obs3.copy())
    out = {
        "generator": "tools/make_wrapper_golden.py (executes the reference's wrapper modules from /root/reference)",
        "cell_distance": rows,
        "normalize_reward": {"scale_2d_ra1e5": float(r2.scale), "scale_3d_ra2500": float(r3.scale),
                             "reward_2d_of_-5": float(r2.reward(-5.0)), "reward_3d_of_-1.8": float(r3.reward(-1.8))},
        "normalize_observation": {"obs_seed": 7, "plain_sum": float(o_plain.astype(np.float64).sum()),
                                  "plain_sample": [float(v) for v in o_plain[:, 3, 5]],
                                  "clip_maxval2_sample": [float(v) for v in o_clip[:, 3, 5]], "clip_min": float(o_clip.min()),
                                  "clip_max": float(o_clip.max()), "u_limit_3d_ra2500": float(w3.max_vals[1]),
                                  "obs3_sample": [float(v) for v in o3[:, 1, 2, 3]], "space_high_2d": float(1.3)},
    }
    path = ROOT / "tests/golden/wrappers_reference.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path, len(rows), "cell-distance cases")


if __name__ == "__main__":
    main()
