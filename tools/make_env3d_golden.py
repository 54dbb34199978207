#!/usr/bin/env python
"""Golden episode of the reference's own `RayleighBenardConvection3DEnv` (`/root/reference/src/rbc_gym/envs/rbc3D.py`,
executed from the mount), with the Julia module replaced by an oracle-backed object that has the API and array layouts
of `src/rbc_gym/sim/rbc_sim3D_api.jl` (`get_state()` is `(4, Nx, Ny, Nz)`, the 8 x 8 action arrives un-transposed,
time advances by `dt * t_ff`).  See tools/make_env_golden.py for the 2D twin and the idea.

    python tools/make_env3d_golden.py   # writes tests/golden/env3d_reference_episode.npz (needs /root/reference)
"""
import importlib.util
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_ENV = Path("/root/reference/src/rbc_gym/envs/rbc3D.py")
RA, N_STEPS = 2500, 3


def initial_bank():
    """Two deterministic episodes (seeded numpy + oracle projection); the test re-creates the same file."""
    from oracle import oracle3d as O3
    from tests.test_oracle3d import random_state
    P = O3.make_params(float(RA))
    eps = [random_state(P, s, amp=0.2) for s in (21, 22)]
    return [np.stack([e[q] for e in eps]) for q in range(4)]


def build_fake_julia():
    from oracle import oracle3d as O3
    from rbc_gym_b200.h5lite import load_checkpoint_3d

    class FakeJulia3D:
        def include(self, path):
            assert str(path).endswith("rbc_sim3D_api.jl")

        def initialize_simulation(self, Ra, Pr, L, grid, T_diff, heaters, heater_limit, dt, dt_solver, seed, checkpoint_path, checkpoint_idx,
                                  use_gpu):
            nx, ny, nz = grid                                                     # Julia order
            lx, ly, lz = L
            self.P = O3.make_params(float(Ra), shape=(nz, ny, nx), domain=(lz, ly, lx), pr=Pr, heaters=heaters, heater_limit=heater_limit,
                                    t_diff=tuple(T_diff))
            self.t_ff = lz ** 2                                                   # rbc_sim3D_api.jl:43
            self.dt, self.dts = dt, O3.substep_schedule(dt, dt_solver, lz)
            # Julia indexes the file 1-based: read(h5, "b")[idx, :, :, :] (rbc_sim3D.jl:191)
            bank = load_checkpoint_3d(checkpoint_path)[checkpoint_idx - 1]        # the golden episode starts from a checkpoint
            nc = nx * ny * nz
            self.b, self.u, self.v = (bank[q * nc:(q + 1) * nc].reshape(nz, ny, nx).copy() for q in range(3))
            self.w = bank[3 * nc:].reshape(nz + 1, ny, nx).copy()
            self.time, self.step, self.actuators = 0.0, 1, (heaters, heaters)

        def step_simulation(self, actuation):
            a = np.asarray(actuation, dtype=np.float64)
            if a.shape != self.actuators:
                raise RuntimeError(f"Action size does not match the number of actuators. Expected {self.actuators}, got {a.shape}.")
            r = O3.step(self.P, self.b, self.u, self.v, self.w, a, self.dts)        # a[i, j] <-> patch i along x, j along y
            self.b, self.u, self.v, self.w = r["b"], r["u"], r["v"], r["w"]
            self.time += self.dt * self.t_ff
            self.step += 1
            return not r["nan"]

        def get_state(self):
            st = np.stack([self.b, self.u, self.v, self.w[:-1]])                  # (4, Nz, Ny, Nx) in C order
            return np.ascontiguousarray(st.transpose(0, 3, 2, 1))                 # Julia's (4, Nx, Ny, Nz)

        def get_info(self):
            return (self.time, self.step)

        def get_nusselt(self):
            return O3.nusselt(self.P, self.b, self.w)

    return FakeJulia3D


def install_stubs(Fake):
    gym = types.ModuleType("gymnasium")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

        def sample(self):
            return np.zeros(self.shape, self.dtype)

    class Env:
        def reset(self, seed=None, options=None):
            self.np_random_seed = 0 if seed is None else seed

        @property
        def unwrapped(self):
            return self

    gym.Env = Env
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = Box
    jc = types.ModuleType("juliacall")
    jc.newmodule = lambda name: Fake()
    jp = types.ModuleType("juliapkg")
    jp.resolve = lambda: None
    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", gym.spaces), ("matplotlib", types.ModuleType("matplotlib")),
                      ("juliacall", jc), ("juliapkg", jp)):
        sys.modules[name] = mod


def main():
    from rbc_gym_b200.h5lite import write_checkpoint_3d
    bank = initial_bank()
    Fake = build_fake_julia()
    install_stubs(Fake)
    spec = importlib.util.spec_from_file_location("_ref_rbc3D", REF_ENV)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / f"3D_ckpt_ra{RA}.h5"
        write_checkpoint_3d(path, *bank, start_seed=42)
        env = ref.RayleighBenardConvection3DEnv(rayleigh_number=RA, checkpoint=str(path), checkpoint_idx=1, episode_length=1.0,
                                                heater_duration=0.125, log_dir=d)
        out["obs_low_sample"], out["obs_high_sample"] = env.observation_space.low[:, ::5, ::11, ::7], env.observation_space.high[:, ::5, ::11, ::7]
        out["action_shape"] = np.array(env.action_space.shape)
        obs, info = env.reset(seed=5)
        out["reset_obs_sample"] = obs[:, ::2, ::4, ::4]
        out["reset_info"] = np.array([info["t"], info["step"], info["nusselt"]])
        assert set(info) == {"t", "step", "nusselt"}
        rng = np.random.default_rng(17)
        acts = rng.uniform(-1, 1, (N_STEPS, 8, 8)).astype(np.float32)
        acts[1, 2:5, :] = 1.0                                                    # asymmetric in (i, j): catches a transposed action
        out["actions"] = acts
        for n in range(N_STEPS):
            obs, reward, terminated, truncated, info = env.step(acts[n])
            out[f"obs_sample{n}"] = obs[:, ::2, ::4, ::4]
            out[f"obs_sum{n}"] = obs.astype(np.float64).sum(axis=(1, 2, 3))
            out[f"scalars{n}"] = np.array([reward, float(terminated), float(truncated), info["t"], info["step"], info["nusselt"]])
        out["obs_last_bottom_level"] = obs[0, 0]                                 # temperature next to the heater patches
    path = ROOT / "tests/golden/env3d_reference_episode.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
