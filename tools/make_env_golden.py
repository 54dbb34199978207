#!/usr/bin/env python
"""Golden episode of the reference's own `RayleighBenardConvection2DEnv` (`/root/reference/src/rbc_gym/envs/rbc2D.py`,
executed from the read-only mount, nothing copied).  Everything above the Julia boundary is the reference's code: spaces,
reset/step, the `[::-1]` shape reversal, `np.array(..., float32)` + `.transpose(0, 2, 1)`, reward sign, `info` keys,
truncation.  Below the boundary the Julia module `RBCGymAPI` (which cannot run here) is replaced by `FakeJulia`, an object
with the functions of `src/rbc_gym/sim/rbc_sim2D_api.jl:17-163` (same names, keyword arguments and ARRAY LAYOUTS:
`get_state()` is `(5, Nx, Nz)`, `get_observation()` `(5, No_x, No_z)`) whose arithmetic is the CPU oracle.
Modules missing in this image (gymnasium, pygame, matplotlib, juliacall, juliapkg) are minimal stubs.

    python tools/make_env_golden.py     # writes tests/golden/env2d_reference_episode.npz (needs /root/reference)
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF_ENV = Path("/root/reference/src/rbc_gym/envs/rbc2D.py")
CKPT = ROOT / "data/checkpoints/train/ckpt_ra100000.h5"
EPISODE_IDX = 7          # the reference draws rand(1:n) with Julia's RNG; the fake pins the draw
N_STEPS = 4


def build_fake_julia():
    from oracle import oracle as O
    from rbc_gym_b200.h5lite import load_checkpoint_2d

    class FakeJulia:
        """The functions of rbc_sim2D_api.jl on top of the CPU oracle (hydrostatic split on, as in Oceananigans)."""

        def include(self, path):
            assert str(path).endswith("rbc_sim2D_api.jl")

        def initialize_simulation(self, Ra, sensors, grid, heaters, heater_limit, dt, seed, checkpoint_path, use_gpu):
            self.N, self.N_obs = list(grid), list(sensors)                      # Julia order: [Nx, Nz]
            self.P = O.make_params(float(Ra), nx=self.N[0], nz=self.N[1], heaters=heaters, heater_limit=heater_limit, split_phy=True)
            self.dt = float(dt)
            assert checkpoint_path is not None, "the golden episode starts from a checkpoint"
            c = load_checkpoint_2d(checkpoint_path)
            self.b, self.u, self.w = c.b[EPISODE_IDX].copy(), c.u[EPISODE_IDX].copy(), c.w[EPISODE_IDX].copy()
            # set! projects the velocities and leaves that projection's pressure in pNHS (rbc_sim2D.jl:173-186; a checkpoint
            # state is already solenoidal, so this is round-off); pHY' follows from b
            self.u, self.w, self.pnhs = O.project(self.P, self.u, self.w)
            self.phy = O.step(self.P, self.b, self.u, self.w, np.zeros(heaters), np.zeros(0), want_pressure=True)["phy"]
            self.time, self.step = 0.0, 1

        def step_simulation(self, actuation):
            a = np.asarray(actuation, dtype=np.float64)
            r = O.step(self.P, self.b, self.u, self.w, a, O.substep_schedule(self.dt), want_pressure=True)
            self.b, self.u, self.w, self.phy, self.pnhs = r["b"], r["u"], r["w"], r["phy"], r["pnhs"]
            self.time += self.dt
            self.step += 1
            return not r["nan"]

        def get_state(self):
            st = np.stack([self.b, self.u, self.w[:-1], self.phy, self.pnhs])     # (5, Nz, Nx) in C order
            return np.ascontiguousarray(st.transpose(0, 2, 1))                   # Julia's (5, Nx, Nz)

        def get_observation(self):
            sx, sz = self.N[0] // self.N_obs[0], self.N[1] // self.N_obs[1]
            return self.get_state()[:, ::sx, ::sz]

        def get_info(self):
            return (self.time, self.step)

        def get_nusselt(self, state=False):
            data = self.get_state() if state else self.get_observation()
            T, uy = data[0], data[2]                                             # (Nx, Nz)
            return O.nusselt(np.ascontiguousarray(T.T), np.ascontiguousarray(uy.T), self.P.kappa)

    return FakeJulia


def install_stubs(FakeJulia):
    gym = types.ModuleType("gymnasium")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
            self._rng = np.random.default_rng(0)

        def sample(self):
            return self._rng.uniform(-1, 1, self.shape).astype(self.dtype)

    class Env:
        def reset(self, seed=None, options=None):
            self.np_random_seed = 0 if seed is None else seed
            self.np_random = np.random.default_rng(seed)

        @property
        def unwrapped(self):
            return self

    gym.Env = Env
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = Box
    gym.logger = types.SimpleNamespace(warn=lambda *a, **k: None)
    jc = types.ModuleType("juliacall")
    jc.newmodule = lambda name: FakeJulia()
    jp = types.ModuleType("juliapkg")
    jp.resolve = lambda: None
    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", gym.spaces), ("pygame", types.ModuleType("pygame")),
                      ("matplotlib", types.ModuleType("matplotlib")), ("juliacall", jc), ("juliapkg", jp)):
        sys.modules[name] = mod


def main():
    FakeJulia = build_fake_julia()              # before the stubs (imports the real package)
    install_stubs(FakeJulia)
    spec = importlib.util.spec_from_file_location("_ref_rbc2D", REF_ENV)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for tag, kw in (("default", dict()), ("pressure_full", dict(pressure=True, observation_shape=[64, 96]))):
        env = ref.RayleighBenardConvection2DEnv(rayleigh_number=100_000, heater_duration=0.3, checkpoint=str(CKPT), episode_length=0.9, **kw)
        out[f"{tag}/obs_low"], out[f"{tag}/obs_high"] = env.observation_space.low, env.observation_space.high
        out[f"{tag}/episode_steps"] = np.array(env.episode_steps)
        obs, info = env.reset(seed=3)
        out[f"{tag}/reset_obs"] = obs if tag == "default" else obs[:, ::8, ::2]
        if tag == "default":
            out[f"{tag}/reset_state"] = info["state"]
        out[f"{tag}/reset_info"] = np.array([info["t"], info["step"], info["nusselt_state"], info["nusselt_obs"]])
        rng = np.random.default_rng(11)
        acts = rng.uniform(-1, 1, (N_STEPS, 12)).astype(np.float32)
        out[f"{tag}/actions"] = acts
        steps = N_STEPS if tag == "default" else 2
        for n in range(steps):
            obs, reward, terminated, truncated, info = env.step(acts[n])
            if tag == "default" or n == steps - 1:        # keep the fixture small: full-grid observations only once
                out[f"{tag}/obs{n}"] = obs
            if n == steps - 1:
                out[f"{tag}/state{n}"] = info["state"]
            out[f"{tag}/scalars{n}"] = np.array([reward, float(terminated), float(truncated), info["t"], info["step"],
                                                info["nusselt_state"], info["nusselt_obs"]])
    path = ROOT / "tests/golden/env2d_reference_episode.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
