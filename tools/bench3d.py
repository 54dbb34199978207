#!/usr/bin/env python
"""Throughput of the 3D environment step (BASELINE config 4 shape: 32x32x16, Ra=1e4, 8x8 heaters, dt_solver=0.01,
heater_duration=0.125 -> 13 RK3 steps).  Not the headline metric; reported in profiles/ for the 3D row of SURVEY §8."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1184)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--precision", type=int, default=32)
ap.add_argument("--shape", default="16x32x32", help="nz x ny x nx; anything but 16x32x32 runs the stage-streaming kernels (rbc3dg)")
ap.add_argument("--heater-duration", type=float, default=0.125)
ap.add_argument("--dt-solver", type=float, default=0.01)
args = ap.parse_args()
shape = tuple(int(v) for v in args.shape.split("x"))

import torch  # noqa: E402
from rbc_gym_b200 import backend  # noqa: E402
from rbc_gym_b200.envs import noise_initial_fields_3d  # noqa: E402

sim = backend.Sim3D(args.envs, ra=1e4, precision=args.precision, state_shape=shape, heater_duration=args.heater_duration, dt_solver=args.dt_solver)
rng = np.random.default_rng(0)
base = np.concatenate([noise_initial_fields_3d(rng, shape, kick=0.05) for _ in range(8)])
sim.reset_from_fields(base[np.arange(args.envs) % 8], project=True)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
acts = torch.rand((args.steps + 3, args.envs, 8, 8), device="cuda", generator=g) * 2 - 1
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0:
    sim.step(acts[0], want_obs=False); torch.cuda.synchronize()
out = {}
for want_obs in (False, True):
    for i in range(3):
        sim.step(acts[i], want_obs=want_obs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        sim.step(acts[3 + i], want_obs=want_obs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    out["with_obs" if want_obs else "no_obs"] = {"ms_per_step": ms, "env_steps_per_s": args.envs / (ms * 1e-3)}
S = sim.nstate
alg = sim.nsub * 10 * S * (args.precision // 8)          # SURVEY §8d generalisation: ceil(dt/dt_solver) x 10 x S x sizeof(real)
out["algorithmic_MB_per_env_step"] = alg / 1e6
out["streaming_equiv_GBps"] = out["no_obs"]["env_steps_per_s"] * alg / 1e9
out["nan"] = int(sim.nan.sum().item())
out["launch"] = sim.launch_info()
print(json.dumps({"metric": f"3D {shape[2]}x{shape[1]}x{shape[0]} Ra=1e4 env-steps/s ({sim.nsub} RK3 steps)", "envs": args.envs, "precision": args.precision, **out}))
