"""Throughput of the cluster kernel on a registered grid (config 3: 192 x 128, Ra=1e6, dt_solver=0.015).
Usage: python tools/bench_grid.py [--envs N] [--steps K] [--dt 1.0] [--grid 128x192] [--cluster 0|1]"""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1056)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--dt", type=float, default=1.0)
ap.add_argument("--grid", default="128x192")
ap.add_argument("--ra", type=float, default=1e6)
ap.add_argument("--dt-solver", type=float, default=0.015)
ap.add_argument("--precision", type=int, default=32)
ap.add_argument("--cluster", type=int, default=0, help="1: force the cluster kernel on the 96 x 64 grid")
ap.add_argument("--cfl", type=float, default=None, help="CFL guard limit of the cluster kernels (0 = off; default: the library's, 1.4 in fp32)")
a = ap.parse_args()
if a.cluster:
    os.environ["RBC_B200_CLUSTER"] = "1"
from rbc_gym_b200 import backend  # noqa: E402
from rbc_gym_b200.envs.rbc2d import noise_initial_fields  # noqa: E402

nz, nx = (int(v) for v in a.grid.split("x"))
sim = backend.Sim2D(a.envs, ra=a.ra, dt_action=a.dt, dt_solver=a.dt_solver, state_shape=(nz, nx), precision=a.precision)
if a.cfl is not None:
    sim.set_cfl_guard(a.cfl)
rng = np.random.default_rng(42)
base = np.concatenate([noise_initial_fields(rng, (nz, nx), kick=0.01) for _ in range(16)])
sim.reset_from_fields(np.tile(base, (a.envs // 16 + 1, 1))[: a.envs], project=True)
g = torch.Generator(device="cuda").manual_seed(1)
acts = [torch.rand((a.envs, 12), device="cuda", generator=g) * 2 - 1 for _ in range(a.steps + a.warmup)]
for i in range(a.warmup):
    sim.step(acts[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    out = sim.step(acts[a.warmup + i])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
nsub = sim.nsub
S = sim.nstate
alg = nsub * 10 * S * (4 if a.precision == 32 else 8)
sps = a.envs / ms * 1e3
print(json.dumps({"metric": f"2D {nx}x{nz} Ra={a.ra:g} env-steps/s (dt={a.dt:g}, {nsub} RK3 steps)", "envs": a.envs, "precision": a.precision,
                  "ms_per_step": ms, "env_steps_per_s": sps, "kernel_ms": sim.last_step_kernel_ms(),
                  "algorithmic_MB_per_env_step": alg / 1e6, "streaming_equiv_GBps": sps * alg / 1e9,
                  "nan": int(out[5].sum().item()), "launch": sim.launch_info(), "cfl_guard": a.cfl, "cfl_extra_rk3_steps": int(sim.cfl_events().sum())}))
