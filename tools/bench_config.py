#!/usr/bin/env python
"""bench.py-style JSON line (value, e2e, roofline, clocks) for the other BASELINE.json configurations on one GPU:
    python tools/bench_config.py --config 3     # 2D 192x128, Ra=1e6, dt_solver 0.015, dt=1 (67 RK3 steps), 1056 envs, noise init
    python tools/bench_config.py --config 4     # 3D 32x32x16, Ra=1e4, heater_duration 0.125 (13 RK3 steps), 1184 envs
Same timing rules as bench.py: 1 s clock pre-warm, W >= 3 warm-up steps, K steps between CUDA events, inputs larger than L2,
nvidia-smi clocks sampled inside the timed region, e2e through the host-buffer C-ABI entry point with pinned memory."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench as B0  # noqa: E402
from rbc_gym_b200 import backend  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, choices=[3, 4], required=True)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--envs", type=int, default=0)
a = ap.parse_args()
K, W = a.steps, max(a.warmup, 3)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
if a.config == 3:
    n = a.envs or 1056
    sim = backend.Sim2D(n, ra=1e6, dt_action=1.0, dt_solver=0.015, state_shape=(128, 192), precision=32)
    # developed flows: 32 states spun up for 40 time units in the fp64 validation mode (see checkpoints.developed_states_2d for
    # why not fp32 from the conduction state), replicated over the batch
    from rbc_gym_b200.checkpoints import developed_states_2d
    base = developed_states_2d(32, 1e6, (128, 192), dt_solver=0.015, spin_up=40)
    sim.reset_from_fields(np.tile(base, (n // 32 + 1, 1))[:n], project=False)
    acts = torch.rand((W + K, n, 12), device=dev, generator=gen) * 2 - 1
    step = lambda x: sim.step(x)
    S, nsub = sim.nstate, sim.nsub
    metric, kernel = "2D 192x128 Ra=1e6 env-steps/s (dt=1, dt_solver=0.015)", "rbc2dx_env_kernel<192x128,cl4,f32>"
    workload = (f"configs[2]: 2D Ra=1e6 192x128 dt=1 ({nsub} RK3 steps), {n} envs, reset from 32 developed states (noise initialisation + 40 time "
                "units of fp64 spin-up), U(-1,1) actions")
    out = sim.alloc_host_outputs(pinned=True)
    step_host = lambda x: sim.step_host(x, out)
    kernel_ms = lambda: float(np.mean(sim.step_kernel_ms_history(min(K, 64))))
else:
    n = a.envs or 1184
    sim = backend.Sim3D(n, ra=1e4, precision=32)
    sim.noise_reset(kick=0.05, generator=gen)
    acts = torch.rand((W + K, n, 8, 8), device=dev, generator=gen) * 2 - 1
    step = lambda x: sim.step(x)
    S, nsub = backend.NSTATE3, sim.nsub
    metric, kernel = "3D 32x32x16 Ra=1e4 env-steps/s (heater_duration=0.125)", "rbc3d_env_kernel<float,tiled>"
    workload = f"configs[3]: 3D Ra=1e4 32x32x16, 8x8 heater patches, {nsub} RK3 steps per action step, {n} envs, noise initialisation, U(-1,1) actions"
    pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
    out = {"obs": pin((n, 4, 16, 32, 32), torch.float32), "reward": pin((n,), torch.float32), "nusselt": pin((n,), torch.float64),
           "truncated": pin((n,), torch.int32), "nan": pin((n,), torch.int32)}
    step_host = lambda x: sim.step_host(x, out)
    ms_c = __import__("ctypes").c_float()
    def kernel_ms():
        sim._check(sim._L.rbc3d_last_step_kernel_ms(sim._h, __import__("ctypes").byref(ms_c)))
        return ms_c.value

sampler = B0.ClockSampler(0)
sampler.start()
t0 = time.perf_counter()
while time.perf_counter() - t0 < 1.0:
    step(acts[0]); torch.cuda.synchronize()
for i in range(W):
    step(acts[i])
torch.cuda.synchronize()
tr0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = sim.launch_info()["launches"]
e0.record()
for i in range(K):
    res = step(acts[W + i])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
clocks = sampler.stop(tr0, time.time())
launches = sim.launch_info()["launches"] - l0
kms = kernel_ms()
nan = int(res[-1].sum().item())
host_acts = [acts[W + i].cpu().pin_memory() for i in range(K)]
step_host(host_acts[0].numpy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(K):
    step_host(host_acts[i].numpy())
torch.cuda.synchronize()
e2e_ms = (time.perf_counter() - t0) * 1e3
per_env = nsub * 10 * S * 4
peak, src = B0.hbm_peak()
achieved = per_env * n / (kms * 1e-3) / 1e9
print(json.dumps({
    "metric": metric, "value": n * K / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": {"workload": workload, "envs_per_gpu": n, "l2": "inputs larger than L2 (%.0f MB of state)" % (n * S * 4 / 1e6)},
    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                 "peak_source": src, "kernel": kernel, "kernel_ms": kms, "algorithmic_bytes_per_launch": per_env * n},
    "e2e": {"value": n * K / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": int(host_acts[0].numel() * 4),
            "d2h_bytes_per_step": int(sum(v.nbytes for v in out.values()))},
    "gpu_launches": launches, "clocks": clocks, "nan_envs": nan}))
sim.close()
