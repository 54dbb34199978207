#!/usr/bin/env python
"""Seed ensemble of the linear-instability phase of the reference's flowstats protocol (`experiments/flowstats/flowstats_ra.py:27-36`)
on the GPU: (32, 64, 64), dt_solver 0.005, heater_duration 0.25, zero action, noise kick 0.01, fp64; `--seeds` noise realisations
per Rayleigh number as ONE batch.  The largest local growth rate of Nu - 1 ("plateau") depends on the noise realisation — the
series saturates before the fastest mode dominates — so one Julia run per Rayleigh number has to be placed inside this
distribution, not compared with a single run.  Writes gpurun_out/growth_ensemble.json.

    python tools/gpu_growth_ensemble.py --ra 500 4000 16000 64000 256000 1000000 --seeds 24
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ra", type=float, nargs="+", default=[500, 4000, 8000, 16000, 64000, 256000, 1000000])
    ap.add_argument("--seeds", type=int, default=24)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out/growth_ensemble.json"))
    a = ap.parse_args()
    import torch
    from rbc_gym_b200 import backend
    res = {}
    for ra in a.ra:
        samples = 45 if ra < 700 else (30 if ra < 3000 else 22 if ra < 10000 else 16)
        sim = backend.Sim3D(a.seeds, ra=ra, state_shape=(32, 64, 64), heater_duration=0.25, dt_solver=0.005, episode_length=1e9, precision=64)
        sim.noise_reset(kick=0.01, generator=torch.Generator(device="cuda").manual_seed(int(ra) + 7))
        act = torch.zeros((a.seeds, 8, 8), device="cuda")
        nus = []
        for s in range(samples):
            _, _, nu, _, nan = sim.step(act, want_obs=False)
            nus.append(nu.clone())
        nus = torch.stack(nus, 1).cpu().numpy()
        assert int(nan.sum()) == 0
        slopes = np.diff(np.log(np.maximum(nus - 1.0, 1e-14)), axis=1)
        plateau = slopes[:, 2:].max(axis=1)
        res[str(int(ra))] = {"nusselt_step": nus.tolist(), "plateau": plateau.tolist()}
        print(f"Ra={ra:g}: plateau mean {plateau.mean():.4f} std {plateau.std():.4f} min {plateau.min():.4f} max {plateau.max():.4f}", flush=True)
        sim.close()
    Path(a.out).parent.mkdir(exist_ok=True)
    Path(a.out).write_text(json.dumps({"seeds": a.seeds, "runs": res}))
