#!/usr/bin/env python
"""Smallest invocation of every step kernel, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py`."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from rbc_gym_b200 import backend

g = torch.Generator(device="cuda").manual_seed(0)
for kw in (dict(), dict(state_shape=(128, 192), dt_solver=0.015), dict(state_shape=(64, 128), obs_shape=(8, 64))):
    for prec in (32, 64):
        sim = backend.Sim2D(5, ra=1e5, dt_action=0.06, precision=prec, **kw)
        sim.noise_reset(kick=0.05, generator=g)
        for _ in range(2):
            out = sim.step(torch.rand((5, 12), device="cuda", generator=g) * 2 - 1)
        sim.observe()
        torch.cuda.synchronize()
        print("2D", kw, prec, float(out[1].mean()), int(out[5].sum()))
        sim.close()
for prec, split in ((32, False), (32, True), (64, False)):
    sim = backend.Sim3D(3, ra=2500, heater_duration=0.02, precision=prec, split=split)
    sim.noise_reset(kick=0.05, generator=g)
    out = sim.step(torch.rand((3, 8, 8), device="cuda", generator=g) * 2 - 1)
    torch.cuda.synchronize()
    print("3D", prec, split, float(out[1].mean()), int(out[4].sum()))
    sim.close()
