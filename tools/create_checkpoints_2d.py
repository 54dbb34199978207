#!/usr/bin/env python
"""CLI twin of `scripts/create_checkpoints_2D.sh`: train/test/val checkpoint files for one Rayleigh number."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rbc_gym_b200.checkpoints import simulate_2d_rb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("ra", type=float)
ap.add_argument("--dir", default="data/checkpoints")
ap.add_argument("--duration", type=float, default=600.0)
ap.add_argument("--grid", default="64x96", help="NZxNX: 64x96 (reference) or 128x192 (config 3)")
ap.add_argument("--dt-solver", type=float, default=None, help="default 0.03, or 0.015 on the 128x192 grid")
args = ap.parse_args()
nz, nx = (int(v) for v in args.grid.split("x"))
dts = args.dt_solver if args.dt_solver is not None else (0.03 if nx <= 96 else 0.015)
for split, seed, n in (("train", 42, 20), ("test", 62, 10), ("val", 72, 10)):     # create_checkpoints_2D.sh:18-20
    path, stats = simulate_2d_rb(Path(args.dir) / split, seed=seed, random_inits=n, ra=args.ra, duration=args.duration, state_shape=(nz, nx),
                                 delta_t=dts)
    print(path, "Nu_state mean", stats["nu_state"].mean())
