#!/usr/bin/env python
"""Golden-vector generator: the 3D CPU oracle run like the reference's `experiments/flowstats/flowstats_ra.py:27-36`
(64 x 64 x 32 grid, dt_solver 0.005, heater_duration 0.25 -> one sample per time unit, zero action, noise initialisation
with kick 0.01), so that its Nusselt series can be compared with the Julia-produced `flowstats_ra.pkl` at the SAME
resolution.  Test infrastructure (oracle only; no GPU).  One process per Rayleigh number, ~9 s of one core per sample.

    python tools/oracle3d_flowstats.py --ra 500 750 1000 1500 2000 4000 8000 16000 32000 64000 128000 256000 512000 1000000 \
        --samples 300 --seed 1000 --procs 7 --out tests/golden/oracle3d_flowstats_64x64x32.json
"""
import argparse
import json
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def run(args):
    ra, samples, seed, log = args
    from oracle import oracle3d as O3
    shape = (32, 64, 64)
    P = O3.make_params(float(ra), shape=shape)
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z = (np.arange(nz) + 0.5) * 2.0 / nz
    b = np.clip(1 + (2 - z)[:, None, None] / 2 + 0.01 * rng.standard_normal(shape), 1, 2)       # rbc_sim3D.jl:169-178
    u, v = 0.01 * rng.standard_normal(shape), 0.01 * rng.standard_normal(shape)
    w = 0.01 * rng.standard_normal((nz + 1, ny, nx))
    w[0] = 0
    w[-1] = 0
    u, v, w = O3.project(P, u, v, w)                                                          # Oceananigans set!
    dts = O3.substep_schedule(0.25, 0.005)
    zero = np.zeros((8, 8))
    nus, wmax = [], []
    t0 = time.time()
    for s in range(samples):
        r = O3.step(P, b, u, v, w, zero, dts)
        b, u, v, w = r["b"], r["u"], r["v"], r["w"]
        assert not r["nan"]
        nus.append(float(O3.nusselt(P, b, w)))
        wmax.append(float(np.abs(w).max()))
        if log:
            with open(log, "a") as f:
                f.write(f"ra={ra} sample={s + 1} nu={nus[-1]:.6f} wmax={wmax[-1]:.4f} elapsed={time.time() - t0:.0f}s\n")
    return str(ra), {"nusselt_step": nus, "uz_max_step": wmax, "seed": seed}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ra", type=float, nargs="+", default=[500, 4000, 16000])
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--seed", type=int, default=42, help="seed of the first Rayleigh number; the i-th one uses seed + i (one noise "
                    "realisation per run, like the reference, whose env draws a fresh seed per reset)")
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--out", default="tests/golden/oracle3d_flowstats_64x64x32.json")
    ap.add_argument("--log", default="")
    a = ap.parse_args()
    with mp.Pool(a.procs or len(a.ra)) as pool:
        res = dict(pool.map(run, [(ra, a.samples, a.seed + i, a.log) for i, ra in enumerate(a.ra)], chunksize=1))
    meta = {"grid": [32, 64, 64], "dt_solver": 0.005, "heater_duration": 0.25, "kick": 0.01, "samples": a.samples,
            "generator": "tools/oracle3d_flowstats.py", "note": "CPU oracle (oracle/rbc3d_oracle.c), one sample per time unit"}
    Path(a.out).write_text(json.dumps({"meta": meta, "runs": res}, indent=0))
    print("wrote", a.out)
