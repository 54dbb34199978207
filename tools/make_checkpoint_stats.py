#!/usr/bin/env python
"""Golden-vector generator: ensemble statistics of ALL 2D checkpoint states the reference ships
(`data/checkpoints/{train,val,test}/ckpt_ra*.h5`: 7 Rayleigh numbers x 40 end states of independent uncontrolled runs of the real
Julia simulation from noise, t = 600, `scripts/create_checkpoints_2D.sh:18-20`), so that long-horizon statistics of the CUDA
kernels can be compared with them where the mount does not exist.  Per state: Nu_state, Nu_obs (`rbc_sim2D_api.jl:142-163`),
kinetic energy, max|u|, max|w| — plain reductions of the stored fields.

    python tools/make_checkpoint_stats.py        # needs /root/reference; writes tests/golden/checkpoint_ensemble_stats.json
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from oracle import oracle as O                      # noqa: E402
from rbc_gym_b200.h5lite import load_checkpoint_2d  # noqa: E402

REF = Path("/root/reference/data/checkpoints")
OUT = Path(__file__).resolve().parent.parent / "tests/golden/checkpoint_ensemble_stats.json"

if __name__ == "__main__":
    out = {}
    for ra in (10000, 30000, 100000, 300000, 1000000, 3000000, 10000000):
        P = O.make_params(float(ra))
        rows = []
        for split in ("train", "val", "test"):
            c = load_checkpoint_2d(REF / split / f"ckpt_ra{ra}.h5")
            for ep in range(c.num_episodes):
                ns, no = O.nusselt_state_obs(P, c.b[ep], c.u[ep], c.w[ep])
                rows.append({"split": split, "ep": ep, "nu_state": ns, "nu_obs": no, "ke": O.kinetic_energy(c.u[ep], c.w[ep]),
                             "max_u": float(np.abs(c.u[ep]).max()), "max_w": float(np.abs(c.w[ep]).max())})
        out[str(ra)] = rows
        a = np.array([r["nu_state"] for r in rows])
        print(ra, len(rows), "Nu_state %.3f +- %.3f" % (a.mean(), a.std()))
    OUT.write_text(json.dumps({"source": "data/checkpoints/{train,val,test}/ckpt_ra*.h5 (reference mount)",
                               "generator": "tools/make_checkpoint_stats.py", "states": out}))
    print("wrote", OUT)
