#!/usr/bin/env python
"""Golden-vector generator: (Nu_state, KE, Nu_obs) of all 40 Ra=1e5 states the reference ships
(`data/checkpoints/{train,val,test}/ckpt_ra100000.h5`, written by the real Julia simulation through
`rbc_sim2D.jl:14-72`), so that the limit-cycle pin of SURVEY.md section 4 (row "Ra=1e5 files") runs without the
reference mount.  The three numbers are plain array reductions of the stored fields (Nusselt formula of
`rbc_sim2D_api.jl:142-163`, KE = mean(u^2)/2 + mean(w^2)/2 over the stored arrays): no time stepping involved.

    python tools/make_ra1e5_triples.py            # needs /root/reference; writes tests/golden/ra1e5_checkpoint_triples.json
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from oracle import oracle as O                      # noqa: E402
from rbc_gym_b200.h5lite import load_checkpoint_2d  # noqa: E402

REF = Path("/root/reference/data/checkpoints")
OUT = Path(__file__).resolve().parent.parent / "tests/golden/ra1e5_checkpoint_triples.json"

if __name__ == "__main__":
    P = O.make_params(1e5)
    rows = []
    for split in ("train", "val", "test"):
        c = load_checkpoint_2d(REF / split / "ckpt_ra100000.h5")
        for ep in range(c.num_episodes):
            ns, no = O.nusselt_state_obs(P, c.b[ep], c.u[ep], c.w[ep])
            rows.append({"split": split, "ep": ep, "nu_state": ns, "ke": O.kinetic_energy(c.u[ep], c.w[ep]), "nu_obs": no})
    OUT.write_text(json.dumps({"source": "data/checkpoints/{train,val,test}/ckpt_ra100000.h5 (reference mount)",
                               "generator": "tools/make_ra1e5_triples.py", "states": rows}, indent=1))
    print("wrote", OUT, len(rows))
