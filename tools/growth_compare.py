#!/usr/bin/env python
"""Transient growth of Nu - 1 in the flowstats protocol: a GPU seed ensemble (tools/gpu_growth_ensemble.py) against the single
Julia run per Rayleigh number (tests/golden/flowstats_julia_64x64x32.json), compared AT EQUAL AMPLITUDE: the local slope of
log(Nu - 1) per time unit, interpolated at fixed levels of Nu - 1 on the rising flank.  Prints, per Rayleigh number and level,
the Julia slope, the ensemble mean +- standard deviation and the Julia run's distance from the ensemble in standard deviations.
Usage: python tools/growth_compare.py gpurun_out/growth_ensemble.json [label]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LEVELS = [1e-3, 3e-3, 1e-2, 3e-2, 1e-1]


def slope_at_levels(nu):
    e = np.log(np.maximum(np.asarray(nu) - 1.0, 1e-14))
    n = int(np.argmax(e[:80]))
    lev, d = 0.5 * (e[1:] + e[:-1])[:n], np.diff(e)[:n]
    if len(lev) < 3:
        return [np.nan] * len(LEVELS)
    lo = int(np.argmin(lev))
    out = []
    for L in np.log(LEVELS):
        out.append(float(np.interp(L, lev[lo:], d[lo:])) if lev[lo:].min() < L < lev[lo:].max() else np.nan)
    return out


if __name__ == "__main__":
    ens = json.loads(Path(sys.argv[1]).read_text())["runs"]
    label = sys.argv[2] if len(sys.argv) > 2 else ""
    julia = json.loads((ROOT / "tests/golden/flowstats_julia_64x64x32.json").read_text())["runs"]
    print(f"# {label}  levels of Nu-1: {LEVELS};  per level: julia | ensemble mean +- std | (julia - mean)/std")
    zs = []
    for ra, run in ens.items():
        sj = slope_at_levels(julia[ra]["nusselt_step"])
        se = np.array([slope_at_levels(s) for s in run["nusselt_step"]])
        cells = []
        for i in range(len(LEVELS)):
            col = se[:, i][np.isfinite(se[:, i])]
            if np.isfinite(sj[i]) and len(col) >= 3:
                z = (sj[i] - col.mean()) / col.std()
                zs.append((float(ra), LEVELS[i], z, sj[i] / col.mean()))
                cells.append(f"{sj[i]:.4f}|{col.mean():.4f}+-{col.std():.4f}|{z:+5.1f}")
            else:
                cells.append("      -      ")
        print(f"Ra={int(float(ra)):>8d}  " + "   ".join(cells))
    hi = [r for r in zs if r[0] >= 8000]
    if hi:
        print(f"# Ra >= 8000: mean ratio julia/ensemble {np.mean([r[3] for r in hi]):.4f}, mean z {np.mean([r[2] for r in hi]):+.2f}, rms z {np.sqrt(np.mean([r[2] ** 2 for r in hi])):.2f}")
    lo = [r for r in zs if r[0] < 8000]
    if lo:
        print(f"# Ra <  8000: mean ratio julia/ensemble {np.mean([r[3] for r in lo]):.4f}, mean z {np.mean([r[2] for r in lo]):+.2f}, rms z {np.sqrt(np.mean([r[2] ** 2 for r in lo])):.2f}")
