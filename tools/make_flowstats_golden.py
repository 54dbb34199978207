#!/usr/bin/env python
"""Golden-vector generator: the Julia-produced 3D flow statistics the reference ships
(`experiments/flowstats/flowstats_ra.pkl`, written by `experiments/flowstats/flowstats_ra.py:27-95`: uncontrolled 64 x 64 x 32
runs, `dt_solver` 0.005, `heater_duration` 0.25, one sample per time unit, 300 samples, 14 Rayleigh numbers) as plain JSON, so
that the 3D oracle and the GPU kernels can be compared with them where the reference mount does not exist.

The pickle comes from an untrusted mount: it is read with a restricted unpickler that only admits numpy array reconstruction.

    python tools/make_flowstats_golden.py        # needs /root/reference; writes tests/golden/flowstats_julia_64x64x32.json
"""
import importlib
import json
import pickle
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PKL = Path("/root/reference/experiments/flowstats/flowstats_ra.pkl")
OUT = ROOT / "tests/golden/flowstats_julia_64x64x32.json"
ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
           ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")}


class NumpyOnlyUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) not in ALLOWED:
            raise pickle.UnpicklingError(f"refusing to load {module}.{name} from an untrusted pickle")
        return getattr(importlib.import_module(module), name)


def load_flowstats(path=PKL):
    with open(path, "rb") as f:
        d = NumpyOnlyUnpickler(f).load()
    assert isinstance(d, dict)
    return {str(k): {kk: np.asarray(vv, dtype=np.float64) for kk, vv in v.items()} for k, v in d.items()}


if __name__ == "__main__":
    d = load_flowstats()
    runs = {ra: {k: [float(x) for x in v] for k, v in sorted(r.items())} for ra, r in sorted(d.items(), key=lambda kv: float(kv[0]))}
    meta = {"source": "experiments/flowstats/flowstats_ra.pkl (reference mount), produced by experiments/flowstats/flowstats_ra.py:27-95",
            "grid": [32, 64, 64], "dt_solver": 0.005, "heater_duration": 0.25, "samples": 300, "generator": "tools/make_flowstats_golden.py"}
    OUT.write_text(json.dumps({"meta": meta, "runs": runs}))
    print("wrote", OUT, OUT.stat().st_size, "bytes,", len(runs), "Rayleigh numbers")
