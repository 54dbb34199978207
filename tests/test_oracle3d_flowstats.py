"""The 3D oracle against the ONLY Julia-produced 3D data the reference ships: `experiments/flowstats/flowstats_ra.pkl`
(uncontrolled 64 x 64 x 32 runs of `experiments/flowstats/flowstats_ra.py:27-36`, one Nusselt sample per time unit, 300 samples,
14 Rayleigh numbers; extracted to `tests/golden/flowstats_julia_64x64x32.json` by `tools/make_flowstats_golden.py` with a
numpy-only unpickler).  The oracle was run offline with the same protocol at the same resolution for all 14 Rayleigh numbers
x 300 samples, one noise realisation per Rayleigh number like the reference (`tools/oracle3d_flowstats.py` ->
`tests/golden/oracle3d_flowstats_64x64x32.json`, ~10 core-hours).

What one realisation against one realisation can and cannot pin (DESIGN.md section 2 has the ensemble study):
  * saturated Nusselt number, Ra >= 8000 (unsteady convection): mean over the last 150 samples within ONE standard deviation
    of the Julia series over those samples (measured: 0.08 ... 0.75 of it, 0.2 ... 1.9 % of Nu, no sign pattern);
  * Ra = 4000: within two; Ra <= 2000: the flow freezes into a steady planform selected by the noise — the Julia table is not
    even monotonic there (Nu(750) > Nu(1000)) — so only a quarter of Nu - 1 is asserted;
  * linear-instability plateau (largest local growth rate of Nu - 1 per time unit): seed scatter is 0.4 % (24-seed GPU
    ensemble); the oracle is within 1.2 % of Julia for Ra <= 4000 and 1.4 ... 2.9 % ABOVE it for Ra >= 8000 — a systematic
    difference in the transient growth that is documented, not hidden: the test asserts the measured envelope."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
PKL = Path("/root/reference/experiments/flowstats/flowstats_ra.pkl")
RAS = ["500", "750", "1000", "1500", "2000", "4000", "8000", "16000", "32000", "64000", "128000", "256000", "512000", "1000000"]


def plateau(nu):
    e = np.log(np.maximum(np.asarray(nu) - 1.0, 1e-12))
    return float(np.diff(e)[3:60].max())


@pytest.fixture(scope="module")
def julia():
    return json.loads((GOLD / "flowstats_julia_64x64x32.json").read_text())["runs"]


@pytest.fixture(scope="module")
def oracle():
    d = json.loads((GOLD / "oracle3d_flowstats_64x64x32.json").read_text())
    assert d["meta"]["grid"] == [32, 64, 64] and d["meta"]["samples"] == 300
    return {str(int(float(k))): v for k, v in d["runs"].items()}


def test_golden_julia_series_match_the_pickle_when_mounted(julia):
    if not PKL.exists():
        pytest.skip("reference mount not present")
    import sys
    sys.path.insert(0, str(GOLD.parent.parent / "tools"))
    from make_flowstats_golden import NumpyOnlyUnpickler, load_flowstats      # restricted: the mount is untrusted content
    d = load_flowstats(PKL)
    assert sorted(d, key=float) == RAS
    for ra in RAS:
        assert np.array_equal(d[ra]["nusselt_step"], np.array(julia[ra]["nusselt_step"])) and len(julia[ra]["nusselt_step"]) == 300
    import io
    import pickle
    with pytest.raises(pickle.UnpicklingError):                                 # anything but numpy arrays is refused
        NumpyOnlyUnpickler(io.BytesIO(pickle.dumps(Path("/")))).load()


def test_baseline_table_numbers(julia):
    # mean +- std over the last 150 samples
    for ra, m, s in (("500", 1.371, 0.006), ("4000", 2.123, 0.032), ("16000", 2.851, 0.072), ("1000000", 9.230, 0.222)):
        nu = np.array(julia[ra]["nusselt_step"])
        assert nu[150:].mean() == pytest.approx(m, abs=1e-3) and nu[150:].std() == pytest.approx(s, abs=1e-3)


@pytest.mark.parametrize("ra", RAS)
def test_oracle_reproduces_julia_flow_statistics_at_64x64x32(julia, oracle, ra):
    o, j = np.array(oracle[ra]["nusselt_step"]), np.array(julia[ra]["nusselt_step"])
    assert len(o) == 300 and np.isfinite(o).all()
    m_o, m_j, s_j = o[150:].mean(), j[150:].mean(), j[150:].std()
    r = float(ra)
    tol = 0.25 * (m_j - 1) if r <= 2000 else (2 * s_j if r < 8000 else s_j)
    assert abs(m_o - m_j) < tol, (ra, m_o, m_j, tol)
    if r >= 8000:
        assert o[150:].std() == pytest.approx(s_j, rel=0.5)                    # the fluctuation level, too
    ratio = plateau(o) / plateau(j)
    lo, hi = (0.985, 1.012) if r <= 4000 else (1.005, 1.035)
    assert lo < ratio < hi, (ra, ratio)
    # the first sample (t = 1) measures the noise level after the set! projection and one time unit: same kick, same decay
    assert o[0] - 1 == pytest.approx(j[0] - 1, rel=0.10)                     # measured 0.97 ... 1.09 (one realisation each)


def test_no_systematic_bias_in_the_saturated_regime(julia, oracle):
    rel = [np.mean(oracle[ra]["nusselt_step"][150:]) / np.mean(julia[ra]["nusselt_step"][150:]) - 1 for ra in RAS if float(ra) >= 8000]
    assert abs(np.mean(rel)) < 0.006 and max(np.abs(rel)) < 0.02, rel          # measured: mean +0.4 %, largest 1.9 %
