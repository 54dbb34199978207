"""The 3D CUDA kernel source (rbc_gym_b200/csrc/rbc3d_core.h) compiled for the host and run as a sequential
one-CTA emulator against the 3D oracle (see test_kernel_core_emulated.py for the idea)."""
import numpy as np
import pytest

from oracle import oracle3d as O3
from tests.emu import emu
from tests.test_oracle3d import random_state


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


@pytest.mark.parametrize("split", [False, True, "tiled"])
def test_emulated_3d_kernel_matches_oracle(split):
    # "tiled": the tendency phase marches from a shared-memory tile per half of the domain (fp32 throughput path)
    P = O3.make_params(5e3, split_phy=(split is True))
    b, u, v, w = random_state(P, 1)
    a = np.random.default_rng(2).uniform(-1, 1, (8, 8)).astype(np.float32)
    r = O3.step(P, b, u, v, w, a.astype(np.float64), O3.substep_schedule())
    for prec, tol in ((64, 1e-13), (32, 2e-6)):
        e = emu.step3(emu.pack3(b[None], u[None], v[None], w[None]), a[None], 5e3, precision=prec, split=split)
        eb, eu, ev, ew = emu.unpack3(e["state"].astype(np.float64))
        assert rel(eb[0], r["b"]) < tol and rel(eu[0], r["u"]) < tol and rel(ev[0], r["v"]) < tol and rel(ew[0], r["w"]) < tol
        nu = O3.nusselt(P, r["b"], r["w"])
        assert e["nusselt"][0] == pytest.approx(nu, rel=1e-10 if prec == 64 else 1e-5)
        assert e["reward"][0] == pytest.approx(-nu, rel=1e-5)
        assert e["t"][0] == 0.5 and e["step"][0] == 2 and e["nan"][0] == 0 and e["truncated"][0] == 0
        if prec == 64:
            np.testing.assert_array_equal(e["obs"][0], np.stack([r["b"], r["u"], r["v"], r["w"][:-1]]).astype(np.float32))


def test_emulated_3d_projection_only_and_truncation():
    P = O3.make_params(2500)
    rng = np.random.default_rng(3)
    b = 1.5 + 0.1 * rng.standard_normal((16, 32, 32))
    u, v = 0.1 * rng.standard_normal((16, 32, 32)), 0.1 * rng.standard_normal((16, 32, 32))
    w = 0.1 * rng.standard_normal((17, 32, 32)); w[0] = 0; w[-1] = 0
    e = emu.step3(emu.pack3(b[None], u[None], v[None], w[None]), np.zeros((1, 8, 8), np.float32), 2500, precision=64,
                  project_first=True, nsub=0, t0=[299.6])
    eb, eu, ev, ew = emu.unpack3(e["state"])
    ou, ov, ow = O3.project(P, u, v, w)                                  # what Oceananigans' set! does
    assert rel(eu[0], ou) < 1e-13 and rel(ev[0], ov) < 1e-13 and rel(ew[0], ow) < 1e-13 and np.array_equal(eb[0], b)
    assert np.abs(O3.divergence(P, eu[0], ev[0], ew[0])).max() < 1e-13
    assert e["t"][0] == pytest.approx(300.1) and e["truncated"][0] == 1
