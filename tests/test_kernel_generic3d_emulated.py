"""The stage-streaming 3D kernels for arbitrary grids (rbc_gym_b200/csrc/rbc3dg_core.h: what serves `state_shape` other than
(16, 32, 32), e.g. the (32, 64, 64) of the reference's flowstats runs) compiled for the host and checked against the 3D oracle."""
import numpy as np
import pytest

from oracle import oracle3d as O3
from tests.emu import emu


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


def random_state(P, seed, amp=0.2):
    rng = np.random.default_rng(seed)
    nz, ny, nx = P.nz, P.ny, P.nx
    z = (np.arange(nz) + 0.5) * P.lz / nz
    b = 1 + (P.lz - z)[:, None, None] / 2 + 0.05 * rng.standard_normal((nz, ny, nx))
    u, v = amp * rng.standard_normal((nz, ny, nx)), amp * rng.standard_normal((nz, ny, nx))
    w = amp * rng.standard_normal((nz + 1, ny, nx)); w[0] = 0; w[-1] = 0
    return (b, *O3.project(P, u, v, w))


def pack(b, u, v, w):
    return np.concatenate([x.reshape(1, -1) for x in (b, u, v, w)], axis=1)


def unpack(st, shape):
    nz, ny, nx = shape
    n = nz * ny * nx
    return st[:n].reshape(shape), st[n:2 * n].reshape(shape), st[2 * n:3 * n].reshape(shape), st[3 * n:].reshape(nz + 1, ny, nx)


@pytest.mark.parametrize("shape", [(16, 32, 32), (8, 16, 32), (12, 32, 16), (7, 8, 16), (32, 64, 64)])      # nz = 7: an unpaired top level in the spectral planes
def test_generic_3d_kernels_match_oracle(shape):
    full = shape == (32, 64, 64)
    P = O3.make_params(5e3, shape=shape, split_phy=False)
    b, u, v, w = random_state(P, 1)
    a = np.random.default_rng(2).uniform(-1, 1, (8, 8)).astype(np.float32)
    dts = O3.substep_schedule(0.02 if full else 0.05, 0.01)                 # 2 (5) RK3 steps, the last one clipped where it applies
    r = O3.step(P, b, u, v, w, a.astype(np.float64), dts)
    for prec, tol in ((64, 2e-13), (32, 3e-6)):
        e = emu.step3g(pack(b, u, v, w), a[None], 5e3, shape, precision=prec, heater_duration=0.02 if full else 0.05)
        eb, eu, ev, ew = unpack(e["state"][0].astype(np.float64), shape)
        assert rel(eb, r["b"]) < tol and rel(eu, r["u"]) < tol and rel(ev, r["v"]) < tol and rel(ew, r["w"]) < tol, prec
        assert np.all(ew[0] == 0) and np.all(ew[-1] == 0)
        assert e["nusselt"][0] == pytest.approx(O3.nusselt(P, r["b"], r["w"]), rel=1e-10 if prec == 64 else 2e-5) and e["nan"][0] == 0
        if prec == 64:
            assert np.abs(O3.divergence(P, eu, ev, ew)).max() < 1e-12


def test_generic_3d_projection_and_per_environment_rayleigh():
    shape = (8, 16, 16)
    P = O3.make_params(2500, shape=shape)
    rng = np.random.default_rng(3)
    u, v = 0.1 * rng.standard_normal(shape), 0.1 * rng.standard_normal(shape)
    w = 0.1 * rng.standard_normal((9, 16, 16)); w[0] = 0; w[-1] = 0
    b = 1.5 + 0.1 * rng.standard_normal(shape)
    e = emu.step3g(pack(b, u, v, w), np.zeros((1, 8, 8), np.float32), 2500, shape, project_first=True, nsub=0)
    eb, eu, ev, ew = unpack(e["state"][0], shape)
    ou, ov, ow = O3.project(P, u, v, w)                                     # what Oceananigans' set! does
    assert rel(eu, ou) < 1e-13 and rel(ev, ov) < 1e-13 and rel(ew, ow) < 1e-13 and np.array_equal(eb, b)
    # two environments with different Rayleigh numbers in one batch == two single runs
    st = np.concatenate([pack(b, ou, ov, ow)] * 2)
    a = rng.uniform(-1, 1, (2, 8, 8)).astype(np.float32)
    both = emu.step3g(st, a, 2500, shape, ra_env=[800.0, 30000.0], heater_duration=0.03)
    for j, ra in enumerate((800.0, 30000.0)):
        one = emu.step3g(st[j:j + 1], a[j:j + 1], ra, shape, heater_duration=0.03)
        assert np.array_equal(both["state"][j], one["state"][0]) and both["nusselt"][j] == one["nusselt"][0]
        r = O3.step(O3.make_params(ra, shape=shape, split_phy=False), b, ou, ov, ow, a[j].astype(np.float64), O3.substep_schedule(0.03, 0.01))
        assert rel(unpack(one["state"][0], shape)[0], r["b"]) < 1e-13
