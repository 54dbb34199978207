"""The cluster kernel for generic 2D grids (rbc_gym_b200/csrc/rbc2dx_core.h) compiled for the host and run
as a sequential emulator of a thread-block cluster (CTAs one after the other inside every phase, remote
shared memory = another slice of one arena), checked against the fp64 oracle.  Exercises the slab
decomposition, halo pushes, the SPIKE-partitioned tridiagonal solve, the N1 x N2 FFT and the cross-CTA
reductions; the `-m gpu` tests repeat the comparison on the device (tests/test_gpu_grid192.py)."""
import numpy as np
import pytest

from oracle import oracle as O
from rbc_gym_b200 import wrappers as W
from tests.emu import emu
from tests.gridstates import smooth_state


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


COS = np.cos(2 * np.pi * np.arange(12) / 12).astype(np.float32)


@pytest.mark.parametrize("cl", [1, 2, 4])
@pytest.mark.parametrize("nxt_global", [False, True])
def test_96x64_any_cluster_size_matches_oracle(ckpt_ra1e5, cl, nxt_global):
    c = ckpt_ra1e5
    eps, acts = [0, 7], np.stack([COS, np.linspace(-1, 1, 12).astype(np.float32)])
    e = emu.stepx(emu.pack(c.b[eps], c.u[eps], c.w[eps]), acts, 1e5, 0.09, cl=cl, precision=64, nxt_global=nxt_global)
    b, u, w = emu.unpackx(e["state"], 96, 64)
    P = O.make_params(1e5, split_phy=False)
    for j, ep in enumerate(eps):
        r = O.step(P, c.b[ep], c.u[ep], c.w[ep], acts[j].astype(np.float64), O.substep_schedule(0.09))
        assert rel(b[j], r["b"]) < 1e-13 and rel(u[j], r["u"]) < 1e-13 and rel(w[j], r["w"]) < 1e-13
        ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"])
        assert e["nu_state"][j] == pytest.approx(ns, abs=1e-10) and e["nu_obs"][j] == pytest.approx(no, abs=1e-10)
        np.testing.assert_array_equal(e["obs"][j], O.observe(O.state_channels(r["b"], r["u"], r["w"])).astype(np.float32))
    assert np.all(e["nan"] == 0) and np.all(e["t"] == 0.09) and np.all(e["step"] == 2)


@pytest.mark.parametrize("nx,nz,cl,obs", [(192, 128, 4, (8, 48)), (192, 128, 8, (16, 96)), (128, 64, 2, (8, 64)), (64, 64, 1, (8, 32)),
                                          (96, 128, 4, (8, 48))])
def test_other_grids_match_oracle(nx, nz, cl, obs):
    """config 3's grid (192 x 128, Ra=1e6, dt_solver=0.015: the heater's cubic blends fire) and a power-of-two
    width (4 x 16 FFT split), fp64 to round-off and fp32 within the stated tolerance."""
    ra, dts = 1e6, 0.015
    P = O.make_params(ra, nx=nx, nz=nz, split_phy=False)
    b, u, w = smooth_state(nx, nz)
    act = np.random.default_rng(1).uniform(-1, 1, 12).astype(np.float32)
    if nx == 192:
        Tb = O.heater_profile(P, act)
        assert len(np.unique(np.round(Tb, 12))) > 12           # not piecewise constant: blends active
    dt = 3 * dts + dts / 3
    r = O.step(P, b, u, w, act.astype(np.float64), O.substep_schedule(dt, dts))
    ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"], obs)
    st = emu.pack(b[None], u[None], w[None])
    for precision, nxt_global, tol in ((64, False, 1e-13), (64, True, 1e-13), (32, False, 2e-6)):
        e = emu.stepx(st, act[None], ra, dt, nx=nx, nz=nz, cl=cl, precision=precision, nxt_global=nxt_global, dt_solver=dts, obs=obs)
        bb, uu, ww = emu.unpackx(e["state"].astype(np.float64), nx, nz)
        assert rel(bb[0], r["b"]) < tol and rel(uu[0], r["u"]) < tol and rel(ww[0], r["w"]) < tol
        assert e["nu_state"][0] == pytest.approx(ns, abs=1e-9 if precision == 64 else 1e-3)
        assert e["nu_obs"][0] == pytest.approx(no, abs=1e-9 if precision == 64 else 1e-3)
        ref_obs = O.observe(O.state_channels(r["b"], r["u"], r["w"]), obs).astype(np.float32)
        np.testing.assert_allclose(e["obs"][0], ref_obs, rtol=0, atol=0 if precision == 64 else 2e-6)
        assert e["nan"][0] == 0 and e["step"][0] == 2


def test_set_projection_of_noise_init_192():
    """reset-from-noise path: project_first (Oceananigans set!) through the SPIKE solve == the oracle projection."""
    nx, nz = 192, 128
    rng = np.random.default_rng(42)
    u = 0.01 * rng.standard_normal((nz, nx))
    w = 0.01 * rng.standard_normal((nz + 1, nx))
    w[0] = 0
    w[-1] = 0
    b = np.ones((nz, nx)) * 1.5
    P = O.make_params(1e6, nx=nx, nz=nz, split_phy=False)
    up, wp, _ = O.project(P, u, w)
    e = emu.stepx(emu.pack(b[None], u[None], w[None]), np.zeros((1, 12), np.float32), 1e6, 0.015, nx=nx, nz=nz, cl=4, precision=64,
                  dt_solver=0.015, project_first=True, nsub=0)
    _, uu, ww = emu.unpackx(e["state"], nx, nz)
    assert rel(uu[0], up) < 1e-12 and rel(ww[0], wp) < 1e-12
    div = (np.roll(uu[0], -1, axis=-1) - uu[0]) / (2 * np.pi / nx) + (ww[0, 1:] - ww[0, :-1]) / (2 / nz)
    assert np.abs(div).max() < 1e-12


def test_cluster_flags_and_fused_wrappers_192():
    nx, nz, ra, dts = 192, 128, 1e6, 0.015
    b, u, w = smooth_state(nx, nz, seed=3)
    st = np.repeat(emu.pack(b[None], u[None], w[None]), 2, axis=0)
    st[1, nx * 100 + 17] = np.nan                          # a NaN in the slab of cluster rank 3
    acts = np.random.default_rng(5).uniform(-1, 1, (2, 12)).astype(np.float32)
    e = emu.stepx(st, acts, ra, 2 * dts, nx=nx, nz=nz, cl=4, precision=64, dt_solver=dts, t0=[299.98, 0.0])
    assert list(e["nan"]) == [0, 1] and list(e["truncated"]) == [1, 0]
    wr = emu.HostWrappers()
    wr.normalize_obs, wr.obs_clip, wr.obs_maxval = 1, 0, 1.0
    for ch, (lo, hi) in enumerate([(1.0, 2.75), (-1.3, 1.3), (-1.3, 1.3), (-1.3, 1.3)]):
        wr.obs_lo[ch], wr.obs_hi[ch] = lo, hi
    wr.normalize_reward, wr.reward_scale = 1, 0.1 * ra ** 0.4
    wr.shaping, wr.shaping_weight = 1, 0.1
    plain = emu.stepx(st[:1], acts[:1], ra, 2 * dts, nx=nx, nz=nz, cl=4, precision=64, dt_solver=dts)
    fused = emu.stepx(st[:1], acts[:1], ra, 2 * dts, nx=nx, nz=nz, cl=4, precision=64, dt_solver=dts, wrappers=wr)
    np.testing.assert_array_equal(fused["state"], plain["state"])
    bb, uu, ww = emu.unpackx(plain["state"], nx, nz)
    state = np.stack([bb[0], uu[0], ww[0, :-1]]).astype(np.float32)
    cd = W.cell_distance(state, size_state=(nz, nx))
    assert fused["cell_dist"][0] == pytest.approx(cd, abs=1e-12) and cd > 0
    rwd = W.shape_reward(W.normalize_reward(float(-plain["nu_obs"][0]), ra), cd, 0.1)
    assert fused["reward"][0] == pytest.approx(rwd, rel=1e-6)
    np.testing.assert_allclose(fused["obs"][0], W.normalize_observation(plain["obs"][0].copy(), 0.75), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("nx,nz,cl,nxt_global,obs", [(96, 64, 1, False, (8, 48)), (96, 64, 2, False, (8, 48)), (96, 64, 4, True, (8, 48)),
                                                     (192, 128, 4, False, (8, 48)), (192, 128, 8, True, (16, 96)), (128, 64, 2, False, (8, 64))])
def test_pressure_split_mode_matches_oracle(nx, nz, cl, nxt_global, obs):
    """pressure-split mode on every cluster size: the hydrostatic column integral crosses the slabs (column totals sent
    downwards), pNHS needs a cluster-wide mean; state and both pressure fields against the oracle's split scheme."""
    ra, dts = 1e6, 0.015
    P = O.make_params(ra, nx=nx, nz=nz, split_phy=True)
    b, u, w = smooth_state(nx, nz)
    act = np.random.default_rng(3).uniform(-1, 1, 12).astype(np.float32)
    dt = 2 * dts + dts / 3
    r = O.step(P, b, u, w, act.astype(np.float64), O.substep_schedule(dt, dts), want_pressure=True)
    st = emu.pack(b[None], u[None], w[None])
    for precision, tol, ptol in ((64, 1e-13, 1e-10), (32, 2e-6, 2e-3)):
        if precision == 32 and nxt_global:
            continue
        e = emu.stepx(st, act[None], ra, dt, nx=nx, nz=nz, cl=cl, precision=precision, nxt_global=nxt_global, dt_solver=dts, obs=obs, split=True)
        bb, uu, ww = emu.unpackx(e["state"].astype(np.float64), nx, nz)
        assert rel(bb[0], r["b"]) < tol and rel(uu[0], r["u"]) < tol and rel(ww[0], r["w"]) < tol
        pr = e["pressure"][0].astype(np.float64)
        assert rel(pr[0], r["phy"]) < max(tol, 1e-13) * 10
        assert rel(pr[1], r["pnhs"]) < ptol
        oz, ox = nz // obs[0], nx // obs[1]
        np.testing.assert_array_equal(e["obs"][0, 3], e["pressure"][0, 0, ::oz, ::ox].astype(np.float32))
        np.testing.assert_array_equal(e["obs"][0, 4], e["pressure"][0, 1, ::oz, ::ox].astype(np.float32))
        assert e["nan"][0] == 0


def test_pressure_split_set_projection_192():
    """reset path in split mode: the set! projection (dtau = 1) leaves pNHS, pHY' follows from b."""
    nx, nz = 192, 128
    rng = np.random.default_rng(7)
    u = 0.01 * rng.standard_normal((nz, nx))
    w = 0.01 * rng.standard_normal((nz + 1, nx))
    w[0] = 0
    w[-1] = 0
    b = 1.5 - 0.5 * (np.arange(nz)[:, None] + 0.5) / nz + 0.01 * rng.standard_normal((nz, nx))
    P = O.make_params(1e6, nx=nx, nz=nz, split_phy=True)
    up, wp, phi = O.project(P, u, w)
    e = emu.stepx(emu.pack(b[None], u[None], w[None]), np.zeros((1, 12), np.float32), 1e6, 0.015, nx=nx, nz=nz, cl=4, precision=64,
                  dt_solver=0.015, project_first=True, nsub=0, split=True)
    _, uu, ww = emu.unpackx(e["state"], nx, nz)
    assert rel(uu[0], up) < 1e-12 and rel(ww[0], wp) < 1e-12
    assert rel(e["pressure"][0, 1], phi - phi.mean()) < 1e-10
    dz = 2.0 / nz
    phy = np.empty((nz, nx))
    phy[-1] = -0.5 * (b[-1] + (2 * 1.0 - b[-1])) * dz
    for k in range(nz - 2, -1, -1):
        phy[k] = phy[k + 1] - 0.5 * (b[k] + b[k + 1]) * dz
    assert rel(e["pressure"][0, 0], phy) < 1e-12


def test_cfl_guard_splits_fast_rk3_steps_and_leaves_slow_ones_alone(ckpt_ra1e5):
    """The CFL guard of the cluster kernels (`cfl_parts` in rbc2dx_core.h): an RK3 step whose max(|w| dt/dz, |u| dt/dx) exceeds
    the limit is taken as ceil(CFL) equal parts; the measurement is repeated every fourth RK3 step.  Emulated kernel (96 x 64 over 2 CTAs) against the oracle stepped with the
    schedule the guard must choose, decided here from the oracle's own states; below the limit nothing changes, bit for bit."""
    P = O.make_params(1e5, split_phy=False)
    c = ckpt_ra1e5
    b, u, w = c.b[3], c.u[3] * 1.9, c.w[3] * 1.9                      # max|w| ~ 1.8: CFL_z = 1.8 * 0.03 * 32 ~ 1.7
    u, w, _ = O.project(P, u, w)
    a = np.linspace(-1, 1, 12).astype(np.float32)
    dz, dx = 2.0 / 64, 2 * np.pi / 96
    rb, ru, rw, parts_seen = b, u, w, []
    for n, dt in enumerate(O.substep_schedule(0.15)):                    # five RK3 steps of 0.03; measured before steps 0 and 4
        if n % 4 == 0:
            cfl = max(np.abs(rw[:-1]).max() * 0.03 / dz, np.abs(ru).max() * 0.03 / dx)
            parts = int(min(8, max(2, np.ceil(cfl)))) if cfl > 1.4 else 1
        parts_seen.append(parts)
        r = O.step(P, rb, ru, rw, a.astype(np.float64), np.full(parts, dt / parts))
        rb, ru, rw = r["b"], r["u"], r["w"]
    assert parts_seen[0] == 2                                            # the guard really has something to do
    st = emu.packx(b[None], u[None], w[None])
    g = emu.stepx(st, a[None], 1e5, 0.15, cl=2, precision=64, nxt_global=True, cfl_limit=1.4)
    gb, gu, gw = emu.unpackx(g["state"], 96, 64)
    assert rel(gb[0], rb) < 1e-12 and rel(gu[0], ru) < 1e-12 and rel(gw[0], rw) < 1e-12
    plain = emu.stepx(st, a[None], 1e5, 0.15, cl=2, precision=64, nxt_global=True)
    assert rel(emu.unpackx(plain["state"], 96, 64)[0][0], rb) > 1e-6    # ... and without it the result differs
    # a developed flow (CFL ~ 0.9) is untouched by the guard, bit for bit, in the fp32 asynchronous variant too
    s0 = emu.packx(c.b[3][None], c.u[3][None], c.w[3][None])
    for prec, nxtg in ((32, False), (64, True)):
        x = emu.stepx(s0, a[None], 1e5, 0.09, cl=2, precision=prec, nxt_global=nxtg, cfl_limit=1.4)
        y = emu.stepx(s0, a[None], 1e5, 0.09, cl=2, precision=prec, nxt_global=nxtg)
        assert np.array_equal(x["state"], y["state"])
