"""The CUDA kernel's source (rbc_gym_b200/csrc/rbc2d_core.h) compiled for the host and run as a
sequential one-CTA emulator, checked against the fp64 oracle.  This exercises the exact phase
logic of the GPU kernel (strip march, in-place 6x8 FFT, untangle, Thomas, epilogue reductions)
on a machine without a GPU; the `-m gpu` tests repeat the comparison on the real device."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.emu import emu

ACT = {"zero": np.zeros(12, np.float32), "cos": np.cos(2 * np.pi * np.arange(12) / 12).astype(np.float32),
       "ramp": np.linspace(-1, 1, 12).astype(np.float32)}


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


def oracle_step(c, ep, act, ra, dt, split):
    P = O.make_params(ra, split_phy=split)
    r = O.step(P, c.b[ep], c.u[ep], c.w[ep], act.astype(np.float64), O.substep_schedule(dt), want_pressure=True)
    r["nu"] = O.nusselt_state_obs(P, r["b"], r["u"], r["w"])
    return r


@pytest.mark.parametrize("split,nxt_global", [(False, False), (False, True), (True, True)])
def test_emulated_kernel_fp64_matches_oracle(ckpt_ra1e5, split, nxt_global):
    c = ckpt_ra1e5
    eps, acts = [0, 7, 16], [ACT["cos"], ACT["ramp"], ACT["zero"]]
    st = emu.pack(c.b[eps], c.u[eps], c.w[eps])
    e = emu.step(st, np.stack(acts), 1e5, 1.0, precision=64, split=split, nxt_global=nxt_global, pressure=split)
    b, u, w = emu.unpack(e["state"])
    for j, (ep, a) in enumerate(zip(eps, acts)):
        r = oracle_step(c, ep, a, 1e5, 1.0, split)
        assert rel(b[j], r["b"]) < 1e-13 and rel(u[j], r["u"]) < 1e-13 and rel(w[j], r["w"]) < 1e-13
        assert e["nu_state"][j] == pytest.approx(r["nu"][0], abs=1e-10)
        assert e["nu_obs"][j] == pytest.approx(r["nu"][1], abs=1e-10)
        assert e["reward"][j] == pytest.approx(-r["nu"][1], rel=1e-6)
        ref_obs = O.observe(O.state_channels(r["b"], r["u"], r["w"])).astype(np.float32)
        np.testing.assert_array_equal(e["obs"][j, :3], ref_obs)
        if split:
            assert rel(e["pressure"][j, 0], r["phy"]) < 1e-13
            assert rel(e["pressure"][j, 1], r["pnhs"]) < 1e-11
    assert np.all(e["nan"] == 0) and np.all(e["t"] == 1.0) and np.all(e["step"] == 2) and np.all(e["truncated"] == 0)


def test_emulated_kernel_fp32_close_to_oracle(ckpt_ra1e5):
    c = ckpt_ra1e5
    st = emu.pack(c.b[:2], c.u[:2], c.w[:2])
    e = emu.step(st, np.stack([ACT["cos"], ACT["ramp"]]), 1e5, 1.5, precision=32)
    b, u, w = emu.unpack(e["state"].astype(np.float64))
    for j, a in enumerate([ACT["cos"], ACT["ramp"]]):
        r = oracle_step(c, j, a, 1e5, 1.5, False)
        assert rel(b[j], r["b"]) < 2e-6 and rel(u[j], r["u"]) < 5e-6 and rel(w[j], r["w"]) < 5e-6
        assert e["nu_state"][j] == pytest.approx(r["nu"][0], rel=2e-5)
        assert e["nu_obs"][j] == pytest.approx(r["nu"][1], rel=2e-5)
    dx, dz = 2 * np.pi / 96, 2 / 64
    div = (np.roll(u, -1, axis=-1) - u) / dx + (w[:, 1:] - w[:, :-1]) / dz
    assert np.abs(div).max() < 2e-5                      # fp32 projection keeps the flow discretely solenoidal


def test_emulated_truncation_and_nan_flag(ckpt_ra1e5):
    c = ckpt_ra1e5
    st = emu.pack(c.b[:2], c.u[:2], c.w[:2])
    st[1, 100] = np.nan
    e = emu.step(st, np.zeros((2, 12), np.float32), 1e5, 0.06, precision=64, t0=[299.95, 0.0], episode_length=300.0)
    assert list(e["truncated"]) == [1, 0] and list(e["nan"]) == [0, 1]


def test_emulated_full_observation_grid(ckpt_ra1e5):
    # config 1 style: sensors = full grid -> obs == state, Nu_obs == Nu_state
    c = ckpt_ra1e5
    st = emu.pack(c.b[:1], c.u[:1], c.w[:1])
    e = emu.step(st, ACT["cos"][None], 1e5, 0.09, precision=64, obs=(64, 96))
    b, u, w = emu.unpack(e["state"])
    np.testing.assert_array_equal(e["obs"][0], np.stack([b[0], u[0], w[0, :-1]]).astype(np.float32))
    assert e["nu_obs"][0] == pytest.approx(e["nu_state"][0], abs=1e-12)
