"""Parity of the CUDA path (through the C ABI) against the fp64 oracle, on a real GPU.

Tolerances (SURVEY §8c "Achievable parity claims"):
  fp64 validation mode: relative L2 per field after one action step <= 1e-10
  fp32 throughput mode: relative L2 per field after one action step <= 1e-5
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

ACT = {"zero": np.zeros(12, np.float32), "cos": np.cos(2 * np.pi * np.arange(12) / 12).astype(np.float32),
       "ramp": np.linspace(-1, 1, 12).astype(np.float32)}
FP64_TOL, FP32_TOL = 1e-10, 1e-5


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


@pytest.fixture(scope="module")
def B():
    from rbc_gym_b200 import backend
    return backend


def make_sim(B, n, ckpt, eps, **kw):
    import torch
    sim = B.Sim2D(n, **kw)
    sim.load_checkpoints(ckpt)
    sim.reset_from_checkpoints(torch.tensor(eps, dtype=torch.int32))
    return sim


def oracle_ref(c, ep, act, ra, dt, split=False):
    P = O.make_params(ra, split_phy=split)
    r = O.step(P, c.b[ep], c.u[ep], c.w[ep], act.astype(np.float64), O.substep_schedule(dt), want_pressure=True)
    r["nu"] = O.nusselt_state_obs(P, r["b"], r["u"], r["w"])
    return r


@pytest.mark.parametrize("precision,tol", [(64, FP64_TOL), (32, FP32_TOL)])
@pytest.mark.parametrize("dt", [1.0, 1.5])
def test_one_action_step_matches_oracle(B, ckpt_ra1e5, precision, tol, dt):
    import torch
    c = ckpt_ra1e5
    eps = [0, 7, 3, 16]
    rng = np.random.default_rng(5)
    acts = np.stack([ACT["zero"], ACT["cos"], ACT["ramp"], rng.uniform(-1, 1, 12).astype(np.float32)])
    sim = make_sim(B, 4, c, eps, ra=1e5, dt_action=dt, precision=precision)
    obs, rew, nus, nuo, trunc, nan = sim.step(torch.from_numpy(acts).cuda())
    b, u, w = B.split_fields(sim.fields())
    t, step = sim.info()
    assert np.all(t == dt) and np.all(step == 2)
    assert not nan.any().item() and not trunc.any().item()
    for j, ep in enumerate(eps):
        r = oracle_ref(c, ep, acts[j], 1e5, dt)
        assert rel(b[j], r["b"]) < tol and rel(u[j], r["u"]) < tol and rel(w[j], r["w"]) < tol
        nu_tol = 1e-9 if precision == 64 else 5e-5
        assert nus[j].item() == pytest.approx(r["nu"][0], rel=nu_tol)
        assert nuo[j].item() == pytest.approx(r["nu"][1], rel=nu_tol)
        assert rew[j].item() == pytest.approx(-r["nu"][1], rel=max(nu_tol, 1e-6))
        ref_obs = O.observe(O.state_channels(r["b"], r["u"], r["w"]))
        np.testing.assert_allclose(obs[j].cpu().numpy(), ref_obs, rtol=0, atol=2e-7 if precision == 64 else 1e-5)
    sim.close()


def test_pressure_split_mode_matches_oracle(B, ckpt_ra1e5):
    import torch
    c = ckpt_ra1e5
    sim = make_sim(B, 2, c, [0, 7], ra=1e5, dt_action=0.3, precision=64, pressure=True)
    acts = np.stack([ACT["cos"], ACT["ramp"]])
    obs, *_ = sim.step(torch.from_numpy(acts).cuda())
    st = sim.get_state().cpu().numpy()
    assert st.shape == (2, 5, 64, 96) and obs.shape == (2, 5, 8, 48)
    b, u, w = B.split_fields(sim.fields())
    for j, ep in enumerate([0, 7]):
        r = oracle_ref(c, ep, acts[j], 1e5, 0.3, split=True)
        assert rel(b[j], r["b"]) < FP64_TOL and rel(u[j], r["u"]) < FP64_TOL and rel(w[j], r["w"]) < FP64_TOL
        np.testing.assert_allclose(st[j, 3], r["phy"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(st[j, 4], r["pnhs"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(st[j, 2], r["w"][:-1], rtol=0, atol=2e-7)
        np.testing.assert_array_equal(obs[j, 3:].cpu().numpy(), st[j, 3:, ::8, ::2])
    sim.close()


def test_multi_step_rollout_fp64_tracks_oracle(B, ckpt_ra1e5):
    """Three consecutive action steps with changing actions (G- handling, clock, state write-back)."""
    import torch
    c = ckpt_ra1e5
    sim = make_sim(B, 1, c, [11], ra=1e5, dt_action=0.3, precision=64)
    P = O.make_params(1e5, split_phy=False)
    b, u, w = c.b[11], c.u[11], c.w[11]
    rng = np.random.default_rng(0)
    for n in range(3):
        a = rng.uniform(-1, 1, (1, 12)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda())
        r = O.step(P, b, u, w, a[0].astype(np.float64), O.substep_schedule(0.3))
        b, u, w = r["b"], r["u"], r["w"]
    gb, gu, gw = B.split_fields(sim.fields())
    assert rel(gb[0], b) < FP64_TOL and rel(gu[0], u) < FP64_TOL and rel(gw[0], w) < FP64_TOL
    t, step = sim.info()
    assert t[0] == pytest.approx(0.9, abs=1e-12) and step[0] == 4
    sim.close()


@pytest.mark.parametrize("precision", [32, 64])
def test_batch_larger_than_grid_is_replica_exact(B, ckpt_ra1e5, precision):
    """More environments than resident CTAs: every replica of the same (state, action) must be bitwise equal."""
    import torch
    n = 333
    eps = [i % 5 for i in range(n)]
    sim = make_sim(B, n, ckpt_ra1e5, eps, ra=1e5, dt_action=0.09, precision=precision)
    acts = np.stack([ACT["cos"] * ((i % 5) / 4.0) for i in range(n)]).astype(np.float32)
    obs, rew, *_ = sim.step(torch.from_numpy(acts).cuda())
    f = sim.fields()
    for i in range(5, n):
        assert np.array_equal(f[i], f[i % 5])
    assert torch.equal(obs[5:10], obs[0:5]) and torch.equal(rew[330:333], rew[0:3])
    sim.close()


def test_host_buffer_entry_point_equals_device_entry_point(B, ckpt_ra1e5):
    import torch
    eps = [0, 1, 2, 3, 4, 5]
    acts = np.random.default_rng(1).uniform(-1, 1, (6, 12)).astype(np.float32)
    s1 = make_sim(B, 6, ckpt_ra1e5, eps, ra=1e5, dt_action=0.15, precision=32)
    s2 = make_sim(B, 6, ckpt_ra1e5, eps, ra=1e5, dt_action=0.15, precision=32)
    obs, rew, nus, nuo, tr, nan = s1.step(torch.from_numpy(acts).cuda())
    out = s2.step_host(acts)
    np.testing.assert_array_equal(out["obs"], obs.cpu().numpy())
    np.testing.assert_array_equal(out["reward"], rew.cpu().numpy())
    np.testing.assert_array_equal(out["nu_state"], nus.cpu().numpy())
    np.testing.assert_array_equal(out["nu_obs"], nuo.cpu().numpy())
    assert np.array_equal(s1.fields(), s2.fields())
    s1.close(); s2.close()


def test_host_step_in_overlapped_chunks_equals_device_step(B, ckpt_ra1e5):
    """rbc2d_step_host launches a multi-wave batch in chunks of whole waves and copies each chunk out on a second stream
    while the next one computes: same outputs as the single device-buffer launch, over two consecutive steps, into pinned
    buffers; the step's kernel time spans the chunks."""
    import torch
    n = 700                                                   # 5 waves of 148 CTAs -> 4 chunks
    eps = [i % 20 for i in range(n)]
    rng = np.random.default_rng(5)
    s1 = make_sim(B, n, ckpt_ra1e5, eps, ra=1e5, dt_action=0.09, precision=32)
    s2 = make_sim(B, n, ckpt_ra1e5, eps, ra=1e5, dt_action=0.09, precision=32)
    out = s2.alloc_host_outputs(pinned=True)
    for _ in range(2):
        acts = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        obs, rew, nus, nuo, tr, nan = s1.step(torch.from_numpy(acts).cuda())
        s2.step_host(acts, out)
        np.testing.assert_array_equal(out["obs"], obs.cpu().numpy())
        np.testing.assert_array_equal(out["reward"], rew.cpu().numpy())
        np.testing.assert_array_equal(out["nu_state"], nus.cpu().numpy())
        np.testing.assert_array_equal(out["nu_obs"], nuo.cpu().numpy())
        np.testing.assert_array_equal(out["truncated"], tr.cpu().numpy())
        np.testing.assert_array_equal(out["nan"], nan.cpu().numpy())
    assert np.array_equal(s1.fields(), s2.fields())
    t1, c1 = s1.info()
    t2, c2 = s2.info()
    assert np.array_equal(t1, t2) and np.array_equal(c1, c2)
    assert 0.0 < s2.last_step_kernel_ms() < 10 * s1.last_step_kernel_ms() + 1.0
    s1.close(); s2.close()


def test_reset_observe_and_state_layout(B, ckpt_ra1e5):
    import torch
    c = ckpt_ra1e5
    sim = make_sim(B, 3, c, [2, 9, 19], ra=1e5, dt_action=1.0, precision=64)
    obs, nus, nuo = sim.observe()
    P = O.make_params(1e5)
    st = sim.get_state().cpu().numpy()
    for j, ep in enumerate([2, 9, 19]):
        ref = O.state_channels(c.b[ep], c.u[ep], c.w[ep])
        np.testing.assert_array_equal(st[j], ref.astype(np.float32))
        np.testing.assert_array_equal(obs[j].cpu().numpy(), O.observe(ref).astype(np.float32))
        ns, no = O.nusselt_state_obs(P, c.b[ep], c.u[ep], c.w[ep])
        assert nus[j].item() == pytest.approx(ns, rel=1e-12) and nuo[j].item() == pytest.approx(no, rel=1e-12)
    # partial reset of env 1 only
    sim.step(torch.zeros(3, 12, device="cuda"))
    sim.reset_from_checkpoints(torch.tensor([5], dtype=torch.int32), env_ids=torch.tensor([1], dtype=torch.int32))
    t, step = sim.info()
    assert list(t) == [1.0, 0.0, 1.0] and list(step) == [2, 1, 2]
    b, u, w = B.split_fields(sim.fields())
    assert np.array_equal(b[1], c.b[5]) and np.array_equal(w[1], c.w[5])
    sim.close()


def test_noise_init_is_projected_like_set(B):
    """initialize_model (rbc_sim2D.jl:163-171): random fields, then set! projects to a solenoidal flow."""
    rng = np.random.default_rng(42)
    kick, z = 0.01, (np.arange(64) + 0.5) * (2 / 64)
    b = np.clip(1 + (2 - z)[:, None] / 2 + kick * rng.standard_normal((64, 96)), 1, 2)
    u = kick * rng.standard_normal((64, 96))
    w = kick * rng.standard_normal((65, 96)); w[0] = 0; w[-1] = 0
    sim = B.Sim2D(1, ra=1e4, dt_action=1.5, precision=64)
    sim.reset_from_fields(B.pack_fields(b[None], u[None], w[None]), project=True)
    gb, gu, gw = B.split_fields(sim.fields())
    P = O.make_params(1e4)
    ou, ow, _ = O.project(P, u, w)
    assert rel(gu[0], ou) < 1e-12 and rel(gw[0], ow) < 1e-12 and np.array_equal(gb[0], b)
    dx, dz = 2 * np.pi / 96, 2 / 64
    div = (np.roll(gu[0], -1, axis=-1) - gu[0]) / dx + (gw[0][1:] - gw[0][:-1]) / dz
    assert np.abs(div).max() < 1e-13
    sim.close()


def test_nan_is_reported_not_raised_and_truncation(B, ckpt_ra1e5):
    import torch
    c = ckpt_ra1e5
    f = B.pack_fields(c.b[:2], c.u[:2], c.w[:2])
    f[1, 1234] = np.nan
    sim = B.Sim2D(2, ra=1e5, dt_action=0.06, precision=32, episode_length=0.1)
    sim.reset_from_fields(f, project=False)
    *_, trunc, nan = sim.step(torch.zeros(2, 12, device="cuda"))
    assert nan.tolist() == [0, 1] and trunc.tolist() == [0, 0]
    *_, trunc, nan = sim.step(torch.zeros(2, 12, device="cuda"))
    assert trunc.tolist() == [1, 1]
    sim.close()


def test_fp32_full_batch_properties(B, ckpt_ra1e5):
    """BASELINE config 2 size (4096 envs): size-independent properties after a dt=1 action step."""
    import torch
    n = 4096
    sim = make_sim(B, n, ckpt_ra1e5, [i % 20 for i in range(n)], ra=1e5, dt_action=1.0, precision=32)
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    acts = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
    obs, rew, nus, nuo, trunc, nan = sim.step(acts)
    assert not nan.any().item()
    b, u, w = B.split_fields(sim.fields())
    dx, dz = 2 * np.pi / 96, 2 / 64
    div = (np.roll(u, -1, axis=-1) - u) / dx + (w[:, 1:] - w[:, :-1]) / dz
    assert np.abs(div).max() < 5e-5                                   # discretely solenoidal (fp32)
    assert np.abs(w[:, 0]).max() == 0 and np.abs(w[:, -1]).max() == 0  # impenetrable walls
    assert np.abs(w.mean(axis=2)).max() < 2e-6                        # zero net vertical mass flux per level
    assert b.min() > 0.9 and b.max() < 2.8 and np.abs(w).max() < 1.5
    # spot-check 3 environments against the oracle
    a = acts.cpu().numpy()
    for j in (0, 2047, 4095):
        r = oracle_ref(ckpt_ra1e5, j % 20, a[j], 1e5, 1.0)
        assert rel(b[j], r["b"]) < FP32_TOL and rel(u[j], r["u"]) < FP32_TOL and rel(w[j], r["w"]) < FP32_TOL
        assert rew[j].item() == pytest.approx(-r["nu"][1], rel=5e-5)
    sim.close()
