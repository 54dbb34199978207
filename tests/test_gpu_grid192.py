"""Config 3 (192 x 128, Ra=1e6, dt_solver=0.015) on the GPU: the thread-block-cluster kernel (one cluster per
environment, slabs in distributed shared memory) through the C ABI against the fp64 oracle.

Tolerances as for the 96 x 64 grid: rel-L2 per field after one action step <= 1e-10 (fp64), <= 1e-5 (fp32)."""
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

from oracle import oracle as O  # noqa: E402
from tests.gridstates import smooth_state  # noqa: E402

NX, NZ, RA, DTS = 192, 128, 1e6, 0.015


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


@pytest.fixture(scope="module")
def B():
    from rbc_gym_b200 import backend
    return backend


@pytest.mark.parametrize("precision,tol", [(64, 1e-10), (32, 1e-5)])
def test_one_action_step_192x128_matches_oracle(B, precision, tol):
    import torch
    P = O.make_params(RA, nx=NX, nz=NZ, split_phy=False)
    states = [smooth_state(NX, NZ, seed=s) for s in (0, 1, 2)]
    rng = np.random.default_rng(7)
    acts = rng.uniform(-1, 1, (3, 12)).astype(np.float32)
    acts[2] = 0
    dt = 0.1                                                  # 6 x 0.015 + 0.01: 7 RK3 steps, clipped last step
    sim = B.Sim2D(3, ra=RA, dt_action=dt, dt_solver=DTS, state_shape=(NZ, NX), obs_shape=(8, 48), precision=precision)
    sim.reset_from_fields(np.concatenate([B.pack_fields(b[None], u[None], w[None]) for b, u, w in states]), project=False)
    obs, rew, nus, nuo, trunc, nan = sim.step(torch.from_numpy(acts).cuda())
    b, u, w = B.split_fields(sim.fields(), (NZ, NX))
    t, step = sim.info()
    assert np.all(t == dt) and np.all(step == 2) and not nan.any().item() and not trunc.any().item()
    for j, (b0, u0, w0) in enumerate(states):
        r = O.step(P, b0, u0, w0, acts[j].astype(np.float64), O.substep_schedule(dt, DTS))
        assert rel(b[j], r["b"]) < tol and rel(u[j], r["u"]) < tol and rel(w[j], r["w"]) < tol
        ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"])
        nu_tol = 1e-8 if precision == 64 else 2e-3
        assert nus[j].item() == pytest.approx(ns, abs=nu_tol) and nuo[j].item() == pytest.approx(no, abs=nu_tol)
        assert rew[j].item() == pytest.approx(-no, abs=max(nu_tol, 1e-4))
        ref_obs = O.observe(O.state_channels(r["b"], r["u"], r["w"]))
        np.testing.assert_allclose(obs[j].cpu().numpy(), ref_obs, rtol=0, atol=2e-7 if precision == 64 else 1e-5)
    st = sim.get_state().cpu().numpy()
    assert st.shape == (3, 3, NZ, NX)
    np.testing.assert_allclose(st[0, 2], w[0, :-1], rtol=0, atol=1e-6)
    info = sim.launch_info()
    assert info["grid"] == 3 * (4 if precision == 32 else 8)   # one cluster per environment
    sim.close()


@pytest.mark.parametrize("precision", [32, 64])
def test_more_envs_than_clusters_is_replica_exact_192(B, precision):
    import torch
    n = 90                                                    # > 33 resident clusters of 4 (or 16 of 8)
    base = [smooth_state(NX, NZ, seed=s) for s in (0, 1, 2)]
    fields = np.concatenate([B.pack_fields(*(a[None] for a in base[i % 3])) for i in range(n)])
    sim = B.Sim2D(n, ra=RA, dt_action=0.045, dt_solver=DTS, state_shape=(NZ, NX), precision=precision)
    sim.reset_from_fields(fields, project=False)
    acts = np.stack([np.cos(2 * np.pi * np.arange(12) / 12) * ((i % 3) / 2.0) for i in range(n)]).astype(np.float32)
    obs, rew, *_ = sim.step(torch.from_numpy(acts).cuda())
    f = sim.fields()
    r = rew.cpu().numpy()
    for i in range(3, n):
        assert np.array_equal(f[i], f[i % 3]) and r[i] == r[i % 3]
    sim.close()


def test_config3_env_noise_init_wrappers_rollout(B):
    """Config 3 end to end: the reference env class at doubled resolution, noise initialisation (kick 0.01, seed 42)
    projected on the device, NormalizeReward + reward shaping fused, a short fp32 rollout that stays solenoidal."""
    import torch
    from rbc_gym_b200.envs.rbc2d import noise_initial_fields
    from rbc_gym_b200 import wrappers as W
    nenv = 8
    sim = B.Sim2D(nenv, ra=RA, dt_action=0.3, dt_solver=DTS, state_shape=(NZ, NX), precision=32)
    rng = np.random.default_rng(42)
    sim.reset_from_fields(np.concatenate([noise_initial_fields(rng, (NZ, NX), kick=0.01) for _ in range(nenv)]), project=True)
    b, u, w = B.split_fields(sim.fields(), (NZ, NX))
    dx, dz = 2 * np.pi / NX, 2.0 / NZ
    div = (np.roll(u, -1, axis=-1) - u) / dx + (w[:, 1:] - w[:, :-1]) / dz
    assert np.abs(div).max() < 5e-5 and np.all(w[:, 0] == 0) and np.all(w[:, -1] == 0)
    sim.set_wrappers(normalize_obs=True, normalize_reward=True, shaping_weight=0.1)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(3):
        a = torch.rand((nenv, 12), device="cuda", generator=g) * 2 - 1
        obs, rew, nus, nuo, trunc, nan = sim.step(a)
    assert not nan.any().item()
    state = sim.get_state().cpu().numpy()
    cds = sim.cell_dist()
    for j in range(nenv):
        cd = W.cell_distance(state[j], size_state=(NZ, NX))
        assert cds[j] == pytest.approx(cd, abs=1e-9)
        r = W.shape_reward(W.normalize_reward(float(-nuo[j].item()), RA), cd, 0.1)
        assert rew[j].item() == pytest.approx(r, rel=1e-5, abs=1e-6)
    b, u, w = B.split_fields(sim.fields(), (NZ, NX))
    div = (np.roll(u, -1, axis=-1) - u) / dx + (w[:, 1:] - w[:, :-1]) / dz
    assert np.abs(div).max() < 1e-4
    sim.close()


def test_env_class_at_192x128(B):
    from rbc_gym_b200.envs.rbc2d import RayleighBenardConvection2DEnv
    env = RayleighBenardConvection2DEnv(rayleigh_number=1_000_000, state_shape=[NZ, NX], observation_shape=[8, 48],
                                        heater_duration=0.15, dt_solver=DTS, precision=64)
    obs, info = env.reset(seed=42)
    assert obs.shape == (3, 8, 48) and info["state"].shape == (3, NZ, NX)
    obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
    assert np.isfinite(reward) and info["t"] == pytest.approx(0.15) and not truncated
    assert info["nusselt_obs"] == pytest.approx(-reward)
    env.close()


def test_env_class_at_192x128_with_pressure_channels(B):
    """`pressure=True` (rbc2D.py:57,94-97: two unbounded extra channels) on the config-3 grid."""
    from rbc_gym_b200.envs.rbc2d import RayleighBenardConvection2DEnv
    env = RayleighBenardConvection2DEnv(rayleigh_number=1_000_000, state_shape=[NZ, NX], observation_shape=[8, 48],
                                        heater_duration=0.15, dt_solver=DTS, pressure=True)
    obs, info = env.reset(seed=42)
    assert obs.shape == (5, 8, 48) and info["state"].shape == (5, NZ, NX) and env.observation_space.shape == (5, 8, 48)
    obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
    assert obs.shape == (5, 8, 48) and np.all(np.isfinite(obs)) and np.isfinite(reward)
    st = info["state"]
    np.testing.assert_array_equal(obs[3:], st[3:, ::16, ::4])
    assert abs(st[4].mean()) < 1e-4                              # pNHS in the zero-mean gauge
    assert np.all(np.diff(st[3].mean(axis=1)) > 0)               # pHY' integrates b > 0 downwards from the top wall
    env.close()


def test_unregistered_grid_is_rejected(B):
    with pytest.raises(RuntimeError, match="registered grids"):
        B.Sim2D(1, ra=1e5, state_shape=(100, 200))


@pytest.mark.parametrize("nx,nz,precision,tol,ptol", [(192, 128, 64, 1e-10, 1e-8), (192, 128, 32, 1e-5, 5e-3), (128, 64, 64, 1e-10, 1e-8),
                                                      (128, 64, 32, 1e-5, 5e-3), (64, 64, 32, 1e-5, 5e-3), (96, 128, 64, 1e-10, 1e-8),
                                                      (192, 64, 32, 1e-5, 5e-3), (128, 128, 64, 1e-10, 1e-8)])
def test_pressure_channels_on_cluster_grids_match_oracle(B, nx, nz, precision, tol, ptol):
    """Pressure-split mode of the cluster kernel: the hydrostatic column integral and the zero-mean gauge of pNHS cross the
    CTAs of the cluster.  State, both pressure fields (get_state channels 3-4) and their sensor sub-sample against the
    oracle's split scheme; more environments than resident clusters."""
    import torch
    n = 40
    P = O.make_params(RA, nx=nx, nz=nz, split_phy=True)
    states = [smooth_state(nx, nz, seed=s) for s in (0, 1)]
    acts = np.random.default_rng(11).uniform(-1, 1, (2, 12)).astype(np.float32)
    dt = 3 * DTS + 0.005
    sim = B.Sim2D(n, ra=RA, dt_action=dt, dt_solver=DTS, state_shape=(nz, nx), obs_shape=(8, 48 if nx in (192, 96) else 64), precision=precision,
                  pressure=True)
    sim.reset_from_fields(np.concatenate([B.pack_fields(*(a[None] for a in states[i % 2])) for i in range(n)]), project=False)
    obs, *_ = sim.step(torch.from_numpy(acts[np.arange(n) % 2]).cuda())
    st = sim.get_state().cpu().numpy()
    assert st.shape == (n, 5, nz, nx) and obs.shape[1] == 5
    b, u, w = B.split_fields(sim.fields(), (nz, nx))
    oz, ox = nz // 8, nx // obs.shape[3]
    for j in (0, 1, n - 2, n - 1):
        b0, u0, w0 = states[j % 2]
        r = O.step(P, b0, u0, w0, acts[j % 2].astype(np.float64), O.substep_schedule(dt, DTS), want_pressure=True)
        assert rel(b[j], r["b"]) < tol and rel(u[j], r["u"]) < tol and rel(w[j], r["w"]) < tol
        assert rel(st[j, 3], r["phy"]) < max(tol, 1e-6) and rel(st[j, 4], r["pnhs"]) < max(ptol, 1e-6)   # get_state is fp32
        np.testing.assert_array_equal(obs[j, 3:].cpu().numpy(), st[j, 3:, ::oz, ::ox])
    sim.close()


def test_pressure_at_reset_on_192x128(B):
    """reset in pressure mode: the set! projection (dtau = 1) leaves pNHS, pHY' follows from b; reset()'s observation
    carries the channels from the stored fields (observe-only launch)."""
    rng = np.random.default_rng(5)
    u = 0.01 * rng.standard_normal((NZ, NX))
    w = 0.01 * rng.standard_normal((NZ + 1, NX))
    w[0] = 0
    w[-1] = 0
    b = 1.5 - 0.5 * (np.arange(NZ)[:, None] + 0.5) / NZ + 0.01 * rng.standard_normal((NZ, NX))
    P = O.make_params(RA, nx=NX, nz=NZ, split_phy=True)
    up, wp, phi = O.project(P, u, w)
    sim = B.Sim2D(2, ra=RA, dt_action=0.045, dt_solver=DTS, state_shape=(NZ, NX), precision=64, pressure=True)
    sim.reset_from_fields(np.concatenate([B.pack_fields(b[None], u[None], w[None])] * 2), project=True)
    st = sim.get_state().cpu().numpy()
    assert rel(st[1, 4], phi - phi.mean()) < 1e-5
    dz = 2.0 / NZ
    phy = np.empty((NZ, NX))
    phy[-1] = -0.5 * (b[-1] + (2 * 1.0 - b[-1])) * dz
    for k in range(NZ - 2, -1, -1):
        phy[k] = phy[k + 1] - 0.5 * (b[k] + b[k + 1]) * dz
    assert rel(st[1, 3], phy) < 1e-6
    obs = sim.observe()[0].cpu().numpy()
    np.testing.assert_array_equal(obs[1, 3:], st[1, 3:, ::16, ::4])
    sim.close()


@pytest.mark.parametrize("precision,tol", [(64, 1e-12), (32, 2e-6)])
def test_generic_split_kernel_on_96x64_agrees_with_dedicated_split_kernel(B, ckpt_ra1e5, precision, tol):
    import torch
    c = ckpt_ra1e5
    acts = np.random.default_rng(4).uniform(-1, 1, (4, 12)).astype(np.float32)
    out = []
    for flag in ("0", "g"):
        os.environ["RBC_B200_CLUSTER"] = flag
        try:
            sim = B.Sim2D(4, ra=1e5, dt_action=0.3, precision=precision, pressure=True)
        finally:
            os.environ.pop("RBC_B200_CLUSTER", None)
        sim.load_checkpoints(c)
        sim.reset_from_checkpoints(torch.tensor([0, 7, 3, 16], dtype=torch.int32))
        obs0 = sim.observe()[0].cpu().numpy()
        obs, *_ = sim.step(torch.from_numpy(acts).cuda())
        out.append((sim.fields(), sim.get_state().cpu().numpy(), obs.cpu().numpy(), obs0, sim.launch_info()["grid"]))
        sim.close()
    assert out[0][4] == 4 and out[1][4] == (4 if precision == 32 else 8)
    assert rel(out[1][0], out[0][0]) < tol
    for ch in (3, 4):
        assert rel(out[1][1][:, ch], out[0][1][:, ch]) < max(tol * 500, 1e-6)
        assert rel(out[1][3][:, ch], out[0][3][:, ch]) < max(tol * 500, 1e-6)
        assert rel(out[1][2][:, ch], out[0][2][:, ch]) < max(tol * 500, 1e-6)


@pytest.mark.parametrize("precision,tol,flag,ctas", [(64, 1e-12, "1", 8), (32, 2e-6, "1", 8), (32, 2e-6, "g", 4)])
def test_cluster_kernel_on_96x64_agrees_with_dedicated_kernel(B, ckpt_ra1e5, precision, tol, flag, ctas):
    """RBC_B200_CLUSTER=1 routes the 96 x 64 grid through the 2-CTA cluster kernel, =g through the same generic code as one
    CTA per environment: same results as the dedicated one-CTA kernel."""
    import torch
    c = ckpt_ra1e5
    acts = np.random.default_rng(3).uniform(-1, 1, (4, 12)).astype(np.float32)
    out = []
    for flag in ("0", flag):
        os.environ["RBC_B200_CLUSTER"] = flag
        try:
            sim = B.Sim2D(4, ra=1e5, dt_action=0.3, precision=precision)
        finally:
            os.environ.pop("RBC_B200_CLUSTER", None)
        sim.load_checkpoints(c)
        sim.reset_from_checkpoints(torch.tensor([0, 7, 3, 16], dtype=torch.int32))
        _, rew, nus, nuo, *_ = sim.step(torch.from_numpy(acts).cuda())
        out.append((sim.fields(), nus.cpu().numpy(), nuo.cpu().numpy(), sim.launch_info()["grid"]))
        sim.close()
    assert out[0][3] == 4 and out[1][3] == ctas
    assert rel(out[1][0], out[0][0]) < tol
    np.testing.assert_allclose(out[1][1], out[0][1], rtol=1e-9 if precision == 64 else 1e-4)
    np.testing.assert_allclose(out[1][2], out[0][2], rtol=1e-9 if precision == 64 else 1e-4)


def test_checkpoint_generator_and_reset_at_192x128(B, tmp_path):
    """Config 3 has no shipped checkpoints: the generator spins up 192 x 128 episodes on the device, writes the reference's
    HDF5 layout, and the env resets from the file exactly (fp64 round trip through the file and the device bank)."""
    import torch
    from rbc_gym_b200.checkpoints import simulate_2d_rb
    from rbc_gym_b200.h5lite import load_checkpoint_2d
    from rbc_gym_b200.envs.rbc2d import RayleighBenardConvection2DEnv
    path, stats = simulate_2d_rb(tmp_path / "train", seed=42, random_inits=3, ra=RA, delta_t=DTS, delta_t_snap=0.3, duration=3.0,
                                 state_shape=(NZ, NX))
    assert path.name == "ckpt_ra1000000_192x128.h5" and np.all(np.isfinite(stats["nu_state"]))
    c = load_checkpoint_2d(path)
    assert c.shape == (NZ, NX) and c.num_episodes == 3 and c.start_seed == 42
    dx, dz = 2 * np.pi / NX, 2.0 / NZ
    div = (np.roll(c.u, -1, axis=-1) - c.u) / dx + (c.w[:, 1:] - c.w[:, :-1]) / dz
    assert np.abs(div).max() < 1e-11                                     # fp64 spin-up: discretely solenoidal
    env = RayleighBenardConvection2DEnv(rayleigh_number=1_000_000, state_shape=[NZ, NX], heater_duration=0.15, dt_solver=DTS,
                                        checkpoint=str(path), checkpoint_idx=1, precision=64)
    obs, info = env.reset(seed=0)
    b, u, w = B.split_fields(env.sim.fields(), (NZ, NX))
    assert np.array_equal(b[0], c.b[1]) and np.array_equal(u[0], c.u[1]) and np.array_equal(w[0], c.w[1])
    np.testing.assert_array_equal(obs[0], c.b[1][::16, ::4].astype(np.float32))
    env.close()


@pytest.mark.parametrize("precision,tol", [(64, 1e-10), (32, 1e-5)])
def test_power_of_two_grid_128x64_matches_oracle(B, precision, tol):
    """A second registered cluster grid (128 x 64: 4 x 16 FFT split, 2 CTAs per environment, 256 threads per CTA)."""
    import torch
    nx, nz = 128, 64
    P = O.make_params(1e5, nx=nx, nz=nz, split_phy=False)
    states = [smooth_state(nx, nz, seed=s) for s in (5, 6)]
    acts = np.random.default_rng(11).uniform(-1, 1, (2, 12)).astype(np.float32)
    sim = B.Sim2D(2, ra=1e5, dt_action=0.2, dt_solver=0.03, state_shape=(nz, nx), obs_shape=(8, 64), precision=precision)
    sim.reset_from_fields(np.concatenate([B.pack_fields(b[None], u[None], w[None]) for b, u, w in states]), project=False)
    obs, rew, nus, nuo, trunc, nan = sim.step(torch.from_numpy(acts).cuda())
    b, u, w = B.split_fields(sim.fields(), (nz, nx))
    assert not nan.any().item() and sim.launch_info()["grid"] == 4
    for j, (b0, u0, w0) in enumerate(states):
        r = O.step(P, b0, u0, w0, acts[j].astype(np.float64), O.substep_schedule(0.2, 0.03))
        assert rel(b[j], r["b"]) < tol and rel(u[j], r["u"]) < tol and rel(w[j], r["w"]) < tol
        ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"], (8, 64))
        assert nuo[j].item() == pytest.approx(no, abs=1e-8 if precision == 64 else 2e-3)
    sim.close()


@pytest.mark.parametrize("precision,tol", [(64, 1e-10), (32, 1e-5)])
@pytest.mark.parametrize("nx,nz", [(64, 32), (64, 64), (96, 32), (96, 128), (128, 32), (128, 128), (192, 64)])
def test_further_registered_grids_match_oracle(B, nx, nz, precision, tol):
    """`state_shape` is a free keyword of the reference environment (`rbc2D.py:43-60`): the further registered decompositions of the
    cluster kernel (`rbc2dx_more.cuh`: widths 64 ... 192 against heights 32 ... 128, one to eight CTAs per environment) against the
    oracle on the same grid, one action step, both precisions, plus the sensor-grid Nusselt number."""
    import torch
    P = O.make_params(1e5, nx=nx, nz=nz, split_phy=False)
    states = [smooth_state(nx, nz, seed=s) for s in (5, 6, 7)]
    acts = np.random.default_rng(11).uniform(-1, 1, (3, 12)).astype(np.float32)
    obs_shape = (8, 32)
    sim = B.Sim2D(3, ra=1e5, dt_action=0.2, dt_solver=0.03, state_shape=(nz, nx), obs_shape=obs_shape, precision=precision)
    sim.reset_from_fields(np.concatenate([B.pack_fields(b[None], u[None], w[None]) for b, u, w in states]), project=False)
    obs, rew, nus, nuo, trunc, nan = sim.step(torch.from_numpy(acts).cuda())
    b, u, w = B.split_fields(sim.fields(), (nz, nx))
    assert not nan.any().item()
    for j, (b0, u0, w0) in enumerate(states):
        r = O.step(P, b0, u0, w0, acts[j].astype(np.float64), O.substep_schedule(0.2, 0.03))
        assert rel(b[j], r["b"]) < tol and rel(u[j], r["u"]) < tol and rel(w[j], r["w"]) < tol, (j, rel(b[j], r["b"]), rel(u[j], r["u"]), rel(w[j], r["w"]))
        ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"], obs_shape)
        assert nuo[j].item() == pytest.approx(no, abs=1e-8 if precision == 64 else 2e-3)
        assert nus[j].item() == pytest.approx(ns, abs=1e-8 if precision == 64 else 2e-3)
    sim.close()


def test_config3_full_batch_properties(B):
    """Config 3 at bench size (1056 envs = 32 per resident cluster): size-independent properties after an action step —
    discretely solenoidal, impenetrable walls, zero net vertical mass flux, bounded temperature, replica-exactness."""
    import torch
    from rbc_gym_b200.envs.rbc2d import noise_initial_fields
    n = 1056
    rng = np.random.default_rng(1)
    base = np.concatenate([noise_initial_fields(rng, (NZ, NX), kick=0.05) for _ in range(4)])
    sim = B.Sim2D(n, ra=RA, dt_action=0.3, dt_solver=DTS, state_shape=(NZ, NX), precision=32)
    sim.reset_from_fields(np.tile(base, (n // 4, 1)), project=True)
    acts = torch.from_numpy(np.tile(np.random.default_rng(2).uniform(-1, 1, (4, 12)).astype(np.float32), (n // 4, 1))).cuda()
    for _ in range(2):
        obs, rew, nus, nuo, trunc, nan = sim.step(acts)
    assert not nan.any().item() and torch.isfinite(rew).all().item()
    f = sim.fields()
    b, u, w = B.split_fields(f, (NZ, NX))
    dx, dz = 2 * np.pi / NX, 2.0 / NZ
    div = (np.roll(u, -1, axis=-1) - u) / dx + (w[:, 1:] - w[:, :-1]) / dz
    assert np.abs(div).max() < 1e-4
    assert np.all(w[:, 0] == 0) and np.all(w[:, -1] == 0)
    assert np.abs(w.sum(axis=-1)).max() < 1e-3                        # no net mass flux through any horizontal plane
    assert b.min() > 0.2 and b.max() < 2.8
    for i in range(4, n):
        assert np.array_equal(f[i], f[i % 4])
    sim.close()


def test_host_step_in_overlapped_chunks_on_the_cluster_grid(B):
    """Three waves of resident clusters -> three chunks; host-buffer step == device-buffer step."""
    import torch
    n = 90
    base = [smooth_state(NX, NZ, seed=s) for s in (0, 1, 2)]
    fields = np.concatenate([B.pack_fields(*(a[None] for a in base[i % 3])) for i in range(n)])
    acts = np.random.default_rng(2).uniform(-1, 1, (n, 12)).astype(np.float32)
    sims = [B.Sim2D(n, ra=RA, dt_action=0.045, dt_solver=DTS, state_shape=(NZ, NX)) for _ in range(2)]
    for sim in sims:
        sim.reset_from_fields(fields, project=False)
    obs, rew, nus, nuo, *_ = sims[0].step(torch.from_numpy(acts).cuda())
    out = sims[1].step_host(acts)
    np.testing.assert_array_equal(out["obs"], obs.cpu().numpy())
    np.testing.assert_array_equal(out["reward"], rew.cpu().numpy())
    np.testing.assert_array_equal(out["nu_obs"], nuo.cpu().numpy())
    assert np.array_equal(sims[0].fields(), sims[1].fields())
    for sim in sims:
        sim.close()


_JITTER_SCRIPT = r"""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, {root!r})
from rbc_gym_b200 import backend
from tests.gridstates import smooth_state
n = 6
sim = backend.Sim2D(n, ra=1e6, dt_action=0.045, dt_solver=0.015, state_shape=(128, 192), precision=32)
st = [smooth_state(192, 128, seed=s) for s in range(n)]
sim.reset_from_fields(backend.pack_fields(*[np.stack([s[k] for s in st]) for k in range(3)]), project=True)
g = torch.Generator(device="cuda").manual_seed(1)
for it in range(100):
    sim.step(torch.rand((n, 12), device="cuda", generator=g) * 2 - 1)
f = sim.fields()
assert np.isfinite(f).all()
print("DIGEST", hashlib.sha256(f.tobytes()).hexdigest(), sim.launch_info()["grid"])
"""


def _run_variant(lib, cluster_env):
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    env.pop("RBC_B200_LIB", None)
    env.pop("RBC_B200_CLUSTER", None)
    if lib is not None:
        env["RBC_B200_LIB"] = str(lib)
    if cluster_env is not None:
        env["RBC_B200_CLUSTER"] = cluster_env
    r = subprocess.run([sys.executable, "-c", _JITTER_SCRIPT.format(root=str(ROOT))], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1].split()
    return line[1], int(line[2])


def test_async_halo_protocol_survives_jitter():
    """Stress test of the `st.async` + mbarrier protocol of the cluster kernel (no cluster barrier and no fence in the stage loop;
    its write-after-read safety is argued in rbc2dx_core.h).  A second build of the library (`-DRBX_JITTER`, build.build_jitter)
    sleeps for pseudo-random, clock-seeded times in front of every remote store and every mbarrier wait, and makes one rank of the
    cluster the laggard for ~30 us at a time.  100 action steps (900 RK3 stages with all four exchange channels) of six
    192 x 128 environments must come out BITWISE identical: at CL = 4 to the normal build, at CL = 8 (an instantiation only the
    stress build has) across three runs with three different delay patterns."""
    from rbc_gym_b200 import build
    jit = build.LIB_JITTER
    if not jit.exists():
        jit = build.build_jitter()
    ref, grid = _run_variant(None, None)
    assert grid == 24                                          # 6 clusters x 4 CTAs
    for _ in range(2):
        assert _run_variant(jit, None)[0] == ref
    runs8 = [_run_variant(jit, "8") for _ in range(3)]
    assert runs8[0][1] == 48 and len({d for d, _ in runs8}) == 1
    assert runs8[0][0] != ref                                  # a different decomposition rounds differently: it really was another kernel


def test_cfl_guard_makes_fp32_noise_initialisation_robust_on_config3():
    """Config 3's own reset path in the throughput mode: 4096 environments, 192 x 128, Ra = 1e6, fp32, the reference's noise
    initialisation (kick 0.01), 60 action steps of dt = 1 with zero action.  The first plume burst overshoots to CFL ~ 1.6, where the
    scheme is linearly unstable: without the guard about one fp32 environment in a thousand ends with NaNs;
    with it (default in fp32) none may, the guard must have acted on a minority of RK3 steps only, and the Nusselt statistics must
    match an fp64 ensemble run without any guard (the reference's fixed time step)."""
    import torch
    from rbc_gym_b200 import backend
    n = 4096
    sim = backend.Sim2D(n, ra=RA, dt_action=1.0, dt_solver=DTS, state_shape=(NZ, NX), precision=32)
    gen = torch.Generator(device="cuda").manual_seed(11)
    for lo in range(0, n, 1024):                                         # noise fields in slices: 1024 x 73 920 doubles at a time
        sim.noise_reset(torch.arange(lo, lo + 1024, dtype=torch.int32), kick=0.01, generator=gen)
    zero = torch.zeros((n, 12), device="cuda")
    nu32 = []
    for step in range(60):
        _, _, nus, _, _, nan = sim.step(zero)
        nu32.append(nus.clone())
    assert int(nan.sum()) == 0
    nu32 = torch.stack(nu32).cpu().numpy()
    assert np.isfinite(nu32).all()
    ev = sim.cfl_events()
    rk3_steps = 60 * sim.nsub
    assert ev.max() > 0 and ev.mean() < 0.05 * rk3_steps, (ev.max(), ev.mean())      # it acted, and only during the burst
    sim.close()
    m = 256
    ref = backend.Sim2D(m, ra=RA, dt_action=1.0, dt_solver=DTS, state_shape=(NZ, NX), precision=64)
    ref.noise_reset(kick=0.01, generator=torch.Generator(device="cuda").manual_seed(12))
    zero = torch.zeros((m, 12), device="cuda")
    nu64 = []
    for step in range(60):
        _, _, nus, _, _, nan = ref.step(zero)
        nu64.append(nus.clone())
    assert int(nan.sum()) == 0 and ref.cfl_events().max() == 0              # the validation mode keeps the reference's fixed dt
    nu64 = torch.stack(nu64).cpu().numpy()
    ref.close()
    # developed phase (steps 40..59): ensemble means within three standard errors, ensemble widths alike
    a, b = nu32[40:].mean(axis=0), nu64[40:].mean(axis=0)
    se = np.sqrt(a.var() / len(a) + b.var() / len(b))
    assert abs(a.mean() - b.mean()) < 3 * se, (a.mean(), b.mean(), se)
    assert 0.8 * b.std() < a.std() < 1.25 * b.std()
    # the burst itself (largest ensemble-mean Nusselt number and when it happens)
    assert abs(int(nu32.mean(axis=1).argmax()) - int(nu64.mean(axis=1).argmax())) <= 1
    assert nu32.mean(axis=1).max() == pytest.approx(nu64.mean(axis=1).max(), rel=0.05)
