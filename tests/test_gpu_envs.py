"""API-surface tests of the environment classes on a GPU (SURVEY §8a a8/a11/a12/a14, §8b "surface kept")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from pathlib import Path  # noqa: E402

from oracle import oracle as O  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
CKPT = str(ROOT / "data/checkpoints/train/ckpt_ra100000.h5")


def test_single_env_reference_api(ckpt_ra1e5):
    import rbc_gym_b200 as R
    env = R.make(rayleigh_number=100_000, heater_duration=0.3, checkpoint=CKPT, checkpoint_idx=7, render_mode="rgb_array")
    assert env.action_space.shape == (12,) and env.observation_space.shape == (3, 8, 48)
    assert env.observation_space.low[0, 0, 0] == 1 and env.observation_space.high[0, 0, 0] == np.float32(2.75)
    assert env.episode_steps == 1000 and env.temperature_difference == [1, 2] and env.state_shape == [64, 96]
    obs, info = env.reset(seed=3)
    assert obs.shape == (3, 8, 48) and obs.dtype == np.float32
    assert set(info) == {"t", "step", "nusselt_state", "nusselt_obs", "state"}
    assert info["t"] == 0.0 and info["step"] == 1 and info["state"].shape == (3, 64, 96) and info["state"].dtype == np.float32
    c = ckpt_ra1e5
    np.testing.assert_array_equal(info["state"], O.state_channels(c.b[7], c.u[7], c.w[7]).astype(np.float32))
    P = O.make_params(1e5)
    ns, no = O.nusselt_state_obs(P, c.b[7], c.u[7], c.w[7])
    assert info["nusselt_state"] == pytest.approx(ns, rel=1e-12) and info["nusselt_obs"] == pytest.approx(no, rel=1e-12)
    a = np.linspace(-1, 1, 12).astype(np.float32)
    obs, reward, terminated, truncated, info = env.step(a)
    r = O.step(P, c.b[7], c.u[7], c.w[7], a.astype(np.float64), O.substep_schedule(0.3))
    ns, no = O.nusselt_state_obs(P, r["b"], r["u"], r["w"])
    assert isinstance(reward, float) and reward == pytest.approx(-no, rel=1e-9)
    assert terminated is False and truncated is False and info["t"] == pytest.approx(0.3) and info["step"] == 2
    np.testing.assert_allclose(info["state"][0], r["b"], atol=2e-7)
    img = env.render()
    assert img.shape == (64, 96, 3) and img.dtype == np.uint8
    with pytest.warns(UserWarning):
        env.step(None)                                                     # zero action + warning (rbc2D.py:164-166)
    env.close()


def test_single_env_truncation_errors_and_noise_init():
    import rbc_gym_b200 as R
    env = R.make(rayleigh_number=10_000, heater_duration=0.06, episode_length=0.1)
    with pytest.raises(RuntimeError):
        env.step(np.zeros(12, np.float32))                                 # not initialised (rbc_sim2D_api.jl:79-81)
    obs, info = env.reset(seed=42)                                         # noise init + projection (set!)
    st = info["state"]
    assert 1.0 <= st[0].min() and st[0].max() <= 2.0 and abs(st[0].mean() - 1.5) < 0.01
    obs2, info2 = env.reset(seed=42)
    np.testing.assert_array_equal(info2["state"], st)                      # seeded reset is reproducible
    _, _, _, tr1, _ = env.step(env.action_space.sample())
    _, _, _, tr2, i2 = env.step(env.action_space.sample())
    assert (tr1, tr2) == (False, True) and i2["t"] >= 0.1
    env.close()
    bad = R.make(checkpoint="/nonexistent/ckpt.h5")
    with pytest.raises(FileNotFoundError):
        bad.reset()
    bad.close()


def test_single_env_nan_raises_like_reference(ckpt_ra1e5):
    import rbc_gym_b200 as R
    from rbc_gym_b200 import backend
    env = R.make(rayleigh_number=100_000, heater_duration=0.03, checkpoint=CKPT, checkpoint_idx=0, precision=32)
    env.reset()
    f = backend.pack_fields(ckpt_ra1e5.b[:1], ckpt_ra1e5.u[:1], ckpt_ra1e5.w[:1])
    f[0, 77] = np.nan
    env.sim.reset_from_fields(f, project=False)
    with pytest.raises(RuntimeError, match="probably NaN values"):
        env.step(np.zeros(12, np.float32))
    env.close()


def test_python_wrappers_compose_like_run_wrapped():
    import rbc_gym_b200 as R
    from rbc_gym_b200 import wrappers as W
    base = R.make(rayleigh_number=100_000, heater_duration=0.3, checkpoint=CKPT, checkpoint_idx=2)
    env = W.RBCRewardShaping(W.RBCNormalizeReward(W.RBCNormalizeObservation(base, heater_limit=base.unwrapped.heater_limit)), 0.1)
    obs, info = env.reset(seed=0)
    assert env.observation_space.shape == (3, 8, 48) and abs(obs).max() <= 1.3
    obs, reward, term, trunc, info = env.step(np.zeros(12, np.float32))
    assert "cell_dist" in info and 0 <= info["cell_dist"] <= np.pi
    raw = -info["nusselt_obs"]
    assert reward == pytest.approx(W.shape_reward(W.normalize_reward(raw, 100_000), info["cell_dist"], 0.1), rel=1e-12)
    env.close()


def test_fused_wrappers_on_device_match_python(ckpt_ra1e5):
    import torch
    from rbc_gym_b200 import backend, wrappers as W
    eps = [0, 7, 16, 3, 11]
    acts = np.random.default_rng(2).uniform(-1, 1, (5, 12)).astype(np.float32)
    sims = []
    for fused in (False, True):
        s = backend.Sim2D(5, ra=1e5, dt_action=0.3, precision=32)
        s.load_checkpoints(ckpt_ra1e5)
        s.reset_from_checkpoints(torch.tensor(eps, dtype=torch.int32))
        if fused:
            s.set_wrappers(normalize_obs=True, normalize_reward=True, shaping_weight=0.1)
        s.step(torch.from_numpy(acts).cuda())
        sims.append(s)
    plain, fused = sims
    st = plain.get_state().cpu().numpy()
    cd = fused.cell_dist()
    for j in range(5):
        ref_cd = W.cell_distance(st[j])
        assert cd[j] == pytest.approx(ref_cd, abs=1e-12)
        r = W.shape_reward(W.normalize_reward(-plain.nu_obs[j].item(), 1e5), ref_cd, 0.1)
        assert fused.reward[j].item() == pytest.approx(r, rel=1e-6)
        np.testing.assert_allclose(fused.obs[j].cpu().numpy(), W.normalize_observation(plain.obs[j].cpu().numpy().copy(), 0.75),
                                   rtol=2e-6, atol=2e-7)
    plain.close(); fused.close()


@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_vector_env_autoreset_semantics(mode):
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv2D
    n = 6
    env = RBCVectorEnv2D(n, rayleigh_number=100_000, heater_duration=0.06, episode_length=0.1, checkpoint=CKPT,
                         autoreset_mode=mode, seed=5, precision=32)
    obs0, info0 = env.reset()
    assert obs0.shape == (n, 3, 8, 48) and obs0.is_cuda
    reset_obs = obs0.clone()
    a = torch.zeros(n, 12, device="cuda")
    o1, r1, te1, tr1, _ = env.step(a)
    assert not tr1.any() and not te1.any()
    o2, r2, te2, tr2, i2 = env.step(a)
    assert tr2.all()
    if mode == "same_step":
        assert "final_obs" in i2 and not torch.equal(i2["final_obs"], o2)
        # the returned observation is already the reset observation of the next episode (a new random checkpoint)
        t, step = env.sim.info()
        assert np.all(t == 0) and np.all(step == 1)
        o3, r3, _, tr3, _ = env.step(a)
        assert not tr3.any() and (r3 != 0).all()
    else:
        t, _ = env.sim.info()
        assert np.all(t > 0.1)
        o3, r3, _, tr3, _ = env.step(a)                  # this call only resets (gymnasium NEXT_STEP)
        assert not tr3.any() and (r3 == 0).all()
        t, step = env.sim.info()
        assert np.all(t == 0) and np.all(step == 1)
    # checkpoint draws are keyed by (seed, global env id, episode): the same env ids elsewhere draw the same episodes
    env2 = RBCVectorEnv2D(3, rayleigh_number=100_000, heater_duration=0.06, episode_length=0.1, checkpoint=CKPT,
                          autoreset_mode=mode, seed=5, precision=32, env_id_offset=3)
    o, _ = env2.reset()
    assert torch.equal(o, reset_obs[3:6])
    env.close(); env2.close()


def test_checkpoint_generator_reaches_reference_fixed_point(tmp_path):
    """Spin-up from noise at Ra = 1e4 (the generator of the reference's fixtures, `rbc_sim2D.jl:14-72`): the flow
    must settle on the reference's near-fixed point — the band spanned by its 40 Ra=1e4 checkpoints widened by
    the residual oscillation amplitude — and the file must read back bit-identically."""
    from rbc_gym_b200.checkpoints import simulate_2d_rb
    from rbc_gym_b200.h5lite import load_checkpoint_2d
    path, stats = simulate_2d_rb(tmp_path, seed=42, random_inits=4, ra=1e4, duration=600.0, precision=64)
    assert path.name == "ckpt_ra10000.h5"
    c = load_checkpoint_2d(path)
    assert c.num_episodes == 4 and c.start_seed == 42 and c.b.shape == (4, 64, 96) and c.w.shape == (4, 65, 96)
    P = O.make_params(1e4)
    for e in range(4):
        ns, no = O.nusselt_state_obs(P, c.b[e], c.u[e], c.w[e])
        assert ns == pytest.approx(stats["nu_state"][e], rel=1e-12)
        assert 3.997490 - 3e-4 <= ns <= 3.997789 + 3e-4, ns
        assert 4.192437 - 3e-4 <= no <= 4.192916 + 3e-4, no
        assert 0.0974366 - 2e-6 <= O.kinetic_energy(c.u[e], c.w[e]) <= 0.0974397 + 2e-6
    dx, dz = 2 * np.pi / 96, 2 / 64
    div = (np.roll(c.u, -1, axis=-1) - c.u) / dx + (c.w[:, 1:] - c.w[:, :-1]) / dz
    assert np.abs(div).max() < 1e-13


def test_device_render_matches_host_colormap(ckpt_ra1e5):
    """rbc2d_render_rgb_dev (whole batch on the device) against the host-side turbo map the single env renders with."""
    import torch
    from rbc_gym_b200.colormap import turbo
    from rbc_gym_b200.envs import RBCVectorEnv2D
    env = RBCVectorEnv2D(5, rayleigh_number=100_000, heater_duration=0.3, checkpoint=CKPT, precision=32)
    env.reset(seed=1)
    env.step(torch.rand((5, 12), device="cuda") * 2 - 1)
    img = env.render().cpu().numpy()
    assert img.shape == (5, 64, 96, 3) and img.dtype == np.uint8
    state = env.get_state().cpu().numpy()
    ref = turbo(state[:, 0, ::-1, :], vmin=1, vmax=2.75)                 # flipped: origin at the top left (rbc2D.py:239-241)
    assert np.abs(img.astype(int) - ref.astype(int)).max() <= 1           # fp32 vs fp64 polynomial: at most one level
    assert img[:, -1].mean() > img[:, 0].mean() - 255                     # sanity: image has structure
    assert img.std() > 10
    env.close()


def test_config1_run_2d_example(ckpt_ra1e4):
    """BASELINE config 1 = example/run_2D.py: the registered defaults (Ra=1e4, heater_duration 1.5) with pressure=True and the
    full grid as observation, random actions; fp64 env against the oracle in hydrostatic-split mode for two action steps."""
    import rbc_gym_b200 as R
    c = ckpt_ra1e4
    env = R.make(R.ENV_ID_2D, pressure=True, observation_shape=[64, 96], checkpoint=str(ROOT / "data/checkpoints/test/ckpt_ra10000.h5"),
                 checkpoint_idx=4)
    assert env.ra == 10_000 and env.heater_duration == 1.5 and env.episode_steps == 200
    assert env.observation_space.shape == (5, 64, 96) and env.action_space.shape == (12,)
    obs, info = env.reset(seed=0)
    assert obs.shape == (5, 64, 96) and info["state"].shape == (5, 64, 96)
    P = O.make_params(1e4, split_phy=True)
    b, u, w = c.b[4], c.u[4], c.w[4]
    env.action_space.seed(5)
    for n in range(2):
        a = env.action_space.sample()
        obs, reward, terminated, truncated, info = env.step(a)
        r = O.step(P, b, u, w, a.astype(np.float64), O.substep_schedule(1.5), want_pressure=True)
        b, u, w = r["b"], r["u"], r["w"]
        ref = np.stack([b, u, w[:-1], r["phy"], r["pnhs"]]).astype(np.float32)
        np.testing.assert_allclose(obs, ref, rtol=0, atol=2e-6)              # full-grid sensors: obs == state, 5 channels
        np.testing.assert_array_equal(obs, info["state"])
        ns, no = O.nusselt_state_obs(P, b, u, w, (64, 96))
        assert info["nusselt_obs"] == pytest.approx(no, rel=1e-9) and info["nusselt_state"] == pytest.approx(ns, rel=1e-9)
        assert reward == pytest.approx(-no, rel=1e-9) and info["t"] == pytest.approx(1.5 * (n + 1)) and not truncated
    env.close()


def test_full_episode_rollout_4096_envs_autoreset():
    """Config 2 as SURVEY 8d describes it: 4096 envs reset from the train checkpoints, U(-1,1) actions, one full episode of
    300 action steps (dt = 1) so that the auto-reset fires exactly once at the end; nothing may blow up on the way."""
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv2D
    n = 4096
    env = RBCVectorEnv2D(n, rayleigh_number=100_000, heater_duration=1.0, episode_length=300, checkpoint=CKPT, precision=32,
                         autoreset_mode="same_step", seed=11)
    obs, info = env.reset(seed=11)
    g = torch.Generator(device="cuda").manual_seed(1234)
    n_trunc, rew_sum, nu_max = 0, 0.0, 0.0
    for step in range(300):
        a = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
        obs, rew, term, trunc, info = env.step(a)
        n_trunc += int(trunc.sum().item())
        if step % 50 == 49 or step == 299:
            assert torch.isfinite(rew).all().item() and torch.isfinite(obs).all().item()
            rew_sum += float(rew.mean().item())
            nu_max = max(nu_max, float(info["nusselt_state"].max().item()))
    assert n_trunc == n                                           # every env truncated exactly once, at step 300
    t, stepc = env.sim.info()
    assert np.all(t == 0.0) and np.all(stepc == 1)                # SAME_STEP autoreset already happened
    assert "final_obs" in info and info["final_info"]["episode_return"].shape == (n,)
    assert -12.0 < rew_sum / 7 < -3.0 and nu_max < 40.0           # Ra=1e5 with random heating: Nu_obs ~ 5-9
    ret = info["final_info"]["episode_return"].cpu().numpy()
    assert np.all(np.isfinite(ret)) and -300 * 12 < ret.min() and ret.max() < -300 * 2
    env.close()


def test_device_noise_reset_matches_reference_initialisation():
    """`Sim2D.noise_reset`: initialize_model (rbc_sim2D.jl:163-171) drawn on the device + the set! projection, no host trip."""
    import torch
    from rbc_gym_b200 import backend
    sim = backend.Sim2D(64, ra=1e5, dt_action=0.3, precision=64)
    gen = torch.Generator(device="cuda").manual_seed(5)
    sim.noise_reset(kick=0.01, generator=gen)
    b, u, w = backend.split_fields(sim.fields())
    dx, dz = 2 * np.pi / 96, 2.0 / 64
    div = (np.roll(u, -1, axis=-1) - u) / dx + (w[:, 1:] - w[:, :-1]) / dz
    assert np.abs(div).max() < 1e-12 and np.all(w[:, 0] == 0) and np.all(w[:, -1] == 0)
    z = (np.arange(64) + 0.5) * dz
    resid = b - (1 + (2 - z)[None, :, None] / 2)
    assert b.min() >= 1.0 and b.max() <= 2.0 and abs(resid[:, 2:-2].std() - 0.01) < 5e-4       # kick 0.01 away from the clamp
    assert 0.004 < u.std() < 0.011 and not np.array_equal(b[0], b[1])
    t, step = sim.info()
    assert np.all(t == 0) and np.all(step == 1)
    ids = torch.tensor([3, 9], dtype=torch.int32)
    before = sim.fields()
    sim.noise_reset(ids, kick=0.01, generator=gen)                       # partial reset leaves the others untouched
    after = sim.fields()
    assert not np.array_equal(after[3], before[3]) and np.array_equal(after[4], before[4])
    sim.close()


@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_vector_env_nan_policy(mode):
    """One failed environment: nan_policy="raise" is the reference's RuntimeError; "reset" re-initialises only that env,
    reports it (truncated + info["nan_reset"], reward 0) and the rest of the batch carries on."""
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import RBCVectorEnv2D

    def poison(env, e):
        f = env.sim.fields()
        f[e, 1000] = np.nan
        env.sim.reset_from_fields(f[e:e + 1], env_ids=[e], project=False)

    env = RBCVectorEnv2D(6, rayleigh_number=100_000, heater_duration=0.15, checkpoint=CKPT, autoreset_mode=mode, nan_policy="raise")
    env.reset(seed=0)
    poison(env, 2)
    with pytest.raises(RuntimeError, match="probably NaN"):
        env.step(torch.zeros((6, 12), device="cuda"))
    env.close()
    env = RBCVectorEnv2D(6, rayleigh_number=100_000, heater_duration=0.15, checkpoint=CKPT, autoreset_mode=mode, nan_policy="reset")
    env.reset(seed=0)
    poison(env, 2)
    obs, rew, term, trunc, info = env.step(torch.zeros((6, 12), device="cuda"))
    assert info["nan_reset"].tolist() == [False, False, True, False, False, False]
    assert trunc.tolist() == [False, False, True, False, False, False] and rew[2].item() == 0 and torch.isfinite(obs).all()
    t, step = env.sim.info()
    assert t[2] == 0.0 and step[2] == 1 and np.all(np.delete(t, 2) == pytest.approx(0.15))
    obs, rew, term, trunc, info = env.step(torch.zeros((6, 12), device="cuda"))        # everybody steps normally again
    assert not trunc.any() and not info["nan_reset"].any() and torch.isfinite(rew).all() and (rew != 0).all()
    t, step = env.sim.info()
    assert t[2] == pytest.approx(0.15) and t[0] == pytest.approx(0.30)
    env.close()
