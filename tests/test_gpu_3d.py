"""3D environment on a GPU: parity of the CUDA path (through the C ABI) with the 3D oracle, API surface."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle3d as O3  # noqa: E402
from tests.test_oracle3d import random_state  # noqa: E402


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("precision,tol", [(64, 1e-10), (32, 1e-5)])
def test_3d_action_step_matches_oracle(precision, tol, split):
    """split=False + fp32 is the tiled throughput kernel; the other three combinations run the global-memory kernel."""
    import torch
    from rbc_gym_b200 import backend
    P = O3.make_params(5e3, split_phy=split)
    states = [random_state(P, s) for s in (1, 2, 3)]
    acts = np.random.default_rng(9).uniform(-1, 1, (3, 8, 8)).astype(np.float32)
    sim = backend.Sim3D(3, ra=5e3, precision=precision, split=split)
    sim.reset_from_fields(np.concatenate([backend.pack_fields3(*(x[None] for x in st)) for st in states]), project=False)
    obs, rew, nu, trunc, nan = sim.step(torch.from_numpy(acts).cuda())
    b, u, v, w = backend.split_fields3(sim.fields())
    t, step = sim.info()
    assert np.all(t == 0.5) and np.all(step == 2) and not nan.any().item() and not trunc.any().item()
    for j, st in enumerate(states):
        r = O3.step(P, *st, acts[j].astype(np.float64), O3.substep_schedule())
        assert rel(b[j], r["b"]) < tol and rel(u[j], r["u"]) < tol and rel(v[j], r["v"]) < tol and rel(w[j], r["w"]) < tol
        ref_nu = O3.nusselt(P, r["b"], r["w"])
        assert nu[j].item() == pytest.approx(ref_nu, rel=1e-9 if precision == 64 else 2e-5)
        assert rew[j].item() == pytest.approx(-ref_nu, rel=2e-5)
        np.testing.assert_allclose(obs[j].cpu().numpy(), np.stack([r["b"], r["u"], r["v"], r["w"][:-1]]), rtol=0,
                                   atol=2e-7 if precision == 64 else 2e-5)
    sim.close()


def test_3d_batch_larger_than_grid_is_replica_exact():
    import torch
    from rbc_gym_b200 import backend
    P = O3.make_params(2500)
    base = [backend.pack_fields3(*(x[None] for x in random_state(P, s))) for s in (4, 5)]
    n = 301
    sim = backend.Sim3D(n, ra=2500, heater_duration=0.02, precision=32)
    sim.reset_from_fields(np.concatenate([base[i % 2] for i in range(n)]), project=False)
    acts = torch.zeros(n, 8, 8, device="cuda"); acts[1::2, 3, 4] = 1.0
    obs, rew, *_ = sim.step(acts)
    f = sim.fields()
    for i in range(2, n):
        assert np.array_equal(f[i], f[i % 2])
    assert torch.equal(rew[2:n:2], rew[0].expand((n - 1) // 2)) and not torch.equal(obs[0], obs[1])
    sim.close()


def test_3d_host_step_in_overlapped_chunks_equals_device_step():
    """rbc3d_step_host: chunks of whole waves, observations (262 KB per environment) copied out on a second stream while the
    next chunk computes; same outputs as the device-buffer step."""
    import torch
    from rbc_gym_b200 import backend
    P = O3.make_params(2500)
    base = [backend.pack_fields3(*(x[None] for x in random_state(P, s))) for s in (4, 5, 6)]
    n = 450                                                   # 4 waves of 148 CTAs
    fields = np.concatenate([base[i % 3] for i in range(n)])
    acts = np.random.default_rng(3).uniform(-1, 1, (n, 8, 8)).astype(np.float32)
    sims = [backend.Sim3D(n, ra=2500, heater_duration=0.02, precision=32) for _ in range(2)]
    for sim in sims:
        sim.reset_from_fields(fields, project=False)
    obs, rew, nu, tr, nan = sims[0].step(torch.from_numpy(acts).cuda())
    out = {"obs": np.zeros((n, 4, 16, 32, 32), np.float32), "reward": np.zeros(n, np.float32), "nusselt": np.zeros(n),
           "truncated": np.zeros(n, np.int32), "nan": np.zeros(n, np.int32)}
    sims[1].step_host(acts, out)
    np.testing.assert_array_equal(out["obs"], obs.cpu().numpy())
    np.testing.assert_array_equal(out["reward"], rew.cpu().numpy())
    np.testing.assert_array_equal(out["nusselt"], nu.cpu().numpy())
    assert np.array_equal(sims[0].fields(), sims[1].fields())
    for sim in sims:
        sim.close()


def test_3d_env_reference_api():
    import rbc_gym_b200 as R
    env = R.make(R.ENV_ID_3D, rayleigh_number=2500)
    assert env.action_space.shape == (8, 8) and env.observation_space.shape == (4, 16, 32, 32)
    obs, info = env.reset(seed=1)                                   # noise init + set! projection
    assert obs.shape == (4, 16, 32, 32) and obs.dtype == np.float32 and set(info) == {"t", "step", "nusselt"}
    assert info["t"] == 0.0 and info["step"] == 1 and abs(info["nusselt"] - 1) < 0.05
    P = O3.make_params(2500)
    f = env.sim.fields()
    from rbc_gym_b200 import backend
    b, u, v, w = backend.split_fields3(f)
    assert np.abs(O3.divergence(P, u[0], v[0], w[0])).max() < 1e-12
    a = env.action_space.sample()
    obs2, reward, term, trunc, info2 = env.step(a)
    r = O3.step(P, b[0], u[0], v[0], w[0], a.astype(np.float64), O3.substep_schedule())
    assert reward == pytest.approx(-O3.nusselt(P, r["b"], r["w"]), rel=1e-9)
    assert info2["t"] == 0.5 and info2["step"] == 2 and term is False and trunc is False
    np.testing.assert_allclose(obs2[0], r["b"], atol=2e-7)
    with pytest.raises(RuntimeError, match="Action size does not match"):
        env.step(np.zeros((4, 4), np.float32))                     # rbc_sim3D.jl:115-117
    env.close()


def test_3d_convection_onset_statistics():
    """Physics check against the reference's flowstats (BASELINE.md §4.2, 64x64x32 runs of the real Julia sim).  With
    H = 2 and g*alpha*dT = 1 the physical Rayleigh number is 8 Ra, so Ra = 150 (1200 < 1708) must decay to conduction
    (Nu -> 1), while Ra = 500 and Ra = 4000 must convect with Nu near the reference's 1.371 +- 0.006 and 2.123 +- 0.032
    (the registered 32x32x16 grid is coarser than the flowstats runs, hence the generous bands)."""
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import noise_initial_fields_3d
    late = {}
    for ra in (150.0, 500.0, 4000.0):
        sim = backend.Sim3D(2, ra=ra, heater_duration=1.0, precision=32)
        rng = np.random.default_rng(7)
        sim.reset_from_fields(np.concatenate([noise_initial_fields_3d(rng, kick=0.05) for _ in range(2)]), project=True)
        zero = torch.zeros(2, 8, 8, device="cuda")
        nus = []
        for it in range(150):                                      # 150 free-fall times
            _, _, nu, _, nan = sim.step(zero, want_obs=False)
            nus.append(nu.cpu().numpy().copy())
        assert not nan.any().item()
        late[ra] = np.array(nus)[-40:].mean()
        sim.close()
    assert abs(late[150.0] - 1.0) < 5e-3, late
    assert 1.25 < late[500.0] < 1.50, late
    assert 1.85 < late[4000.0] < 2.45, late


@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_vector_env_3d_autoreset_and_batch_consistency(mode):
    """RBCVectorEnv3D (the on-device replacement of run_sarl.py's SubprocVecEnv): shapes, reward = -Nu, truncation and
    autoreset semantics, and every env of the batch behaves like the single-env class."""
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv3D
    env = RBCVectorEnv3D(6, rayleigh_number=2500, heater_duration=0.125, episode_length=1.0, autoreset_mode=mode, seed=3)
    obs, info = env.reset(seed=3)
    assert obs.shape == (6, 4, 16, 32, 32) and info["nusselt"].shape == (6,)
    acts = torch.rand((6, 8, 8), device="cuda") * 2 - 1
    obs, rew, term, trunc, info = env.step(acts)                      # t = 0.5
    assert torch.allclose(rew.double(), -info["nusselt"], atol=1e-6) and not trunc.any() and not term.any()
    obs1 = obs.clone()
    obs, rew, term, trunc, info = env.step(acts)                      # t = 1.0 >= episode_length: truncation
    assert trunc.all()
    t, step = env.sim.info()
    if mode == "same_step":
        assert "final_obs" in info and not torch.equal(info["final_obs"], obs)
        assert np.all(t == 0.0) and np.all(step == 1)                 # already reset
    else:
        assert np.all(t == 1.0)
        obs, rew, term, trunc, info = env.step(acts)                  # this call only resets
        t, step = env.sim.info()
        assert np.all(t == 0.0) and torch.all(rew == 0) and not trunc.any()
    with pytest.raises(RuntimeError, match="Action size"):
        env.step(torch.zeros((6, 4, 4), device="cuda"))
    assert not torch.equal(obs1[0], obs1[1])
    env.close()


def test_3d_flowstats_protocol_reproduces_julia_series_at_ra500():
    """The reference's flowstats protocol (zero action, kick 0.01, heater_duration 0.25 -> one sample per time unit,
    dt_solver 0.005; `experiments/flowstats/flowstats_ra.py:27-36`) at Ra = 500, where the registered 32 x 32 x 16 grid
    resolves the flow as well as the reference's 64 x 64 x 32 run.  Numbers extracted from the Julia-produced
    `flowstats_ra.pkl`: growth rate of (Nu - 1) in the linear phase 0.3164 per time unit, first-burst peak 1.417, mean
    over samples 50..99 1.4035 (std 0.0145).  At Ra = 4000 / 16000 the half-resolution grid grows 8 % / 11 % slower
    (0.70 / 0.74 vs 0.767 / 0.839) — the matched-resolution comparison runs on the GPU in test_gpu_3d_generic.py."""
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import noise_initial_fields_3d

    def growth_rate(nu, lo=1e-3, hi=5e-2):
        """slope of log(Nu - 1) over the first run of samples with lo < Nu - 1 < hi"""
        e = np.asarray(nu) - 1.0
        idx = [i for i in range(min(len(e), 80)) if lo < e[i] < hi]
        run = [idx[0]]
        for i in idx[1:]:
            if i != run[-1] + 1:
                break
            run.append(i)
        return float(np.polyfit(np.array(run), np.log(e[run]), 1)[0])

    sim = backend.Sim3D(4, ra=500.0, heater_duration=0.25, dt_solver=0.005, precision=32)
    rng = np.random.default_rng(42)
    sim.reset_from_fields(np.concatenate([noise_initial_fields_3d(rng, kick=0.01) for _ in range(4)]), project=True)
    zero = torch.zeros(4, 8, 8, device="cuda")
    nus = []
    for _ in range(100):
        _, _, nu, _, nan = sim.step(zero, want_obs=False)
        nus.append(nu.cpu().numpy().copy())
    assert not nan.any().item()
    nus = np.array(nus)
    for e in range(4):
        assert growth_rate(nus[:, e]) == pytest.approx(0.3164, rel=0.03)
        assert nus[:60, e].max() == pytest.approx(1.417, rel=0.02)
        assert nus[50:, e].mean() == pytest.approx(1.4035, abs=0.02)
    sim.close()


def test_sb3_vecenv_facade_over_the_3d_batch():
    """What `experiments/run_sarl.py:130-153` builds (SubprocVecEnv of Monitor + RBCNormalizeObservation 3D envs), as one
    on-device batch behind the SB3 VecEnv contract: numpy in/out, reset inside the truncating step, terminal_observation,
    TimeLimit.truncated, Monitor's episode record, the `nusselt` key NusseltCallback reads."""
    from rbc_gym_b200 import wrappers as W
    from rbc_gym_b200.envs import RBCSB3VecEnv, RBCVectorEnv3D
    venv = RBCVectorEnv3D(4, rayleigh_number=2500, heater_duration=0.125, episode_length=1.0, autoreset_mode="same_step", seed=1)
    env = RBCSB3VecEnv(venv)
    assert env.num_envs == 4 and env.observation_space.shape == (4, 16, 32, 32) and env.action_space.shape == (8, 8)
    assert float(env.observation_space.high.max()) == pytest.approx(1.3)
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (4, 4, 16, 32, 32) and obs.dtype == np.float32
    raw = venv.sim.obs.cpu().numpy()
    ref = W.normalize_observation(raw[0].copy(), 0.9, u_limit=W.RBCNormalizeObservation._get_u_limit_3d(2500))
    np.testing.assert_allclose(obs[0], ref, rtol=1e-5, atol=1e-6)
    acts = np.random.default_rng(0).uniform(-1, 1, (4, 8, 8)).astype(np.float32)
    obs, rew, dones, infos = env.step(acts)
    assert rew.shape == (4,) and not dones.any() and set(infos[0]) == {"nusselt", "t", "step"} and infos[0]["nusselt"] == pytest.approx(-rew[0], rel=1e-5)
    ret = rew.astype(np.float64).copy()
    env.step_async(acts)
    obs, rew, dones, infos = env.step_wait()                               # t = 1.0 >= episode_length
    ret += rew
    assert dones.all() and all(i["TimeLimit.truncated"] for i in infos)
    assert infos[2]["terminal_observation"].shape == (4, 16, 32, 32) and not np.array_equal(infos[2]["terminal_observation"], obs[2])
    assert infos[1]["episode"]["l"] == 2 and infos[1]["episode"]["r"] == pytest.approx(ret[1], rel=1e-5)
    assert all(i["t"] == 0.0 and i["step"] == 1 for i in infos)            # already reset: the returned obs is the reset observation
    assert env.get_attr("ra") == [2500] * 4 and env.env_is_wrapped(object) == [False] * 4
    env.close()


def test_3d_rgb_array_render_volume_rendering():
    """`render("rgb_array")` of the 3D env (`rbc3D.py:247-318`: PyVista volume rendering of the temperature, turbo colormap, clim =
    temperature_difference, 800 x 608): ray-marched on the device.  The conduction state must show the hot (red) plate at the
    bottom of the box and the cold (blue) one on top; a plume changes the picture; the batch renderer agrees with the env class."""
    import torch
    from pathlib import Path
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import RayleighBenardConvection3DEnv
    sim = backend.Sim3D(2, ra=2500, precision=32)
    z = (np.arange(16) + 0.5) / 8
    b = np.repeat((2 - z / 2)[None, :, None, None], 2, 0) * np.ones((2, 16, 32, 32))
    b[1, :, 10:16, 4:10] = 2.0                                              # a hot column in environment 1
    zeros = np.zeros((2, 16, 32, 32))
    sim.reset_from_fields(backend.pack_fields3(b, zeros, zeros, np.zeros((2, 17, 32, 32))), project=False)
    img = sim.render_rgb().cpu().numpy()
    assert img.shape == (2, 608, 800, 3) and img.dtype == np.uint8
    assert (img[0, :20] == 255).all() and (img[0, :, :8] == 255).all()       # white background around the box
    cols = img[0, :, 380:420].astype(int)
    inside = np.where((cols != 255).any(axis=(1, 2)))[0]
    top, bottom = cols[inside[:40]].mean(axis=(0, 1)), cols[inside[-40:]].mean(axis=(0, 1))
    assert top[2] > top[0] + 20 and bottom[0] > bottom[2] + 20               # cold (blue) lid above, hot (red) plate below
    assert np.abs(img[0].astype(int) - img[1].astype(int)).max() > 40        # the plume is visible
    env = RayleighBenardConvection3DEnv(rayleigh_number=2500, render_mode="rgb_array")
    env.reset(seed=0)
    frame = env.render()
    assert frame.shape == (608, 800, 3) and frame.std() > 10
    out = Path(__file__).resolve().parent.parent / "gpurun_out"
    if out.exists():
        with open(out / "render3d.ppm", "wb") as f:
            f.write(b"P6 800 608 255\n" + img[1].tobytes())
    env.close(); sim.close()
