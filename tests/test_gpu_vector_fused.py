"""The vector environments step in ONE kernel launch with the auto-reset fused into it (`rbc2d_vec_step_dev` /
`rbc3d_vec_step_dev`).  Reference semantics = the same class with `fused=False`, which drives the resets from Python the way the
gymnasium / SB3 vector wrappers drive single environments (`example/run_vectorized.py:11-31`, `experiments/run_sarl.py:130-153`):
separate reset + observe launches, host reads of the flags.  The two must agree bitwise, step by step."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CKPT = str(ROOT / "data/checkpoints/train/ckpt_ra100000.h5")


def same(a, b):
    import torch
    return torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all())


@pytest.mark.parametrize("mode", ["next_step", "same_step", "disabled"])
@pytest.mark.parametrize("precision,pressure", [(32, False), (64, False), (32, True)])
def test_fused_step_equals_python_driven_autoreset_2d(mode, precision, pressure):
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv2D
    n = 300                                               # two waves of the persistent grid
    kw = dict(rayleigh_number=100_000, heater_duration=0.09, episode_length=0.27, checkpoint=CKPT, autoreset_mode=mode, seed=5,
              precision=precision, pressure=pressure, env_id_offset=1000)
    fused, plain = RBCVectorEnv2D(n, **kw), RBCVectorEnv2D(n, fused=False, **kw)
    of, inf = fused.reset()
    op, inp = plain.reset()
    assert torch.equal(of, op) and torch.equal(inf["nusselt_obs"], inp["nusselt_obs"])
    assert set(inf) >= {"t", "step", "nusselt_state", "nusselt_obs"} and (inf["step"] == 1).all() and (inf["t"] == 0).all()
    g = torch.Generator(device="cuda").manual_seed(3)
    n_trunc = 0
    for it in range(8):
        a = torch.rand((n, 12), device="cuda", generator=g) * 2 - 1
        launches = fused.sim.launch_info()["launches"]
        f, p = fused.step(a), plain.step(a)
        assert fused.sim.launch_info()["launches"] == launches + 1       # one kernel launch of ours per vector step
        for k in range(4):
            assert same(f[k], p[k]), (it, k)
        for k in ("nusselt_state", "nusselt_obs", "t", "step", "episode_return", "nan"):
            assert same(f[4][k], p[4][k]), (it, k)
        tr = p[3]
        if mode == "same_step" and tr.any():
            assert same(f[4]["final_obs"][tr], p[4]["final_obs"][tr])
            for k in ("nusselt_state", "nusselt_obs", "episode_return"):
                assert same(f[4]["final_info"][k][tr], p[4]["final_info"][k][tr]), (it, k)
        n_trunc += int(tr.sum())
        assert np.array_equal(fused.sim.fields(), plain.sim.fields())
    assert n_trunc >= 2 * n
    fused.close(); plain.close()


def test_fused_step_on_the_cluster_kernel_grid():
    """192 x 128 (thread-block-cluster kernel): bank of developed states, both modes, against the Python-driven path."""
    import torch
    from rbc_gym_b200.checkpoints import developed_states_2d
    from rbc_gym_b200.envs import RBCVectorEnv2D
    from rbc_gym_b200.h5lite import Checkpoint2D
    from tests.gridstates import smooth_state
    states = [smooth_state(192, 128, seed=s) for s in range(3)]
    bank = Checkpoint2D(np.stack([s[0] for s in states]), np.stack([s[1] for s in states]), np.stack([s[2] for s in states]), 3, 0)
    for mode in ("next_step", "same_step"):
        kw = dict(rayleigh_number=1e6, heater_duration=0.045, dt_solver=0.015, episode_length=0.09, state_shape=(128, 192),
                  autoreset_mode=mode, seed=2, precision=32)
        fused, plain = RBCVectorEnv2D(40, **kw), RBCVectorEnv2D(40, fused=False, **kw)
        for e in (fused, plain):
            e.sim.load_checkpoints(bank)
        of, _ = fused.reset(); op, _ = plain.reset()
        assert torch.equal(of, op)
        g = torch.Generator(device="cuda").manual_seed(1)
        for it in range(5):
            a = torch.rand((40, 12), device="cuda", generator=g) * 2 - 1
            f, p = fused.step(a), plain.step(a)
            for k in range(4):
                assert same(f[k], p[k]), (mode, it, k)
            assert same(f[4]["t"], p[4]["t"]) and same(f[4]["nusselt_obs"], p[4]["nusselt_obs"])
        assert np.array_equal(fused.sim.fields(), plain.sim.fields())
        fused.close(); plain.close()


@pytest.mark.parametrize("shape", [(16, 32, 32), (8, 16, 32), (16, 64, 64)])      # dedicated kernel; stage-streaming kernels (per-cell / tiled tendency)
@pytest.mark.parametrize("mode", ["next_step", "same_step"])
@pytest.mark.parametrize("precision", [32, 64])
def test_fused_step_equals_python_driven_autoreset_3d(mode, shape, precision):
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv3D
    rng = np.random.default_rng(0)
    n, nb = 20, 3
    kw = dict(rayleigh_number=2500, heater_duration=0.03, dt_solver=0.01, episode_length=0.2, autoreset_mode=mode, seed=9, precision=precision,
              state_shape=shape)
    fused, plain = RBCVectorEnv3D(n, **kw), RBCVectorEnv3D(n, fused=False, **kw)
    # a bank of projected noisy states
    fused.sim.noise_reset(kick=0.05, generator=torch.Generator(device="cuda").manual_seed(4))
    bank = fused.sim.fields()[:nb]
    for e in (fused, plain):
        e.sim.load_checkpoints(bank)
    of, inf = fused.reset(); op, inp = plain.reset()
    assert torch.equal(of, op) and set(inf) >= {"t", "step", "nusselt"}
    g = torch.Generator(device="cuda").manual_seed(3)
    n_trunc = 0
    for it in range(6):
        a = torch.rand((n, 8, 8), device="cuda", generator=g) * 2 - 1
        f, p = fused.step(a), plain.step(a)
        for k in range(4):
            assert same(f[k], p[k]), (it, k)
        for k in ("nusselt", "t", "step", "episode_return"):
            assert same(f[4][k], p[4][k]), (it, k)
        tr = p[3]
        if mode == "same_step" and tr.any():
            assert same(f[4]["final_obs"][tr], p[4]["final_obs"][tr]) and same(f[4]["final_info"]["nusselt"][tr], p[4]["final_info"]["nusselt"][tr])
        n_trunc += int(tr.sum())
    assert n_trunc >= n
    fused.close(); plain.close()


@pytest.mark.parametrize("shape", [(16, 32, 32), (8, 16, 32)])      # dedicated kernel epilogue; batch kernels of the stage-streaming path
@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_nan_reset_policy_3d(mode, shape):
    """`nan_policy="reset"` in 3D: the one environment whose fields went NaN is re-initialised from the bank inside the step (truncated,
    reward 0, clock 0, finite observation, counted), the others step on; the step after that everybody steps normally."""
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv3D
    n = 5
    env = RBCVectorEnv3D(n, rayleigh_number=2500, heater_duration=0.03, dt_solver=0.01, episode_length=10.0, autoreset_mode=mode, seed=2,
                         precision=32, state_shape=shape, nan_policy="reset")
    env.sim.noise_reset(kick=0.05, generator=torch.Generator(device="cuda").manual_seed(8))
    env.sim.load_checkpoints(env.sim.fields()[:2])
    env.reset()
    f = env.sim.fields()
    f[3, 777] = np.nan
    env.sim.reset_from_fields(f[3:4], env_ids=[3], project=False)
    a = torch.zeros((n, 8, 8), device="cuda")
    obs, rew, term, trunc, info = env.step(a)
    want = [False, False, False, True, False]
    assert info["nan_reset"].tolist() == want and trunc.tolist() == want
    assert rew[3].item() == 0 and torch.isfinite(obs).all() and (rew[[0, 1, 2, 4]] != 0).all()
    t, step = env.sim.info()
    dt = 0.03 * 4.0                                   # heater_duration in free-fall units of the height-2 domain
    assert t[3] == 0.0 and step[3] == 1 and np.allclose(np.delete(t, 3), dt)
    assert env.sim.vec_nan_count() == 1
    obs, rew, term, trunc, info = env.step(a)
    assert not trunc.any() and not info["nan_reset"].any() and torch.isfinite(rew).all() and (rew != 0).all()
    t, step = env.sim.info()
    assert t[3] == pytest.approx(dt) and t[0] == pytest.approx(2 * dt)
    env.close()


def test_info_state_and_deferred_nan_check():
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv2D
    env = RBCVectorEnv2D(6, rayleigh_number=100_000, heater_duration=0.15, checkpoint=CKPT, info_state=True, nan_policy="raise_deferred")
    obs, info = env.reset(seed=0)
    assert info["state"].shape == (6, 3, 64, 96)                                 # rbc2D.py:211, opt-in for the batch
    # the attribute surface of gymnasium.vector.VectorEnv that callers of gym.make_vec read
    assert env.num_envs == 6 and env.unwrapped is env and not env.closed and env.metadata["autoreset_mode"] == "next_step"
    assert env.single_observation_space.shape == (3, 8, 48) and env.observation_space.shape == (6, 3, 8, 48)
    assert env.single_action_space.shape == (12,) and env.action_space.shape == (6, 12) and env.render_mode == "rgb_array"
    a = torch.zeros((6, 12), device="cuda")
    obs, rew, term, trunc, info = env.step(a)
    assert torch.equal(info["state"][:, :, ::8, ::2], obs) and info["t"].tolist() == [0.15] * 6 and info["step"].tolist() == [2] * 6
    f = env.sim.fields()
    f[4, 77] = np.nan
    env.sim.reset_from_fields(f[4:5], env_ids=[4], project=False)
    env.step(a)                                                                  # the failing step itself does not synchronise ...
    with pytest.raises(RuntimeError, match="probably NaN"):
        for _ in range(6):                                                       # ... a later call raises the reference's error: the first one
            env.step(a)                                                          # whose predecessor's counter copy has completed (never blocks
    env.close()                                                                  # on the device; at most five copies are kept in flight)
    # once the device has caught up it is exactly the next call
    env = RBCVectorEnv2D(6, rayleigh_number=100_000, heater_duration=0.15, checkpoint=CKPT, nan_policy="raise_deferred")
    env.reset(seed=0)
    env.sim.reset_from_fields(f[4:5], env_ids=[4], project=False)
    env.step(a)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="probably NaN"):
        env.step(a)
    env.close()


def test_reset_indices_are_validated():
    import torch
    from rbc_gym_b200 import backend
    sim = backend.Sim2D(4, ra=1e5, dt_action=0.3)
    n = sim.load_checkpoints(CKPT)
    with pytest.raises(IndexError):
        sim.reset_from_checkpoints(torch.tensor([0, 1, n, 2], dtype=torch.int32))        # the reference: Julia BoundsError
    with pytest.raises(IndexError):
        sim.reset_from_checkpoints(torch.tensor([0], dtype=torch.int32), env_ids=torch.tensor([4], dtype=torch.int32))
    with pytest.raises(IndexError):
        sim.reset_from_fields(np.zeros((1, sim.nstate)), env_ids=[-1])
    sim.close()


@pytest.mark.parametrize("mode", ["next_step", "same_step"])
def test_host_buffer_vector_step_equals_device_vector_step(mode):
    """`rbc2d_vec_step_host` (chunked launches, copies overlapped on a second stream) against `rbc2d_vec_step_dev`."""
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv2D
    n = 700
    kw = dict(rayleigh_number=100_000, heater_duration=0.09, episode_length=0.27, checkpoint=CKPT, autoreset_mode=mode, seed=8)
    dev, host = RBCVectorEnv2D(n, **kw), RBCVectorEnv2D(n, **kw)
    dev.reset(); host.reset()
    out = host.alloc_host_outputs()
    rng = np.random.default_rng(2)
    for it in range(7):
        a = rng.uniform(-1, 1, (n, 12)).astype(np.float32)
        d = dev.step(torch.from_numpy(a).cuda())
        h = host.step_host(a, out)
        assert np.array_equal(d[0].cpu().numpy(), h[0]) and np.array_equal(d[1].cpu().numpy(), h[1]) and np.array_equal(d[3].cpu().numpy(), h[3])
        for k in ("nusselt_state", "nusselt_obs", "t", "step", "episode_return"):
            assert np.array_equal(d[4][k].cpu().numpy(), h[4][k]), (it, k)
        if mode == "same_step" and h[3].any():
            m = h[3]
            assert np.array_equal(d[4]["final_obs"].cpu().numpy()[m], h[4]["final_obs"][m])
            assert np.array_equal(d[4]["final_info"]["episode_return"].cpu().numpy()[m], h[4]["final_info"]["episode_return"][m])
    dev.close(); host.close()
