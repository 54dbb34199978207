"""Our `RayleighBenardConvection2DEnv` (CUDA backend, fp64) against a golden episode produced by the REFERENCE'S OWN env
class (`tools/make_env_golden.py` executes `/root/reference/src/rbc_gym/envs/rbc2D.py` with the Julia module replaced by
an oracle-backed object that has the API and array layouts of `rbc_sim2D_api.jl`).  This pins everything above the
simulation boundary — spaces, shape reversal / transposes, channel order, reward sign, info keys, time accumulation and
truncation — to the reference's code rather than to a reading of it."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
G = np.load(ROOT / "tests/golden/env2d_reference_episode.npz")
CKPT = str(ROOT / "data/checkpoints/train/ckpt_ra100000.h5")


@pytest.mark.parametrize("tag,kw,steps", [("default", dict(), 4), ("pressure_full", dict(pressure=True, observation_shape=[64, 96]), 2)])
def test_env_class_reproduces_the_reference_episode(tag, kw, steps):
    from rbc_gym_b200.envs import RayleighBenardConvection2DEnv
    env = RayleighBenardConvection2DEnv(rayleigh_number=100_000, heater_duration=0.3, checkpoint=CKPT, checkpoint_idx=7,
                                        episode_length=0.9, precision=64, **kw)
    np.testing.assert_array_equal(env.observation_space.low, G[f"{tag}/obs_low"])
    np.testing.assert_array_equal(env.observation_space.high, G[f"{tag}/obs_high"])
    assert env.episode_steps == int(G[f"{tag}/episode_steps"])
    obs, info = env.reset(seed=3)
    ref_reset = G[f"{tag}/reset_obs"]
    np.testing.assert_allclose(obs if tag == "default" else obs[:, ::8, ::2], ref_reset, rtol=0, atol=1e-6)
    if tag == "default":
        np.testing.assert_array_equal(info["state"], G["default/reset_state"])
    t, step, nus, nuo = G[f"{tag}/reset_info"]
    assert info["t"] == t and info["step"] == step
    assert info["nusselt_state"] == pytest.approx(nus, rel=1e-12) and info["nusselt_obs"] == pytest.approx(nuo, rel=1e-12)
    acts = G[f"{tag}/actions"]
    for n in range(steps):
        obs, reward, terminated, truncated, info = env.step(acts[n])
        ref_reward, ref_term, ref_trunc, ref_t, ref_step, ref_nus, ref_nuo = G[f"{tag}/scalars{n}"]
        assert reward == pytest.approx(ref_reward, rel=1e-9) and terminated == bool(ref_term) and truncated == bool(ref_trunc)
        assert info["t"] == ref_t and info["step"] == ref_step                         # bitwise: same accumulation in double
        assert info["nusselt_state"] == pytest.approx(ref_nus, rel=1e-9) and info["nusselt_obs"] == pytest.approx(ref_nuo, rel=1e-9)
        assert set(info) == {"t", "step", "nusselt_state", "nusselt_obs", "state"}
        if f"{tag}/obs{n}" in G.files:
            np.testing.assert_allclose(obs, G[f"{tag}/obs{n}"], rtol=0, atol=2e-6)
        if f"{tag}/state{n}" in G.files:
            np.testing.assert_allclose(info["state"], G[f"{tag}/state{n}"], rtol=0, atol=2e-6)
            assert info["state"].dtype == np.float32 and info["state"].shape == G[f"{tag}/state{n}"].shape
    env.close()


def test_env3d_class_reproduces_the_reference_episode(tmp_path):
    """The 3D twin (`tools/make_env3d_golden.py`): our `RayleighBenardConvection3DEnv` against the reference's own 3D env class
    over an oracle-backed Julia stand-in — transposes `(4, Nx, Ny, Nz)` -> `(4, Nz, Ny, Nx)`, the un-transposed 8 x 8 action,
    free-fall time units, reward = -Nu, info keys, truncation."""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    from make_env3d_golden import initial_bank, RA, N_STEPS
    from rbc_gym_b200.envs import RayleighBenardConvection3DEnv
    from rbc_gym_b200.h5lite import write_checkpoint_3d
    G3 = np.load(ROOT / "tests/golden/env3d_reference_episode.npz")
    path = tmp_path / f"3D_ckpt_ra{RA}.h5"
    write_checkpoint_3d(path, *initial_bank(), start_seed=42)
    env = RayleighBenardConvection3DEnv(rayleigh_number=RA, checkpoint=str(path), checkpoint_idx=1, episode_length=1.0,
                                        heater_duration=0.125, precision=64)
    np.testing.assert_array_equal(env.observation_space.low[:, ::5, ::11, ::7], G3["obs_low_sample"])
    np.testing.assert_array_equal(env.observation_space.high[:, ::5, ::11, ::7], G3["obs_high_sample"])
    assert tuple(env.action_space.shape) == tuple(G3["action_shape"])
    obs, info = env.reset(seed=5)
    np.testing.assert_allclose(obs[:, ::2, ::4, ::4], G3["reset_obs_sample"], rtol=0, atol=1e-6)
    t, step, nu = G3["reset_info"]
    assert info["t"] == t and info["step"] == step and info["nusselt"] == pytest.approx(nu, rel=1e-10)
    acts = G3["actions"]
    for n in range(N_STEPS):
        obs, reward, terminated, truncated, info = env.step(acts[n])
        ref_reward, ref_term, ref_trunc, ref_t, ref_step, ref_nu = G3[f"scalars{n}"]
        assert reward == pytest.approx(ref_reward, rel=1e-9) and terminated == bool(ref_term) and truncated == bool(ref_trunc)
        assert info["t"] == ref_t and info["step"] == ref_step and info["nusselt"] == pytest.approx(ref_nu, rel=1e-9)
        assert set(info) == {"t", "step", "nusselt"}
        np.testing.assert_allclose(obs[:, ::2, ::4, ::4], G3[f"obs_sample{n}"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(obs.astype(np.float64).sum(axis=(1, 2, 3)), G3[f"obs_sum{n}"], rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(obs[0, 0], G3["obs_last_bottom_level"], rtol=0, atol=2e-6)
    env.close()
