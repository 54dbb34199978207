"""environments shard across GPUs with no data-path collective, and "the test is bitwise equality with
single-GPU slices".  Needs two visible devices (skipped otherwise); the single-process form is enough because ranks never
communicate inside a step — each device simply owns a contiguous slice of the global batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_device_shards_equal_single_device_slices(ckpt_ra1e5):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rbc_gym_b200.envs import RBCVectorEnv2D
    from rbc_gym_b200.sharding import shard_range
    from tests.conftest import ROOT
    ckpt = str(ROOT / "data/checkpoints/train/ckpt_ra100000.h5")
    n = 300
    g = torch.Generator().manual_seed(3)
    acts = [torch.rand((n, 12), generator=g) * 2 - 1 for _ in range(3)]
    single = RBCVectorEnv2D(n, rayleigh_number=100_000, heater_duration=0.3, checkpoint=ckpt, device=0, seed=5)
    obs, _ = single.reset(seed=5)
    for a in acts:
        obs, rew, *_ = single.step(a.cuda(0))
    ref_fields, ref_rew = single.sim.fields(), rew.cpu().numpy()
    single.close()
    for rank in range(2):
        lo, hi = shard_range(n, 2, rank)
        env = RBCVectorEnv2D(hi - lo, rayleigh_number=100_000, heater_duration=0.3, checkpoint=ckpt, device=rank, seed=5, env_id_offset=lo)
        with torch.cuda.device(rank):
            env.reset(seed=5)
            for a in acts:
                obs, rew, *_ = env.step(a[lo:hi].cuda(rank))
            assert np.array_equal(env.sim.fields(), ref_fields[lo:hi])          # bitwise, including the checkpoint draws
            assert np.array_equal(rew.cpu().numpy(), ref_rew[lo:hi])
        env.close()
