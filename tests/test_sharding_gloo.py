"""World-size-2 CPU (gloo) test of the N>1 host logic: contiguous env sharding and the statistics all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rbc_gym_b200.sharding import EpisodeStats, global_env_ids, shard_range


def test_shard_range_partitions_exactly():
    for n, g in ((8192, 2), (32768, 8), (10, 3), (5, 8)):
        cuts = [shard_range(n, g, r) for r in range(g)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(g - 1))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = global_env_ids(10, world, rank)
    # synthetic per-env step outputs that depend only on the GLOBAL env id -> result independent of the sharding
    st = EpisodeStats("cpu")
    for step in range(3):
        rew = -(ids.to(torch.float32) + step)
        st.accumulate(rew, -rew.double(), -rew.double() * 2, (ids % 4 == 0).to(torch.int32))
    out = st.reduce(world)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_episode_stats_allreduce_world2_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ids = torch.arange(10)
    st = EpisodeStats("cpu")
    for step in range(3):
        rew = -(ids.to(torch.float32) + step)
        st.accumulate(rew, -rew.double(), -rew.double() * 2, (ids % 4 == 0).to(torch.int32))
    ref = st.reduce(1)
    assert out == ref
    assert out["env_steps"] == 30 and out["nan_envs"] == 9 and out["max_abs_reward"] == 11.0
