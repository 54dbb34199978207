#!/usr/bin/env python
"""Config 5 of BASELINE.json: an env-sharded rollout over the GPUs of one box, the replacement for the
process-per-env `SubprocVecEnv` rollout of `experiments/run_sarl.py:130-173`.

One process per GPU (torchrun), each owning a contiguous slice of the global batch of 2D environments; the
policy runs on the same device and stream as the environments, so observations and actions never leave the GPU;
NCCL only all-reduces the episode statistics once per logging interval.

    python examples/rollout_sharded.py --envs-per-gpu 4096 --steps 20                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        examples/rollout_sharded.py --envs-per-gpu 4096 --steps 20                          # 32768 envs on 8 GPUs
"""
import argparse
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rbc_gym_b200.envs import RBCVectorEnv2D  # noqa: E402
from rbc_gym_b200.sharding import EpisodeStats, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--ra", type=float, default=1e5)
    ap.add_argument("--checkpoint", default=str(ROOT / "data/checkpoints/train/ckpt_ra100000.h5"))
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_range(a.envs_per_gpu * world, world, rank)
    env = RBCVectorEnv2D(hi - lo, rayleigh_number=a.ra, heater_duration=1.5, checkpoint=a.checkpoint, device=local,
                         autoreset_mode="same_step", seed=0, env_id_offset=lo)
    obs, info = env.reset(seed=0)
    # a fixed random linear actor (stand-in for the PPO actor of experiments/run_sarl.py), same weights on every rank
    g = torch.Generator(device=dev).manual_seed(7)
    W = torch.randn((obs[0].numel(), env.heater_segments), device=dev, generator=g) * 0.05
    stats = EpisodeStats(dev)
    t0 = time.perf_counter()
    for step in range(1, a.steps + 1):
        actions = torch.tanh((obs.reshape(obs.shape[0], -1) - 1.5) @ W)          # stays on the device
        obs, reward, terminated, truncated, info = env.step(actions)
        stats.accumulate(reward, info["nusselt_obs"], info["nusselt_state"], torch.zeros_like(truncated, dtype=torch.int32))
        if step % a.log_every == 0 or step == a.steps:
            tot = stats.reduce(world)                                            # the only collective: a few scalars
            torch.cuda.synchronize()
            if rank == 0:
                el = time.perf_counter() - t0
                print(f"step {step:4d}  global envs {a.envs_per_gpu * world}  mean reward {tot['mean_reward']:+.4f}  "
                      f"mean Nu_obs {tot['mean_nu_obs']:.4f}  {tot['env_steps'] / el:,.0f} env-steps/s", flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
