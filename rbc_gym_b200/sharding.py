"""Environment sharding across GPUs and the only collective of the path: episode statistics.

Environments are fully independent (no halo, no shared state), so the batch is cut into contiguous
ranges, rank r owning global environments [r*B/G, (r+1)*B/G) — the replacement for the reference's
process-per-env vectorisation (`example/run_vectorized.py:11-20`, `experiments/run_sarl.py:130-153`).
Nothing is exchanged inside a step; a handful of scalars is all-reduced per logging interval
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n_global: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `n_global` environments owned by `rank` (remainder to the low ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_env_ids(n_global: int, world: int, rank: int, device=None) -> torch.Tensor:
    lo, hi = shard_range(n_global, world, rank)
    return torch.arange(lo, hi, dtype=torch.int64, device=device)


class EpisodeStats:
    """Running sums of what the reference logs per step (`callbacks.py:29-43` logs info["nusselt"]):
    sum reward, sum Nu_obs, sum Nu_state, env-steps, NaN count, max |reward|.  `reduce` all-reduces them."""

    FIELDS = ("sum_reward", "sum_nu_obs", "sum_nu_state", "env_steps", "nan_envs")

    def __init__(self, device):
        self.acc = torch.zeros(len(self.FIELDS), dtype=torch.float64, device=device)
        self.max_abs_reward = torch.zeros(1, dtype=torch.float64, device=device)

    def accumulate(self, reward, nu_obs, nu_state, nan) -> None:
        self.acc[0] += reward.sum(dtype=torch.float64)
        self.acc[1] += nu_obs.sum(dtype=torch.float64)
        self.acc[2] += nu_state.sum(dtype=torch.float64)
        self.acc[3] += reward.numel()
        self.acc[4] += nan.sum(dtype=torch.float64)
        self.max_abs_reward = torch.maximum(self.max_abs_reward, reward.abs().max().to(torch.float64).reshape(1))

    def reduce(self, world: int) -> Dict[str, float]:
        acc, mx = self.acc.clone(), self.max_abs_reward.clone()
        if world > 1 and dist.is_initialized():
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        out = {k: float(v) for k, v in zip(self.FIELDS, acc.tolist())}
        n = max(out["env_steps"], 1.0)
        out["mean_reward"] = out["sum_reward"] / n
        out["mean_nu_obs"] = out["sum_nu_obs"] / n
        out["mean_nu_state"] = out["sum_nu_state"] / n
        out["max_abs_reward"] = float(mx.item())
        return out
