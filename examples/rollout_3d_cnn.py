#!/usr/bin/env python
"""The caller side of `experiments/run_sarl.py` on this backend: a batch of 3D environments stepped on the device by
a small periodic-padding 3D CNN actor (same layer plan as `src/rbc_gym/models/CNN.py:34-57`: periodic pad -> Conv3d(4,8,3)
-> GELU -> MaxPool3d(2), twice, then a linear head onto the 8 x 8 heater patches), running on the same device and stream
as the environments — observations never leave the GPU.  Random weights: this shows the data path, not a trained policy.

    python examples/rollout_3d_cnn.py --envs 256 --steps 10
"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from rbc_gym_b200.envs import RBCVectorEnv3D  # noqa: E402


class PeriodicPad3D(nn.Module):
    """x, y periodic; z zero-padded.  Input (B, C, Nz, Ny, Nx)."""

    def forward(self, x):
        x = F.pad(x, (1, 1, 1, 1, 0, 0), mode="circular")
        return F.pad(x, (0, 0, 0, 0, 1, 1), mode="constant", value=0.0)


class Actor(nn.Module):
    def __init__(self, heaters=8):
        super().__init__()
        self.cnn = nn.Sequential(PeriodicPad3D(), nn.Conv3d(4, 8, 3), nn.GELU(), nn.MaxPool3d(2),
                                 PeriodicPad3D(), nn.Conv3d(8, 8, 3), nn.GELU(), nn.MaxPool3d(2), nn.Flatten())
        self.head = nn.Linear(8 * 4 * 8 * 8, heaters * heaters)
        self.heaters = heaters

    def forward(self, obs):
        return torch.tanh(self.head(self.cnn(obs))).reshape(-1, self.heaters, self.heaters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--ra", type=float, default=2500)
    a = ap.parse_args()
    env = RBCVectorEnv3D(a.envs, rayleigh_number=a.ra, autoreset_mode="same_step", seed=0)
    torch.manual_seed(0)
    actor = Actor().to(env.device).eval()
    obs, info = env.reset(seed=0)
    t0 = time.perf_counter()
    with torch.no_grad():
        for step in range(a.steps):
            obs_n = obs.clone()
            obs_n[:, 0] -= 1.5                                  # centre the temperature channel
            obs, reward, terminated, truncated, info = env.step(actor(obs_n))
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    print(f"{a.envs} envs x {a.steps} steps: {a.envs * a.steps / el:,.0f} env-steps/s incl. the actor; "
          f"mean Nu {info['nusselt'].mean().item():.4f}, mean reward {reward.mean().item():+.4f}")
    env.close()


if __name__ == "__main__":
    main()
