"""Synthetic initial states for grids the reference ships no checkpoints for (test helper)."""
import numpy as np

from oracle import oracle as O


def smooth_state(nx, nz, seed=0, amp=0.3):
    """A divergence-free, moderately rough flow with a perturbed conduction profile on an nx x nz grid
    (projected by the oracle).  Returns (P-independent) b[nz,nx], u[nz,nx], w[nz+1,nx]."""
    rng = np.random.default_rng(seed)
    z = (np.arange(nz) + 0.5) * 2 / nz
    x = (np.arange(nx) + 0.5) * 2 * np.pi / nx
    b = 1 + (2 - z)[:, None] / 2 + 0.1 * np.sin(3 * x)[None, :] * np.sin(np.pi * z / 2)[:, None] + 0.02 * rng.standard_normal((nz, nx))
    u = amp * rng.standard_normal((nz, nx))
    w = amp * rng.standard_normal((nz + 1, nx))
    for f in (u, w):
        f[:] = 0.25 * (np.roll(f, 1, 1) + np.roll(f, -1, 1)) + 0.5 * f
    w[0] = 0
    w[-1] = 0
    P = O.make_params(1e5, nx=nx, nz=nz, split_phy=False)
    u, w, _ = O.project(P, u, w)
    return b, u, w
