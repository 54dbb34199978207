"""Checkpoint generation: the reference's `simulate_2d_rb` (`src/rbc_gym/sim/rbc_sim2D.jl:14-72`, arguments in
`scripts/create_checkpoints_2D.sh:18-20`) as one batched on-device spin-up.

For each of `random_inits` episodes: noise initialisation (kick 0.02), `duration / delta_t_snap` runs of
`delta_t_snap` time units with zero action (bottom T = 2), NaN check per snapshot, final `b, u, w` written to a
reference-compatible HDF5 file (`h5lite.write_checkpoint`).  All episodes advance together as one batch.
Julia's RNG streams cannot be reproduced, so files are statistically — not bitwise — equivalent to the
reference's; episode i is seeded with `seed + i + 1` like `Random.seed!(seed + i)` (1-based i).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import backend
from .envs.rbc2d import noise_initial_fields
from .envs.rbc3d import noise_initial_fields_3d
from .h5lite import write_checkpoint, write_checkpoint_3d


def simulate_2d_rb(directory, seed: int = 42, random_inits: int = 20, ra: float = 1e5, pr: float = 0.7, random_kick: float = 0.02,
                   delta_t: float = 0.03, delta_t_snap: float = 0.3, duration: float = 600.0, precision: int = 64, device: int = 0,
                   progress=None, state_shape=(64, 96)):
    """`state_shape` = (Nz, Nx): any registered grid, e.g. (128, 192) with `delta_t=0.015` for config 3 — the reference
    ships 96 x 64 files only (`N = [96, 64]` default, `rbc_sim2D.jl:17`)."""
    import torch

    n = int(random_inits)
    shape = (int(state_shape[0]), int(state_shape[1]))
    sim = backend.Sim2D(n, ra=float(ra), dt_action=float(delta_t_snap), pr=pr, dt_solver=delta_t, episode_length=1e30,
                        precision=precision, device=device, state_shape=shape)
    fields = np.concatenate([noise_initial_fields(np.random.default_rng(seed + i + 1), shape, kick=random_kick) for i in range(n)])
    sim.reset_from_fields(fields, project=True)
    zero = torch.zeros((n, sim.heaters), device=sim.device)
    total = int(duration // delta_t_snap)
    for it in range(total):
        *_, nan = sim.step(zero)
        if (it % 100 == 99 or it == total - 1) and bool(nan.any()):
            raise RuntimeError("[ERROR] NaN values found!")             # rbc_sim2D.jl:196-199
        if progress and it % 200 == 0:
            progress(it, total)
    b, u, w = backend.split_fields(sim.fields(), shape)
    _, nus, nuo = sim.observe()
    stats = {"nu_state": nus.cpu().numpy().copy(), "nu_obs": nuo.cpu().numpy().copy()}
    sim.close()
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    ra_tag = int(ra) if float(ra).is_integer() else ra
    grid_tag = "" if shape == (64, 96) else f"_{shape[1]}x{shape[0]}"
    path = directory / f"ckpt_ra{ra_tag}{grid_tag}.h5"                  # rbc_sim2D.jl:36 (grid tag for non-default grids)
    write_checkpoint(path, b, u, w, start_seed=seed)
    return path, stats


def simulate_3d_rb(directory, seed: int = 42, random_inits: int = 20, ra: float = 2500, pr: float = 0.7, random_kick: float = 0.2,
                   delta_t: float = 0.01, delta_t_snap: float = 0.25, duration: float = 200.0, precision: int = 64, device: int = 0):
    """3D twin (`src/rbc_gym/sim/rbc_sim3D.jl:13-96`, arguments in `scripts/create_checkpoints_3D.sh`): the reference's 3D
    checkpoint files are missing from the mount, this regenerates statistically equivalent ones on the device.
    Times are in free-fall units like the reference's 3D API (`dt_solver`, `heater_duration`)."""
    import torch

    n = int(random_inits)
    sim = backend.Sim3D(n, ra=float(ra), pr=pr, heater_duration=float(delta_t_snap), dt_solver=delta_t, episode_length=1e30,
                        precision=precision, device=device)
    fields = np.concatenate([noise_initial_fields_3d(np.random.default_rng(seed + i + 1), kick=random_kick) for i in range(n)])
    sim.reset_from_fields(fields, project=True)
    zero = torch.zeros((n, sim.heaters, sim.heaters), device=sim.device)
    total = int(duration // delta_t_snap)
    for it in range(total):
        *_, nan = sim.step(zero, want_obs=False)
        if (it % 50 == 49 or it == total - 1) and bool(nan.any()):
            raise RuntimeError("[ERROR] NaN values found!")
    b, u, v, w = backend.split_fields3(sim.fields())
    _, nu = sim.observe()
    stats = {"nusselt": nu.cpu().numpy().copy()}
    sim.close()
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    ra_tag = int(ra) if float(ra).is_integer() else ra
    path = directory / f"3D_ckpt_ra{ra_tag}.h5"
    write_checkpoint_3d(path, b, u, v, w, start_seed=seed)
    return path, stats


def developed_states_2d(n_states: int, ra: float, state_shape=(64, 96), dt_solver: float = 0.03, spin_up: float = 40.0, seed: int = 42,
                        device: int = 0) -> np.ndarray:
    """`n_states` statistically developed flows `[n, nstate]` (float64) for grids without shipped checkpoints: noise
    initialisation + `spin_up` time units of zero-action evolution in the fp64 validation mode.

    fp64 on purpose: from the conduction state the first plumes overshoot (at Ra = 1e6 on 192 x 128: Nu > 100, max|w| ~ 1.7,
    i.e. CFL ~ 1.6 at dt_solver = 0.015), and at CFL > 1.5 the RK3 / 5th-order-upwind scheme amplifies grid-scale round-off by
    ~1.5x per step for a few dozen steps.  A 1e-16 seed survives that burst (as in the reference, which is fp64), a 1e-7 seed
    does not always (1 fp32 environment in ~1000 ends with the NaN flag set); developed flows have max|w| ~ 1.1."""
    import torch

    shape = (int(state_shape[0]), int(state_shape[1]))
    sim = backend.Sim2D(int(n_states), ra=float(ra), dt_action=1.0, dt_solver=dt_solver, episode_length=1e30, precision=64,
                        device=device, state_shape=shape)
    gen = torch.Generator(device=sim.device).manual_seed(seed)
    sim.noise_reset(kick=0.01, generator=gen)
    zero = torch.zeros((sim.B, sim.heaters), device=sim.device)
    for _ in range(int(round(spin_up))):
        *_, nan = sim.step(zero)
    if bool(nan.any()):
        raise RuntimeError("[ERROR] NaN values found!")
    f = sim.fields()
    sim.close()
    return f
