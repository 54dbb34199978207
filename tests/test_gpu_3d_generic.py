"""The 3D environment on grids other than the registered 32 x 32 x 16 (`state_shape` is a free keyword of the reference env,
`rbc3D.py:43-60`; its flowstats experiment, the only Julia-produced 3D data, ran at (32, 64, 64)): the stage-streaming kernels of
`rbc3dg_lib.cu` through the C ABI, against the 3D oracle, against the dedicated kernel on the registered grid, and — the like-for-like
physics check — the reference's own flowstats protocol for all 14 Rayleigh numbers against the Julia-produced series
(`tests/golden/flowstats_julia_64x64x32.json`, extracted from `experiments/flowstats/flowstats_ra.pkl`)."""
import json
import os
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def rel(x, y):
    return np.linalg.norm(x - y) / np.linalg.norm(y)


def projected_state(P, seed, amp=0.2):
    from oracle import oracle3d as O3
    rng = np.random.default_rng(seed)
    nz, ny, nx = P.nz, P.ny, P.nx
    z = (np.arange(nz) + 0.5) * P.lz / nz
    b = 1 + (P.lz - z)[:, None, None] / 2 + 0.05 * rng.standard_normal((nz, ny, nx))
    u, v = amp * rng.standard_normal((nz, ny, nx)), amp * rng.standard_normal((nz, ny, nx))
    w = amp * rng.standard_normal((nz + 1, ny, nx)); w[0] = 0; w[-1] = 0
    return (b, *O3.project(P, u, v, w))


@pytest.mark.parametrize("shape", [(32, 64, 64), (8, 16, 32), (7, 8, 16)])      # the last: per-cell tendency, odd number of levels
def test_generic_grid_matches_oracle(shape):
    import torch
    from oracle import oracle3d as O3
    from rbc_gym_b200 import backend
    P = O3.make_params(5e3, shape=shape, split_phy=False)
    states = [projected_state(P, s) for s in (1, 2)]
    acts = np.random.default_rng(2).uniform(-1, 1, (2, 8, 8)).astype(np.float32)
    refs = [O3.step(P, *st, acts[j].astype(np.float64), O3.substep_schedule(0.02, 0.01)) for j, st in enumerate(states)]
    for prec, tol in ((64, 1e-10), (32, 1e-5)):
        sim = backend.Sim3D(2, ra=5e3, state_shape=shape, heater_duration=0.02, precision=prec)
        sim.reset_from_fields(np.concatenate([backend.pack_fields3(*[x[None] for x in st]) for st in states]), project=False)
        obs, rew, nu, trunc, nan = sim.step(torch.from_numpy(acts).cuda())
        gb, gu, gv, gw = backend.split_fields3(sim.fields(), shape)
        for j, r in enumerate(refs):
            assert rel(gb[j], r["b"]) < tol and rel(gu[j], r["u"]) < tol and rel(gv[j], r["v"]) < tol and rel(gw[j], r["w"]) < tol, (prec, j)
            assert nu[j].item() == pytest.approx(O3.nusselt(P, r["b"], r["w"]), rel=1e-9 if prec == 64 else 1e-4)
            assert rew[j].item() == pytest.approx(-nu[j].item(), rel=1e-6) and nan[j].item() == 0
        assert obs.shape == (2, 4, *shape)
        np.testing.assert_array_equal(obs[0].cpu().numpy(), np.stack([gb[0], gu[0], gv[0], gw[0][:-1]]).astype(np.float32))
        t, step = sim.info()
        assert np.all(t == pytest.approx(0.08)) and np.all(step == 2)
        sim.close()


def test_generic_kernels_agree_with_dedicated_kernel_on_registered_grid(monkeypatch):
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import noise_initial_fields_3d
    f = np.concatenate([noise_initial_fields_3d(np.random.default_rng(s), kick=0.05) for s in range(3)])
    a = torch.rand((3, 8, 8), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 2 - 1
    out = {}
    for name, env in (("dedicated", "0"), ("generic", "1")):
        monkeypatch.setenv("RBC_B200_3D_GENERIC", env)
        sim = backend.Sim3D(3, ra=1e4, precision=64)
        sim.reset_from_fields(f, project=True)
        for _ in range(2):
            _, _, nu, _, _ = sim.step(a)
        out[name] = (sim.fields(), nu.cpu().numpy())
        sim.close()
    assert rel(out["generic"][0], out["dedicated"][0]) < 1e-11 and np.allclose(out["generic"][1], out["dedicated"][1], rtol=1e-10)


@pytest.mark.parametrize("shape", [(16, 32, 64), (32, 64, 64), (6, 16, 32)])
def test_tiled_tendency_equals_the_per_cell_kernel(monkeypatch, shape):
    """The tendency of the generic path has two device forms: one thread per cell gathering its windows from global memory
    (also what the host emulator runs), and — for nx % 32 == 0, ny % 16 == 0 — a 32 x 16 patch of columns per CTA marching up
    through a ring of level planes with halo in shared memory.  Both evaluate `tendency_from_windows` on the same values, so a
    rollout must agree to round-off of the fused-multiply-add contraction (in practice: bit for bit), in both precisions,
    including grids where a patch is its own periodic neighbour and the shallowest supported column (nz = 6)."""
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import noise_initial_fields_3d
    f = np.concatenate([noise_initial_fields_3d(np.random.default_rng(s), shape, kick=0.05) for s in range(2)])
    a = torch.rand((2, 8, 8), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 2 - 1
    for prec, tol in ((64, 1e-13), (32, 2e-6)):
        out = {}
        for name, sw in (("tiled", "1"), ("cell", "0")):
            monkeypatch.setenv("RBC_B200_G3_TILED", sw)
            sim = backend.Sim3D(2, ra=2e4, state_shape=shape, heater_duration=0.05, precision=prec)
            sim.reset_from_fields(f, project=True)
            for _ in range(2):
                _, _, nu, _, nan = sim.step(a)
            assert int(nan.sum()) == 0
            out[name] = (sim.fields(), nu.cpu().numpy())
            sim.close()
        assert rel(out["tiled"][0], out["cell"][0]) < tol, (prec, rel(out["tiled"][0], out["cell"][0]))
        assert np.allclose(out["tiled"][1], out["cell"][1], rtol=100 * tol)


@pytest.mark.parametrize("shape,envs", [((32, 64, 64), 33), ((8, 32, 32), 5)])
def test_chains_and_compile_time_plane_extents_change_no_bit(monkeypatch, shape, envs):
    """Two scheduling / code-generation choices of the generic path must be invisible in the results: (i) the batch runs as up to eight
    independent chains on separate streams (`RBC_B200_G3_STREAMS`; uneven cuts here: 33 environments over eight chains, 5 over two) —
    environments never interact; (ii) the FFT plane kernels instantiated with compile-time plane extents for 64 x 64 and 32 x 32 columns
    (`RBC_B200_G3_FIXED_PLANE`) run the same butterflies in the same order as the run-time-extent kernels.  Bitwise equality of a
    two-step rollout, states and Nusselt numbers, in the throughput precision."""
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import noise_initial_fields_3d
    f = np.concatenate([noise_initial_fields_3d(np.random.default_rng(s), shape, kick=0.05) for s in range(4)])[np.arange(envs) % 4]
    a = torch.rand((envs, 8, 8), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 2 - 1
    ra = np.linspace(5e3, 4e4, envs)
    out = {}
    for name, streams, fixed in (("product", None, None), ("one-chain", "1", None), ("two-chains", "2", None), ("run-time-planes", None, "0")):
        for key, val in (("RBC_B200_G3_STREAMS", streams), ("RBC_B200_G3_FIXED_PLANE", fixed)):
            if val is None:
                monkeypatch.delenv(key, raising=False)
            else:
                monkeypatch.setenv(key, val)
        sim = backend.Sim3D(envs, ra=1e4, state_shape=shape, heater_duration=0.05, precision=32)
        sim.set_rayleigh(ra)
        sim.reset_from_fields(f, project=True)
        for _ in range(2):
            _, _, nu, _, nan = sim.step(a)
        assert int(nan.sum()) == 0
        out[name] = (sim.fields().copy(), nu.cpu().numpy().copy())
        sim.close()
    for name in ("one-chain", "two-chains", "run-time-planes"):
        assert np.array_equal(out[name][0], out["product"][0]), name
        assert np.array_equal(out[name][1], out["product"][1]), name


def test_vector_env_on_a_generic_grid_and_per_environment_rayleigh():
    import torch
    from rbc_gym_b200.envs import RBCVectorEnv3D
    env = RBCVectorEnv3D(4, rayleigh_number=2500, state_shape=(8, 16, 16), heater_duration=0.05, episode_length=0.35, autoreset_mode="same_step", seed=3)
    env.sim.set_rayleigh([500.0, 2500.0, 2500.0, 1e5])
    obs, info = env.reset()
    assert obs.shape == (4, 4, 8, 16, 16) and (info["step"] == 1).all()
    a = torch.zeros((4, 8, 8), device="cuda")
    for it in range(2):
        obs, rew, term, trunc, info = env.step(a)
    assert trunc.all() and (info["step"] == 1).all() and "final_obs" in info          # noise re-initialisation driven from Python
    assert torch.isfinite(obs).all()
    # environments 1 and 2 share the Rayleigh number but not the noise; 0 and 3 differ in diffusivity: reward = -Nu differs
    assert rew[1].item() != rew[2].item() and abs(info["final_info"]["nusselt"][0].item() - 1) < 0.05
    env.close()


def growth_plateau(nu):
    e = np.log(np.maximum(np.asarray(nu) - 1.0, 1e-12))
    return float(np.diff(e)[2:60].max())


def test_flowstats_protocol_reproduces_the_julia_runs_at_64x64x32():
    """`experiments/flowstats/flowstats_ra.py:27-36` on the GPU, like for like: (32, 64, 64), dt_solver 0.005, heater_duration 0.25
    (one sample per time unit), zero action, noise initialisation with kick 0.01, 300 samples, all 14 Rayleigh numbers as ONE
    batch (per-environment Rayleigh numbers), fp64 like the reference.  Compared with the Julia-produced Nusselt series:
      * Ra >= 8000 (unsteady convection): saturated mean over the last 150 samples within the Julia series' own scatter (its
        standard deviation over those samples; measured differences are 0.0-0.6 %, i.e. <= 0.3 of that scatter); Ra = 4000:
        within twice that scatter;
      * Ra <= 2000: the flow settles on a steady planform whose Nusselt number depends on the pattern the noise selects — the
        Julia table itself is not monotonic there (Nu(750) = 1.517 > Nu(1000) = 1.510, i.e. a quarter of Nu - 1 is pattern
        scatter) — so only 25 % of Nu - 1 is asserted;
      * the linear-instability plateau (largest local growth rate of Nu - 1) within 3.5 %.  DESIGN.md section 2 has the full
        picture: a 24-seed GPU ensemble puts the Julia run inside the seed scatter at Ra = 500 and 1.1-2.2 % below the ensemble
        mean for Ra >= 4000 (a real, still unexplained difference in the transient growth; the saturated statistics agree)."""
    import torch
    from rbc_gym_b200 import backend
    gold = json.loads((ROOT / "tests/golden/flowstats_julia_64x64x32.json").read_text())["runs"]
    ras = sorted(gold, key=float)
    n, shape = len(ras), (32, 64, 64)
    sim = backend.Sim3D(n, ra=2500, state_shape=shape, heater_duration=0.25, dt_solver=0.005, episode_length=1e9, precision=64)
    sim.set_rayleigh([float(r) for r in ras])
    sim.noise_reset(kick=0.01, generator=torch.Generator(device="cuda").manual_seed(2024))
    a = torch.zeros((n, 8, 8), device="cuda")
    series, wmax = [], []
    for s in range(300):
        obs, rew, nu, trunc, nan = sim.step(a)
        series.append(nu.clone())
        wmax.append(obs[:, 3].abs().amax(dim=(1, 2, 3)))
    nus = torch.stack(series, 1).cpu().numpy()
    wm = torch.stack(wmax, 1).cpu().numpy()
    assert int(nan.sum()) == 0 and np.isfinite(nus).all()
    out = ROOT / "gpurun_out"
    if out.exists():
        (out / "flowstats_gpu_64x64x32.json").write_text(json.dumps({"ra": ras, "nusselt_step": nus.tolist(), "uz_max_step": wm.tolist(),
                                                                     "precision": 64, "kernel_ms_per_sample": sim.last_step_kernel_ms()}))
    for j, ra in enumerate(ras):
        jn = np.array(gold[ra]["nusselt_step"])
        m_j, s_j, m_g = jn[150:].mean(), jn[150:].std(), nus[j, 150:].mean()
        tol = 0.25 * (m_j - 1) if float(ra) <= 2000 else (2 * s_j if float(ra) < 8000 else s_j)
        assert abs(m_g - m_j) < tol, (ra, m_g, m_j, tol)
        assert growth_plateau(nus[j]) == pytest.approx(growth_plateau(jn), rel=0.035), ra
    sim.close()
