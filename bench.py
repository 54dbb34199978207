#!/usr/bin/env python
"""Benchmark of the hot path: batched 2D Ra=1e5 environment action steps (dt = 1) on B200.

Contract (one JSON line on rank 0):
  python bench.py --gpus N --steps K --warmup W            -> this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  -> the reference's CPU algorithm (oracle port) on host cores
A "step" is one pass of the hot path over one batch: every environment of the batch advances by
one action step (34 RK3 steps = 102 projected stages at dt = 1, dt_solver = 0.03).

Workload = BASELINE.json configs[1]: 4096 envs per GPU, 96x64, 12 heaters, obs (3,8,48), Ra = 1e5, Pr = 0.7,
initial state of env e = episode (e mod 20) of data/checkpoints/train/ckpt_ra100000.h5, actions ~ U(-1,1)
from a device Philox generator seeded 1234 + rank.  Envs shard across ranks with no data-path collective
(weak scaling); NCCL only reduces episode statistics and the timing.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ENVS_PER_GPU = 4096
RA, DT_ACTION, DT_SOLVER = 1e5, 1.0, 0.03
CKPT = ROOT / "data/checkpoints/train/ckpt_ra100000.h5"
S_VALUES = 2 * 96 * 64 + 96 * 65                  # 18528 stored values per env
METRIC = "2D Ra=1e5 env-steps/s (dt=1)"
UNIT = "env-steps/s"


def algorithmic_bytes_per_env_step(nsub: int, real_bytes: int) -> float:
    """stage-streaming formulation, 10 * S values per RK3 step."""
    return nsub * 10 * S_VALUES * real_bytes


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc, self.thread = gpu_index, [], None, None
        self._stop = threading.Event()

    def start(self):
        """In-process NVML polling every 100 ms (clock + event-reason bitmask: two light queries).  An `nvidia-smi -lms 50` child
        process was measured to perturb the launches of the timed loop on some boxes (device-resident `value` 6 % below the
        host-buffer `e2e` of the same run); nvidia-smi at 250 ms is kept as the fallback when NVML cannot be loaded."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = (("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap))

            def poll():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.time(), ["", sm, mx, "", ""] + ["Active" if r & m else "Not Active" for _, m in bits]))
                    except Exception:
                        pass
                    self._stop.wait(0.1)
            self.proc = "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        """Summary of the samples that arrived inside [t_begin, t_end] (the timed region); nvidia-smi takes a few
        hundred ms to start, so the sampler is launched before the warm-up and the region is cut out afterwards."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self._stop.set()
        if self.proc != "nvml":
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if ts < t_begin or ts > t_end:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(sample_env_steps: int, threads: int):
    """Time the CPU oracle (fp64 port of the reference scheme) on `sample_env_steps` env action-steps."""
    from oracle import oracle as O
    from rbc_gym_b200.h5lite import load_checkpoint_2d
    c = load_checkpoint_2d(CKPT)
    P = O.make_params(RA, split_phy=True)
    dts = O.substep_schedule(DT_ACTION, DT_SOLVER)
    n = sample_env_steps
    idx = np.arange(n) % c.num_episodes
    b, u, w = c.b[idx].copy(), c.u[idx].copy(), c.w[idx].copy()
    acts = np.random.default_rng(1234).uniform(-1, 1, (n, 12))
    O.step_batch(P, b[:threads].copy(), u[:threads].copy(), w[:threads].copy(), acts[:threads], dts[:2], threads)  # warm
    t0 = time.perf_counter()
    bad = O.step_batch(P, b, u, w, acts, dts, threads)
    el = time.perf_counter() - t0
    assert not bad
    return n / el, el


def secondary_configs(device: int):
    """Side measurements of the other BASELINE.json configs on this GPU (not the headline; 2 timed steps each):
    config 3 = 2D 192x128 Ra=1e6 dt_solver=0.015 (cluster kernel), config 4 = 3D 32x32x16 Ra=1e4 (13 RK3 steps)."""
    import torch
    from rbc_gym_b200 import backend
    from rbc_gym_b200.envs import noise_initial_fields_3d
    out = {}
    peak, _ = hbm_peak()

    def timed(step_fn, n_steps=2):
        step_fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            step_fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n_steps

    try:
        B3 = 1056
        sim = backend.Sim2D(B3, ra=1e6, dt_action=1.0, dt_solver=0.015, state_shape=(128, 192), precision=32, device=device)
        from rbc_gym_b200.checkpoints import developed_states_2d
        base = developed_states_2d(16, 1e6, (128, 192), dt_solver=0.015, spin_up=40, device=device)   # fp64 spin-up, ~1 s
        sim.reset_from_fields(np.tile(base, (B3 // 16, 1)), project=False)
        a = torch.rand((B3, 12), device=f"cuda:{device}") * 2 - 1
        ms = timed(lambda: sim.step(a))
        alg = sim.nsub * 10 * sim.nstate * 4
        out["config3_2d_192x128_ra1e6"] = {"env_steps_per_s": B3 / ms * 1e3, "envs": B3, "rk3_steps": sim.nsub, "ms_per_step": ms,
                                           "streaming_equiv_GBps": B3 / ms * 1e3 * alg / 1e9, "frac_of_hbm_peak": B3 / ms * 1e3 * alg / 1e9 / peak,
                                           "kernel": "rbc2dx_env_kernel<192x128,cl4,f32>", "launch": sim.launch_info()}
        sim.close()
    except Exception as e:  # a side measurement must never take the headline down
        out["config3_2d_192x128_ra1e6"] = {"error": str(e)[:200]}
    try:
        B4 = 1184
        sim = backend.Sim3D(B4, ra=1e4, precision=32, device=device)
        rng = np.random.default_rng(0)
        base = np.concatenate([noise_initial_fields_3d(rng, kick=0.05) for _ in range(8)])
        sim.reset_from_fields(base[np.arange(B4) % 8], project=True)
        a = torch.rand((B4, 8, 8), device=f"cuda:{device}") * 2 - 1
        ms = timed(lambda: sim.step(a, want_obs=False))
        alg = 13 * 10 * 66560 * 4
        out["config4_3d_32x32x16_ra1e4"] = {"env_steps_per_s": B4 / ms * 1e3, "envs": B4, "rk3_steps": 13, "ms_per_step": ms,
                                            "streaming_equiv_GBps": B4 / ms * 1e3 * alg / 1e9, "frac_of_hbm_peak": B4 / ms * 1e3 * alg / 1e9 / peak,
                                            "kernel": "rbc3d_env_kernel<float, tiled>", "launch": sim.launch_info()}
        sim.close()
    except Exception as e:
        out["config4_3d_32x32x16_ra1e4"] = {"error": str(e)[:200]}
    try:
        # the resolution of the reference's own 3D runs (experiments/flowstats/flowstats_ra.py:27-36): 64 x 64 x 32, 50 RK3 steps per
        # action; the stage-streaming kernels (rbc3dg) really move the roofline's bytes through HBM
        B5, shape = 64, (32, 64, 64)
        sim = backend.Sim3D(B5, ra=1e4, state_shape=shape, heater_duration=0.25, dt_solver=0.005, precision=32, device=device)
        rng = np.random.default_rng(0)
        base = np.concatenate([noise_initial_fields_3d(rng, shape, kick=0.05) for _ in range(4)])
        sim.reset_from_fields(base[np.arange(B5) % 4], project=True)
        a = torch.rand((B5, 8, 8), device=f"cuda:{device}") * 2 - 1
        ms = timed(lambda: sim.step(a, want_obs=False), n_steps=1)
        alg = sim.nsub * 10 * sim.nstate * 4
        out["flowstats_3d_64x64x32_ra1e4"] = {"env_steps_per_s": B5 / ms * 1e3, "envs": B5, "rk3_steps": sim.nsub, "ms_per_step": ms,
                                              "streaming_equiv_GBps": B5 / ms * 1e3 * alg / 1e9, "frac_of_hbm_peak": B5 / ms * 1e3 * alg / 1e9 / peak,
                                              "kernel": "rbc3dg stage-streaming kernels (tendency, div+FFT, Thomas, inverse, correct)", "launch": sim.launch_info()}
        sim.close()
    except Exception as e:
        out["flowstats_3d_64x64x32_ra1e4"] = {"error": str(e)[:200]}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; Julia cannot run here),
    all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(threads * 8, 64)                       # a bounded sample: well under a second of CPU work per step on 16 cores
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(threads, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, _ = cpu_reference_rate(per_step, threads)
        rates.append(r)
    el = time.perf_counter() - t0
    value = per_step * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2D Ra=1e5 96x64 dt=1 (34 RK3 steps), reset from train checkpoints, U(-1,1) actions",
                   "sample_envs_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} env action-steps per step x {args.steps} steps, oracle fp64 C port, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Julia/Oceananigans reference cannot run here (no julia); CPU oracle port of the same scheme. "
                "Reference's published figure: 8.3 env-steps/s per process (README.md:62, Apple silicon).",
    }
    print(json.dumps(line), flush=True)


def source_hash() -> str:
    """Hash of the kernel sources the committed ncu figures (profiles/traffic.json) were measured on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("rbc2d_core.h", "rbc2d_lib.cu"):
        h.update((ROOT / "rbc_gym_b200" / "csrc" / f).read_bytes())
    return h.hexdigest()[:16]


def ncu_figures():
    """Per-launch DRAM bytes and fp32 flop of the dominant kernel from the committed ncu capture (tools/gpu_ncu.sh writes
    profiles/traffic.json with the hash of the sources it ran); a capture of other sources is refused, not reported."""
    tj = ROOT / "profiles" / "traffic.json"
    if not tj.exists():
        return None, None, "no ncu capture committed"
    try:
        d = json.loads(tj.read_text())
    except Exception:
        return None, None, "unreadable profiles/traffic.json"
    if d.get("source_hash") != source_hash():
        return None, None, f"stale: ncu capture is of sources {d.get('source_hash')}, this build is {source_hash()}"
    return d.get("dram_bytes_per_launch"), d.get("fp32_flop_per_launch"), f"ncu --set full, {d.get('envs_per_launch')} envs per launch"


def uniform_actions(torch, step: int, gids, heaters: int, seed: int = 1234):
    """U(-1,1) actions keyed by (seed, GLOBAL env id, step, heater): rank r of an N-GPU run feeds its environments exactly what
    the single-GPU run feeds the same global ids, so the runs can be compared bit for bit (slice_checksum)."""
    from rbc_gym_b200.envs.vector import _as_i64, _mix63, _M63
    k = torch.arange(heaters, device=gids.device, dtype=torch.int64)[None, :]
    x = (gids[:, None] * _as_i64(0x9E3779B97F4A7C15) + k * _as_i64(0xC2B2AE3D27D4EB4F) + _as_i64((seed * 1_000_003 + step) * 0x165667B19E3779F9)) & _M63
    x = _mix63(torch, _mix63(torch, x))
    return ((x >> 10).to(torch.float64) * (2.0 / (1 << 53)) - 1.0).to(torch.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=18.0, help="CPU work of the cpu_baseline sample")
    ap.add_argument("--no-secondary", action="store_true", help="skip the side measurements (configs 3/4/5, fp64, 300-step episode)")
    ap.add_argument("--episode-steps", type=int, default=300, help="length of the full-episode leg (auto-reset fires inside it)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import hashlib
    import torch
    import torch.distributed as dist
    from rbc_gym_b200.envs import RBCVectorEnv2D
    from rbc_gym_b200.sharding import EpisodeStats, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.envs_per_gpu
    lo, hi = shard_range(B * world, world, rank)           # this rank's slice of the global batch
    # the path RL callers use: the vector environment (auto-reset fused into the step kernel, gymnasium's next_step mode)
    env = RBCVectorEnv2D(B, rayleigh_number=RA, heater_duration=DT_ACTION, dt_solver=DT_SOLVER, episode_length=300, checkpoint=str(CKPT),
                         precision=args.precision, device=local_rank, autoreset_mode="next_step", seed=1234, env_id_offset=lo)
    sim = env.sim
    n_ep = sim.n_episodes
    gid = torch.arange(lo, hi, device=dev, dtype=torch.int64)
    start_idx = (gid % n_ep).to(torch.int32)               # initial state of env e = episode (e mod 20) of the train file
    K, W = args.steps, args.warmup
    actions = torch.stack([uniform_actions(torch, i, gid, sim.heaters) for i in range(W + K)])
    stats = EpisodeStats(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    # untimed pre-warm: bring the SM clock up from idle (120 MHz) before the W warm-up steps
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("RBC_BENCH_NO_CLOCKS"):
        sampler.start()
    env.reset(options={"checkpoint_idx": start_idx})
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 1.0:
        env.step(actions[0])
        torch.cuda.synchronize()
    env.reset(options={"checkpoint_idx": start_idx})
    # warm-up with exactly the statements of the timed loop: besides torch's lazily loaded kernels this brings the caching
    # allocator to its steady state — the results of step i are still referenced while step i + 1 allocates its copies, and a
    # loop that first reaches that peak inside the timed region pays a cudaMalloc there (measured: a 2 ... 120 ms host stall in
    # the SECOND timed step with an empty launch queue behind it, i.e. up to 10 % of a 20-step region)
    for i in range(W):
        obs, rew, term, trunc, info = env.step(actions[i])
        stats.accumulate(rew, info["nusselt_obs"], info["nusselt_state"], info["nan"])
    stats = EpisodeStats(dev)
    # the queue is empty after the barrier: a host pause in the first steps of the region (a cyclic-GC pass costs tens of ms)
    # would be GPU idle time, so collect now and keep the collector off while timing
    gc.collect()
    gc.disable()
    barrier()
    t_region0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sim.launch_info()["launches"]
    e0.record()
    diag = [] if os.environ.get("RBC_BENCH_DIAG") else None      # per-step device and host timeline (debugging aid, off by default)
    for i in range(K):
        th = time.perf_counter()
        obs, rew, term, trunc, info = env.step(actions[W + i])
        stats.accumulate(rew, info["nusselt_obs"], info["nusselt_state"], info["nan"])
        if diag is not None:
            ev = torch.cuda.Event(enable_timing=True); ev.record()
            diag.append((ev, (time.perf_counter() - th) * 1e3))
    e1.record()
    barrier()
    gc.enable()
    elapsed_ms = e0.elapsed_time(e1)
    if diag is not None and rank == 0:
        prev, out_ = e0, []
        for ev, host_ms in diag:
            out_.append(f"{prev.elapsed_time(ev):.2f}/{host_ms:.2f}")
            prev = ev
        print("diag device ms per step / host ms per step: " + " ".join(out_), file=sys.stderr)
    launches = sim.launch_info()["launches"] - launches0
    clocks = sampler.stop(t_region0, time.time()) if rank == 0 else None
    # per-launch duration of the dominant kernel over the timed region: the library records a CUDA event pair around every
    # step kernel on its launch stream (ring of 64); they are read here, after the region, so nothing synchronised inside it
    kernel_ms = sim.step_kernel_ms_history(min(K, 64))
    # bitwise fingerprint of the first 64 environments of the GLOBAL batch after W + K steps (rank 0 owns them at every N):
    # equal values on the N = 1 and N = 8 lines prove that sharding does not change a single bit
    slice_checksum = hashlib.sha256(sim.get_state()[:64].cpu().numpy().tobytes()).hexdigest()[:16] if rank == 0 else None
    tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = tmax.item()
    totals = stats.reduce(world)                         # NCCL all_reduce of a handful of scalars
    value = world * B * K / (elapsed_ms * 1e-3)

    # ---------------- end to end through the host-buffer C-ABI entry point (`e2e`) ----------------
    host_actions = [torch.empty((B, sim.heaters), dtype=torch.float32).pin_memory() for _ in range(K)]
    for i in range(K):
        host_actions[i].copy_(actions[W + i].cpu())
    out = env.alloc_host_outputs(pinned=True)
    env.step_host(host_actions[0].numpy(), out)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        env.step_host(host_actions[i].numpy(), out)       # H2D actions, fused vector step, D2H obs/reward/nu/flags/t/step/returns, sync
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    tmax = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (tmax.item() * 1e-3)
    h2d = B * sim.heaters * 4
    d2h = sum(v.nbytes for v in out.values())

    # ---------------- side measurements (every rank runs the sharded ones; rank 0 the single-GPU ones) ----------------
    secondary = {}
    if not args.no_secondary:
        def timed_loop(fn, n):
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(n):
                fn(i)
            a1.record()
            barrier()
            tm = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return tm.item()

        # (1) one full episode: 300 vector steps from reset, same_step mode, so that the fused auto-reset of every environment
        #     fires inside the timed region (SURVEY 8d protocol)
        T = args.episode_steps
        ep_env = RBCVectorEnv2D(B, rayleigh_number=RA, heater_duration=DT_ACTION, dt_solver=DT_SOLVER, episode_length=T * DT_ACTION,
                                checkpoint=str(CKPT), precision=args.precision, device=local_rank, autoreset_mode="same_step", seed=1234,
                                env_id_offset=lo)
        ep_env.reset(options={"checkpoint_idx": start_idx})
        n_reset = torch.zeros((), device=dev, dtype=torch.int64)
        def ep_step(i):
            nonlocal n_reset
            o = ep_env.step(uniform_actions(torch, 10_000 + i, gid, sim.heaters))
            n_reset += o[3].sum()
        ms = timed_loop(ep_step, T)
        tot = n_reset.to(torch.float64).reshape(1)
        if world > 1:
            dist.all_reduce(tot)
        secondary["episode_rollout"] = {"env_steps_per_s": world * B * T / (ms * 1e-3), "steps": T, "autoreset_mode": "same_step",
                                        "environments_auto_reset_inside_the_timed_region": int(tot.item()), "ms_per_step": ms / T,
                                        "note": "vector env, action generation on the device inside the loop, every env truncates at the last step"}
        ep_env.close()

        # (2) config 5: a policy in the loop on the same stream (fixed random linear map obs -> 12 actions, tanh), next_step mode
        Wp = torch.randn((sim.channels * 8 * 48, sim.heaters), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 0.05
        pol = {"obs": env.reset(options={"checkpoint_idx": start_idx})[0]}
        def pol_step(i):
            a = torch.tanh(pol["obs"].flatten(1) @ Wp)
            pol["obs"] = env.step(a)[0]
        pol_step(0)
        ms = timed_loop(pol_step, K)
        secondary["config5_policy_in_loop"] = {"env_steps_per_s": world * B * K / (ms * 1e-3), "global_envs": world * B, "steps": K,
                                               "policy": "tanh(linear(obs)) on the env's stream, actions never leave the device",
                                               "ms_per_step": ms / K}

        # (3) end to end INCLUDING info["state"] (the reference returns the full state every step, rbc2D.py:211: 74 KB per env)
        st_host = torch.empty((B, sim.channels, 64, 96), dtype=torch.float32).pin_memory()
        def e2e_state(i):
            env.step_host(host_actions[i % K].numpy(), out)
            st_host.copy_(sim.get_state(), non_blocking=True)
            torch.cuda.synchronize()
        e2e_state(0)
        barrier()
        t0 = time.perf_counter()
        n_s = min(K, 5)
        for i in range(n_s):
            e2e_state(i)
        tm = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        secondary["e2e_with_state"] = {"env_steps_per_s": world * B * n_s / (tm.item() * 1e-3), "d2h_bytes_per_step": d2h + st_host.numel() * 4,
                                       "steps": n_s}
        if rank == 0:
            # (4) the fp64 validation mode (the reference's arithmetic) on the same workload
            try:
                e64 = RBCVectorEnv2D(B, rayleigh_number=RA, heater_duration=DT_ACTION, dt_solver=DT_SOLVER, checkpoint=str(CKPT), precision=64,
                                     device=local_rank, seed=1234, env_id_offset=lo)
                e64.reset(options={"checkpoint_idx": start_idx})
                e64.step(actions[0]); torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for i in range(3):
                    e64.step(actions[1 + i])
                a1.record(); torch.cuda.synchronize()
                ms = a0.elapsed_time(a1) / 3
                secondary["fp64_validation_mode"] = {"env_steps_per_s": B / ms * 1e3, "ms_per_step": ms, "dtype": "f64", "envs": B,
                                                     "streaming_equiv_GBps": B / ms * 1e3 * algorithmic_bytes_per_env_step(sim.nsub, 8) / 1e9}
                e64.close()
            except Exception as e:
                secondary["fp64_validation_mode"] = {"error": str(e)[:200]}
            secondary.update(secondary_configs(local_rank))
    if rank == 0:
        real_bytes = args.precision // 8
        per_env = algorithmic_bytes_per_env_step(sim.nsub, real_bytes)
        k_ms = float(np.mean(kernel_ms))
        achieved = per_env * B / (k_ms * 1e-3) / 1e9
        peak, peak_src = hbm_peak()
        traffic, flop, ncu_note = ncu_figures()
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r0, _ = cpu_reference_rate(2 * threads, threads)            # calibrate, then ~10 s of CPU work
            n = int(min(max(args.cpu_seconds * r0, 4 * threads), 20000))
            n -= n % threads
            rate, el = cpu_reference_rate(n, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{n} env action-steps of the same workload (oracle fp64 C port, {threads} threads, {el:.1f} s)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 2D Ra=1e5 96x64 dt=1 (34 RK3 steps / 102 projected stages), "
                                   f"{B} envs per GPU, reset from train checkpoints, U(-1,1) actions, through the vector env "
                                   "(RBCVectorEnv2D.step: auto-reset fused into the step kernel, info t/step/nusselt)",
                       "envs_per_gpu": B, "global_envs": world * B, "parallelism": f"env-sharded x{world}",
                       "l2": "inputs larger than L2 (%.0f MB of state per GPU)" % (B * S_VALUES * real_bytes / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "rbc2d_env_kernel",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": per_env * B,
                         "note": "algorithmic bytes = stage-streaming formulation (SURVEY 8d); the kernel keeps each env "
                                 "on-chip for the whole action step, so measured DRAM traffic is far below this figure",
                         "ncu": ncu_note,
                         "fp32_tflops": None if flop is None else flop * (B / 4096.0) / (k_ms * 1e-3) / 1e12},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "rbc2d_vec_step_host: host actions in, obs/reward/nusselt/flags/t/step/returns out; info['state'] "
                            "(74 KB per env, rbc2D.py:211) is opt-in for the batch: see secondary.e2e_with_state"},
            "gpu_launches": launches,
            "clocks": clocks,
            "episode_stats": totals,
            "slice_checksum": slice_checksum,
            "secondary": secondary or None,
        }
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
