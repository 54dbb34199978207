"""Host-side timing of the vector-step loop: how far the host runs ahead of the GPU, and whether any call stalls it (a cudaMalloc
of the caching allocator, a blocking event wait).  `RBC_BENCH_DIAG=1 python bench.py` prints the same timeline for the timed
region of the benchmark itself.  Usage: python tools/diag_host_gaps.py"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from rbc_gym_b200.envs import RBCVectorEnv2D
from rbc_gym_b200.sharding import EpisodeStats
env = RBCVectorEnv2D(4096, rayleigh_number=100_000, heater_duration=1.0, checkpoint=str(ROOT / "data/checkpoints/train/ckpt_ra100000.h5"), autoreset_mode="same_step")
env.reset(seed=0)
a = torch.rand((4096, 12), device="cuda") * 2 - 1
stats = EpisodeStats(torch.device("cuda"))
for _ in range(40):
    env.step(a)
torch.cuda.synchronize()
for rep in range(6):
    ht = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        t0 = time.perf_counter()
        o = env.step(a)
        t1 = time.perf_counter()
        stats.accumulate(o[1], o[4]["nusselt_obs"], o[4]["nusselt_state"], o[4]["nan"])
        ht.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
    e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: {e0.elapsed_time(e1) / 20:.3f} ms/step; host step ms: " + " ".join(f"{x:.1f}" for x, _ in ht) + " | acc max %.2f" % max(y for _, y in ht))
