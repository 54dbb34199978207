#!/usr/bin/env python
"""Per-launch DRAM traffic and fp32 flop count of the headline kernel from one `ncu --set full` capture, written as
profiles/traffic.json together with the hash of the kernel sources it was measured on (bench.py refuses the file when the
hash differs from the build it runs).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep 4096 [out.json]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "inst": 1.0, "": 1.0}

if __name__ == "__main__":
    rep, envs = sys.argv[1], int(sys.argv[2])
    out = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "profiles" / "traffic.json"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    M = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}

    def g(k):
        if k not in M:
            return None
        return float(M[k][0].replace(",", "")) * UNIT.get(M[k][1], 1.0)

    from bench import source_hash
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    fadd, fmul, ffma = (g(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum") for o in ("fadd", "fmul", "ffma"))
    flop = None if None in (fadd, fmul, ffma) else fadd + fmul + 2 * ffma
    d = {"kernel": M.get("Kernel Name", ("rbc2d_env_kernel", ""))[0], "envs_per_launch": envs, "source_hash": source_hash(),
         "source": "ncu --set full --clock-control none (tools/gpu_ncu.sh)",
         "gpu_time_ms": None if g("gpu__time_duration.sum") is None else g("gpu__time_duration.sum") * (1e-6 if M["gpu__time_duration.sum"][1] == "ns" else 1e-3 if M["gpu__time_duration.sum"][1] == "us" else 1.0),
         "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": None if rd is None else rd + wr,
         "dram_bytes_per_env_step": None if rd is None else (rd + wr) / envs,
         "fp32_thread_inst": {"fadd": fadd, "fmul": fmul, "ffma": ffma},
         "fp32_flop_per_launch": flop, "fp32_flop_per_env_step": None if flop is None else flop / envs}
    out.write_text(json.dumps(d, indent=1))
    print(json.dumps(d, indent=1))
