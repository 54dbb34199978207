#!/usr/bin/env python
"""Summarise an .ncu-rep of the env kernel: key raw metrics + samples/instructions per source function.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [envs_in_launch [stages_per_env [cells_per_env [core_header]]]]"""
import collections, csv, io, re, subprocess, sys
from pathlib import Path

rep = sys.argv[1]
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 592
stages_per_env = int(sys.argv[3]) if len(sys.argv) > 3 else 102      # RK3 stages per env action step
cells = int(sys.argv[4]) if len(sys.argv) > 4 else 6144               # cells per env
core = sys.argv[5] if len(sys.argv) > 5 else "rbc2d_core.h"           # source file whose functions are attributed
ROOT = Path(__file__).resolve().parent.parent
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k):
    return float(M[k][0].replace(",", ""))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum"]
for k in keys:
    if k in M: print(f"{k:75s} {M[k][0]:>18s} {M[k][1]}")
stages = envs * stages_per_env
print(f"per stage: cycles {g('sm__cycles_elapsed.avg') * 148 / stages:.0f}  warp-instr {g('smsp__inst_executed.sum') / stages:.0f} "
      f"(thread-instr/cell {g('smsp__inst_executed.sum') / stages * 32 / cells:.1f})  smem wavefronts {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / stages:.0f} "
      f"conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / stages:.0f}")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
def gb(k):
    return g(k) * UNIT[M[k][1]]
print(f"dram bytes per env-step: read {gb('dram__bytes_read.sum') / envs:.0f} write {gb('dram__bytes_write.sum') / envs:.0f}")
for k, (v, u) in M.items():
    if "average_warps_issue_stalled" in k and "_per_issue_active" in k and float(v) > 0.05:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per')[0]:22s} {float(v):.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[2]
iS, iI = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
ci = {h: i for i, h in enumerate(hdr)}
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = (ROOT / "rbc_gym_b200/csrc" / core).read_text().split("\n")
linefunc, cur = {}, None
for n, l in enumerate(lines, 1):
    m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:RBC_HD|__device__ __forceinline__|inline)\s+[\w<>:,\s\*&]+?\s+(\w+)(?:<[^>]*>)?\(", l)
    if m: cur = m.group(1)
    linefunc[n] = cur
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
curfile = ""
for r in rows:
    if r and r[0] == "File Path": curfile = r[1]; continue
    if len(r) < len(hdr) or not r[0].isdigit(): continue
    f = linefunc.get(int(r[0]), "?") if curfile.endswith("/" + core) else Path(curfile).name
    try:
        s_, ins = int(r[iS] or 0), int(r[iI] or 0)
    except ValueError:          # a source line whose quotes broke the CSV row (inline asm)
        continue
    agg[f][0] += s_; agg[f][1] += ins
    for c in stallcols:
        v = r[ci[c]]
        if v and v != "0": agg[f][2][c] += int(v)
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("samples", tot, "warp instr", toti)
for f, (s_, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if s_ < tot * 0.002: continue
    top = ", ".join(f"{k[6:]}={v / s_:.2f}" for k, v in st.most_common(4))
    print(f"{str(f):26s} samples {s_ / tot:6.3f}  instr {ins / toti:6.3f} ({ins / stages:7.0f}/stage)  {top}")
