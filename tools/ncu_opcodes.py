#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode of the kernel in an .ncu-rep (source page, per-instruction execution counts).
Usage: python tools/ncu_opcodes.py gpurun_out/prof_3d.ncu-rep [cell_stages_in_launch]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
cell_stages = float(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Instructions Executed" in r)
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
warp, thread = collections.Counter(), collections.Counter()
for r in rows:
    if len(r) < len(hdr) or r is hdr:
        continue
    try:
        n, t = int(r[iI] or 0), int(r[iT] or 0)
    except ValueError:
        continue
    toks = r[iS].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    key = base if base not in ("LDS", "STS", "LDG", "STG", "LD", "ST", "SHFL", "BAR", "ATOMS") else ".".join(op.split(".")[:2]) if base in ("SHFL",) else base
    warp[key] += n; thread[key] += t
tot = sum(warp.values())
print(f"total warp instructions {tot}" + (f"  thread-instr per cell-stage {sum(thread.values()) / cell_stages:.1f}" if cell_stages else ""))
for k, v in warp.most_common(40):
    print(f"  {k:14s} {v / tot:6.3f}" + (f"  {thread[k] / cell_stages:7.1f} /cell-stage" if cell_stages else ""))
