#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch table (count, total, mean, share).
Usage: python tools/launch_list.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rN_launch_list.txt"""
import collections
import csv
import sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= max(iK, iV) or r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[iU].replace("second", "s").replace("nsecond", "ns"), v / 1e3)
    a = agg.setdefault(r[iK], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {cmd}")
print("# per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes")
print(f"# {n} launches; total {tot / 1e3:.3f} ms\n")
print(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'mean_us':>10s} {'share':>8s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {c:8d} {t / 1e3:10.3f} {t / c:10.1f} {t / tot:8.4f}")
