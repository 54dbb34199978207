#!/usr/bin/env python
"""Per-kernel launch count, average duration, share of the captured time and DRAM bytes from an ncu CSV launch list
(`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file ...`).
Usage: python tools/ncu_launch_shares.py gpurun_out/launches.csv"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
SCALE = {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows[i0 + 1:]:
    if len(r) < len(hdr):
        continue
    k = r[iK].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48]
    agg[k][r[iM]] += float(r[iV].replace(",", "")) * SCALE.get(r[iU], 1)
    if r[iM] == "gpu__time_duration.sum":
        cnt[k] += 1
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    t = a["gpu__time_duration.sum"]
    b = a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]
    print(f"{k:50s} n={cnt[k]:4d}  {t / cnt[k]:9.1f} us/launch  share {t / tot:5.3f}  dram {b / cnt[k] / 1e6:8.1f} MB/launch  {b / max(t, 1e-9) / 1e3:7.1f} GB/s")
