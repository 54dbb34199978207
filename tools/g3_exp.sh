set -u
B="python tools/bench3d.py --steps 2 --envs 64 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005"
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_3d_generic.py -q -x -k "chains_and or oracle" 2>&1 | tail -3
echo "== cluster projection"; timeout 200 $B | cut -c1-200
echo "== four kernels"; RBC_B200_G3_CLUSTER=0 timeout 200 $B | cut -c1-200
echo "== cluster projection, one chain"; RBC_B200_G3_STREAMS=1 timeout 200 $B | cut -c1-200
echo "== cluster projection, 148 envs"; timeout 200 ${B/--envs 64/--envs 148} | cut -c1-200
for sw in 1 0; do
RBC_B200_G3_CLUSTER=$sw RBC_B200_G3_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__cluster_max_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 3000 -c 12 --csv --log-file gpurun_out/g3_l$sw.csv python tools/bench3d.py --envs 64 --steps 1 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005 > gpurun_out/g3_ncu.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/g3_l$sw.csv")))
hdr=[r for r in rows if "Metric Name" in r][0]
iN,iV,iK=hdr.index("Metric Name"),hdr.index("Metric Value"),hdr.index("Kernel Name")
seen=set()
for r in rows:
    if len(r)>iV:
        k=(r[iK].split("(")[0][-40:],r[iN])
        if k not in seen: seen.add(k); print(k[0],k[1],r[iV])
PY
done
