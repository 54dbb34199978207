#!/bin/bash
# ncu passes on a small batch (592 envs = 4 waves of 148 CTAs): launch list, then one full capture.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --envs-per-gpu ${1:-592}"
echo "== plain"; timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/plain.log | cut -c1-400; tail -3 gpurun_out/ncu_list.log | cut -c1-300
echo "== ncu full"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:rbc2d_env_kernel -s 4 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
