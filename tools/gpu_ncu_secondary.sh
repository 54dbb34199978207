#!/bin/bash
# ncu full captures of the cluster kernel (config 3 grid) and the 3D kernel, each after a plain run of the same command.
# Usage: bash tools/gpu_ncu_secondary.sh
set -u
mkdir -p gpurun_out
C3="python tools/bench_grid.py --envs 33 --steps 1 --warmup 1 --dt 0.15"
C4="python tools/bench3d.py --envs 148 --steps 1"
echo "== cluster"; timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && tail -1 gpurun_out/plain_c3.log | cut -c1-200 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rbc2dx_env_kernel -s 1 -c 1 -f -o gpurun_out/prof_cluster $C3 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log | cut -c1-200
echo "== 3d"; timeout 300 $C4 > gpurun_out/plain_c4.log 2>&1 && tail -1 gpurun_out/plain_c4.log | cut -c1-200 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rbc3d_env_kernel -s 1 -c 1 -f -o gpurun_out/prof_3d $C4 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log | cut -c1-200
ls -la gpurun_out | head -30
