#!/bin/bash
# ncu passes at the bench configuration: launch list of the whole command, then one full capture of the env kernel.
# Usage: bash tools/gpu_ncu.sh [envs_per_gpu]
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --envs-per-gpu ${1:-4096}"
echo "== plain"; timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-300; tail -2 gpurun_out/ncu_list.log | cut -c1-300
echo "== ncu full"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum --clock-control none --import-source on -k regex:rbc2d_env_kernel -s 12 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
python tools/ncu_traffic.py gpurun_out/prof.ncu-rep ${1:-4096} gpurun_out/traffic.json > /dev/null 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
