#!/bin/bash
# One GPU-box pass: smoke, GPU parity tests, bench, ncu launch list + one full capture of the env kernel.
# Usage (from the repo root, on the GPU box): bash tools/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
if [ "${1:-}" != "quick" ]; then
  echo "== ncu launch list"
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --envs-per-gpu 592"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  tail -3 gpurun_out/ncu_list.log
  echo "== ncu full"
  timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rbc2d_env_kernel -s 3 -c 1 -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
