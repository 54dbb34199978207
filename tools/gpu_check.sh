#!/bin/bash
# One GPU-box pass: smoke, GPU parity tests, bench (+ reference arm).  Usage: bash tools/gpu_check.sh
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
PY
tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 300 gpurun_out/bench_ref.json
