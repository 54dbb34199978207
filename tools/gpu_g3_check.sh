#!/bin/bash
# Generic-grid 3D path (rbc3dg): parity tests, the flowstats-protocol bench (64 envs x 64x64x32, 50 RK3 steps per action), the ncu
# launch list of one action step and a full capture of the tendency kernel.
set -u
mkdir -p gpurun_out
B="python tools/bench3d.py --envs 64 --steps 2 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005"
echo "== pytest generic 3D"
timeout 900 python -m pytest tests/test_gpu_3d_generic.py -q -x > gpurun_out/g3_pytest.log 2>&1; echo "exit $?"; tail -4 gpurun_out/g3_pytest.log
echo "== bench"; timeout 300 $B > gpurun_out/g3_plain.log 2>&1; echo "exit $?"; tail -1 gpurun_out/g3_plain.log | cut -c1-400
echo "== bench, one chain"; RBC_B200_G3_STREAMS=1 timeout 300 $B 2>&1 | tail -1 | cut -c1-200
echo "== bench, 148 envs"; timeout 300 ${B/--envs 64/--envs 148} 2>&1 | tail -1 | cut -c1-200
echo "== launch list (one chain: ncu serialises kernels, so whole-batch launches are the meaningful ones)"
RBC_B200_G3_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 60 --csv --log-file gpurun_out/g3_launches.csv \
  python tools/bench3d.py --envs 64 --steps 1 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005 > gpurun_out/g3_ncu.log 2>&1; echo "exit $?"
python tools/ncu_launch_shares.py gpurun_out/g3_launches.csv 2>&1 | tail -12
[ "${G3_FULL:-0}" = 1 ] || exit 0
echo "== ncu full, tendency"
RBC_B200_G3_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:g3_tendency_tiled -s 40 -c 1 -f -o gpurun_out/prof_g3t \
  python tools/bench3d.py --envs 64 --steps 1 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005 > gpurun_out/g3_ncu_full.log 2>&1; echo "exit $?"
