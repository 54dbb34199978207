set -u
B="python tools/bench3d.py --steps 2 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005"
r() { echo "== $1"; shift; env "$@" timeout 300 $B --envs ${ENVS:-64} 2>&1 | tail -1 | cut -c1-190; }
for c in 1 2 3 4; do r chains$c RBC_B200_G3_STREAMS=$c; done
for c in 2 4; do ENVS=148 r envs148-chains$c RBC_B200_G3_STREAMS=$c; done
for c in 2 4; do ENVS=14 r envs14-chains$c RBC_B200_G3_STREAMS=$c; done
