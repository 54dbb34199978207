set -u
B="python tools/bench3d.py --steps 2 --shape 32x64x64 --heater-duration 0.25 --dt-solver 0.005"
r() { echo "== $1"; shift; env "$@" timeout 300 $B --envs ${ENVS:-64} 2>&1 | tail -1 | cut -c1-190; }
r base A=1
r tiled16 RBC_B200_G3_TILED=16
r fft512 RBC_B200_G3_FFT_THREADS=512
r fft128 RBC_B200_G3_FFT_THREADS=128
ENVS=111 r envs111 A=1
ENVS=74 r envs74 A=1
ENVS=148 r envs148 A=1
