#!/usr/bin/env python
"""Memory / UB check of the kernel sources without a GPU: the host emulators (tests/emu) compiled with
-fsanitize=address,undefined and driven through every kernel variant (dedicated 2D incl. split and fp64, cluster 2D for
all registered grids and cluster sizes incl. the set! projection, 3D global / split / tiled).  The emulator buffers have
exactly the sizes of the shared-memory regions and per-CTA scratch, so an out-of-bounds index in the shared phase logic
is reported here.  (compute-sanitizer is closed on the GPU pool.)

    bash -c 'cd tests/emu && for f in "emu_rbc2dx.cpp -o /tmp/libemu_asan_x.so" "emu_rbc2d.cpp emu_rbc3d.cpp -o /tmp/libemu_asan.so"; do \
        g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer $f; done'
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_emulators.py
"""
import sys, numpy as np
sys.path.insert(0,'/root/repo')
from tests.emu import emu
from pathlib import Path
emu._SO = Path('/tmp/libemu_asan.so'); emu._SOX = Path('/tmp/libemu_asan_x.so')
emu.build = lambda: emu._SO
emu.build_x = lambda: emu._SOX
from rbc_gym_b200.h5lite import load_checkpoint_2d
from tests.gridstates import smooth_state
from tests.test_oracle3d import random_state
from oracle import oracle3d as O3
c = load_checkpoint_2d('/root/repo/data/checkpoints/train/ckpt_ra100000.h5')
act = np.linspace(-1,1,12).astype(np.float32)
st = emu.pack(c.b[:1], c.u[:1], c.w[:1])
for split, ng in ((False, False), (True, True)):
    for prec in (32, 64):
        emu.step(st, act[None], 1e5, 0.06, precision=prec, split=split, nxt_global=ng, pressure=split)
print('2d dedicated ok')
for cl in (1, 2, 4):
    for ng in (False, True):
        emu.stepx(st, act[None], 1e5, 0.06, cl=cl, precision=32, nxt_global=ng)
b,u,w = smooth_state(192,128)
for cl in (4, 8):
    emu.stepx(emu.pack(b[None],u[None],w[None]), act[None], 1e6, 0.03, nx=192, nz=128, cl=cl, precision=32, dt_solver=0.015, project_first=True)
b,u,w = smooth_state(128,64)
emu.stepx(emu.pack(b[None],u[None],w[None]), act[None], 1e5, 0.06, nx=128, nz=64, cl=2, precision=64, obs=(8,64))
print('2d cluster ok')
P = O3.make_params(5e3)
bb,uu,vv,ww = random_state(P, 1)
a = np.random.default_rng(2).uniform(-1,1,(8,8)).astype(np.float32)
for split in (False, True, 'tiled'):
    emu.step3(emu.pack3(bb[None],uu[None],vv[None],ww[None]), a[None], 5e3, precision=32, split=split, heater_duration=0.03)
print('3d ok')
