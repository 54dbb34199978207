"""ctypes wrapper of the CPU emulator of the CUDA kernel core (test infrastructure)."""
import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_ROOT = _HERE.parent.parent
_SO = _HERE / "libemu_rbc2d.so"


class HostConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ra", "pr", "lx", "lz", "b_top", "heater_limit", "dt_action", "dt_solver", "episode_length")] + \
               [(n, C.c_int) for n in ("heaters", "obs_nz", "obs_nx", "channels")] + [("cfl_limit", C.c_double)]


class HostWrappers(C.Structure):
    _fields_ = [("normalize_obs", C.c_int), ("obs_clip", C.c_int), ("obs_lo", C.c_float * 4), ("obs_hi", C.c_float * 4),
                ("obs_maxval", C.c_float), ("normalize_reward", C.c_int), ("reward_scale", C.c_double), ("shaping", C.c_int),
                ("shaping_weight", C.c_double)]


_SOX = _HERE / "libemu_rbc2dx.so"


def _compile(so, cpps, deps):
    if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in list(cpps) + list(deps)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(so), *map(str, cpps)], check=True)
    return so


def build():
    csrc = _ROOT / "rbc_gym_b200" / "csrc"
    return _compile(_SO, [_HERE / "emu_rbc2d.cpp", _HERE / "emu_rbc3d.cpp"], [csrc / "rbc2d_core.h", csrc / "rbc3d_core.h"])


def build_x():
    """The cluster kernel for generic grids (rbc2dx_core.h), emulated."""
    csrc = _ROOT / "rbc_gym_b200" / "csrc"
    return _compile(_SOX, [_HERE / "emu_rbc2dx.cpp"], [csrc / "rbc2d_core.h", csrc / "rbc2dx_core.h"])


_SOG = _HERE / "libemu_rbc3dg.so"


def build_g():
    """The stage-streaming 3D kernels for arbitrary grids (rbc3dg_core.h), emulated."""
    csrc = _ROOT / "rbc_gym_b200" / "csrc"
    return _compile(_SOG, [_HERE / "emu_rbc3dg.cpp"], [csrc / "rbc2d_core.h", csrc / "rbc3dg_core.h"])


def step3g(state, actions, ra, shape, precision=64, heater_duration=0.125, dt_solver=0.01, heaters=8, heater_limit=0.9, project_first=False,
           nsub=-1, ra_env=None, domain=(2.0, 4 * math.pi, 4 * math.pi)):
    """state: [B, nstate] (b,u,v,w flattened, each [z][y][x]) on the grid `shape` = (nz, ny, nx); actions [B, 8, 8]."""
    lib = C.CDLL(str(build_g()))
    nz, ny, nx = shape
    B = state.shape[0]
    dt = np.float64 if precision == 64 else np.float32
    st = np.array(state, dtype=dt, order="C")
    h = HostConfig3(ra, 0.7, domain[2], domain[1], domain[0], 1.0, 1.0, heater_limit, heater_duration, dt_solver, 300.0, heaters)
    a = np.ascontiguousarray(actions, dtype=np.float32)
    nu, nf = np.zeros(B), np.zeros(B, np.int32)
    re = None if ra_env is None else np.ascontiguousarray(ra_env, dtype=np.float64)
    vp = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
    rc = lib.emu_rbc3dg_step(C.byref(h), nx, ny, nz, precision, B, vp(st), vp(a), vp(re), vp(nu), vp(nf), int(project_first), nsub)
    assert rc == 0, rc
    return dict(state=st, nusselt=nu, nan=nf)


class HostConfig3(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ra", "pr", "lx", "ly", "lz", "b_top", "delta_b", "heater_limit", "heater_duration",
                                          "dt_solver", "episode_length")] + [("heaters", C.c_int)]


def step3(state, actions, ra, precision=64, split=False, heater_duration=0.125, dt_solver=0.01, heaters=8, heater_limit=0.9,
          episode_length=300.0, project_first=False, nsub=-1, t0=None):
    """state: [B, 66560] (b,u,v,w flattened, each [z][y][x]); actions [B, 8, 8]."""
    lib = C.CDLL(str(build()))
    B = state.shape[0]
    dt = np.float64 if precision == 64 else np.float32
    st = np.array(state, dtype=dt, order="C")
    h = HostConfig3(ra, 0.7, 4 * math.pi, 4 * math.pi, 2.0, 1.0, 1.0, heater_limit, heater_duration, dt_solver, episode_length, heaters)
    a = np.ascontiguousarray(actions, dtype=np.float32)
    ob = np.zeros((B, 4, 16, 32, 32), np.float32)
    rew = np.zeros(B, np.float32)
    nu = np.zeros(B)
    t = np.zeros(B) if t0 is None else np.array(t0, dtype=np.float64)
    sc, tr, nf = np.ones(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    rc = lib.emu_rbc3d_step(C.byref(h), precision, 2 if split == "tiled" else int(split), B, vp(st), vp(a), vp(ob), vp(rew), vp(nu), vp(t), vp(sc), vp(tr), vp(nf),
                            int(project_first), nsub)
    assert rc == 0
    return dict(state=st, obs=ob, reward=rew, nusselt=nu, t=t, step=sc, truncated=tr, nan=nf)


def pack3(b, u, v, w):
    return np.concatenate([x.reshape(x.shape[0], -1) for x in (b, u, v, w)], axis=1)


def unpack3(st):
    B = st.shape[0]
    n = 16 * 32 * 32
    return (st[:, :n].reshape(B, 16, 32, 32), st[:, n:2 * n].reshape(B, 16, 32, 32), st[:, 2 * n:3 * n].reshape(B, 16, 32, 32),
            st[:, 3 * n:].reshape(B, 17, 32, 32))


def step(state, actions, ra, dt_action, precision=64, split=False, nxt_global=False, dt_solver=0.03,
         obs=(8, 48), heaters=12, heater_limit=0.75, episode_length=300.0, pressure=False, t0=None, wrappers=None):
    """state: [B, 18528] (b,u,w flattened) in the given precision; returns dict."""
    lib = C.CDLL(str(build()))
    B = state.shape[0]
    dt = np.float64 if precision == 64 else np.float32
    st = np.array(state, dtype=dt, order="C")
    ch = 5 if pressure else 3
    h = HostConfig(ra, 0.7, 2 * math.pi, 2.0, 1.0, heater_limit, dt_action, dt_solver, episode_length, heaters, obs[0], obs[1], ch)
    a = np.ascontiguousarray(actions, dtype=np.float32)
    ob = np.zeros((B, ch, obs[0], obs[1]), np.float32)
    rew = np.zeros(B, np.float32)
    nus, nuo = np.zeros(B), np.zeros(B)
    t = np.zeros(B) if t0 is None else np.array(t0, dtype=np.float64)
    sc, tr, nf = np.ones(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
    pr = np.zeros((B, 2, 64, 96), dt) if pressure else None
    vp = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
    cd = np.zeros(B)
    rc = lib.emu_rbc2d_step(C.byref(h), C.byref(wrappers) if wrappers is not None else None, vp(cd), precision, int(split), int(nxt_global), B, vp(st), vp(a), vp(ob), vp(rew), vp(nus), vp(nuo),
                            vp(t), vp(sc), vp(tr), vp(nf), vp(pr))
    assert rc == 0
    return dict(state=st, obs=ob, reward=rew, nu_state=nus, nu_obs=nuo, t=t, step=sc, truncated=tr, nan=nf, pressure=pr, cell_dist=cd)


class VecEmu2D:
    """B environments stepped by the emulated kernel WITH the fused vector-env semantics (`VecIO` in rbc2d_core.h): the CPU
    twin of `rbc2d_vec_step_dev`.  Holds the per-environment arrays the library handle owns on the device."""

    def __init__(self, B, bank, ra, dt_action, mode, nan_reset=False, seed=0, id_offset=0, precision=32, episode_length=300.0,
                 dt_solver=0.03, pressure=False, cluster=0):
        """cluster = CL > 0 runs the emulated CLUSTER kernel (rbc2dx_core.h, 96 x 64 split over CL CTAs) instead of the dedicated one."""
        self.lib = C.CDLL(str(build()))
        self.cluster = cluster
        self.libx = C.CDLL(str(build_x())) if cluster else None
        self.B, self.mode, self.nan_reset, self.seed, self.id_offset, self.precision = B, mode, int(nan_reset), seed, id_offset, precision
        self.dtype = np.float64 if precision == 64 else np.float32
        self.bank = np.ascontiguousarray(bank, dtype=np.float64)
        self.ch = 5 if pressure else 3
        self.split = int(pressure)
        self.h = HostConfig(ra, 0.7, 2 * math.pi, 2.0, 1.0, 0.75, dt_action, dt_solver, episode_length, 12, 8, 48, self.ch)
        self.state = np.zeros((B, 18528), self.dtype)
        self.t, self.step_count = np.zeros(B), np.ones(B, np.int32)
        self.pending, self.episode, self.ep_return = np.zeros(B, np.int32), np.zeros(B, np.int64), np.zeros(B)
        self.nan_count = np.zeros(1, np.int32)
        self.pressure = np.zeros((B, 2, 64, 96), self.dtype) if pressure else None
        self.lib.emu_checkpoint_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]

    def draw(self, env, episode):
        return self.lib.emu_checkpoint_draw(self.seed, self.id_offset + env, episode, self.bank.shape[0])

    def reset(self):
        for e in range(self.B):
            self.state[e] = self.bank[self.draw(e, 0)].astype(self.dtype)
        self.t[:] = 0; self.step_count[:] = 1; self.pending[:] = 0; self.episode[:] = 1; self.ep_return[:] = 0

    def step(self, actions):
        B = self.B
        a = np.ascontiguousarray(actions, dtype=np.float32)
        o = dict(obs=np.zeros((B, self.ch, 8, 48), np.float32), reward=np.zeros(B, np.float32), nu_state=np.zeros(B), nu_obs=np.zeros(B),
                 truncated=np.zeros(B, np.int32), nan=np.zeros(B, np.int32), final_obs=np.zeros((B, self.ch, 8, 48), np.float32),
                 final_nu_state=np.zeros(B), final_nu_obs=np.zeros(B), final_return=np.zeros(B))
        vp = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
        if self.cluster:
            self.libx.emu_rbc2dx_set_vec(self.mode, self.nan_reset, vp(self.bank), self.bank.shape[0], C.c_uint64(self.seed),
                                         C.c_uint64(self.id_offset), vp(self.pending), vp(self.episode), vp(self.ep_return),
                                         vp(o["final_obs"]), vp(o["final_nu_state"]), vp(o["final_nu_obs"]), vp(o["final_return"]),
                                         vp(self.nan_count))
            rc = self.libx.emu_rbc2dx_step(C.byref(self.h), None, None, 96, 64, self.cluster, self.precision, int(self.precision == 64), B,
                                           vp(self.state), vp(a), vp(o["obs"]), vp(o["reward"]), vp(o["nu_state"]), vp(o["nu_obs"]),
                                           vp(self.t), vp(self.step_count), vp(o["truncated"]), vp(o["nan"]), 0, -1, self.split, vp(self.pressure))
            self.libx.emu_rbc2dx_set_vec(-1, 0, None, 0, C.c_uint64(0), C.c_uint64(0), None, None, None, None, None, None, None, None)
            assert rc == 0
            o.update(t=self.t.copy(), step=self.step_count.copy(), episode_return=self.ep_return.copy())
            return o
        rc = self.lib.emu_rbc2d_vec_step(C.byref(self.h), self.precision, self.split, B, vp(self.state), vp(a), vp(o["obs"]), vp(o["reward"]),
                                         vp(o["nu_state"]), vp(o["nu_obs"]), vp(self.t), vp(self.step_count), vp(o["truncated"]), vp(o["nan"]),
                                         vp(self.pressure), self.mode, self.nan_reset, vp(self.bank), self.bank.shape[0],
                                         C.c_uint64(self.seed), C.c_uint64(self.id_offset), vp(self.pending), vp(self.episode),
                                         vp(self.ep_return), vp(o["final_obs"]), vp(o["final_nu_state"]), vp(o["final_nu_obs"]),
                                         vp(o["final_return"]), vp(self.nan_count))
        assert rc == 0
        o.update(t=self.t.copy(), step=self.step_count.copy(), episode_return=self.ep_return.copy())
        return o


def pack(b, u, w):
    return np.concatenate([b.reshape(b.shape[0], -1), u.reshape(u.shape[0], -1), w.reshape(w.shape[0], -1)], axis=1)


def unpack(st):
    B = st.shape[0]
    return st[:, :6144].reshape(B, 64, 96), st[:, 6144:12288].reshape(B, 64, 96), st[:, 12288:].reshape(B, 65, 96)


def stepx(state, actions, ra, dt_action, nx=96, nz=64, cl=1, precision=64, nxt_global=False, dt_solver=0.03, obs=(8, 48), heaters=12,
          heater_limit=0.75, episode_length=300.0, t0=None, wrappers=None, project_first=False, nsub=-1, split=False, cfl_limit=0.0):
    """One action step through the emulated CLUSTER kernel (rbc2dx_core.h): grid nx x nz split over `cl` CTAs.
    state: [B, 2*nx*nz + nx*(nz+1)] (b,u,w flattened).  split=True: pressure-split mode with the two pressure channels."""
    lib = C.CDLL(str(build_x()))
    B = state.shape[0]
    dt = np.float64 if precision == 64 else np.float32
    st = np.array(state, dtype=dt, order="C")
    ch = 5 if split else 3
    h = HostConfig(ra, 0.7, 2 * math.pi, 2.0, 1.0, heater_limit, dt_action, dt_solver, episode_length, heaters, obs[0], obs[1], ch, cfl_limit)
    a = np.ascontiguousarray(actions, dtype=np.float32)
    ob = np.zeros((B, ch, obs[0], obs[1]), np.float32)
    rew = np.zeros(B, np.float32)
    nus, nuo = np.zeros(B), np.zeros(B)
    t = np.zeros(B) if t0 is None else np.array(t0, dtype=np.float64)
    sc, tr, nf = np.ones(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
    pr = np.zeros((B, 2, nz, nx), dt) if split else None
    vp = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
    cd = np.zeros(B)
    rc = lib.emu_rbc2dx_step(C.byref(h), C.byref(wrappers) if wrappers is not None else None, vp(cd), nx, nz, cl, precision, int(nxt_global), B,
                             vp(st), vp(a), vp(ob), vp(rew), vp(nus), vp(nuo), vp(t), vp(sc), vp(tr), vp(nf), int(project_first), nsub, int(split), vp(pr))
    assert rc == 0, rc
    return dict(state=st, obs=ob, reward=rew, nu_state=nus, nu_obs=nuo, t=t, step=sc, truncated=tr, nan=nf, cell_dist=cd, pressure=pr)


def packx(b, u, w):
    return pack(b, u, w)


def unpackx(st, nx, nz):
    B = st.shape[0]
    n = nx * nz
    return st[:, :n].reshape(B, nz, nx), st[:, n:2 * n].reshape(B, nz, nx), st[:, 2 * n:].reshape(B, nz + 1, nx)
