"""ctypes binding + numpy helpers for the CPU oracle (`oracle/rbc2d_oracle.c`).

TEST INFRASTRUCTURE ONLY — see the header of the C file.  Importable from ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``; never from the
product package ``rbc_gym_b200``.

Parity pin: fixtures shipped by the reference (checkpoint HDF5 files); against Julia output
itself parity is unpinned (the reference cannot run here).
"""
from __future__ import annotations

import ctypes as C
import math
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "librbc_oracle.so"


class Params(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("nz", C.c_int),
        ("lx", C.c_double), ("lz", C.c_double),
        ("nu", C.c_double), ("kappa", C.c_double),
        ("b_top", C.c_double),
        ("heaters", C.c_int), ("heater_limit", C.c_double),
        ("split_phy", C.c_int), ("poisson_mode", C.c_int),
    ]


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only, no external libraries)."""
    src = _HERE / "rbc2d_oracle.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB))
        dp = C.POINTER(C.c_double)
        L.rbc2d_oracle_heater_profile.argtypes = [C.POINTER(Params), dp, dp]
        L.rbc2d_oracle_step.argtypes = [C.POINTER(Params), dp, dp, dp, dp, C.c_int, dp, dp, dp]
        L.rbc2d_oracle_step.restype = C.c_int
        L.rbc2d_oracle_step_batch.argtypes = [C.POINTER(Params), C.c_int, dp, dp, dp, dp, C.c_int, dp, C.c_int]
        L.rbc2d_oracle_step_batch.restype = C.c_int
        L.rbc2d_oracle_tendencies.argtypes = [C.POINTER(Params), dp, dp, dp, dp, dp, dp, dp]
        L.rbc2d_oracle_project.argtypes = [C.POINTER(Params), dp, dp, dp]
        L.rbc2d_oracle_nusselt.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.rbc2d_oracle_nusselt.restype = C.c_double
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(ra: float, nx: int = 96, nz: int = 64, pr: float = 0.7, heaters: int = 12,
                heater_limit: float = 0.75, split_phy: bool = True, poisson_mode: int = 0) -> Params:
    """Constants hard-coded in `rbc_sim2D_api.jl:28-41`: L=[2pi,2], Pr=0.7, min_b=1."""
    return Params(nx, nz, 2 * math.pi, 2.0, math.sqrt(pr / ra), 1.0 / math.sqrt(pr * ra), 1.0,
                  heaters, heater_limit, int(split_phy), poisson_mode)


def substep_schedule(dt_action: float, dt_solver: float = 0.03) -> np.ndarray:
    """run!'s `dt' = min(dt_solver, stop_time - t)` (SURVEY §8a "run! semantics"), from t = 0.

    dt=1 -> 33 x 0.03 + one clipped step; dt=1.5 -> 50 x 0.03.  A remainder below 1e-10 is dropped
    (documented deviation: Julia may take one extra ~1e-16 step from clock round-off).
    """
    out, t = [], 0.0
    while dt_action - t > 1e-10:
        d = min(dt_solver, dt_action - t)
        out.append(d)
        t += d
    return np.asarray(out, dtype=np.float64)


def heater_profile(P: Params, action) -> np.ndarray:
    a = np.ascontiguousarray(action, dtype=np.float64)
    out = np.empty(P.nx)
    lib().rbc2d_oracle_heater_profile(C.byref(P), _p(a), _p(out))
    return out


def step(P: Params, b, u, w, action, dts, want_pressure: bool = False):
    """One action step on copies of (b[nz,nx], u[nz,nx], w[nz+1,nx]); returns dict of new fields."""
    b, u, w = (np.array(x, dtype=np.float64, order="C") for x in (b, u, w))
    a = np.ascontiguousarray(action, dtype=np.float64)
    dts = np.ascontiguousarray(dts, dtype=np.float64)
    phy = np.empty((P.nz, P.nx)) if want_pressure else None
    pn = np.empty((P.nz, P.nx)) if want_pressure else None
    bad = lib().rbc2d_oracle_step(C.byref(P), _p(b), _p(u), _p(w), _p(a), len(dts), _p(dts),
                                  _p(phy) if want_pressure else None, _p(pn) if want_pressure else None)
    return {"b": b, "u": u, "w": w, "nan": bool(bad), "phy": phy, "pnhs": pn}


def step_batch(P: Params, b, u, w, actions, dts, threads: int = 1) -> bool:
    """In-place batched step over env axis 0; Python threads (ctypes drops the GIL) over env chunks."""
    B = b.shape[0]
    dts = np.ascontiguousarray(dts, dtype=np.float64)
    L = lib()

    def run(lo, hi):
        return L.rbc2d_oracle_step_batch(C.byref(P), hi - lo, _p(b[lo:hi]), _p(u[lo:hi]), _p(w[lo:hi]),
                                         _p(actions[lo:hi]), len(dts), _p(dts), 1)

    if threads <= 1:
        return bool(run(0, B))
    edges = np.linspace(0, B, threads + 1).astype(int)
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(lambda i: run(edges[i], edges[i + 1]) if edges[i + 1] > edges[i] else 0, range(threads)))
    return bool(any(res))


def tendencies(P: Params, b, u, w, action):
    b, u, w = (np.ascontiguousarray(x, dtype=np.float64) for x in (b, u, w))
    a = np.ascontiguousarray(action, dtype=np.float64)
    Gb, Gu, Gw = np.empty_like(b), np.empty_like(u), np.empty_like(w)
    lib().rbc2d_oracle_tendencies(C.byref(P), _p(b), _p(u), _p(w), _p(a), _p(Gb), _p(Gu), _p(Gw))
    return Gb, Gu, Gw


def project(P: Params, u, w):
    u, w = (np.array(x, dtype=np.float64, order="C") for x in (u, w))
    p = np.empty((P.nz, P.nx))
    lib().rbc2d_oracle_project(C.byref(P), _p(u), _p(w), _p(p))
    return u, w, p


def nusselt(T, wv, kappa: float, db: float = 1.0, H: float = 2.0) -> float:
    """`get_nusselt` (`rbc_sim2D_api.jl:142-163`) on arrays laid out [z, x]."""
    T = np.ascontiguousarray(T, dtype=np.float64)
    wv = np.ascontiguousarray(wv, dtype=np.float64)
    return lib().rbc2d_oracle_nusselt(_p(T), _p(wv), T.shape[0], T.shape[1], kappa, db, H)


def state_channels(b, u, w):
    """`get_state` (`rbc_sim2D_api.jl:102-118`) in the Python layout (C, z, x): b, u, w[faces 0..nz-1]."""
    return np.stack([b, u, w[:-1]], axis=0)


def observe(state, obs_shape=(8, 48)):
    """`get_observation` (`rbc_sim2D_api.jl:123-129`): strided sub-sample 1:N/No:N in each direction."""
    nz, nx = state.shape[-2:]
    return state[..., :: nz // obs_shape[0], :: nx // obs_shape[1]]


def nusselt_state_obs(P: Params, b, u, w, obs_shape=(8, 48)):
    s = state_channels(b, u, w)
    o = observe(s, obs_shape)
    return nusselt(s[0], s[2], P.kappa), nusselt(o[0], o[2], P.kappa)


def kinetic_energy(u, w) -> float:
    return 0.5 * float(np.mean(u * u)) + 0.5 * float(np.mean(w * w))
