"""ctypes binding of the 3D CPU oracle (`oracle/rbc3d_oracle.c`).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import oracle as O2


class Params3(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double),
                ("nu", C.c_double), ("kappa", C.c_double), ("b_top", C.c_double), ("delta_b", C.c_double),
                ("heaters", C.c_int), ("heater_limit", C.c_double), ("split_phy", C.c_int)]


_ready = False


def lib():
    global _ready
    L = O2.lib()
    if not _ready:
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(Params3)
        L.rbc3d_oracle_preprocess_action.argtypes = [pp, dp, dp]
        L.rbc3d_oracle_heater_profile.argtypes = [pp, dp, dp]
        L.rbc3d_oracle_step.argtypes = [pp, dp, dp, dp, dp, dp, C.c_int, dp]
        L.rbc3d_oracle_step.restype = C.c_int
        L.rbc3d_oracle_tendencies.argtypes = [pp] + [dp] * 9
        L.rbc3d_oracle_project.argtypes = [pp, dp, dp, dp]
        L.rbc3d_oracle_nusselt.argtypes = [pp, dp, dp]
        L.rbc3d_oracle_nusselt.restype = C.c_double
        _ready = True
    return L


_p = O2._p


def make_params(ra: float, shape=(16, 32, 32), domain=(2.0, 4 * math.pi, 4 * math.pi), pr: float = 0.7, heaters: int = 8,
                heater_limit: float = 0.9, t_diff=(1.0, 2.0), split_phy: bool = True) -> Params3:
    """shape/domain in the Python order of the reference env (`rbc3D.py:43-51`): (z, y, x)."""
    nz, ny, nx = shape
    lz, ly, lx = domain
    return Params3(nx, ny, nz, lx, ly, lz, math.sqrt(pr / ra), 1.0 / math.sqrt(pr * ra), t_diff[0], t_diff[1] - t_diff[0],
                   heaters, heater_limit, int(split_phy))


def substep_schedule(heater_duration: float = 0.125, dt_solver: float = 0.01, lz: float = 2.0) -> np.ndarray:
    """Simulation-time substeps: dt = dt_solver * t_ff, stop = heater_duration * t_ff, t_ff = Lz^2
    (`rbc_sim3D_api.jl:43,65`): defaults give 12 x 0.04 + 0.02."""
    t_ff = lz ** 2
    return O2.substep_schedule(heater_duration * t_ff, dt_solver * t_ff)


def preprocess_action(P: Params3, action) -> np.ndarray:
    a = np.ascontiguousarray(action, dtype=np.float64)
    out = np.empty_like(a)
    lib().rbc3d_oracle_preprocess_action(C.byref(P), _p(a), _p(out))
    return out


def heater_profile(P: Params3, action) -> np.ndarray:
    a = np.ascontiguousarray(action, dtype=np.float64)
    out = np.empty((P.ny, P.nx))
    lib().rbc3d_oracle_heater_profile(C.byref(P), _p(a), _p(out))
    return out


def step(P: Params3, b, u, v, w, action, dts):
    b, u, v, w = (np.array(x, dtype=np.float64, order="C") for x in (b, u, v, w))
    a = np.ascontiguousarray(action, dtype=np.float64)
    dts = np.ascontiguousarray(dts, dtype=np.float64)
    bad = lib().rbc3d_oracle_step(C.byref(P), _p(b), _p(u), _p(v), _p(w), _p(a), len(dts), _p(dts))
    return {"b": b, "u": u, "v": v, "w": w, "nan": bool(bad)}


def tendencies(P: Params3, b, u, v, w, action):
    b, u, v, w = (np.ascontiguousarray(x, dtype=np.float64) for x in (b, u, v, w))
    a = np.ascontiguousarray(action, dtype=np.float64)
    G = [np.empty_like(b), np.empty_like(u), np.empty_like(v), np.empty_like(w)]
    lib().rbc3d_oracle_tendencies(C.byref(P), _p(b), _p(u), _p(v), _p(w), _p(a), *[_p(g) for g in G])
    return G


def project(P: Params3, u, v, w):
    u, v, w = (np.array(x, dtype=np.float64, order="C") for x in (u, v, w))
    lib().rbc3d_oracle_project(C.byref(P), _p(u), _p(v), _p(w))
    return u, v, w


def nusselt(P: Params3, b, w) -> float:
    b, w = (np.ascontiguousarray(x, dtype=np.float64) for x in (b, w))
    return lib().rbc3d_oracle_nusselt(C.byref(P), _p(b), _p(w))


def divergence(P: Params3, u, v, w):
    dx, dy, dz = P.lx / P.nx, P.ly / P.ny, P.lz / P.nz
    return (np.roll(u, -1, axis=2) - u) / dx + (np.roll(v, -1, axis=1) - v) / dy + (w[1:] - w[:-1]) / dz
