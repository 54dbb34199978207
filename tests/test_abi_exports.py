"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/rbc_b200.h declares (no compute calls without a GPU), and fails loudly without one."""
import ctypes
import re
from pathlib import Path

import pytest

from rbc_gym_b200 import backend, build

ROOT = Path(__file__).resolve().parent.parent


def test_library_builds_and_exports_every_declared_symbol():
    lib = build.build()
    assert lib.exists()
    L = ctypes.CDLL(str(lib))
    header = (ROOT / "include" / "rbc_b200.h").read_text()
    declared = set(re.findall(r"\b(rbc(?:2d|3d)?_[a-z0-9_]+)\s*\(", header))
    declared = {d for d in declared if not d.endswith("_config")}
    assert declared == set(backend.ABI_SYMBOLS), declared ^ set(backend.ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), f"{sym} not exported"
    assert L.rbc_abi_version() == backend.ABI_VERSION == 2


def test_config_struct_matches_header():
    header = (ROOT / "include" / "rbc_b200.h").read_text()
    body = header[header.index("typedef struct rbc2d_config {"):header.index("} rbc2d_config;")]
    names = []
    for typ, decl in re.findall(r"^\s*(int32_t|double)\s+([^;]+);", body, flags=re.M):
        names += [(n.strip(), typ) for n in decl.split(",")]
    fields = [(n, "int32_t" if t is ctypes.c_int32 else "double") for n, t in backend.Rbc2dConfig._fields_]
    assert names == fields


@pytest.mark.parametrize("name,cls", [("rbc_autoreset", "RbcAutoreset"), ("rbc2d_vec_out", "Rbc2dVecOut"), ("rbc3d_vec_out", "Rbc3dVecOut")])
def test_vector_step_structs_match_header(name, cls):
    header = (ROOT / "include" / "rbc_b200.h").read_text()
    body = header[header.index(f"typedef struct {name} {{"):header.index(f"}} {name};")]
    names = [n.strip() for _, n in re.findall(r"^\s*(int32_t|int64_t|float\*|double\*|int32_t\*)\s+([a-z_]+);", body, flags=re.M)]
    assert names == [n for n, _ in getattr(backend, cls)._fields_]


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(backend.BackendUnavailable):
        backend.Sim2D(4, 1e5)
    # and at the ABI level
    L = backend.load_library()
    cfg = backend.Rbc2dConfig(4, 96, 64, 48, 8, 12, 0.75, 1e5, 0.7, 1.0, 0.03, 300.0, 32, 0, 0)
    h = ctypes.c_void_p()
    assert L.rbc2d_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CUDA device" in L.rbc_last_error()


def test_product_package_never_touches_the_oracle():
    """Nothing in the product package may import, link or execute anything under oracle/ or tests/."""
    pat = re.compile(r"^\s*(from|import)\s+(oracle|tests)\b|#include\s+\"[^\"]*(oracle|tests)/|CDLL\([^)]*oracle", re.M)
    for p in (ROOT / "rbc_gym_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".h", ".cuh"):
            assert not pat.search(p.read_text()), p
