"""In-tree build of the sm_100a shared library (`rbc_gym_b200/csrc/librbc_b200.so`).

nvcc cross-compiles without a GPU; the built `.so` is git-ignored but travels with the repo
snapshot to the GPU box.  Run as ``python -m rbc_gym_b200.build [-v]``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "librbc_b200.so"
SOURCES = [CSRC / "rbc2d_lib.cu", CSRC / "rbc2dx_lib.cu", CSRC / "rbc2dx_split.cu", CSRC / "rbc2dx_more.cu", CSRC / "rbc2dx_more_split.cu", CSRC / "rbc3d_lib.cu",
           CSRC / "rbc3dg_lib.cu"]
HEADERS = [CSRC / "rbc2d_core.h", CSRC / "rbc2dx_core.h", CSRC / "rbc2dx_api.h", CSRC / "rbc2dx_kernel.cuh", CSRC / "rbc2dx_more.cuh", CSRC / "rbc3d_core.h", CSRC / "rbc3dg_core.h", CSRC / "rbc3dg_api.h", CSRC / "rbc_common.h", CSRC.parent.parent / "include" / "rbc_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = CSRC / "build"


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build rbc_gym_b200")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Build the library if a source is newer than it.  Safe to call from several processes at once (one rank per GPU
    under torchrun): the check and the build run under an exclusive file lock, objects and the library are written to
    temporary names and renamed into place, so no process can ever dlopen a half-written file."""
    import fcntl
    OBJ_DIR.mkdir(exist_ok=True)
    with open(OBJ_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    if not force and not needs_build():
        return LIB
    base = [find_nvcc(), *NVCC_FLAGS]
    if verbose:
        base += ["-Xptxas", "-v"]
    # $CC/$CXX in this image point at a wrapper compiler; let nvcc use the system g++
    host = shutil.which("g++")
    if host:
        base += ["-ccbin", host]
    newest_header = max(p.stat().st_mtime for p in HEADERS)

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, newest_header):
            return obj
        tmp = obj.with_suffix(f".{os.getpid()}.tmp.o")
        cmd = base + ["-c", "-o", str(tmp), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"nvcc failed ({res.returncode}): {' '.join(cmd)}")
        os.replace(tmp, obj)
        return obj

    # one translation unit per kernel family, compiled concurrently (the unrolled kernels take ~30-60 s each)
    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp_lib = LIB.with_suffix(f".{os.getpid()}.tmp.so")
    cmd = base + ["-shared", "-o", str(tmp_lib), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        tmp_lib.unlink(missing_ok=True)
        raise RuntimeError(f"nvcc link failed ({res.returncode}): {' '.join(cmd)}")
    os.replace(tmp_lib, LIB)
    return LIB


LIB_JITTER = CSRC / "librbc_b200_jitter.so"


def build_jitter(force: bool = False) -> Path:
    """The stress build of the cluster kernels (`-DRBX_JITTER`: random delays in front of every DSMEM push and mbarrier wait,
    see rbc2dx_core.h) linked with the normal objects of the other translation units.  Test infrastructure for
    tests/test_gpu_grid192.py; select it with RBC_B200_LIB=<path> before the package is imported."""
    import fcntl
    build()
    OBJ_DIR.mkdir(exist_ok=True)
    with open(OBJ_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            newest = max(p.stat().st_mtime for p in SOURCES + HEADERS)
            if not force and LIB_JITTER.exists() and LIB_JITTER.stat().st_mtime > newest:
                return LIB_JITTER
            base = [find_nvcc(), *NVCC_FLAGS]
            host = shutil.which("g++")
            if host:
                base += ["-ccbin", host]
            dx = [s for s in SOURCES if s.name in ("rbc2dx_lib.cu", "rbc2dx_split.cu")]      # the grids the stress test runs

            def one(src: Path) -> Path:
                obj = OBJ_DIR / (src.stem + ".jitter.o")
                res = subprocess.run(base + ["-DRBX_JITTER=1", "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError(f"nvcc failed on the jitter build of {src.name}")
                return obj

            with ThreadPoolExecutor(len(dx)) as ex:
                jobjs = list(ex.map(one, dx))
            others = [OBJ_DIR / (s.stem + ".o") for s in SOURCES if s not in dx]
            tmp = LIB_JITTER.with_suffix(f".{os.getpid()}.tmp.so")
            res = subprocess.run(base + ["-shared", "-o", str(tmp), *map(str, jobjs + others)], capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc link failed for the jitter build")
            os.replace(tmp, LIB_JITTER)
            return LIB_JITTER
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
