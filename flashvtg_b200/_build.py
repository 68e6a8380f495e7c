"""In-tree nvcc build of libflashvtg_b200.so (sm_100a only).

The library has no torch / pybind dependency: it is a plain C-ABI shared object
(include/flashvtg_b200.h) linked against the shared CUDA runtime (the one torch already loaded), so the same
file serves the Python host code (ctypes) and any other FFI.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libflashvtg_b200.so"
OBJ_DIR = PKG_DIR / "build"
STAMP = PKG_DIR / "build" / "sources.sha256"
# the debug flavour: the same sources with -DFVTG_DEBUG_HOOKS (fvtg_dbg_* test / trace hooks + csrc/probe.cu
# micro-benchmarks); loaded only by the kernel-level unit tests and tools/, never by the product path
DBG_LIB_PATH = PKG_DIR / "libflashvtg_b200_dbg.so"
DBG_OBJ_DIR = PKG_DIR / "build" / "dbg"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else "nvcc"


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) +
                    [PKG_DIR.parent / "include" / "flashvtg_b200.h", PKG_DIR.parent / "include" / "flashvtg_b200_dbg.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return (LIB_PATH.exists() and DBG_LIB_PATH.exists() and STAMP.exists()
            and STAMP.read_text().strip() == _digest())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libflashvtg_b200.so (the product) and
    libflashvtg_b200_dbg.so (product + debug hooks + probes) in-tree."""
    if not force and is_fresh():
        return LIB_PATH
    OBJ_DIR.mkdir(exist_ok=True)
    DBG_OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    srcs = _sources()

    def compile_one(job) -> Path:
        src, dbg = job
        odir = DBG_OBJ_DIR if dbg else OBJ_DIR
        obj = odir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(["-DFVTG_DEBUG_HOOKS"] if dbg else []), "-Xptxas", "-v", "-c", str(src),
               "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        (odir / (src.stem + ".ptxas.log")).write_text(r.stderr)
        if verbose and not dbg:
            sys.stderr.write(r.stderr)
        return obj

    jobs = [(s, False) for s in srcs if s.stem != "probe"] + [(s, True) for s in srcs]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        objs = list(ex.map(compile_one, jobs))
    for lib, sel in ((LIB_PATH, [o for o, j in zip(objs, jobs) if not j[1]]),
                     (DBG_LIB_PATH, [o for o, j in zip(objs, jobs) if j[1]])):
        cmd = [nvcc, "-shared", "-o", str(lib), *map(str, sel), "-cudart", "shared"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    STAMP.write_text(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
