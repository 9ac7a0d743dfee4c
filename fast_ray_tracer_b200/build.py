"""Build libfrt_b200.so in-tree (sm_100a only).  Run as `python -m fast_ray_tracer_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libfrt_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return [CSRC / "frt_core.cu", CSRC / "frt_blob.c"]


def headers():
    return sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(INCLUDE.glob("*.h"))


def up_to_date() -> bool:
    if not LIB.exists():
        return False
    t = LIB.stat().st_mtime
    return all(p.stat().st_mtime <= t for p in sources() + headers())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    blob_o = PKG / "csrc" / "frt_blob.o"
    subprocess.run(["gcc", "-O2", "-fPIC", "-std=c11", "-Wall", "-I", str(INCLUDE), "-c", str(CSRC / "frt_blob.c"), "-o", str(blob_o)],
                   check=True)
    extra = os.environ.get("FRT_NVCC_EXTRA", "").split()
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-o", str(LIB), str(CSRC / "frt_core.cu"), str(blob_o)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
