"""In-tree build of the CUDA library (sm_100a only).  ``python -m openpystruct_b200.build``."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libopenpystruct_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("beamopt_kernels.cu", "beamopt_lanes.cu", "beamopt_wide.cu")]
HEADERS = [os.path.join(_PKG, "csrc", f) for f in ("beamopt_core.cuh", "beamopt_flex.cuh", "beamopt_lanes.cuh", "beamopt_wide.cuh",
                                                   "beamopt_internal.cuh", "fastmath.cuh")] + \
          [os.path.join(os.path.dirname(_PKG), "include", "openpystruct_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # one IEEE rounding per written operation; FMAs are explicit (torch parity)
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
