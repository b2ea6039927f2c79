"""In-tree build of the CUDA library (sm_100a only).  ``python -m openpystruct_b200.build``."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libopenpystruct_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("beamopt_kernels.cu", "beamopt_lanes.cu", "beamopt_lanes_tm.cu", "beamopt_wide.cu", "frameopt.cu", "sampler_host.cpp")]
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


def build(force: bool = False, verbose: bool = False, out: str = LIB_PATH, defines=(), sources=None) -> str:
    """Compiles the translation units in parallel (one nvcc per .cu) and links them into ``out``.
    ``defines``: extra -D macros (A/B variants of the kernels, loaded with OPS_B200_LIB)."""
    if not force and out == LIB_PATH and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj_" + os.path.splitext(os.path.basename(out))[0])
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = nvcc_path()
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines]

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        cmd = [nvcc, *flags, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, sources or SOURCES))
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs], check=True)
    return out


if __name__ == "__main__":
    # python -m openpystruct_b200.build [--force] [--out lib/variant.so] [-DNAME[=VALUE] ...]
    _out = LIB_PATH
    if "--out" in sys.argv:
        _out = os.path.abspath(sys.argv[sys.argv.index("--out") + 1])
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv or _out != LIB_PATH, verbose="--quiet" not in sys.argv, out=_out, defines=_defs))
