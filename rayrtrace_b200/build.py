"""Builds the C-ABI CUDA library in-tree (replaces the reference's setup.py:5-12).

    python -m rayrtrace_b200.build          # nvcc, sm_100a only

The output, rayrtrace_b200/_lib/libsidtrace_b200.so, is git-ignored but travels
to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "sidtrace_b200.cu")
SRC_RK = os.path.join(_HERE, "csrc", "rk_b200.cu")      # adaptive-RK physics mode (extension)
SRC_PLAN = os.path.join(_HERE, "csrc", "plan_b200.cu")  # exactly-sized trajectory arenas (CUB sort + scan)
SRC_COMM = os.path.join(_HERE, "csrc", "comm_b200.cu")  # multi-GPU entry points (NCCL via dlopen)
SRC_FIELD = os.path.join(_HERE, "csrc", "field_b200.cu")  # O(rays) tail of Beam.field_r / field_result
SOURCES = [SRC, SRC_RK, SRC_PLAN, SRC_COMM, SRC_FIELD]
DEPS = SOURCES + [os.path.join(_HERE, "csrc", h) for h in ("trace_kernels.cuh", "launch_kernels.cuh", "rk_kernels.cuh")] + [
    os.path.join(os.path.dirname(_HERE), "include", h) for h in ("sidtrace_b200.h", "sidtrace_b200_rk.h", "sidtrace_b200_comm.h")]
LIB_DIR = os.path.join(_HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libsidtrace_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # parity: no FMA contraction anywhere (kernels use explicit _rn intrinsics too)
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
    "--threads", "5",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the sidtrace_b200 CUDA library cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    # SIDTRACE_B200_NVCC_EXTRA / SIDTRACE_B200_LIB: tuning sweeps build variants (-DSTB_...=...) side by side
    extra = os.environ.get("SIDTRACE_B200_NVCC_EXTRA", "").split()
    out = os.environ.get("SIDTRACE_B200_LIB", LIB) if extra else LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    # the image exports CC=/opt/gcc/bin/gcc; nvcc picks its host compiler from PATH
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
