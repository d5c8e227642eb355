"""In-tree build of librappas_b200.so (CUDA, sm_100a only).  `python -m rappas_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librappas_b200.so")
SOURCES = ["rp_db.cu", "rp_place.cu", "rp_dbbuild.cu", "rp_synthdb.cu", "rp_xchg.cu", "rp_ingest.cpp"]
HEADERS = ["rp_common.h", "rp_device.cuh", "rp_xchg_plan.h", "rp_synth.h", "rp_dbbuild_core.h", "rp_dbbuild_merge.h", os.path.join("..", "..", "include", "rappas_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden", "--expt-relaxed-constexpr",
    # bit-exact f32: the kernels spell every rounding with __f*_rn intrinsics; additionally forbid
    # mul+add contraction and keep IEEE division / sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not stale():
        return OUT
    cmd = [nvcc(), *NVCC_FLAGS, *extra, "-shared", "-o", OUT, *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    print(OUT)
