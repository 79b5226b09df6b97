"""Build the sm_100a CUDA library in-tree (nvcc cross-compiles without a GPU).

    python -m encodec_pytorch_b200.build

The output ``librvq_b200.so`` sits next to this file so it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librvq_b200.so")
SOURCES = ["rvq_abi.cu", "rvq_simt.cu", "rvq_bits.cu", "rvq_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "rvq_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    defs = [f"-D{d}" for d in os.environ.get("RVQ_NVCC_DEFS", "").split() if d]   # e.g. RVQ_TC_TRACE RVQ_TC_TIMERS
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if os.environ.get("RVQ_TC_SRC"):                       # A/B runs: another revision of the tcgen05 kernel file
        srcs[-1] = os.path.join(CSRC, os.environ["RVQ_TC_SRC"])
    cmd = [nvcc, *NVCC_FLAGS, *defs, "-o", LIB, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
