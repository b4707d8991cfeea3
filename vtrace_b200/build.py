"""Builds vtrace_b200/librender.so (sm_100a only) with nvcc, in-tree.

    python -m vtrace_b200.build [--force] [--verbose]

Flags that matter for parity: --fmad=false (no FMA contraction), IEEE
division / square root, no flush-to-zero, no fast-math; the host compiler gets
-ffp-contract=off for the per-frame uniform matrices.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "librender.so")
SOURCES = ["kernels.cu", "render_abi.cu"]
HEADERS = [os.path.join(CSRC, "kernels.h"), os.path.join(CSRC, "paths_wave.cuh"), os.path.join(CSRC, "bricks.cuh"), os.path.join(CSRC, "world_grid.cuh"), os.path.join(ROOT, "include", "vtrace_abi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
    "-shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building librender.so")
    with open(os.path.join(PKG, "librender.build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
