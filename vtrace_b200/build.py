"""Builds vtrace_b200/librender.so and vtrace_b200/librender.a (sm_100a only) with nvcc, in-tree.

    python -m vtrace_b200.build [--force] [--verbose]

Flags that matter for parity: --fmad=false (no FMA contraction), IEEE
division / square root, no flush-to-zero, no fast-math; the host compiler gets
-ffp-contract=off for the per-frame uniform matrices.

Every source is compiled to its own object (in parallel, only when stale); the objects are linked
into the shared library the ctypes / C++ hosts load, and archived into a static `librender.a` — the
form the reference's build script links (`cargo:rustc-link-lib=static=render`, build.rs:96-97).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
LIB = os.path.join(PKG, "librender.so")
STATIC_LIB = os.path.join(PKG, "librender.a")
SOURCES = ["kernels.cu", "render_abi.cu", "peaks.cu"]


def headers() -> list[str]:
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))] + [
        os.path.join(ROOT, "include", "vtrace_abi.h")]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    common = headers() + [os.path.abspath(__file__)]
    return any(_stale(_obj(s), [os.path.join(CSRC, s)] + common) for s in SOURCES) or \
        any(_stale(out, [_obj(s) for s in SOURCES if os.path.exists(_obj(s))]) for out in (LIB, STATIC_LIB))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    common = headers() + [os.path.abspath(__file__)]
    log = []

    def compile_one(src: str):
        out = _obj(src)
        if not force and not _stale(out, [os.path.join(CSRC, src)] + common):
            return 0, f"(up to date) {src}\n"
        cmd = [nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-c", "-o", out, os.path.join(CSRC, src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return proc.returncode, " ".join(cmd) + "\n" + proc.stdout + proc.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = False
    for rc, text in results:
        log.append(text)
        failed = failed or rc != 0
    if not failed:
        objs = [_obj(s) for s in SOURCES]
        link = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-shared", "-o", LIB, *objs]
        proc = subprocess.run(link, capture_output=True, text=True)
        log.append(" ".join(link) + "\n" + proc.stdout + proc.stderr)
        failed = proc.returncode != 0
        if not failed:
            if os.path.exists(STATIC_LIB):
                os.remove(STATIC_LIB)
            ar = ["ar", "rcs", STATIC_LIB, *objs]
            proc = subprocess.run(ar, capture_output=True, text=True)
            log.append(" ".join(ar) + "\n" + proc.stdout + proc.stderr)
            failed = proc.returncode != 0
    text = "".join(log)
    if verbose or failed:
        sys.stderr.write(text)
    if failed:
        raise RuntimeError("nvcc failed building librender")
    with open(os.path.join(PKG, "librender.build.log"), "w") as f:
        f.write(text)
    return LIB


def build_variant(name: str, defs: list[str]) -> str:
    """Side-by-side build with extra -D flags into variants/NAME/librender.so (git-ignored; for A/B timing via VT_LIBRENDER)."""
    out_dir = os.path.join(ROOT, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    objs = []

    def compile_one(src: str):
        out = os.path.join(out_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, *defs, "-ccbin", "/usr/bin/g++", "-c", "-o", out, os.path.join(CSRC, src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(proc.stdout + proc.stderr)
        return out, proc.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    with open(os.path.join(out_dir, "build.log"), "w") as f:
        f.write("".join(t for _, t in results))
    lib = os.path.join(out_dir, "librender.so")
    subprocess.run([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-shared", "-o", lib, *objs], check=True)
    for o in objs:
        os.remove(o)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
