"""Builds libgten_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

-fmad=false is part of the numeric contract (gtb_dev.cuh): nvcc must never contract a*b+c, the
reference's AVX build has no FMA.  Explicit fmaf()/fma() calls stay fused.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libgten_b200.so"
SOURCES = ["gtb_api.cu", "gtb_ops.cu", "gtb_engine.cu", "gtb_prefill.cu", "gtb_xrows.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "gten_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> Path:
    if not force and not needs_build():
        return LIB
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    procs = []
    bdir = HERE / "build"
    bdir.mkdir(exist_ok=True)
    for s in SOURCES:
        o = bdir / (s + ".o")
        cmd = [_nvcc(), "-ccbin", ccbin, *NVCC_FLAGS, *(extra or []), "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-ccbin", ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout, r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    extra = (["-Xptxas", "-v"] if "--ptxas" in sys.argv else []) + (["-DGTB_FD_TRACE"] if "--trace" in sys.argv else [])
    build(force="--force" in sys.argv or "--trace" in sys.argv, verbose=True, extra=extra or None)
    print("built", LIB)
