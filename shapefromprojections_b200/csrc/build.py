"""Build libctdr_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m shapefromprojections_b200.csrc.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/ctdr_b200.h): no torch headers, no Python.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libctdr_b200.so")
SOURCES = ["ctdr_b200.cu"]
DEPS = ["ctdr_common.cuh", "ctdr_raster.cuh", "ctdr_vertex.cuh", "ctdr_raster64.cuh", "ctdr_vertex64.cuh", "ctdr_peer.cuh", os.path.join(REPO, "include", "ctdr_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC",
              "--use_fast_math=false"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(HERE, s) for s in SOURCES] + [d if os.path.isabs(d) else os.path.join(HERE, d) for d in DEPS]
    return any(os.path.getmtime(f) > t for f in files + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [nvcc_path(), *flags, "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", LIB, *[os.path.join(HERE, s) for s in SOURCES]]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libctdr_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
