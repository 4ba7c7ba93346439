"""oracle/build_ref.py — TEST INFRASTRUCTURE.  Compile the reference's own CUDA extension.

Builds ``/root/reference/ctdr/cuda/{diff_fp.cpp,diff_fp_for.cu,diff_fp_back.cu}`` *from where they
lie* into ``oracle/_ref/cuda_diff_fp.so`` (git-ignored, travels to the GPU box with the snapshot)
for sm_100a.  Nothing is copied into the repository; the only adaptation is the force-included
``oracle/ref_compat.h`` (see its header).  The reference's own build system (``build.py`` /
``setup.py``) is not run.

The module is a torch extension: it can only execute on a CUDA device (``diff_fp.cpp:4`` asserts
``is_cuda``), so it is used by the ``-m gpu`` tests as the ground-truth projector and by
``oracle/gen_ref_golden.py`` to produce ``tests/golden/ref_raster_*.npz``.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUDA = "/root/reference/ctdr/cuda"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "cuda_diff_fp.so")


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Returns the path of the built module, or None when /root/reference is not present."""
    if not os.path.isdir(REF_CUDA):
        return OUT_SO if os.path.exists(OUT_SO) else None
    srcs = [os.path.join(REF_CUDA, n) for n in ("diff_fp.cpp", "diff_fp_for.cu", "diff_fp_back.cu")]
    compat = os.path.join(HERE, "ref_compat.h")
    if (not force and os.path.exists(OUT_SO)
            and all(os.path.getmtime(OUT_SO) > os.path.getmtime(s) for s in srcs + [compat, __file__])):
        return OUT_SO
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    common = ["-DTORCH_EXTENSION_NAME=cuda_diff_fp", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-std=c++17", "-O3"]
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", s, "-o", o, "-include", compat, *inc, *common,
                   "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-w"]
        else:
            cmd = ["g++", "-c", s, "-o", o, *inc, *common, "-fPIC", "-w"]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode(errors="replace")[-4000:])
            raise RuntimeError("reference build failed: " + " ".join(cmd[:4]))
        if verbose:
            print(" ".join(cmd[:4]), "ok")
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", OUT_SO, *objs, *[f"-L{d}" for d in libdirs],
            *[f"-Wl,-rpath,{d}" for d in libdirs],
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return OUT_SO


def load():
    """Import the built reference module (needs torch imported first; GPU needed only to call it)."""
    import importlib.util
    import torch  # noqa: F401  (registers libtorch symbols before dlopen)
    if not os.path.exists(OUT_SO):
        return None
    spec = importlib.util.spec_from_file_location("cuda_diff_fp", OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
