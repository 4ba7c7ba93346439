"""ctypes binding of the C-ABI library ``csrc/libctdr_b200.so`` (include/ctdr_b200.h).

There is no CPU path and no fallback: if the shared library is missing or an entry point fails,
the caller gets an exception.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTDR_LIB") or os.path.join(_HERE, "csrc", "libctdr_b200.so")   # CTDR_LIB: A/B builds (tools/)

FLAG_NO_PREFILTER = 1
FLAG_DETERMINISTIC = 2
ABI_VERSION = 5          # include/ctdr_b200.h: CTDR_ABI_VERSION
STAT_NONPOSITIVE_LEN, STAT_FRAGMENTS, STAT_CANDIDATES, STAT_LIST_SEGMENTS, STAT_COUNT = 0, 1, 2, 3, 8
GEOM_PARALLEL3D, GEOM_PARALLEL3D_VEC, GEOM_CONE_VEC = 0, 1, 2
GEOM_KINDS = {"parallel3d": GEOM_PARALLEL3D, "parallel3d_vec": GEOM_PARALLEL3D_VEC, "cone_vec": GEOM_CONE_VEC}

_vp, _i, _sz, _u, _d, _f = (ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint,
                            ctypes.c_double, ctypes.c_float)

# name -> (restype, argtypes); must list every symbol declared in include/ctdr_b200.h
SIGNATURES = {
    "ctdr_abi_version": (_i, []),
    "ctdr_last_error": (ctypes.c_char_p, []),
    "ctdr_launch_count": (ctypes.c_ulonglong, []),
    "ctdr_tuning_reload": (None, []),
    "ctdr_selftest_quotient": (_i, [ctypes.c_ulonglong, _u, _vp, _vp]),
    "ctdr_raster_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ctdr_raster_fragments_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ctdr_raster_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_raster_forward_multi": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_raster_fragments": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp]),
    "ctdr_raster_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_vertex_forward": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "ctdr_vertex_backward": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "ctdr_raster_forward_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_raster_backward_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_vertex_forward_f64": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "ctdr_vertex_backward_f64": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "ctdr_project_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "ctdr_project_forward": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _d, _d, _d,
                                  _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_project_forward_multi": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _d, _d, _d,
                                        _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_project_backward": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _d, _d, _d,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_project_residual_step": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _d, _d, _d,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp]),
    "ctdr_project_residual_step_timed": (_i, [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _d, _d, _d,
                                              _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _u, _vp, _vp]),
    "ctdr_peer_block_bytes": (_sz, [_sz]),
    "ctdr_peer_block_create": (_i, [_sz, _vp, _vp]),
    "ctdr_peer_block_open": (_i, [_vp, _vp]),
    "ctdr_peer_block_close": (_i, [_vp]),
    "ctdr_peer_block_destroy": (_i, [_vp]),
    "ctdr_peer_block_status": (_i, [_vp, _vp]),
    "ctdr_peer_allreduce": (_i, [_vp, _i, _i, _sz, _vp, _sz, _sz, _vp]),
}
STAGES = ("vertex_forward", "raster_forward", "memset", "raster_step", "vertex_backward")   # CTDR_STAGE_*

_lib = None


def lib() -> ctypes.CDLL:
    """Load the library once; raise (never fall back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m shapefromprojections_b200.csrc.build` "
                "(there is no CPU fallback for the projector)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.ctdr_abi_version() != ABI_VERSION:
            raise RuntimeError("libctdr_b200.so: unexpected ABI version")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ctdr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_workspaces: dict = {}


def workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer, ONE per device (the C ABI never allocates).  Launches that use it are ordered
    by the stream they run on; callers that overlap projector calls on several streams of one device must
    pass distinct workspaces through the C ABI themselves.  ``release_workspaces()`` frees them."""
    key = torch.device(device).index
    if key is None:
        key = torch.cuda.current_device()
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        _workspaces.pop(key, None)
        buf = None                                        # free the old one before allocating the larger one
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    """Drop every cached scratch buffer (they are re-created on demand)."""
    _workspaces.clear()


def project_workspace_bytes(B: int, V: int, F: int, H: int, W: int, need_backward: bool) -> int:
    """Scratch size for the fused projector: the library's own suggestion (angle windows up to 8 GiB), capped by
    ``CTDR_WORKSPACE_MB`` when set — the projector then simply works through more, smaller angle windows."""
    L = lib()
    want = int(L.ctdr_project_workspace_bytes(B, V, F, H, W, int(need_backward)))
    cap = os.environ.get("CTDR_WORKSPACE_MB")
    if cap:
        one_angle = int(L.ctdr_project_workspace_bytes(1, V, F, H, W, int(need_backward)))
        want = max(min(want, int(float(cap) * (1 << 20))), one_angle)
    return want


def reload_tuning() -> None:
    """Make the library re-read its CTDR_* tuning variables (it reads the environment once, not per launch)."""
    lib().ctdr_tuning_reload()


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")          # diff_fp.cpp:4
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: this entry point is the float32 instantiation (got {t.dtype}); float64 runs through "
                        "the staged operators (Rasterizer / FP(..., dtype=torch.float64, ...))")
    return t.contiguous()                                            # rasterizer.py:40-43


def require_cuda_float(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    """float32 or float64 (AT_DISPATCH_FLOATING_TYPES, diff_fp_for.cu:251), CUDA, contiguous; ``dtype`` pins which."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")          # diff_fp.cpp:4
    if t.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"{name} must be float32 or float64 (got {t.dtype})")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype} like the other operands (got {t.dtype})")
    return t.contiguous()


def require_numel(t: torch.Tensor, n: int, name: str) -> torch.Tensor:
    if t.numel() != n:
        raise RuntimeError(f"{name} must hold {n} values, got {t.numel()} (shape {tuple(t.shape)})")
    return t


def require_cuda_i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be a CUDA int32 tensor (got {t.dtype} on {t.device})")   # rasterizer.py:8
    return t.contiguous()


def require_out(t: torch.Tensor, dtype, n: int, name: str) -> torch.Tensor:
    """Caller-provided output buffer: right device kind, dtype, size, and contiguous (it is written in place)."""
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous CUDA {dtype} tensor")
    return require_numel(t, n, name)
