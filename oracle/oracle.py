"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT THE PRODUCT.

NumPy/ctypes front end of the CPU restatement in ``oracle/ctdr_oracle.c``.  It replays the host
logic of the reference operator (``ctdr/function/rasterizer.py:13-107`` of
jakeoung/ShapeFromProjections) around the restated kernels, so tests can compare the CUDA path
with "what the reference does" on a machine without a GPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.

Parity pin: see the header of ``ctdr_oracle.c`` (golden vectors produced by the reference's own
kernels on a B200, ``tests/golden/ref_raster_*.npz``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libctdr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = [os.path.join(_HERE, n) for n in ("ctdr_oracle.c", "ctdr_oracle_impl.inc", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libctdr_oracle.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_abi_version.restype = ctypes.c_int
    return _lib


def _suffix(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _real(dtype, v):
    return ctypes.c_float(v) if np.dtype(dtype) == np.float32 else ctypes.c_double(v)


def bbox(p6: np.ndarray) -> np.ndarray:
    """rasterizer.py:23-32 — [B,F,6] -> [B,F,4] (xmin, ymin, xmax, ymax)."""
    p6 = np.ascontiguousarray(p6)
    out = np.empty(p6.shape[:2] + (4,), dtype=p6.dtype)
    getattr(lib(), "oracle_bbox_" + _suffix(p6.dtype))(
        _p(p6), _p(out), ctypes.c_longlong(p6.shape[0] * p6.shape[1]))
    return out


def forward_pass(p6, ff, bbox4, len3, labels, mus, imwei, imidx, cnt, firstidx, im,
                 first_pass: int, msv: float, *, fast: bool = False,
                 angles=None, rows=None):
    """One launch of the reference forward kernel (diff_fp_for.cu:226-274 signature order).

    ``fast=False`` is the literal pixel-parallel loop over all faces (the reference algorithm, and
    the one timed as the CPU baseline); ``fast=True`` is the face-major equivalent used to check
    larger cases.  ``angles=(b0,b1)`` / ``rows=(y0,y1)`` restrict the literal version to a sample.
    Returns (n_nonpositive_len, n_candidates or None).
    """
    B, F = p6.shape[:2]
    H, W = im.shape[1:3]
    dt = p6.dtype
    sfx = _suffix(dt)
    nonpos = ctypes.c_longlong(0)
    ncand = ctypes.c_longlong(0)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    if fast:
        assert angles is None and rows is None
        getattr(lib(), "oracle_forward_fast_" + sfx)(
            _p(p6), _p(ff), _p(bbox4), _p(len3), _p(labels), _p(mus),
            _p(imwei), _p(imidx), _p(cnt), _p(firstidx), _p(im),
            B, H, W, F, int(first_pass), _real(dt, msv),
            ctypes.byref(nonpos), ctypes.byref(ncand))
        return nonpos.value, ncand.value
    b0, b1 = angles if angles is not None else (0, B)
    y0, y1 = rows if rows is not None else (0, H)
    getattr(lib(), "oracle_forward_literal_" + sfx)(
        _p(p6), _p(ff), _p(bbox4), _p(len3), _p(labels), _p(mus),
        _p(imwei), _p(imidx), _p(cnt), _p(firstidx), _p(im),
        B, H, W, F, int(first_pass), _real(dt, msv),
        int(b0), int(b1), int(y0), int(y1), ctypes.byref(nonpos))
    return nonpos.value, None


def rasterizer_forward(p6, len3, ff, labels, mus, H, W, msv, backprop=True, *, fast=False):
    """``Rasterizer.forward`` (rasterizer.py:13-86) on the CPU oracle.

    Returns a dict with ``im`` [B,H,W,1], ``cnt`` [B,H,W,1] int32 and, when ``backprop``,
    ``firstidx`` [B,H,W,1] int32, ``imidx`` [vf] int32 (face id + 1, ascending per pixel),
    ``imwei`` [3*vf]; plus ``n_nonpositive_len`` (replaces the device assert, for.cu:183).
    """
    p6 = np.ascontiguousarray(p6)
    dt = p6.dtype
    len3 = np.ascontiguousarray(len3, dtype=dt)
    ff = np.ascontiguousarray(ff, dtype=dt)
    mus = np.ascontiguousarray(mus, dtype=dt)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    B, F = p6.shape[:2]
    bb = bbox(p6)                                                   # :23-32
    im = np.zeros((B, H, W, 1), dtype=dt)                           # :36
    cnt = np.zeros((B, H, W, 1), dtype=np.int32)                    # :38
    nonpos, ncand = forward_pass(p6, ff, bb, len3, labels, mus, None, None, cnt, None, im,
                                 1, msv, fast=fast)                 # :51-55
    out = {"im": im, "cnt": cnt, "bbox": bb, "n_nonpositive_len": nonpos, "n_candidates": ncand}
    if not backprop:                                                # :57-58
        return out
    cumsum = np.cumsum(cnt.reshape(-1), dtype=np.int32)             # :63
    vf = int(cumsum[-1]) if cumsum.size else 0                      # :64
    imwei = np.zeros(3 * vf, dtype=dt)                              # :67
    imidx = np.zeros(vf, dtype=np.int32)                            # :68
    firstidx = np.zeros(B * H * W, dtype=np.int32)                  # :71-72
    firstidx[1:] = cumsum[:-1]
    firstidx = firstidx.reshape(B, H, W, 1)
    forward_pass(p6, ff, bb, len3, labels, mus, imwei, imidx, cnt, firstidx, im, 0, msv,
                 fast=fast)                                         # :76-80
    out.update(firstidx=firstidx, imidx=imidx, imwei=imwei, vf=vf)
    return out


def rasterizer_backward(grad_im, fwd, p6, len3, ff, labels, mus, *, angles=None, rows=None):
    """``Rasterizer.backward`` (rasterizer.py:89-107): returns (dldp [B,F,6], dldf [B,F,3], dldmu [n])."""
    p6 = np.ascontiguousarray(p6)
    dt = p6.dtype
    len3 = np.ascontiguousarray(len3, dtype=dt)
    ff = np.ascontiguousarray(ff, dtype=dt)
    mus = np.ascontiguousarray(mus, dtype=dt)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    grad_im = np.ascontiguousarray(grad_im, dtype=dt)
    B, F = p6.shape[:2]
    H, W = fwd["im"].shape[1:3]
    dldp = np.zeros_like(p6)                                        # :97-99
    dldf = np.zeros_like(len3)
    dldmu = np.zeros_like(mus)
    b0, b1 = angles if angles is not None else (0, B)
    y0, y1 = rows if rows is not None else (0, H)
    getattr(lib(), "oracle_backward_" + _suffix(dt))(
        _p(grad_im), _p(fwd["im"]), _p(fwd["imidx"]), _p(fwd["cnt"]), _p(fwd["firstidx"]),
        _p(fwd["imwei"]), _p(p6), _p(len3), _p(labels), _p(mus), _p(ff),
        _p(dldp), _p(dldf), _p(dldmu),
        B, H, W, F, int(mus.shape[0]), int(b0), int(b1), int(y0), int(y1))   # :101-105
    return dldp, dldf, dldmu


def valid_mask(phat: np.ndarray, msv: float) -> np.ndarray:
    """``get_valid_mask`` (ctdr/nn/fp_vecs.py:10-13), comparisons in phat's dtype like torch."""
    dt = phat.dtype
    return (phat > dt.type(-1e-6)) & (phat <= dt.type(msv - 1e-6))
