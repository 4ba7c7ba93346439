"""``cuda_diff_fp``-compatible module: the reference's native entry points on the B200 kernels.

The reference's own ``ctdr/function/rasterizer.py`` does ``import cuda_diff_fp`` (``:3-5``) and calls
``forward`` twice (``:51-55,76-80``) and ``backward`` once (``:101-105``) with the argument lists of
``ctdr/cuda/diff_fp.cpp:21-26,77-82``.  Putting this module on ``sys.path`` under that name
(``sys.modules['cuda_diff_fp'] = shapefromprojections_b200.cuda_diff_fp``) runs the untouched
reference Python on the new kernels.

Differences: ``im``/``cntvisible`` are overwritten rather than accumulated into (they
arrive zeroed in the reference, ``rasterizer.py:36-38``); ``backward`` recomputes coverage from
``points2d`` and ``image`` and ignores the CSR arrays it is handed (they hold exactly that
coverage); faces with ``label_in < label_out`` are skipped in both passes.
"""
from __future__ import annotations

import torch

from . import _lib


def _check(t, name, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")            # diff_fp.cpp:4
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")              # diff_fp.cpp:5
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype} like points2d (got {t.dtype})")


def _dim(t, shape, msg):
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError(msg)


def forward(points2d_bxfx6, pointsdirect_bxfx1, pointsbbox_bxfx4, pointslen_bxfx3, labels_fx2, mus_n,
            imwei_3vf, imidx_vf, cntvisible_bxhxwx1, firstidx_bxhxwx1, im_bxhxwx1,
            is_the_first_pass, max_sino_val):
    """``dr_forward_batch`` (``diff_fp.cpp:21-67``).  ``pointsbbox_bxfx4`` is validated but not read:
    the kernels derive the bounding boxes from ``points2d`` themselves (same values, rasterizer.py:23-32)."""
    dt = points2d_bxfx6.dtype                     # AT_DISPATCH_FLOATING_TYPES(points2d.type(), ...) (diff_fp_for.cu:251)
    if dt not in (torch.float32, torch.float64):
        raise TypeError(f"points2d_bxfx6 must be float32 or float64 (got {dt})")
    for t, n in ((points2d_bxfx6, "points2d_bxfx6"), (pointsdirect_bxfx1, "pointsdirect_bxfx1"),
                 (pointsbbox_bxfx4, "pointsbbox_bxfx4"), (pointslen_bxfx3, "pointslen_bxfx3"),
                 (mus_n, "mus_n"), (im_bxhxwx1, "im_bxhxwx1")):
        _check(t, n, dt)
    _check(labels_fx2, "labels_fx2", torch.int32)
    _check(cntvisible_bxhxwx1, "cntvisible_bxhxwx1", torch.int32)
    _check(firstidx_bxhxwx1, "firstidx_bxhxwx1")
    B, F = points2d_bxfx6.shape[0], points2d_bxfx6.shape[1]
    H, W, D = im_bxhxwx1.shape[1], im_bxhxwx1.shape[2], im_bxhxwx1.shape[3]
    _dim(points2d_bxfx6, (B, F, 6), "points2d_bxfx6 must be same point size")
    _dim(pointsdirect_bxfx1, (B, F, 1), "pointsdirect_bxfx1 must be same point size")
    _dim(pointsbbox_bxfx4, (B, F, 4), "pointsbbox_bxfx4 must be same point size")
    _dim(pointslen_bxfx3, (B, F, 3), "pointslen_bxfx3 must be same point size")
    _dim(cntvisible_bxhxwx1, (B, H, W, 1), "cntvisible_bxhxwx1 must be same im size")
    _dim(im_bxhxwx1, (B, H, W, D), "im_bxhxwx1 must be same im size")
    if D != 1:
        raise RuntimeError("im_bxhxwx1: only one channel is supported (dnum is always 1 in the reference)")
    dev = points2d_bxfx6.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        if is_the_first_pass:
            ws = _lib.workspace(dev, L.ctdr_raster_workspace_bytes(B, F, H, W))
            entry = L.ctdr_raster_forward if dt == torch.float32 else L.ctdr_raster_forward_f64
            rc = entry(points2d_bxfx6.data_ptr(), pointsdirect_bxfx1.data_ptr(),
                                       pointslen_bxfx3.data_ptr(), labels_fx2.data_ptr(), mus_n.data_ptr(),
                                       mus_n.numel(), B, F, H, W, float(max_sino_val), im_bxhxwx1.data_ptr(),
                                       cntvisible_bxhxwx1.data_ptr(), ws.data_ptr(), ws.numel(), None, 0,
                                       _lib.stream_ptr(dev))
            _lib.check(rc, "ctdr_raster_forward")
        else:
            if dt != torch.float32:
                raise TypeError("pass 2 (the CSR fragment list) exists in float32 only here: it is a parity/debug output, "
                                "the backward recomputes coverage (DESIGN.md)")
            _check(imwei_3vf, "imwei_3vf", torch.float32)
            _check(imidx_vf, "imidx_vf", torch.int32)
            if firstidx_bxhxwx1.dtype != torch.int32 or firstidx_bxhxwx1.numel() != B * H * W:
                raise RuntimeError("firstidx_bxhxwx1 must be int32 [B,H,W,1]")
            if imidx_vf.numel() == 0:
                return
            ws = _lib.workspace(dev, L.ctdr_raster_fragments_workspace_bytes(B, F, H, W))
            rc = L.ctdr_raster_fragments(points2d_bxfx6.data_ptr(), pointsdirect_bxfx1.data_ptr(),
                                         pointslen_bxfx3.data_ptr(), labels_fx2.data_ptr(), mus_n.data_ptr(),
                                         mus_n.numel(), B, F, H, W, float(max_sino_val), im_bxhxwx1.data_ptr(),
                                         firstidx_bxhxwx1.data_ptr(), imidx_vf.data_ptr(), imwei_3vf.data_ptr(),
                                         ws.data_ptr(), ws.numel(), 0, _lib.stream_ptr(dev))
            _lib.check(rc, "ctdr_raster_fragments")


def backward(grad_sino_bxhxwxd, image_bxhxwxd, imidx_vf, cntvisible_bxhxwx1, firstidx_bxhxwx1, imwei_3vf,
             points2d_bxfx6, len_bxfx3, labels_fx2, mus_n, front_facing_bxfx1,
             grad_points2d_bxfx6, grad_len_bxfx3, grad_mus_n):
    """``dr_backward_batch`` (``diff_fp.cpp:77-129``); accumulates into the three ``grad_*`` tensors."""
    dt = grad_sino_bxhxwxd.dtype                  # diff_fp_back.cu:326 dispatches float and double
    if dt not in (torch.float32, torch.float64):
        raise TypeError(f"grad_sino_bxhxwxd must be float32 or float64 (got {dt})")
    for t, n in ((grad_sino_bxhxwxd, "grad_sino_bxhxwxd"), (image_bxhxwxd, "image_bxhxwxd"),
                 (points2d_bxfx6, "points2d_bxfx6"), (len_bxfx3, "len_bxfx3"), (mus_n, "mus_n"),
                 (front_facing_bxfx1, "front_facing_bxfx1"), (grad_points2d_bxfx6, "grad_points2d_bxfx6"),
                 (grad_len_bxfx3, "grad_len_bxfx3"), (grad_mus_n, "grad_mus_n")):
        _check(t, n, dt)
    for t, n in ((imidx_vf, "imidx_vf"), (cntvisible_bxhxwx1, "cntvisible_bxhxwx1"),
                 (firstidx_bxhxwx1, "firstidx_bxhxwx1"), (imwei_3vf, "imwei_3vf")):
        _check(t, n)
    _check(labels_fx2, "labels_fx2", torch.int32)
    B, H, W, D = grad_sino_bxhxwxd.shape
    F = points2d_bxfx6.shape[1]
    _dim(image_bxhxwxd, (B, H, W, D), "image_bxhxwxd must be same im size")
    _dim(cntvisible_bxhxwx1, (B, H, W, 1), "cntvisible_bxhxwx1 must be same im size")
    _dim(points2d_bxfx6, (B, F, 6), "points2d_bxfx6 must be same point size")
    _dim(len_bxfx3, (B, F, 3), "len_bxfx3 must be same point size")
    _dim(front_facing_bxfx1, (B, F, 1), "front_facing_bxfx1 must be same point size")
    _dim(grad_points2d_bxfx6, (B, F, 6), "grad_points2d_bxfx6 must be same point size")
    _dim(grad_len_bxfx3, (B, F, 3), "grad_len_bxfx3 must be same point size")
    dev = points2d_bxfx6.device
    L = _lib.lib()
    # max_sino_val is not an argument of the reference backward: pixels skipped by pass 2 have no
    # recorded fragments (imidx == 0, diff_fp_back.cu:97-98).  An upper bound of +inf reproduces the
    # lower test; the upper test (`im > max_sino_val`) is recovered from the fragment list: a valid
    # pixel with cnt > 0 has imidx[firstidx] != 0.
    msv = float("inf")
    grad = grad_sino_bxhxwxd
    if imidx_vf.numel() > 0:
        first = firstidx_bxhxwx1.reshape(-1).long().clamp_(max=imidx_vf.numel() - 1)
        recorded = (imidx_vf[first] != 0) | (cntvisible_bxhxwx1.reshape(-1) == 0)
        grad = grad_sino_bxhxwxd * recorded.reshape(B, H, W, 1).to(grad_sino_bxhxwxd.dtype)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, L.ctdr_raster_workspace_bytes(B, F, H, W))
        entry = L.ctdr_raster_backward if dt == torch.float32 else L.ctdr_raster_backward_f64
        rc = entry(grad.data_ptr(), image_bxhxwxd.data_ptr(), points2d_bxfx6.data_ptr(),
                                    front_facing_bxfx1.data_ptr(), len_bxfx3.data_ptr(), labels_fx2.data_ptr(),
                                    mus_n.data_ptr(), mus_n.numel(), B, F, H, W, msv,
                                    grad_points2d_bxfx6.data_ptr(), grad_len_bxfx3.data_ptr(), grad_mus_n.data_ptr(),
                                    ws.data_ptr(), ws.numel(), None, 0, _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_raster_backward")
