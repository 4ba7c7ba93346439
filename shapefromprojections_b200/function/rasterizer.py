"""The differentiable mesh rasteriser as torch autograd functions.

* :class:`Rasterizer` keeps the operator surface of ``ctdr/function/rasterizer.py:10-107``
  (jakeoung/ShapeFromProjections): same ``apply`` arguments, same output ``im_bxhxwx1`` and the same
  gradient pattern (gradients for ``p_bxfx6``, ``len_bxfx3`` and ``mus_n`` only, ``:107``).
  What differs underneath: one tile-binned sm_100a launch instead of two ``B*H*W*F`` passes, no
  ``cumsum``/``vf`` host sync (``:63-64``), no CSR fragment list (backward recomputes coverage).
* :class:`ProjectMesh` is the fused operator (vertex stage + rasteriser): mesh and geometry in,
  sinogram out, gradients for the vertices and the attenuations.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from .. import _lib

dtype_int = torch.int32          # rasterizer.py:8


def default_flags() -> int:
    flags = 0
    if os.environ.get("CTDR_DETERMINISTIC", "0") not in ("", "0"):
        flags |= _lib.FLAG_DETERMINISTIC
    if os.environ.get("CTDR_NO_PREFILTER", "0") not in ("", "0"):
        flags |= _lib.FLAG_NO_PREFILTER
    return flags


def raster_forward(p6, len3, ff, labels, mus, H, W, max_sino_val, *, want_cnt=False, want_stats=False, flags=None):
    """Pass 1 of the reference (``diff_fp_for.cu`` with ``is_the_first_pass=1``) on the B200 kernels.

    Returns ``(im [B,H,W,1], cnt [B,H,W,1] int32 | None, stats int64[8] | None)``.
    """
    p6 = _lib.require_cuda_float(p6, "p_bxfx6")
    dt = p6.dtype                                     # float32 or float64 (diff_fp_for.cu:251 dispatches both)
    len3 = _lib.require_cuda_float(len3, "len_bxfx3", dt)
    ff = _lib.require_cuda_float(ff, "front_facing_bxfx1", dt)
    mus = _lib.require_cuda_float(mus, "mus_n", dt)
    if labels.dtype != dtype_int or not labels.is_cuda:
        raise RuntimeError("labels_fx2 must be a CUDA int32 tensor")
    labels = labels.contiguous()
    B, F = p6.shape[0], p6.shape[1]
    if p6.shape != (B, F, 6) or len3.shape != (B, F, 3) or ff.numel() != B * F or labels.shape != (F, 2):
        raise RuntimeError("p_bxfx6 / len_bxfx3 / front_facing_bxfx1 / labels_fx2 must be same point size")  # diff_fp.cpp:48-51
    dev = p6.device
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.int64, device=dev) if want_stats else None
    if B == 0 or F == 0:        # empty mesh / no angles: the reference's loops run zero times
        return (torch.zeros(B, H, W, 1, dtype=dt, device=dev),
                torch.zeros(B, H, W, 1, dtype=dtype_int, device=dev) if want_cnt else None, stats)
    im = torch.empty(B, H, W, 1, dtype=dt, device=dev)
    cnt = torch.empty(B, H, W, 1, dtype=dtype_int, device=dev) if want_cnt else None
    L = _lib.lib()
    nbytes = L.ctdr_raster_workspace_bytes(B, F, H, W)
    ws = _lib.workspace(dev, nbytes)
    entry = L.ctdr_raster_forward if dt == torch.float32 else L.ctdr_raster_forward_f64
    with torch.cuda.device(dev):
        rc = entry(p6.data_ptr(), ff.data_ptr(), len3.data_ptr(), labels.data_ptr(), mus.data_ptr(),
                                   mus.numel(), B, F, H, W, float(max_sino_val), im.data_ptr(), _lib.ptr(cnt),
                                   ws.data_ptr(), ws.numel(), _lib.ptr(stats),
                                   default_flags() if flags is None else flags, _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_raster_forward")
    return im, cnt, stats


def raster_backward(grad_im, im, p6, len3, ff, labels, mus, max_sino_val, *, flags=None, want_stats=False):
    """``diff_fp_back.cu`` on the B200 kernels: returns ``(dldp [B,F,6], dldf [B,F,3], dldmu [n])``."""
    B, F = p6.shape[0], p6.shape[1]
    H, W = im.shape[1], im.shape[2]
    dev = p6.device
    grad_im = _lib.require_cuda_float(grad_im, "dldI_bxhxwx1", p6.dtype)
    dldp = torch.zeros_like(p6)                                    # rasterizer.py:97-99
    dldf = torch.zeros_like(len3)
    dldmu = torch.zeros_like(mus)
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.int64, device=dev) if want_stats else None
    if B == 0 or F == 0:
        return (dldp, dldf, dldmu, stats) if want_stats else (dldp, dldf, dldmu)
    L = _lib.lib()
    ws = _lib.workspace(dev, L.ctdr_raster_workspace_bytes(B, F, H, W))
    entry = L.ctdr_raster_backward if p6.dtype == torch.float32 else L.ctdr_raster_backward_f64
    with torch.cuda.device(dev):
        rc = entry(grad_im.data_ptr(), im.data_ptr(), p6.data_ptr(), ff.data_ptr(), len3.data_ptr(),
                                    labels.data_ptr(), mus.data_ptr(), mus.numel(), B, F, H, W, float(max_sino_val),
                                    dldp.data_ptr(), dldf.data_ptr(), dldmu.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _lib.ptr(stats), default_flags() if flags is None else flags,
                                    _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_raster_backward")
    if want_stats:
        return dldp, dldf, dldmu, stats
    return dldp, dldf, dldmu


class Rasterizer(Function):
    """Drop-in for ``ctdr.function.rasterizer.Rasterizer`` (``rasterizer.py:10``)."""

    @staticmethod
    def forward(ctx, p_bxfx6, len_bxfx3, front_facing_bxfx1, labels_fx2, mus_n, H, W, max_sino_val, backprop):
        im, _, _ = raster_forward(p_bxfx6, len_bxfx3, front_facing_bxfx1, labels_fx2, mus_n, H, W, max_sino_val)
        if backprop is False:                                      # rasterizer.py:57-58
            return im
        ctx.max_sino_val = float(max_sino_val)
        ctx.save_for_backward(im, p_bxfx6.contiguous(), len_bxfx3.contiguous(), labels_fx2.contiguous(),
                              mus_n.contiguous(), front_facing_bxfx1.contiguous())
        return im

    @staticmethod
    def backward(ctx, dldI_bxhxwx1):
        im, p6, len3, labels, mus, ff = ctx.saved_tensors
        dldp, dldf, dldmu = raster_backward(dldI_bxhxwx1, im, p6, len3, ff, labels, mus, ctx.max_sino_val)
        return dldp, dldf, None, None, dldmu, None, None, None, None   # rasterizer.py:107


class GeometryDevice:
    """Device copy of what the reference FP constructors upload (fp_opengl.py:63, fp_vecs.py:64)."""

    def __init__(self, proj_geom: dict, device):
        kind = proj_geom["type"]
        if kind not in _lib.GEOM_KINDS:
            raise ValueError(f"unsupported projection geometry type {kind!r} (supported: {sorted(_lib.GEOM_KINDS)})")
        self.kind = _lib.GEOM_KINDS[kind]
        self.H = int(proj_geom["DetectorRowCount"])
        self.W = int(proj_geom["DetectorColCount"])
        self.spacing_x = float(proj_geom["DetectorSpacingX"])
        self.spacing_y = float(proj_geom.get("DetectorSpacingY", proj_geom["DetectorSpacingX"]))
        if self.kind == _lib.GEOM_PARALLEL3D:
            self.max_sino_val = self.W * proj_geom["DetectorSpacingX"] * 2          # fp_opengl.py:42
            table = torch.tensor(proj_geom["ProjectionAngles"], dtype=torch.float32)  # :63
        else:
            scale = 3 if self.kind == _lib.GEOM_PARALLEL3D_VEC else 100            # fp_vecs.py:59-62
            self.max_sino_val = self.W * proj_geom["DetectorSpacingX"] * scale
            table = torch.tensor(proj_geom["Vectors"], dtype=torch.float32)         # :64
        self.table = table.contiguous().to(device)
        self.nangles = int(self.table.shape[0])
        self._table64 = None
        self._source = proj_geom["ProjectionAngles"] if self.kind == _lib.GEOM_PARALLEL3D else proj_geom["Vectors"]

    @property
    def table64(self):
        """The geometry table as the reference builds it for ``dtype=torch.float64`` (``torch.tensor(..., dtype=dtype)``:
        the float32 ProjectionAngles widened, the float64 Vectors as they are)."""
        if self._table64 is None:
            self._table64 = torch.tensor(self._source, dtype=torch.float64).contiguous().to(self.table.device)
        return self._table64


_adjacency_cache: dict = {}       # (data_ptr, F, version, V) -> (faces tensor kept alive, start, corner); a few meshes


def vertex_adjacency(faces_i32, num_verts):
    """CSR vertex -> incident face corners (3*face + corner), grouped by vertex in ascending corner
    order (stable sort): what the deterministic backward gathers over.  Cached per faces tensor
    (several meshes may alternate: coarse-to-fine stages, config-5 style batches)."""
    key = (faces_i32.data_ptr(), int(faces_i32.shape[0]), faces_i32._version, int(num_verts))
    hit = _adjacency_cache.get(key)
    if hit is None:
        flat = faces_i32.reshape(-1).to(torch.int64)
        order = torch.sort(flat, stable=True).indices.to(torch.int32).contiguous()
        counts = torch.bincount(flat, minlength=int(num_verts))
        start = torch.zeros(int(num_verts) + 1, dtype=torch.int32, device=faces_i32.device)
        start[1:] = torch.cumsum(counts, 0).to(torch.int32)
        if len(_adjacency_cache) >= 8:
            _adjacency_cache.pop(next(iter(_adjacency_cache)))
        hit = (faces_i32, start.contiguous(), order)       # keeping the faces tensor alive pins its data_ptr
        _adjacency_cache[key] = hit
    return hit[1], hit[2]


def _adjacency_ptrs(flags, faces_i32, num_verts):
    if flags & _lib.FLAG_DETERMINISTIC:
        start, corner = vertex_adjacency(faces_i32, num_verts)
        return start.data_ptr(), corner.data_ptr()
    return None, None


def _device_idx(idx_angles, device, nangles=None):
    """int64 angle indices on the device.  Host indices are range-checked first (the reference raises IndexError
    through ``self.vecs[idx_angles]``, fp_vecs.py:101); device tensors are not (no sync) — an out-of-range row
    there turns that angle's output into NaN (ctdr_vertex.cuh::angle_const)."""
    t = torch.as_tensor(idx_angles)
    if nangles is not None and not t.is_cuda and t.numel():
        lo, hi = int(t.min()), int(t.max())
        if lo < -nangles or hi >= nangles:
            raise IndexError(f"angle index out of range for a geometry with {nangles} angles: [{lo}, {hi}]")
        if lo < 0:
            t = torch.where(t < 0, t + nangles, t)          # negative indices like torch indexing
    return t.to(device=device, dtype=torch.int64).contiguous()


def _check_mesh(geom, verts, faces_i32, labels, mus, idx):
    """Device, dtype, contiguity and shape of everything the fused entry points hand to the C ABI (which only
    checks for NULL): returns the contiguous tensors."""
    verts = _lib.require_cuda_f32(verts, "verts_vx3")
    mus = _lib.require_cuda_f32(mus, "mus_n")
    faces_i32 = _lib.require_cuda_i32(faces_i32, "faces_fx3")
    labels = _lib.require_cuda_i32(labels, "labels_fx2")
    if verts.dim() != 2 or verts.shape[1] != 3 or faces_i32.dim() != 2 or faces_i32.shape[1] != 3:
        raise RuntimeError("verts must be [V,3] and faces [F,3]")
    if labels.shape != (faces_i32.shape[0], 2):
        raise RuntimeError("labels_fx2 must be [F,2]")
    if not idx.is_cuda or idx.dtype != torch.int64:
        raise RuntimeError("idx_angles must be a CUDA int64 tensor here (use _device_idx)")
    return verts, faces_i32, labels, mus, idx.contiguous()


def vertex_forward(geom: GeometryDevice, verts, faces_i32, idx_angles):
    """p_bxfx6, len_bxfx3, front_facing_bxfx1 exactly as the reference FP modules build them."""
    verts = _lib.require_cuda_float(verts, "verts_vx3")
    faces_i32 = _lib.require_cuda_i32(faces_i32, "faces_fx3")
    dev = verts.device
    idx = _device_idx(idx_angles, dev, geom.nangles)
    B, V, F = idx.numel(), verts.shape[0], faces_i32.shape[0]
    f64 = verts.dtype == torch.float64
    p6 = torch.empty(B, F, 6, dtype=verts.dtype, device=dev)
    len3 = torch.empty(B, F, 3, dtype=verts.dtype, device=dev)
    ff = torch.empty(B, F, 1, dtype=verts.dtype, device=dev)
    table = geom.table64 if f64 else geom.table
    entry = _lib.lib().ctdr_vertex_forward_f64 if f64 else _lib.lib().ctdr_vertex_forward
    with torch.cuda.device(dev):
        rc = entry(geom.kind, table.data_ptr(), geom.nangles, idx.data_ptr(), B,
                                            verts.data_ptr(), V, faces_i32.data_ptr(), F, geom.H, geom.W,
                                            geom.spacing_x, geom.spacing_y, float(geom.max_sino_val),
                                            p6.data_ptr(), len3.data_ptr(), ff.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_vertex_forward")
    return p6, len3, ff


class VertexStage(Function):
    """Unfused vertex stage with its transposed Jacobian (compat mode of the FP modules)."""

    @staticmethod
    def forward(ctx, verts, faces_i32, idx_angles, geom):
        p6, len3, ff = vertex_forward(geom, verts, faces_i32, idx_angles)
        ctx.geom = geom
        ctx.save_for_backward(verts.contiguous(), faces_i32, _device_idx(idx_angles, verts.device, geom.nangles))
        ctx.mark_non_differentiable(ff)                            # rasterizer.py:107: no gradient
        return p6, len3, ff

    @staticmethod
    def backward(ctx, gp6, glen3, _gff):
        verts, faces_i32, idx = ctx.saved_tensors
        geom = ctx.geom
        dev = verts.device
        gp6 = _lib.require_cuda_float(gp6, "grad p6", verts.dtype)
        glen3 = _lib.require_cuda_float(glen3, "grad len3", verts.dtype)
        gverts = torch.zeros_like(verts)
        f64 = verts.dtype == torch.float64
        table = geom.table64 if f64 else geom.table
        entry = _lib.lib().ctdr_vertex_backward_f64 if f64 else _lib.lib().ctdr_vertex_backward
        with torch.cuda.device(dev):
            rc = entry(geom.kind, table.data_ptr(), geom.nangles, idx.data_ptr(),
                                                 idx.numel(), verts.data_ptr(), verts.shape[0], faces_i32.data_ptr(),
                                                 faces_i32.shape[0], geom.H, geom.W, geom.spacing_x, geom.spacing_y,
                                                 float(geom.max_sino_val), gp6.data_ptr(), glen3.data_ptr(),
                                                 gverts.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "ctdr_vertex_backward")
        return gverts, None, None, None


def project_forward(geom: GeometryDevice, verts, faces_i32, labels, mus, idx, *, want_cnt=False, want_stats=False,
                    flags=None, need_backward=False):
    verts, faces_i32, labels, mus, idx = _check_mesh(geom, verts, faces_i32, labels, mus, idx)
    dev = verts.device
    B, V, F = idx.numel(), verts.shape[0], faces_i32.shape[0]
    im = torch.empty(B, geom.H, geom.W, dtype=torch.float32, device=dev)
    cnt = torch.empty(B, geom.H, geom.W, dtype=dtype_int, device=dev) if want_cnt else None
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.int64, device=dev) if want_stats else None
    L = _lib.lib()
    ws = _lib.workspace(dev, _lib.project_workspace_bytes(B, V, F, geom.H, geom.W, need_backward))
    with torch.cuda.device(dev):
        rc = L.ctdr_project_forward(geom.kind, geom.table.data_ptr(), geom.nangles, idx.data_ptr(), B,
                                    verts.data_ptr(), V, faces_i32.data_ptr(), F, labels.data_ptr(), mus.data_ptr(),
                                    mus.numel(), geom.H, geom.W, geom.spacing_x, geom.spacing_y,
                                    float(geom.max_sino_val), im.data_ptr(), _lib.ptr(cnt), ws.data_ptr(), ws.numel(),
                                    _lib.ptr(stats), default_flags() if flags is None else flags,
                                    _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_project_forward")
    return im, cnt, stats


def _gram_matrices(gram, K):
    """(A [K,K] symmetric, rhs [K]) from the callee's upper-triangle layout."""
    A = gram[:K * K].view(K, K)
    A = torch.triu(A) + torch.triu(A, 1).T
    return A, gram[K * K:]


def raster_forward_multi(p6, len3, ff, labels, mus_kxn, H, W, max_sino_val, *, target=None, want_gram=False,
                         want_stats=False, flags=None):
    """``K = mus_kxn.shape[0]`` sinograms from one raster pass (``ctdr_raster_forward_multi``).

    Returns ``(im [K,B,H,W], A [K,K] | None, rhs [K] | None, stats)``: output ``k`` is bit-identical to
    ``raster_forward(..., mus_kxn[k])``; ``A[i,j] = sum im_i*im_j`` and ``rhs[i] = sum im_i*target``
    (float64) are the normal equations of ``Model.estimate_mu`` (``vanilla.py:208-216``).
    """
    p6 = _lib.require_cuda_f32(p6, "p_bxfx6")
    len3 = _lib.require_cuda_f32(len3, "len_bxfx3")
    ff = _lib.require_cuda_f32(ff, "front_facing_bxfx1")
    mus_kxn = _lib.require_cuda_f32(mus_kxn, "mus_kxn")
    B, F, K = p6.shape[0], p6.shape[1], mus_kxn.shape[0]
    dev = p6.device
    im = torch.empty(K, B, H, W, dtype=torch.float32, device=dev)
    want_gram = want_gram or target is not None
    gram = torch.empty(K * K + K, dtype=torch.float64, device=dev) if want_gram else None
    if target is not None:
        target = _lib.require_numel(_lib.require_cuda_f32(target, "target"), B * H * W, "target")
    labels = _lib.require_cuda_i32(labels, "labels_fx2")
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.int64, device=dev) if want_stats else None
    L = _lib.lib()
    ws = _lib.workspace(dev, L.ctdr_raster_workspace_bytes(B, F, H, W))
    with torch.cuda.device(dev):
        rc = L.ctdr_raster_forward_multi(p6.data_ptr(), ff.data_ptr(), len3.data_ptr(), labels.data_ptr(),
                                         mus_kxn.data_ptr(), K, mus_kxn.shape[1], B, F, H, W, float(max_sino_val),
                                         im.data_ptr(), _lib.ptr(target), _lib.ptr(gram), ws.data_ptr(), ws.numel(),
                                         _lib.ptr(stats), default_flags() if flags is None else flags,
                                         _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_raster_forward_multi")
    A, rhs = _gram_matrices(gram, K) if want_gram else (None, None)
    return im, A, rhs, stats


def project_forward_multi(geom: GeometryDevice, verts, faces_i32, labels, mus_kxn, idx, *, target=None,
                          want_gram=False, want_stats=False, flags=None):
    """Fused projector, ``K`` attenuation vectors in one sweep (``ctdr_project_forward_multi``); see
    :func:`raster_forward_multi` for the outputs."""
    mus_kxn = _lib.require_cuda_f32(mus_kxn, "mus_kxn")
    verts, faces_i32, labels, _, idx = _check_mesh(geom, verts, faces_i32, labels, mus_kxn.reshape(-1), idx)
    dev = verts.device
    B, V, F, K = idx.numel(), verts.shape[0], faces_i32.shape[0], mus_kxn.shape[0]
    im = torch.empty(K, B, geom.H, geom.W, dtype=torch.float32, device=dev)
    want_gram = want_gram or target is not None
    gram = torch.empty(K * K + K, dtype=torch.float64, device=dev) if want_gram else None
    if target is not None:
        target = _lib.require_cuda_f32(target, "target")
        if target.numel() != B * geom.H * geom.W:
            raise RuntimeError("target must hold B*H*W values")
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.int64, device=dev) if want_stats else None
    L = _lib.lib()
    ws = _lib.workspace(dev, _lib.project_workspace_bytes(B, V, F, geom.H, geom.W, False))
    with torch.cuda.device(dev):
        rc = L.ctdr_project_forward_multi(geom.kind, geom.table.data_ptr(), geom.nangles, idx.data_ptr(), B,
                                          verts.data_ptr(), V, faces_i32.data_ptr(), F, labels.data_ptr(),
                                          mus_kxn.data_ptr(), K, mus_kxn.shape[1], geom.H, geom.W, geom.spacing_x,
                                          geom.spacing_y, float(geom.max_sino_val), im.data_ptr(), _lib.ptr(target),
                                          _lib.ptr(gram), ws.data_ptr(), ws.numel(), _lib.ptr(stats),
                                          default_flags() if flags is None else flags, _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_project_forward_multi")
    A, rhs = _gram_matrices(gram, K) if want_gram else (None, None)
    return im, A, rhs, stats


def project_backward(geom: GeometryDevice, verts, faces_i32, labels, mus, idx, grad_im, im, *, flags=None,
                     grad_verts=None, grad_mus=None):
    verts, faces_i32, labels, mus, idx = _check_mesh(geom, verts, faces_i32, labels, mus, idx)
    dev = verts.device
    B, V, F = idx.numel(), verts.shape[0], faces_i32.shape[0]
    npix = B * geom.H * geom.W
    grad_im = _lib.require_numel(_lib.require_cuda_f32(grad_im, "grad_im"), npix, "grad_im")
    im = _lib.require_numel(_lib.require_cuda_f32(im, "im"), npix, "im")
    grad_verts = torch.zeros_like(verts) if grad_verts is None else _lib.require_out(grad_verts, torch.float32, 3 * V, "grad_verts")
    grad_mus = torch.zeros_like(mus) if grad_mus is None else _lib.require_out(grad_mus, torch.float32, mus.numel(), "grad_mus")
    L = _lib.lib()
    flags = default_flags() if flags is None else flags
    adj0, adj1 = _adjacency_ptrs(flags, faces_i32, V)
    ws = _lib.workspace(dev, _lib.project_workspace_bytes(B, V, F, geom.H, geom.W, True))
    with torch.cuda.device(dev):
        rc = L.ctdr_project_backward(geom.kind, geom.table.data_ptr(), geom.nangles, idx.data_ptr(), B,
                                     verts.data_ptr(), V, faces_i32.data_ptr(), F, labels.data_ptr(), mus.data_ptr(),
                                     mus.numel(), geom.H, geom.W, geom.spacing_x, geom.spacing_y,
                                     float(geom.max_sino_val), grad_im.data_ptr(), im.data_ptr(),
                                     grad_verts.data_ptr(), grad_mus.data_ptr(), adj0, adj1, ws.data_ptr(), ws.numel(),
                                     None, flags, _lib.stream_ptr(dev))
    _lib.check(rc, "ctdr_project_backward")
    return grad_verts, grad_mus


def project_residual_step(geom: GeometryDevice, verts, faces_i32, labels, mus, idx, target, *, flags=None,
                          phat=None, grad_verts=None, grad_mus=None, out2=None, want_stats=False, stage_ms=None):
    """One projection + masked-L2 residual + backward sweep (``ctdr/optimize.py:162-168,190``).

    Returns ``(phat, grad_verts, grad_mus, out2)`` with ``out2 = [sum_valid (target-phat)^2, n_valid]``
    (float64, device) and the gradients of that *sum* (divide by ``n_valid`` for the mean).
    Nothing synchronises with the host — unless ``stage_ms`` (a dict) is given: then the profiling entry point
    ``ctdr_project_residual_step_timed`` runs instead, synchronises, and fills ``stage_ms`` with the device
    milliseconds of every stage (``_lib.STAGES``).
    """
    verts, faces_i32, labels, mus, idx = _check_mesh(geom, verts, faces_i32, labels, mus, idx)
    dev = verts.device
    B, V, F = idx.numel(), verts.shape[0], faces_i32.shape[0]
    npix = B * geom.H * geom.W
    # the C ABI only checks for NULL: a float64 sinogram (SinoDataset(dtype=float64)) or a short batch would be
    # read as garbage / out of bounds by the kernels
    target = _lib.require_numel(_lib.require_cuda_f32(target, "target sinogram (p_batch)"), npix, "target sinogram (p_batch)")
    phat = (torch.empty(B, geom.H, geom.W, dtype=torch.float32, device=dev) if phat is None
            else _lib.require_out(phat, torch.float32, npix, "phat"))
    grad_verts = torch.zeros_like(verts) if grad_verts is None else _lib.require_out(grad_verts, torch.float32, 3 * V, "grad_verts")
    grad_mus = torch.zeros_like(mus) if grad_mus is None else _lib.require_out(grad_mus, torch.float32, mus.numel(), "grad_mus")
    out2 = torch.zeros(2, dtype=torch.float64, device=dev) if out2 is None else _lib.require_out(out2, torch.float64, 2, "out2")
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.int64, device=dev) if want_stats else None
    L = _lib.lib()
    flags = default_flags() if flags is None else flags
    adj0, adj1 = _adjacency_ptrs(flags, faces_i32, V)
    ws = _lib.workspace(dev, _lib.project_workspace_bytes(B, V, F, geom.H, geom.W, True))
    args = (geom.kind, geom.table.data_ptr(), geom.nangles, idx.data_ptr(), B,
            verts.data_ptr(), V, faces_i32.data_ptr(), F, labels.data_ptr(),
            mus.data_ptr(), mus.numel(), geom.H, geom.W, geom.spacing_x, geom.spacing_y,
            float(geom.max_sino_val), target.data_ptr(), phat.data_ptr(),
            grad_verts.data_ptr(), grad_mus.data_ptr(), out2.data_ptr(), adj0, adj1,
            ws.data_ptr(), ws.numel(), _lib.ptr(stats), flags, _lib.stream_ptr(dev))
    with torch.cuda.device(dev):
        if stage_ms is None:
            rc = L.ctdr_project_residual_step(*args)
        else:
            import ctypes
            host = (ctypes.c_float * len(_lib.STAGES))()
            rc = L.ctdr_project_residual_step_timed(*args, ctypes.cast(host, ctypes.c_void_p))
            for name, v in zip(_lib.STAGES, host):
                stage_ms[name] = stage_ms.get(name, 0.0) + float(v)
    _lib.check(rc, "ctdr_project_residual_step")
    if want_stats:
        return phat, grad_verts, grad_mus, out2, stats
    return phat, grad_verts, grad_mus, out2


class HostStepper:
    """The residual step with HOST buffers in and out (mini-batch semantics of the reference loop:
    ``p_batch = p_batch.cuda()`` from a pinned loader every iteration, ``optimize.py:157-158``).

    The target sinogram is streamed host->device in angle windows on a copy stream, double
    buffered, while the projector works on the previous window; mesh, attenuations and angle
    indices are uploaded first; loss, valid count and gradients come back to pinned host memory.
    """

    def __init__(self, geom: GeometryDevice, faces_i32, labels, num_verts, num_mus, max_angles, windows=8):
        dev = faces_i32.device
        self.geom, self.faces, self.labels, self.dev = geom, faces_i32, labels, dev
        self.win = max(1, -(-max_angles // windows))
        self.stage = [torch.empty(self.win, geom.H, geom.W, dtype=torch.float32, device=dev) for _ in range(2)]
        self.phat = torch.empty(max_angles, geom.H, geom.W, dtype=torch.float32, device=dev)
        self.copy_stream = torch.cuda.Stream(dev)
        self.flat = torch.zeros(3 * num_verts + num_mus, dtype=torch.float32, device=dev)
        self.grad_verts = self.flat[:3 * num_verts].view(num_verts, 3)
        self.grad_mus = self.flat[3 * num_verts:]
        self.out2 = torch.zeros(2, dtype=torch.float64, device=dev)
        self.h_flat = torch.empty(self.flat.shape, dtype=torch.float32).pin_memory()
        self.h_out2 = torch.empty(2, dtype=torch.float64).pin_memory()
        self.h2d_bytes = self.d2h_bytes = 0

    def step(self, h_verts, h_mus, h_idx, h_target, reduce_fn=None):
        """All arguments are pinned host tensors; returns (h_flat, h_out2) after synchronising."""
        dev, cur = self.dev, torch.cuda.current_stream(self.dev)
        verts = h_verts.to(dev, non_blocking=True)
        mus = h_mus.to(dev, non_blocking=True)
        idx = h_idx.to(dev, non_blocking=True)
        self.flat.zero_(); self.out2.zero_()
        B = h_idx.numel()
        done = [None, None]
        for k, a in enumerate(range(0, B, self.win)):
            n = min(self.win, B - a)
            buf = self.stage[k & 1]
            with torch.cuda.stream(self.copy_stream):
                if done[k & 1] is not None:
                    self.copy_stream.wait_event(done[k & 1])       # the buffer's previous consumer finished
                buf[:n].copy_(h_target[a:a + n], non_blocking=True)
                ready = torch.cuda.Event(); ready.record(self.copy_stream)
            cur.wait_event(ready)
            project_residual_step(self.geom, verts, self.faces, self.labels, mus, idx[a:a + n], buf[:n],
                                  phat=self.phat[a:a + n], grad_verts=self.grad_verts, grad_mus=self.grad_mus,
                                  out2=self.out2)
            ev = torch.cuda.Event(); ev.record(cur); done[k & 1] = ev
        if reduce_fn is not None:
            reduce_fn(self.flat, self.out2)
        self.h_flat.copy_(self.flat, non_blocking=True)
        self.h_out2.copy_(self.out2, non_blocking=True)
        cur.synchronize()
        self.h2d_bytes = h_target[:B].numel() * 4 + h_verts.numel() * 4 + h_mus.numel() * 4 + h_idx.numel() * 8
        self.d2h_bytes = self.h_flat.numel() * 4 + 16
        return self.h_flat, self.h_out2


class ProjectMesh(Function):
    """Fused projector: ``(verts [V,3], mus [n]) -> im [B,H,W]`` with gradients for both."""

    @staticmethod
    def forward(ctx, verts, mus, faces_i32, labels, idx_angles, geom, backprop):
        verts = _lib.require_cuda_f32(verts, "verts_vx3")
        mus = _lib.require_cuda_f32(mus, "mus_n")
        idx = _device_idx(idx_angles, verts.device, geom.nangles)
        im, _, _ = project_forward(geom, verts, faces_i32, labels, mus, idx, need_backward=bool(backprop))
        if backprop:
            ctx.geom = geom
            ctx.save_for_backward(verts, mus, faces_i32, labels, idx, im)
        return im

    @staticmethod
    def backward(ctx, grad_im):
        verts, mus, faces_i32, labels, idx, im = ctx.saved_tensors
        grad_im = _lib.require_cuda_f32(grad_im, "dldI")
        gverts, gmus = project_backward(ctx.geom, verts, faces_i32, labels, mus, idx, grad_im, im)
        return gverts, gmus, None, None, None, None, None
