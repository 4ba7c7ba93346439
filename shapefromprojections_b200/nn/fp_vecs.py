"""Forward-projection module for vector geometries (``parallel3d_vec`` / ``cone_vec``).

Host-side mirror of ``ctdr/nn/fp_vecs.py:10-206`` (jakeoung/ShapeFromProjections): same constructor
``FP(proj_geom, labels, dtype)``, same ``forward(verts_vx3, faces_fx3, idx_angles, mus_n,
backprop=True) -> (phat, mask_valid)``, same public fields (``max_sino_val``, ``height``, ``width``,
``vecs``).  The ~40 torch ops of the reference's vertex stage and its two-pass rasteriser are one
fused device path here (``ProjectMesh``); ``fused=False`` keeps the reference's staging
(vertex stage -> ``Rasterizer.apply``) and its side-effect attributes ``proj/len/front_facing``.

Deviation: the reference indexes ``vecs`` by ``idx_angles`` but not ``det_pos`` (``fp_vecs.py:147,161``),
so it only works for ``idx_angles == arange(nangles)``; here every per-angle quantity follows
``idx_angles`` (identical results for that one working case).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from ..function.rasterizer import GeometryDevice, ProjectMesh, Rasterizer, VertexStage, project_forward_multi


def get_valid_mask(phat, max_val):
    """``fp_vecs.py:10-26`` without the host sync: the dead-pixel report is opt-in (CTDR_VERBOSE=1)."""
    mask_pos = phat > -1e-6
    mask_valid = mask_pos * (phat <= max_val - 1e-6)
    if os.environ.get("CTDR_VERBOSE", "0") not in ("", "0"):
        ndead = int((~mask_pos).sum().item())
        if ndead >= 1:
            print("ndead nneg phat.min and max_v:", ndead, ndead,
                  f"{phat.min().item():.2f}, {phat[mask_valid].max().item():.2f}")
    return mask_valid


class _ProjectorBase(nn.Module):
    """Shared plumbing of the two FP modules (geometry upload, face/label caches, dispatch)."""

    def __init__(self, proj_geom, labels, dtype=torch.float32, fused=True):
        super().__init__()
        if dtype not in (torch.float32, torch.float64):
            raise TypeError(f"FP: dtype must be torch.float32 or torch.float64 (got {dtype})")
        if dtype == torch.float64:
            fused = False      # float64 is the staged path (vertex stage -> Rasterizer), both in double on the device
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 projector needs a CUDA device (there is no CPU fallback)")
        self.proj_geom = proj_geom
        self.dtype = dtype
        self.fused = fused
        self.labels_fx2 = labels.cuda().to(torch.int32).contiguous()
        # one host sync at construction: every forward then checks mus_n against it (mus[label] is read, and in the
        # backward written, by label; the reference would index out of bounds silently)
        self._label_max = int(self.labels_fx2.max()) if self.labels_fx2.numel() else 0
        if self.labels_fx2.numel() and int(self.labels_fx2.min()) < 0:
            raise ValueError("labels must be >= 0")
        self.geom = GeometryDevice(proj_geom, self.labels_fx2.device)
        self.max_sino_val = self.geom.max_sino_val
        self.height = self.geom.H
        self.width = self.geom.W
        self.rasterize = Rasterizer.apply
        self._faces_src = None
        self._faces_i32 = None

    def _faces32(self, faces_fx3):
        # faces are constant across iterations: convert the reference's int64 tensor once
        if self._faces_src is not faces_fx3 or self._faces_i32 is None or self._faces_i32.shape[0] != faces_fx3.shape[0]:
            self._faces_i32 = faces_fx3.to(device=self.labels_fx2.device, dtype=torch.int32).contiguous()
            self._faces_src = faces_fx3
        return self._faces_i32

    def forward(self, verts_vx3, faces_fx3, idx_angles, mus_n, backprop=True):
        if mus_n.numel() <= self._label_max:
            raise IndexError(f"mus_n has {mus_n.numel()} entries but the face labels go up to {self._label_max}")
        faces32 = self._faces32(faces_fx3)
        if self.dtype == torch.float64 and (verts_vx3.dtype != torch.float64 or mus_n.dtype != torch.float64):
            raise TypeError("FP(dtype=torch.float64) expects float64 vertices and attenuations (Model(..., dtype=torch.float64))")
        if self.fused:
            im = ProjectMesh.apply(verts_vx3, mus_n, faces32, self.labels_fx2, idx_angles, self.geom, backprop)
        else:
            proj, length, facing = VertexStage.apply(verts_vx3, faces32, idx_angles, self.geom)
            self.proj, self.len, self.front_facing = proj, length, facing        # fp_vecs.py:198
            im = self.rasterize(proj, length, facing, self.labels_fx2, mus_n, self.height, self.width,
                                self.max_sino_val, backprop)
        phat = im.squeeze()                                                      # fp_vecs.py:203 (B==1 drops the axis)
        return phat, get_valid_mask(phat, self.max_sino_val)


    @torch.no_grad()
    def forward_multi(self, verts_vx3, faces_fx3, idx_angles, mus_kxn, target=None):
        """``K`` renders that differ only in ``mus`` from ONE sweep over the mesh (no gradients):
        ``(phat [K,B,H,W], A [K,K], rhs [K])`` with ``A[i,j] = sum phat_i*phat_j`` and
        ``rhs[i] = sum phat_i*target`` in float64 — what ``Model.estimate_mu`` (``vanilla.py:186-265``)
        obtains from 2-3 full renders plus torch reductions."""
        from ..function.rasterizer import _device_idx
        idx = _device_idx(idx_angles, self.labels_fx2.device, self.geom.nangles)
        # float32 kernels also for a float64 model: the normal equations are accumulated in float64 either way
        im, A, rhs, _ = project_forward_multi(self.geom, verts_vx3.detach().to(torch.float32).contiguous(), self._faces32(faces_fx3),
                                              self.labels_fx2, mus_kxn.to(torch.float32), idx, target=target, want_gram=True)
        return im, A, rhs


class FP(_ProjectorBase):
    """Forward projection by ASTRA vectors (``fp_vecs.py:29-206``)."""

    def __init__(self, proj_geom, labels, dtype=torch.float32, fused=True):
        if proj_geom["type"] not in ("parallel3d_vec", "cone_vec"):
            raise ValueError("fp_vecs.FP expects a *_vec geometry, got " + str(proj_geom["type"]))
        super().__init__(proj_geom, labels, dtype, fused)
        self.vecs = self.geom.table64 if dtype == torch.float64 else self.geom.table   # fp_vecs.py:64
        self.det_pos = self.vecs[:, 3:6]                                         # :72-74
        if proj_geom["type"] == "parallel3d_vec":
            self.det_pos = -self.max_sino_val * self.vecs[:, 0:3]
