"""Forward-projection module for the ``parallel3d`` geometry (the one ``run/ours.py`` exercises).

Host-side mirror of ``ctdr/nn/fp_opengl.py:9-168``: ``FP(proj_geom, labels, dtype)`` and
``forward(verts_vx3, faces_fx3, idx_angles, mus_n, backprop=True) -> (phat, mask_valid)`` with the
public fields ``max_sino_val``, ``height``, ``width``, ``angles``, ``P``.  The rotate / view /
orthographic-projection chain (``:87-158``) runs inside the fused device path.
"""
from __future__ import annotations

import torch

from ..utils.projection import compute_P
from .fp_vecs import _ProjectorBase, get_valid_mask  # noqa: F401  (re-exported like the reference, :7)


class FP(_ProjectorBase):
    def __init__(self, proj_geom, labels, dtype=torch.float32, fused=True):
        if proj_geom["type"] != "parallel3d":
            raise ValueError("fp_opengl.FP expects type 'parallel3d', got " + str(proj_geom["type"]))
        super().__init__(proj_geom, labels, dtype, fused)
        self.angles = self.geom.table64 if dtype == torch.float64 else self.geom.table   # fp_opengl.py:63
        self.P = torch.tensor(compute_P(proj_geom), dtype=dtype, device=self.angles.device)  # :44-45
