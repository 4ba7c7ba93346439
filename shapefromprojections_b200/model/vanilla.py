"""Template-mesh model: vertices = template - center + displace, attenuations ``mus`` per label.

Host-side mirror of ``ctdr/model/vanilla.py:11-310``: same constructor (positional order kept:
``Model(ftemplate, proj_geom, nmaterials, mus, mus_fixed_no, use_disp_param, use_center_param,
wlap, wflat, dtype)``), ``forward(idx_angles, wedge, backprop) -> (phat, mask_valid, edge, lap, flat)``,
``render``, ``estimate_mu``, ``set_mu0``, ``get_vf_np``, ``get_displaced_vertices``.  The projector
behind ``self.fp`` is the fused B200 path; regularisers are the sparse versions.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..nn import softras_loss
from ..utils import util_fun, util_mesh


class Model(nn.Module):
    def __init__(self, ftemplate, proj_geom, nmaterials, mus=None, mus_fixed_no=1, use_disp_param=True,
                 use_center_param=False, wlap=0., wflat=0., dtype=torch.float32):
        super().__init__()
        self.p_min = 0.
        self.dtype = dtype
        self.nmaterials = nmaterials
        if isinstance(ftemplate, (tuple, list)):       # (vertices, faces, labels_v, labels_f) in memory
            vertices, faces, labels_v, labels_f = ftemplate
        else:
            vertices, faces, labels_v, labels_f = util_mesh.load_template(ftemplate)
        self.mus_fixed_no = abs(mus_fixed_no)
        self.register_buffer("vertices", torch.tensor(np.asarray(vertices), dtype=dtype))
        self.register_buffer("faces", torch.tensor(np.asarray(faces), dtype=torch.int64))
        self.register_buffer("labels", torch.tensor(np.asarray(labels_f), dtype=torch.int32))
        self.labels_v = self.labels_v_np = labels_v
        self.use_lap_loss = wlap > 0.
        self.use_flat_loss = wflat > 0.
        self.init_register(mus, mus_fixed_no, use_center_param, use_disp_param)
        if proj_geom["type"] == "cone":
            raise AssertionError("proj_geom type=cone is not supported. It should be cone_vec")   # vanilla.py:62-63
        if proj_geom["type"][-3:] == "vec":                                                      # :66-73
            from ..nn.fp_vecs import FP
            self.nangles = proj_geom["Vectors"].shape[0]
        else:
            from ..nn.fp_opengl import FP
            self.nangles = len(proj_geom["ProjectionAngles"])
        self.fp = FP(proj_geom, self.labels, self.dtype)

    def init_register(self, mus, mus_fixed_no, use_center_param, use_disp_param=True):
        zeros = lambda shape: torch.zeros(shape, dtype=self.dtype)
        if use_center_param:
            self.register_parameter("center", nn.Parameter(zeros([1, 3])))
        else:
            self.register_buffer("center", zeros([1, 3]))
        if use_disp_param:
            self.register_parameter("displace", nn.Parameter(zeros(self.vertices.shape)))
        else:
            self.register_buffer("displace", torch.zeros_like(self.vertices))
        if mus_fixed_no == self.nmaterials:
            self.register_buffer("mus", torch.tensor(mus, dtype=self.dtype))
        else:
            print("set mu as parameters")
            self.register_parameter("mus", nn.Parameter(torch.tensor(mus, dtype=self.dtype)))
        if self.use_lap_loss:
            self.laplacian_loss = softras_loss.LaplacianLoss(self.vertices.cpu(), self.faces.cpu())
        if self.use_flat_loss:
            self.flat_loss = softras_loss.FlattenLoss(self.faces)

    def register_scale(self):
        self.register_parameter("scale", nn.Parameter(torch.ones([1, 3], dtype=self.dtype)))

    def pretransform(self, translation=False, rot=False):
        vert = self.vertices
        if translation:
            vert -= self.center            # in place on the buffer, like the reference (:123-125)
        if hasattr(self, "scale"):
            vert = vert * self.scale
        return vert

    def deform(self, verts, disp):
        return verts + (self.displace if disp is None else disp)

    def get_displaced_vertices(self):
        return self.deform(self.pretransform(True), None)

    def forward(self, idx_angles, wedge=0., backprop=True):
        vertices = self.get_displaced_vertices()
        if self.mus_fixed_no == 1:
            self.mus.data[0] = self.p_min                                                # :148-149
        phat, mask_valid = self.fp.forward(vertices, self.faces, idx_angles, self.mus, backprop)
        edge_loss = util_fun.compute_edge_length(vertices, self.faces) if wedge > 0. else 0.
        lap_loss = self.laplacian_loss(vertices).mean() if self.use_lap_loss else 0.
        flat_loss = self.flat_loss(vertices) if self.use_flat_loss else 0.
        return phat, mask_valid, edge_loss, lap_loss, flat_loss

    def render(self):
        idx = torch.arange(self.nangles, device=self.vertices.device)
        return self.fp.forward(self.get_displaced_vertices(), self.faces, idx, self.mus)

    def set_mu0(self, p_cuda):
        self.p_min = p_cuda.min()

    def estimate_mu(self, p_cuda):
        """Least-squares attenuations from one-hot renders (``vanilla.py:186-265``): the projector is
        linear in ``mus``, so ``p ~= sum_k mu_k * render(mu_0 = p_min, mu_k = 1)``; the reference renders
        the mesh once per material and reduces with torch, here the basis sinograms and the normal
        equations ``A mu = b`` come out of ONE multi-output raster pass (``ctdr_project_forward_multi``)."""
        print("@ estimate mus")
        self.p_min = p_cuda.min()
        n = self.nmaterials
        if n < 3:
            return
        k = n - 1 if n == 3 else 3
        with torch.no_grad():
            self.mus.data[0] = self.p_min
            basis_mus = torch.zeros(k, n, dtype=torch.float32, device=self.mus.device)
            basis_mus[:, 0] = self.p_min
            for j in range(1, k + 1):
                basis_mus[j - 1, j] = 1.0
            idx = torch.arange(self.nangles, device=self.vertices.device)
            _, A, rhs = self.fp.forward_multi(self.get_displaced_vertices(), self.faces, idx, basis_mus,
                                              target=p_cuda.to(torch.float32).contiguous())
            est = torch.pinverse(A.to(self.dtype).cpu()).matmul(rhs.to(self.dtype).cpu())
            self.mus.data[1:] = 0
            for j in range(1, k + 1):
                self.mus.data[j] = est[j - 1]

    def get_vf_np(self, return_as_np=True):
        vertices = self.get_displaced_vertices().detach()
        if return_as_np:
            return vertices.cpu().numpy(), self.faces.cpu().numpy(), self.labels_v_np, self.labels.cpu().numpy()
