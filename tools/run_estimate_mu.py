"""Timing of the multi-output raster pass behind Model.estimate_mu at config-3 size:
cone 720 x 1024 x 1024, torus of 100k triangles enclosing a smaller torus (three materials).
Compares K separate renders + torch reductions (the reference's recipe, vanilla.py:186-265) with
one ctdr_project_forward_multi call that also forms the normal equations.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapefromprojections_b200.function import rasterizer as R  # noqa: E402
from shapefromprojections_b200.utils import meshgen  # noqa: E402
from shapefromprojections_b200.workloads import make_workload  # noqa: E402


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def main(name="cfg3"):
    dev = torch.device("cuda", 0)
    w = make_workload(name)
    gd = R.GeometryDevice(w["geom"], dev)
    # outer torus (label 1 in 0) + inner torus (label 2 in 1)
    iv, iface = meshgen.torus(0.10, 0.48, 100, 160)
    verts = np.concatenate([w["verts"], (iv + 1e-4).astype(np.float32)])
    faces = np.concatenate([w["faces"], iface + w["verts"].shape[0]])
    labels = np.concatenate([w["labels"], np.tile(np.array([2, 1], np.int32), (iface.shape[0], 1))])
    v = torch.as_tensor(verts).to(dev); f = torch.as_tensor(faces).to(torch.int32).to(dev); lab = torch.as_tensor(labels).to(dev)
    idx = torch.arange(w["nangles"], device=dev)
    true_mus = torch.tensor([0.0, 0.8, 1.9], device=dev)
    p, _, _ = R.project_forward(gd, v, f, lab, true_mus, idx)
    basis_mus = torch.tensor([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], device=dev)
    K = basis_mus.shape[0]

    def separate():
        ims = [R.project_forward(gd, v, f, lab, basis_mus[j].contiguous(), idx)[0] for j in range(K)]
        A = torch.stack([torch.stack([(a * b).sum() for b in ims]) for a in ims])
        rhs = torch.stack([(a * p).sum() for a in ims])
        return A, rhs

    def fused():
        _, A, rhs, _ = R.project_forward_multi(gd, v, f, lab, basis_mus, idx, target=p)
        return A, rhs

    ms_sep, (A0, r0) = timed(separate)
    ms_one, (A1, r1) = timed(fused)
    ms_single, _ = timed(lambda: R.project_forward(gd, v, f, lab, true_mus, idx))
    est0 = torch.linalg.pinv(A0.double().cpu()) @ r0.double().cpu()
    est1 = torch.linalg.pinv(A1.cpu()) @ r1.cpu()
    print(json.dumps({"config": f"{name} + inner torus: V={verts.shape[0]} F={faces.shape[0]}, {K} basis renders",
                      "ms_separate_renders_plus_torch_sums": ms_sep, "ms_single_multi_output_pass": ms_one,
                      "ms_one_plain_render": ms_single, "speedup": ms_sep / ms_one,
                      "mus_estimated_separate": est0.tolist(), "mus_estimated_fused": est1.tolist()}))


if __name__ == "__main__":
    main(*sys.argv[1:])
