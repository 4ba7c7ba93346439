"""GPU: vertex gradients pinned to the REFERENCE'S OWN autograd.

tests/golden/vertex_stage_*_f64.npz hold ``grad_verts`` obtained by calling ``torch.autograd.grad`` through the
reference's unmodified ``ctdr/nn/fp_opengl.py:87-158`` / ``ctdr/nn/fp_vecs.py:85-196`` (float64, seeded
cotangents on p_bxfx6 / len_bxfx3; generator: tests/golden/make_vertex_golden.py).  ``ctdr_vertex_backward``
(the hand-written transposed Jacobian, fp32) must agree within north_star's 1e-4 for all three geometry kinds;
the achieved error is printed.  The fused path (``ctdr_project_backward`` / ``_residual_step``) shares these
per-vertex formulas (``vertex_vjp``) and is compared with the staged path in test_gpu_project.py and with the
reference kernels chained through this transpose in test_gpu_baseline_shapes.py.
"""
import glob
import os

import numpy as np
import pytest
import torch

from test_gpu_raster import dev
from test_oracle_golden import _geom_from_npz, golden_cotangents

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "vertex_stage_*_f64.npz"))))
def test_vertex_backward_vs_reference_autograd(cuda_lib, path):
    from shapefromprojections_b200.function import rasterizer as R
    g = np.load(path)
    geom = _geom_from_npz(g)
    gd = R.GeometryDevice(geom, "cuda")
    gp6, gl3 = golden_cotangents(g)
    v = dev(g["verts"], torch.float32).requires_grad_(True)
    p6, len3, ff = R.VertexStage.apply(v, dev(g["faces"], torch.int32), g["idx_angles"], gd)
    ((p6 * dev(gp6)).sum() + (len3 * dev(gl3)).sum()).backward()
    torch.cuda.synchronize()
    got, ref = v.grad.cpu().numpy().astype(np.float64), g["grad_verts"]
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    g32 = np.load(path.replace("_f64", "_f32"))["grad_verts"].astype(np.float64)     # the reference's own float32 run
    rel_ref32 = np.linalg.norm(g32 - ref) / np.linalg.norm(ref)
    print(f"[{geom['type']}] grad_verts vs reference fp64 autograd: rel {rel:.2e} (reference fp32 vs fp64: {rel_ref32:.2e})")
    assert rel <= 1e-4, rel
    assert rel <= max(2e-6, 20 * rel_ref32), (rel, rel_ref32)       # same league as the reference's own fp32 rounding


def _rule(name, g, g64, g32):
    g, g64, g32 = (np.asarray(x, dtype=np.float64).ravel() for x in (g, g64, g32))
    err, n64 = np.linalg.norm(g - g64), np.linalg.norm(g64)
    tol = max(1e-4 * n64, 2 * np.linalg.norm(g32 - g64))
    assert err <= tol, f"{name}: |g-g64|={err:.3e} > tol={tol:.3e} (|g64|={n64:.3e}, |g32-g64|={np.linalg.norm(g32 - g64):.3e})"
    return err / n64


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "pipeline_*_f64.npz"))))
def test_fused_gradients_vs_reference_pipeline(cuda_lib, path):
    """END TO END: verts.grad / mus.grad of the fused device path vs ``loss.backward()`` through the reference's
    own FP modules (torch autograd) + the reference kernels' arithmetic (C oracle, pinned bit-exactly to them):
    tests/golden/make_pipeline_golden.py.  Losses: sum(w * phat) and the masked L2 of ctdr/optimize.py:168.
    Rule: ||g - g64|| <= max(1e-4 ||g64||, 2 ||g_ref32 - g64||); achieved errors printed,
    together with how far the fused sinogram is from the reference's (north_star: 1e-5 relative forward)."""
    import cases
    from pipeline_golden_seeds import target, weights
    from shapefromprojections_b200.function import rasterizer as R
    g64, g32 = np.load(path), np.load(path.replace("_f64", "_f32"))
    geom = _geom_from_npz(g64)
    gd = R.GeometryDevice(geom, "cuda")
    verts, faces, labels, mus = cases.two_material_mesh(2)
    idx = torch.as_tensor(g64["idx_angles"]).cuda()
    B, H, W = idx.numel(), gd.H, gd.W
    f32, lab = dev(faces, torch.int32), dev(labels)
    v = dev(verts).requires_grad_(True); m = dev(mus).requires_grad_(True)
    im = R.ProjectMesh.apply(v, m, f32, lab, idx, gd, True)
    (im * dev(weights(B, H, W))).sum().backward()
    torch.cuda.synchronize()
    # forward: distance of the fused sinogram from the reference modules' (different fp32 op chain upstream of
    # the rasteriser: knife-edge pixels may flip, everything else agrees to rounding)
    ref_im = g32["phat"]
    d = np.abs(im.detach().cpu().numpy() - ref_im)
    scale = np.abs(ref_im).max()
    far = d > 1e-5 * np.abs(ref_im) + 1e-5 * scale
    e_w = {"grad_verts": _rule("grad_verts (sum w*phat)", v.grad.cpu().numpy(), g64["grad_verts_w"], g32["grad_verts_w"]),
           "grad_mus": _rule("grad_mus (sum w*phat)", m.grad.cpu().numpy(), g64["grad_mus_w"], g32["grad_mus_w"])}
    tgt = dev(target(B, H, W))
    phat, gv, gm, out2 = R.project_residual_step(gd, v.detach(), f32, lab, m.detach(), idx, tgt)
    torch.cuda.synchronize()
    n = out2[1].item()
    assert abs(n - int(g64["n_valid"])) <= 2e-4 * n
    assert abs(out2[0].item() / n - float(g64["loss_l2"])) <= 2e-5 * float(g64["loss_l2"])
    e_l2 = {"grad_verts": _rule("grad_verts (masked L2)", (gv / n).cpu().numpy(), g64["grad_verts_l2"], g32["grad_verts_l2"]),
            "grad_mus": _rule("grad_mus (masked L2)", (gm / n).cpu().numpy(), g64["grad_mus_l2"], g32["grad_mus_l2"])}
    print(f"[{geom['type']}] fused vs reference pipeline: sinogram max|d|={d.max():.2e} (scale {scale:.2f}), median rel "
          f"{np.median(d[ref_im != 0] / np.abs(ref_im[ref_im != 0])):.1e}, pixels beyond 1e-5: {int(far.sum())} of {far.size}; "
          f"sum-w grads {e_w}; masked-L2 grads {e_l2}")
    assert far.mean() <= 2e-4
