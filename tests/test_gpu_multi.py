"""GPU tests of the multi-output raster pass (Model.estimate_mu from one sweep).

Every output must equal, bit for bit, the single-output forward run with that output's mus — which is
itself pinned to the oracle / the reference kernels by test_gpu_raster.py — and the fused normal
equations must equal float64 sums over those sinograms.
"""
import numpy as np
import pytest
import torch

import cases
from oracle import oracle
from test_gpu_raster import CASES, bits, dev

pytestmark = pytest.mark.gpu


def mus_variants(mus, k, seed=0):
    """k attenuation vectors: the case's own, then one-hot style rows with a non-zero background."""
    rng = np.random.default_rng(seed)
    rows = [mus]
    for j in range(1, k):
        r = np.zeros_like(mus)
        r[0] = 0.37                       # estimate_mu renders with mus[0] = p_min != 0
        r[1 + (j - 1) % (len(mus) - 1)] = 1.0
        if j == 3:
            r = rng.uniform(-1, 2, size=len(mus)).astype(np.float32)
        rows.append(r)
    return np.stack(rows).astype(np.float32)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("name", ["soup_small", "soup_ragged", "soup_many_faces", "two_material_cone", "torus_axis_aligned"])
def test_multi_outputs_equal_single_forward_bitwise(cuda_lib, name, k):
    from shapefromprojections_b200.function import rasterizer as R
    c = CASES[name]()
    mk = mus_variants(c["mus"], k)
    rng = np.random.default_rng(5)
    B = c["p6"].shape[0]
    target = rng.normal(size=(B, c["H"], c["W"])).astype(np.float32)
    args = (dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]))
    im, A, rhs, stats = R.raster_forward_multi(*args, dev(mk), c["H"], c["W"], c["msv"], target=dev(target), want_stats=True)
    torch.cuda.synchronize()
    im = im.cpu().numpy()
    singles = []
    for j in range(k):
        one, cnt, st1 = R.raster_forward(*args, dev(mk[j]), c["H"], c["W"], c["msv"], want_cnt=True, want_stats=True)
        one = one.cpu().numpy().reshape(B, c["H"], c["W"])
        singles.append(one)
        neq = bits(im[j]) != bits(one)
        assert not neq.any(), f"output {j}: {neq.sum()} pixels differ, first {np.argwhere(neq)[:4]}"
    assert stats[1].item() == st1[1].item() and stats[0].item() == st1[0].item()      # fragments, non-positive lengths
    # output 0 uses the case's own mus: the oracle's sinogram
    ref = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"], False, fast=True)
    assert np.array_equal(bits(im[0]), bits(ref["im"].reshape(B, c["H"], c["W"])))
    # normal equations in float64
    S = np.stack(singles).astype(np.float64).reshape(k, -1)
    A64 = S @ S.T
    r64 = S @ target.astype(np.float64).reshape(-1)
    A, rhs = A.cpu().numpy(), rhs.cpu().numpy()
    assert np.allclose(A, A64, rtol=1e-11, atol=1e-9 * max(1.0, np.abs(A64).max()))
    assert np.allclose(rhs, r64, rtol=1e-11, atol=1e-9 * max(1.0, np.abs(r64).max()))
    assert np.array_equal(A, A.T)


def test_project_multi_windows_and_gram(cuda_lib, monkeypatch):
    """Fused path: mesh in, K sinograms + normal equations out; angle windows must not change anything."""
    from shapefromprojections_b200 import _lib
    from shapefromprojections_b200.function import rasterizer as R
    geom = cases.cone_geometry(7, 160)
    verts, faces, labels, mus = cases.two_material_mesh(2)
    g = R.GeometryDevice(geom, "cuda")
    v, f, lab = dev(verts), dev(faces, torch.int32), dev(labels)
    idx = torch.arange(7, device="cuda")
    mk = mus_variants(mus, 3)
    target = torch.rand(7, 160, 160, device="cuda")
    im, A, rhs, _ = R.project_forward_multi(g, v, f, lab, dev(mk), idx, target=target)
    for j in range(3):
        one, _, _ = R.project_forward(g, v, f, lab, dev(mk[j]), idx)
        assert torch.equal(im[j], one)
    S = im.double().reshape(3, -1)
    assert torch.allclose(A, S @ S.T, rtol=1e-11) and torch.allclose(rhs, S @ target.double().reshape(-1), rtol=1e-11)
    # two-angle windows: shrink the workspace the binding hands to the library
    L = _lib.lib()
    small = L.ctdr_project_workspace_bytes(2, verts.shape[0], faces.shape[0], 160, 160, 0)
    monkeypatch.setattr(_lib, "workspace", lambda device, nbytes: torch.empty(small, dtype=torch.uint8, device=device))
    im2, A2, rhs2, _ = R.project_forward_multi(g, v, f, lab, dev(mk), idx, target=target)
    assert torch.equal(im2, im)
    assert torch.allclose(A2, A, rtol=1e-12) and torch.allclose(rhs2, rhs, rtol=1e-12)


def test_multi_argument_errors(cuda_lib):
    from shapefromprojections_b200.function import rasterizer as R
    c = CASES["soup_one_face"]()
    args = (dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]))
    with pytest.raises(RuntimeError, match="n_out"):
        R.raster_forward_multi(*args, dev(np.zeros((5, len(c["mus"])), np.float32)), c["H"], c["W"], c["msv"])


def test_estimate_mu_single_pass_matches_per_material_renders(cuda_lib):
    """Model.estimate_mu (vanilla.py:186-265): attenuations recovered from a sinogram of known mus; the
    one-pass normal equations give the same estimate as the reference's render-per-material recipe."""
    from ctdr.model.vanilla import Model
    geom = cases.cone_geometry(12, 128)
    verts, faces, labels, _ = cases.two_material_mesh(2)
    labels_v = np.ones(verts.shape[0], dtype=np.int32)
    true_mus = [0.0, 0.8, 1.9]
    model = Model((verts, faces, labels_v, labels), geom, 3, true_mus, 0, wlap=0.0, wflat=0.0).cuda()
    with torch.no_grad():
        p = model.render()[0].clone()
    # the reference's recipe, written out with separate renders
    with torch.no_grad():
        pmin = p.min()
        basis = []
        for j in (1, 2):
            model.mus.data[:] = 0
            model.mus.data[0] = pmin
            model.mus.data[j] = 1
            basis.append(model.render()[0].double())
        A = torch.stack([torch.stack([(a * b).sum() for b in basis]) for a in basis]).cpu()
        rhs = torch.stack([(a * p.double()).sum() for a in basis]).cpu()
        expect = torch.pinverse(A).matmul(rhs)
    model.mus.data[:] = torch.tensor([0.0, 0.1, 0.1], device="cuda")
    model.estimate_mu(p)
    got = model.mus.detach().cpu().double()
    assert torch.allclose(got[1:], expect, rtol=1e-5, atol=1e-6)
    assert abs(got[1].item() - 0.8) < 1e-3 and abs(got[2].item() - 1.9) < 1e-3
    assert got[0].item() == pmin.item()
