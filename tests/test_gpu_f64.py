"""The float64 instantiation on the device (``*_f64`` entry points): the reference dispatches its kernels over float AND
double (``diff_fp_for.cu:251``, ``diff_fp_back.cu:326``) and ``FP`` / ``Model`` take a ``dtype``
(``fp_vecs.py:53-56``, ``vanilla.py:27,31``).

Checked against (a) the goldens the reference's own double kernels produced on a B200 (``ref_raster_*.npz``: ``im64``,
``cnt64``, ``dldp64`` ...), (b) the reference extension live in double, (c) the reference's FP modules / autograd in
float64 (``vertex_stage_*_f64.npz``, ``pipeline_*_f64.npz``).
"""
import glob
import os
import sys

import numpy as np
import pytest
import torch

import cases
from test_gpu_raster import dev
from test_oracle_golden import _geom_from_npz, golden_cotangents

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
from gen_ref_golden import GOLDEN_CASES  # noqa: E402  (seeded inputs; only outputs are stored)


def _as64(c):
    d = torch.float64
    return (dev(c["p6"], d), dev(c["len3"], d), dev(c["ff"], d).reshape(c["p6"].shape[0], -1, 1), dev(c["labels"]), dev(c["mus"], d))


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_raster_f64_vs_reference_double_kernels_golden(cuda_lib, name):
    from shapefromprojections_b200.function import rasterizer as R
    g = np.load(os.path.join(HERE, "golden", f"ref_raster_{name}.npz"))
    c = GOLDEN_CASES[name]()
    p6, len3, ff, labels, mus = _as64(c)
    im, cnt, stats = R.raster_forward(p6, len3, ff, labels, mus, c["H"], c["W"], c["msv"], want_cnt=True, want_stats=True)
    torch.cuda.synchronize()
    assert im.dtype == torch.float64
    assert np.array_equal(cnt.cpu().numpy(), g["cnt64"])                       # coverage: exact
    ref = g["im64"]
    err = np.abs(im.cpu().numpy() - ref).max()
    same = np.array_equal(im.cpu().numpy().view(np.uint64), ref.view(np.uint64))
    print(f"[{name}] float64 sinogram vs reference double kernels: max abs diff {err:.2e} (bit-identical: {same})")
    assert np.allclose(im.cpu().numpy(), ref, rtol=1e-14, atol=1e-13)
    assert int(stats[1]) == int(g["cnt64"].sum())                              # fragments == sum of coverage counts
    # backward: the reference adds per fragment with double atomics (order undefined) -> reassociation noise only
    grads = R.raster_backward(dev(g["grad_im"], torch.float64), dev(ref, torch.float64), p6, len3, ff, labels, mus, c["msv"])
    for a, key in zip(grads, ("dldp64", "dldf64", "dldmu64")):
        r = g[key].ravel()
        rel = np.linalg.norm(a.cpu().numpy().ravel() - r) / (np.linalg.norm(r) + 1e-300)
        print(f"[{name}] {key}: rel {rel:.2e}")
        assert rel <= 1e-11, (key, rel)


@pytest.mark.parametrize("name", ["sphere_parallel", "two_material_cone", "torus_axis_aligned"])
def test_raster_f64_vs_reference_extension_live(cuda_lib, name):
    from oracle import ref_driver
    from shapefromprojections_b200 import cuda_diff_fp as compat
    from test_gpu_reference_ext import REF_CASES
    if not ref_driver.available():
        pytest.fail("oracle/_ref/cuda_diff_fp.so is missing: run `python oracle/build_ref.py` before gpurun")
    c = REF_CASES[name]()
    r = ref_driver.forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"], torch.float64)
    # through the cuda_diff_fp-compatible module, like the reference's rasterizer.py:51-55 would call it
    im = torch.zeros_like(r["im"]); cnt = torch.zeros_like(r["cnt"])
    compat.forward(r["p6"], r["ff"], r["bbox"], r["len3"], r["labels"], r["mus"], r["imwei"], r["imidx"], cnt,
                   r["firstidx"], im, 1, c["msv"])
    torch.cuda.synchronize()
    assert torch.equal(cnt, r["cnt"])
    assert torch.allclose(im, r["im"], rtol=1e-14, atol=1e-13)
    g = np.random.default_rng(5).normal(size=(c["p6"].shape[0], c["H"], c["W"], 1))
    gref = ref_driver.backward(r, g)
    dp = torch.zeros_like(r["p6"]); dl = torch.zeros_like(r["len3"]); dm = torch.zeros_like(r["mus"])
    compat.backward(dev(g, torch.float64), r["im"], r["imidx"], r["cnt"], r["firstidx"], r["imwei"], r["p6"], r["len3"],
                    r["labels"], r["mus"], r["ff"], dp, dl, dm)
    torch.cuda.synchronize()
    for mine, ref, key in zip((dp, dl, dm), gref, ("grad_p6", "grad_len3", "grad_mus")):
        rel = float((mine - ref).norm() / (ref.norm() + 1e-300))
        print(f"[{name}] {key} float64 vs reference extension: rel {rel:.2e}")
        assert rel <= 1e-11, (key, rel)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "vertex_stage_*_f64.npz"))))
def test_vertex_stage_f64_vs_reference_modules(cuda_lib, path):
    """Forward and autograd of the float64 vertex stage vs the reference's FP modules run in float64."""
    from shapefromprojections_b200.function import rasterizer as R
    g = np.load(path)
    geom = _geom_from_npz(g)
    gd = R.GeometryDevice(geom, "cuda")
    v = dev(g["verts"], torch.float64).requires_grad_(True)
    p6, len3, ff = R.VertexStage.apply(v, dev(g["faces"], torch.int32), g["idx_angles"], gd)
    assert p6.dtype == torch.float64 and ff.shape == (p6.shape[0], p6.shape[1], 1)
    e_px, e_len = np.abs(p6.detach().cpu().numpy() - g["p6"]).max(), np.abs(len3.detach().cpu().numpy() - g["len3"]).max()
    print(f"[{geom['type']}] float64 vertex stage vs reference modules: max |dp| = {e_px:.2e} px, max |dlen| = {e_len:.2e}")
    assert e_px <= 1e-10 and e_len <= 1e-12
    big = np.abs(g["ff"]) > 1e-14
    assert np.array_equal((ff.detach().cpu().numpy() < 0)[big], (g["ff"] < 0)[big])
    gp6, gl3 = golden_cotangents(g)
    ((p6 * dev(gp6, torch.float64)).sum() + (len3 * dev(gl3, torch.float64)).sum()).backward()
    torch.cuda.synchronize()
    ref = g["grad_verts"]
    rel = np.linalg.norm(v.grad.cpu().numpy() - ref) / np.linalg.norm(ref)
    print(f"[{geom['type']}] float64 grad_verts vs reference autograd: rel {rel:.2e}")
    assert rel <= 1e-10, rel


@pytest.mark.parametrize("kind", ["parallel3d", "cone_vec"])
def test_model_in_float64_end_to_end(cuda_lib, kind):
    """``Model(..., dtype=torch.float64)`` (vanilla.py:27,31): forward + loss.backward() in double on the device, and the
    float32 model agrees with it to float32 rounding."""
    from shapefromprojections_b200.model.vanilla import Model
    geom = cases.parallel_geometry(5, 96) if kind == "parallel3d" else cases.cone_geometry(5, 96)
    verts, faces, labels, mus = cases.sphere_mesh(2, 0.45)
    lv = np.ones(verts.shape[0], np.int32)
    out = {}
    for dt in (torch.float64, torch.float32):
        model = Model((verts, faces, lv, labels), geom, 2, [0.0, 1.0], 1, dtype=dt).cuda()
        assert model.fp.dtype == dt and (dt == torch.float32 or model.fp.fused is False)
        phat, mask, edge, lap, flat = model(torch.arange(5), 1.0)
        assert phat.dtype == dt and mask.dtype == torch.bool
        w = torch.linspace(0.5, 1.5, phat.numel(), device="cuda", dtype=dt).reshape(phat.shape)
        ((phat * w)[mask].sum() + edge).backward()
        out[dt] = (phat.detach().double().cpu().numpy(), model.displace.grad.double().cpu().numpy(), model.mus.grad.double().cpu().numpy())
    p64, gv64, gm64 = out[torch.float64]
    p32, gv32, gm32 = out[torch.float32]
    assert np.abs(p64).max() > 0.1 and np.linalg.norm(gv64) > 0
    far = np.abs(p32 - p64) > 1e-5 * np.abs(p64) + 1e-5
    assert far.mean() <= 2e-4, f"{far.sum()} pixels of the float32 sinogram off the float64 one"      # knife-edge pixels only
    assert np.linalg.norm(gv32 - gv64) <= 2e-3 * np.linalg.norm(gv64)      # float32 coverage differs on knife-edge pixels
    assert np.linalg.norm(gm32 - gm64) <= 1e-4 * np.linalg.norm(gm64)
