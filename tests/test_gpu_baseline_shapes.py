"""GPU parity AT THE BASELINE SHAPES against the reference's own compiled CUDA kernels (oracle/_ref).

Every case is one of BASELINE.json's configurations at its real mesh and detector size (on a subset of
angles where the reference's O(pixels x faces) loop, ctdr/cuda/diff_fp_for.cu:83, would take minutes):

  cfg1       30 x 192 x 192 parallel3d, icosphere(4)                      -- in full
  cfg3       cone 1024 x 1024, torus 100k faces                           -- face-on, 45 degrees, edge-on
  northstar  cone 1024 x 1024, torus 1M faces (fine-mesh DIRECT variant, RT = 4, list segments) -- 1 angle
  cfg4       parallel3d 2048 x 2048, torus 1M faces                       -- 1 angle (+ 64-bit pixel offsets)
  cfg5       cone 512 x 512, one of the 64 rotated 20k-face tori          -- 4 angles

For each: the reference runs on the p6/len3/ff our vertex stage produces (identical inputs at the
Rasterizer boundary, SURVEY.md section 7 hard part 1); `im` and `cnt` must be BIT-EXACT both through the
Rasterizer-boundary entry point and through the fused ones (ctdr_project_forward, ctdr_project_residual_step);
gradients follow ||g - g64|| <= max(1e-4 ||g64||, 2 ||g_ref32 - g64||) with g64 the reference kernel
instantiated in double (diff_fp_back.cu:326), for grad_p6 / grad_len3 / grad_mus and — chained through the
vertex-stage transpose, itself pinned to the reference's autograd in test_gpu_vertex_grad.py — grad_verts.
"""
import numpy as np
import pytest
import torch

from oracle import ref_driver
from test_gpu_raster import dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_driver.available():
        pytest.fail("oracle/_ref/cuda_diff_fp.so is missing: run `python oracle/build_ref.py` before gpurun")
    return ref_driver


def load(name):
    from shapefromprojections_b200 import workloads
    from shapefromprojections_b200.function import rasterizer as R
    wl = workloads.make_workload(name)
    gd = R.GeometryDevice(wl["geom"], "cuda")
    return wl, gd, dev(wl["verts"]), dev(wl["faces"], torch.int32), dev(wl["labels"]), dev(wl["mus"])


def rule(name, g, g64, g32):
    """The stated gradient tolerance; returns the achieved relative error."""
    g, g64, g32 = (x.double().reshape(-1) for x in (g, g64, g32))
    err, n64 = float((g - g64).norm()), float(g64.norm())
    tol = max(1e-4 * n64, 2 * float((g32 - g64).norm()))
    assert err <= tol + 1e-30, f"{name}: |g-g64|={err:.3e} > tol={tol:.3e} (|g64|={n64:.3e})"
    return err / max(n64, 1e-300)


def rule32(name, g, g32, rel=2e-4, cond=None):
    """No double referee (too slow at this size): compare with the reference's float instantiation, whose own
    float atomics carry reassociation noise.  ``cond`` = the same gradient for |cotangent| (no cancellation):
    grad_mus is a sum over ~1e6 fragments of either sign, so its noise scales with that, not with the result."""
    err, n = float((g.double() - g32.double()).norm()), float(g32.double().norm())
    tol = rel * n + (1e-5 * float(cond.double().norm()) if cond is not None else 0.0)
    assert err <= tol + 1e-30, f"{name}: |g-g32|={err:.3e} > {tol:.3e} (|g32|={n:.3e})"
    return err / max(n, 1e-300)


def vertex_transpose(gd, v, f32, idx, gp6, glen3):
    """grad_verts from Rasterizer-boundary cotangents through ctdr_vertex_backward."""
    from shapefromprojections_b200 import _lib
    gv = torch.zeros_like(v)
    gp6 = gp6.float().contiguous(); glen3 = glen3.float().contiguous()
    rc = _lib.lib().ctdr_vertex_backward(gd.kind, gd.table.data_ptr(), gd.nangles, idx.data_ptr(), idx.numel(),
                                         v.data_ptr(), v.shape[0], f32.data_ptr(), f32.shape[0], gd.H, gd.W,
                                         gd.spacing_x, gd.spacing_y, float(gd.max_sino_val), gp6.data_ptr(),
                                         glen3.data_ptr(), gv.data_ptr(), _lib.stream_ptr(v.device))
    _lib.check(rc, "ctdr_vertex_backward")
    return gv


def check_case(ref, name, angles, *, fp64=True, expect_segments=False, expect_direct=False):
    from shapefromprojections_b200.function import rasterizer as R
    wl, gd, v, f32, lab, mus = load(name)
    idx = torch.as_tensor(angles, dtype=torch.int64, device="cuda")
    B, H, W, msv = idx.numel(), gd.H, gd.W, gd.max_sino_val
    F = f32.shape[0]
    if expect_direct:
        assert F >= 0.4 * H * W            # the per-call heuristic picks the fine-mesh (DIRECT) kernels
    p6, len3, ff = R.vertex_forward(gd, v, f32, idx)

    # ---- forward: Rasterizer boundary and both fused entry points, bit-exact vs the reference kernels ----
    r32 = ref.forward(p6, len3, ff, lab, mus, H, W, msv)
    im, cnt, stats = R.raster_forward(p6, len3, ff, lab, mus, H, W, msv, want_cnt=True, want_stats=True)
    assert torch.equal(cnt, r32["cnt"]), f"{name}: cnt differs on {(cnt != r32['cnt']).sum().item()} pixels"
    assert torch.equal(im.view(torch.int32), r32["im"].view(torch.int32)), \
        f"{name}: {(im != r32['im']).sum().item()} sinogram pixels differ, max {(im - r32['im']).abs().max().item():.3e}"
    assert int(stats[1]) == r32["vf"] and int(stats[0]) == 0          # fragments == the reference's vf (rasterizer.py:64)
    if expect_segments:
        assert int(stats[3]) > 0                                      # region chunk lists took several segments
    im_lean, _, _ = R.raster_forward(p6, len3, ff, lab, mus, H, W, msv)        # product instantiation (no counters)
    assert torch.equal(im_lean.view(torch.int32), r32["im"].view(torch.int32))
    imf, cntf, _ = R.project_forward(gd, v, f32, lab, mus, idx, want_cnt=True)
    assert torch.equal(cntf, r32["cnt"].view(B, H, W))
    assert torch.equal(imf.view(torch.int32), r32["im"].view(B, H, W).view(torch.int32))
    target = torch.rand(B, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(11)) * 1.2
    phat, gv_step, gm_step, out2 = R.project_residual_step(gd, v, f32, lab, mus, idx, target)
    assert torch.equal(phat.view(torch.int32), r32["im"].view(B, H, W).view(torch.int32))
    imr = r32["im"].view(B, H, W)
    mask = (imr > -1e-6) & (imr <= msv - 1e-6)                        # fp_vecs.py:10-13
    assert int(out2[1].item()) == int(mask.sum())
    loss_ref = float(((target - imr).double() ** 2)[mask].sum())
    assert abs(out2[0].item() - loss_ref) <= 1e-9 * loss_ref

    # ---- backward at the Rasterizer boundary -----------------------------------------------------------
    g = torch.randn(B, H, W, 1, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    g_resid = (2.0 * (imr - target) * mask).unsqueeze(-1).contiguous()           # optimize.py:168 cotangent
    achieved = {}
    same_cover, r64 = False, None
    if fp64:        # the double instantiation as referee, usable when it decides every knife-edge pixel like the float one
        r64 = ref.forward(p6, len3, ff, lab, mus, H, W, msv, torch.float64)
        bad32 = (r32["im"].double() < -1e-5) | (r32["im"] > msv)
        bad64 = (r64["im"] < -1e-5) | (r64["im"] > msv)
        same_cover = torch.equal(r64["cnt"], r32["cnt"]) and torch.equal(bad32, bad64)
    # With the double referee the stated rule applies as is: where the double instantiation decides a knife-edge
    # pixel differently (same_cover False) that shows up in ||g_ref32 - g64|| and the rule loosens by itself.
    for tag, gg in (("randn", g), ("residual", g_resid)):
        g32 = ref.backward(r32, gg)
        mine = R.raster_backward(gg, r32["im"], p6, len3, ff, lab, mus, msv)
        if fp64:
            g64 = ref.backward(r64, gg)
            cond = None
        else:
            cond = ref.backward(r32, gg.abs())[2]
        names = ("grad_p6", "grad_len3", "grad_mus")
        for k, nme in enumerate(names):
            achieved[f"{tag}/{nme}"] = (rule(nme, mine[k], g64[k], g32[k]) if fp64
                                        else rule32(nme, mine[k], g32[k], cond=cond if k == 2 else None))
        # ---- fused backward: grad_verts / grad_mus vs (reference raster backward) o (vertex transpose) ----
        gv32 = vertex_transpose(gd, v, f32, idx, g32[0], g32[1])
        if tag == "randn":
            gv, gm = R.project_backward(gd, v, f32, lab, mus, idx, gg.view(B, H, W).contiguous(), imf)
        else:
            gv, gm = gv_step, gm_step
        if fp64:
            gv64 = vertex_transpose(gd, v, f32, idx, g64[0], g64[1])
            achieved[f"{tag}/grad_verts"] = rule("grad_verts (fused)", gv, gv64, gv32)
            achieved[f"{tag}/grad_mus_fused"] = rule("grad_mus (fused)", gm, g64[2], g32[2])
        else:
            achieved[f"{tag}/grad_verts"] = rule32("grad_verts (fused)", gv, gv32)
            achieved[f"{tag}/grad_mus_fused"] = rule32("grad_mus (fused)", gm, g32[2], cond=cond)
        assert float(gv.norm()) > 0
    print(f"[{name}] B={B} F={F} {H}x{W} vf={r32['vf']} same_cover_fp64={same_cover} achieved rel. errors: "
          + ", ".join(f"{k}={e:.2e}" for k, e in achieved.items()))


def test_cfg1_full(cuda_lib, ref):
    check_case(ref, "cfg1", list(range(30)))


def test_cfg3_face_on_45deg_edge_on(cuda_lib, ref):
    check_case(ref, "cfg3", [0, 90, 180])              # theta = 0, pi/4, pi/2 of the 720-angle orbit


def test_northstar_1m_faces_one_angle(cuda_lib, ref):
    check_case(ref, "northstar", [45], expect_segments=True, expect_direct=True)


def test_cfg4_1m_faces_2048_one_angle(cuda_lib, ref):
    check_case(ref, "cfg4", [225], fp64=False)


def test_cfg5_one_mesh(cuda_lib, ref):
    check_case(ref, "cfg5", [0, 45, 90, 200])


def test_cfg5_other_meshes_forward(cuda_lib, ref):
    """Two more of the 64 meshes (different radii and rotations), forward only, like simulation_cone_beam.py."""
    from shapefromprojections_b200 import workloads
    from shapefromprojections_b200.function import rasterizer as R
    wl, gd, _, _, lab, mus = load("cfg5")
    idx = torch.tensor([10, 130], device="cuda")
    for i in (17, 63):
        vv, ff_ = workloads.cfg5_mesh(i)
        v, f32 = dev(vv.astype(np.float32)), dev(ff_, torch.int32)
        p6, len3, ff = R.vertex_forward(gd, v, f32, idx)
        r = ref.forward(p6, len3, ff, lab, mus, gd.H, gd.W, gd.max_sino_val)
        im, cnt, _ = R.project_forward(gd, v, f32, lab, mus, idx, want_cnt=True)
        assert torch.equal(cnt, r["cnt"].view_as(cnt)) and torch.equal(im.view(torch.int32), r["im"].view_as(im).view(torch.int32))


def test_cfg4_pixel_offsets_beyond_int32(cuda_lib):
    """cfg4's full run has 7.55e9 pixels: the reference's `int` pixel index overflows (diff_fp_for.cu:51).
    600 angles x 2048 x 2048 = 2.5e9 pixels in ONE call: the slices beyond 2^31 must equal single-angle calls
    bit for bit, and the residual step's sums must add up over two half calls."""
    from shapefromprojections_b200.function import rasterizer as R
    free, _ = torch.cuda.mem_get_info()
    if free < 70 * 2**30:
        pytest.skip("needs ~70 GB of device memory")
    wl, gd, v, f32, lab, mus = load("cfg4")
    B = 600
    idx = torch.arange(0, 1800, 3, device="cuda")
    assert B * gd.H * gd.W > 2**31
    im, _, _ = R.project_forward(gd, v, f32, lab, mus, idx)
    for b in (0, 511, 512, 599):                       # 512 * 2048^2 = 2^31 exactly
        one, _, _ = R.project_forward(gd, v, f32, lab, mus, idx[b:b + 1].contiguous())
        assert torch.equal(one[0], im[b]), f"angle slot {b}"
    assert float(im[599].max()) > 0.4                  # the tube's chord, 2r = 0.48
    target = im * 0.9
    phat, gv, gm, out2 = R.project_residual_step(gd, v, f32, lab, mus, idx, target)
    assert torch.equal(phat, im)
    del im
    acc_v, acc_m, acc_o = torch.zeros_like(gv), torch.zeros_like(gm), torch.zeros_like(out2)
    for a, b in ((0, 300), (300, 600)):
        _, gvh, gmh, o2 = R.project_residual_step(gd, v, f32, lab, mus, idx[a:b].contiguous(), target[a:b].contiguous())
        acc_v += gvh; acc_m += gmh; acc_o += o2
    assert acc_o[1].item() == out2[1].item() and abs(acc_o[0].item() - out2[0].item()) <= 1e-9 * out2[0].item()
    assert float((acc_v - gv).norm()) <= 1e-4 * float(gv.norm()) and float(gv.norm()) > 0
    assert float((acc_m - gm).norm()) <= 1e-4 * float(gm.norm())
