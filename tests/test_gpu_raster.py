"""GPU parity at the Rasterizer boundary: CUDA kernels (through the C ABI) vs the CPU oracle and,
when oracle/_ref is present, vs the reference's own compiled CUDA extension.

Tolerances: integer results and the ordered-mode sinogram are BIT-EXACT;
gradients: ||g - g64|| <= max(1e-4 ||g64||, 2 ||g_ref32 - g64||) per tensor, where g64 is the
reference algorithm in double (oracle) and g_ref32 its float instantiation.
"""
import numpy as np
import pytest
import torch

import cases
from oracle import oracle

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def run_forward(c, flags=None, want_stats=True):
    from shapefromprojections_b200.function import rasterizer as R
    im, cnt, stats = R.raster_forward(dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"]),
                                      c["H"], c["W"], c["msv"], want_cnt=True, want_stats=want_stats, flags=flags)
    torch.cuda.synchronize()
    return im.cpu().numpy(), cnt.cpu().numpy(), (stats.cpu().numpy() if stats is not None else None)


def mesh_case(geom, mesh, idx=None):
    from oracle import vertex_stage as vs
    verts, faces, labels, mus = mesh
    n = len(geom["ProjectionAngles"]) if "ProjectionAngles" in geom else geom["Vectors"].shape[0]
    idx = np.arange(n) if idx is None else idx
    p6, len3, ff, msv = vs.vertex_stage(geom, verts, faces, idx, np.float32)
    return dict(p6=p6, len3=len3, ff=ff, labels=labels, mus=mus, H=geom["DetectorRowCount"],
                W=geom["DetectorColCount"], msv=msv)


CASES = {
    "soup_small": lambda: cases.soup_case(0, B=3, F=1500, H=96, W=96),
    "soup_ragged": lambda: cases.soup_case(1, B=2, F=777, H=70, W=101, n_mus=4),      # W % 4 != 0, partial tiles
    "soup_many_faces": lambda: cases.soup_case(2, B=1, F=60000, H=64, W=64),          # > LIST_CAP chunks: list segments
    "soup_one_face": lambda: cases.soup_case(3, B=1, F=1, H=16, W=16),
    "soup_big_triangles": lambda: cases.soup_case(4, B=2, F=400, H=300, W=340, max_size=250.0),   # spans far from vertex a
    "soup_slivers": lambda: _slivers(),
    "sphere_parallel": lambda: mesh_case(cases.parallel_geometry(6, 192), cases.sphere_mesh(3)),
    "two_material_cone": lambda: mesh_case(cases.cone_geometry(5, 160), cases.two_material_mesh(2)),
    "torus_axis_aligned": lambda: mesh_case(cases.parallel_geometry(4, 256), cases.torus_mesh(40, 50)),  # knife edges
}


def _slivers():
    """Long thin triangles in all orientations, vertices on and off pixel centres (row-span stress)."""
    c = cases.soup_case(6, B=2, F=1200, H=128, W=128)
    rng = np.random.default_rng(9)
    p = c["p6"].reshape(2, 1200, 3, 2)
    direction = rng.normal(size=(2, 1200, 1, 2)); direction /= np.linalg.norm(direction, axis=-1, keepdims=True)
    length = rng.uniform(5, 90, size=(2, 1200, 1, 1))
    p[:, :, 1:2] = p[:, :, 0:1] + direction * length
    normal = np.concatenate([-direction[..., 1:2], direction[..., 0:1]], axis=-1)
    p[:, :, 2:3] = p[:, :, 0:1] + direction * length * rng.uniform(0, 1, size=(2, 1200, 1, 1)) + normal * rng.uniform(-0.7, 0.7, size=(2, 1200, 1, 1))
    snap = rng.random((2, 1200)) < 0.3
    p[snap] = np.round(p[snap])
    c["p6"] = p.reshape(2, 1200, 6).astype(np.float32)
    return c


@pytest.mark.parametrize("name", list(CASES))
def test_forward_bit_exact_vs_oracle(cuda_lib, name):
    c = CASES[name]()
    ref = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"],
                                    False, fast=True)
    im, cnt, stats = run_forward(c)
    assert np.array_equal(cnt, ref["cnt"]), f"cnt differs at {np.argwhere(cnt != ref['cnt'])[:5]}"
    neq = bits(im) != bits(ref["im"])
    assert not neq.any(), (f"{neq.sum()} pixels differ; first {np.argwhere(neq)[:5]}, "
                           f"max abs {np.abs(im - ref['im']).max()}")
    assert stats[1] == ref["cnt"].sum()                       # fragments == the reference's vf
    # candidates: per-row spans inside the bounding boxes — a superset of the fragments, a subset of the bbox pixels
    assert stats[1] <= stats[2] <= ref["n_candidates"]
    assert stats[0] == ref["n_nonpositive_len"]


@pytest.mark.parametrize("name", ["soup_small", "torus_axis_aligned", "soup_many_faces"])
def test_prefilter_changes_nothing(cuda_lib, name):
    from shapefromprojections_b200 import _lib
    c = CASES[name]()
    a = run_forward(c, flags=0)
    b = run_forward(c, flags=_lib.FLAG_NO_PREFILTER)
    assert np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("name", ["soup_small", "soup_ragged", "two_material_cone", "soup_many_faces"])
def test_fragment_list_exact_vs_oracle(cuda_lib, name):
    """cuda_diff_fp-compatible pass 2: ordered fragment list (face ids exact, weights bit-exact)."""
    from shapefromprojections_b200 import cuda_diff_fp as compat
    c = CASES[name]()
    ref = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"],
                                    True, fast=True)
    B, F = c["p6"].shape[:2]
    H, W = c["H"], c["W"]
    p6, len3, ff, labels, mus = dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"])
    bbox = dev(ref["bbox"])
    im = torch.zeros(B, H, W, 1, device="cuda")
    cnt = torch.zeros(B, H, W, 1, dtype=torch.int32, device="cuda")
    e_f = torch.zeros(0, device="cuda"); e_i = torch.zeros(0, dtype=torch.int32, device="cuda")
    compat.forward(p6, ff, bbox, len3, labels, mus, e_f, e_i, cnt, e_i, im, 1, c["msv"])
    cumsum = torch.cumsum(cnt.reshape(-1), 0, dtype=torch.int32)          # rasterizer.py:63-73
    vf = int(cumsum[-1])
    assert vf == ref["vf"]
    imwei = torch.zeros(3 * vf, device="cuda"); imidx = torch.zeros(vf, dtype=torch.int32, device="cuda")
    first = torch.zeros(B * H * W, dtype=torch.int32, device="cuda"); first[1:] = cumsum[:-1]
    first = first.reshape(B, H, W, 1).contiguous()
    compat.forward(p6, ff, bbox, len3, labels, mus, imwei, imidx, cnt, first, im, 0, c["msv"])
    torch.cuda.synchronize()
    assert np.array_equal(cnt.cpu().numpy(), ref["cnt"])
    assert np.array_equal(imidx.cpu().numpy(), ref["imidx"])
    assert np.array_equal(bits(imwei.cpu().numpy()), bits(ref["imwei"]))


def grad_close(name, g, g64, g32):
    g, g64, g32 = (np.asarray(x, dtype=np.float64).ravel() for x in (g, g64, g32))
    err = np.linalg.norm(g - g64)
    tol = max(1e-4 * np.linalg.norm(g64), 2 * np.linalg.norm(g32 - g64))
    assert err <= tol + 1e-30, f"{name}: |g-g64|={err:.3e} > tol={tol:.3e} (|g64|={np.linalg.norm(g64):.3e})"


def oracle_grads(c, g, dtype):
    cc = {k: (v.astype(dtype) if isinstance(v, np.ndarray) and v.dtype == np.float32 else v) for k, v in c.items()}
    fwd = oracle.rasterizer_forward(cc["p6"], cc["len3"], cc["ff"], cc["labels"], cc["mus"], cc["H"], cc["W"],
                                    cc["msv"], True, fast=True)
    return fwd, oracle.rasterizer_backward(g.astype(dtype), fwd, cc["p6"], cc["len3"], cc["ff"], cc["labels"], cc["mus"])


@pytest.mark.parametrize("name", ["soup_small", "soup_ragged", "sphere_parallel", "two_material_cone", "soup_many_faces",
                                  "soup_big_triangles", "soup_slivers"])
@pytest.mark.parametrize("deterministic", [False, True])
def test_backward_vs_oracle(cuda_lib, name, deterministic):
    from shapefromprojections_b200 import _lib
    from shapefromprojections_b200.function import rasterizer as R
    c = CASES[name]()
    B, H, W = c["p6"].shape[0], c["H"], c["W"]
    g = np.random.default_rng(7).normal(size=(B, H, W, 1)).astype(np.float32)
    fwd32, (dp32, dl32, dm32) = oracle_grads(c, g, np.float32)
    # double referee uses the float coverage decisions: feed it the float-rounded inputs
    fwd64, (dp64, dl64, dm64) = oracle_grads(c, g, np.float64)
    same_cover = np.array_equal(fwd64["cnt"], fwd32["cnt"]) and np.array_equal(
        (fwd64["im"].astype(np.float32) < -1e-5) | (fwd64["im"] > c["msv"]), (fwd32["im"] < -1e-5) | (fwd32["im"] > c["msv"]))
    flags = _lib.FLAG_DETERMINISTIC if deterministic else 0
    p6, len3, ff, labels, mus = dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"])
    im = dev(fwd32["im"])
    dp, dl, dm = R.raster_backward(dev(g), im, p6, len3, ff, labels, mus, c["msv"], flags=flags)
    torch.cuda.synchronize()
    dp, dl, dm = dp.cpu().numpy(), dl.cpu().numpy(), dm.cpu().numpy()
    if same_cover:
        grad_close("grad_p6", dp, dp64, dp32); grad_close("grad_len3", dl, dl64, dl32); grad_close("grad_mus", dm, dm64, dm32)
    else:   # knife-edge pixels decided differently in double: compare with the float restatement only
        for nme, a, b in (("grad_p6", dp, dp32), ("grad_len3", dl, dl32), ("grad_mus", dm, dm32)):
            err = np.linalg.norm(a.astype(np.float64) - b); ref = np.linalg.norm(b.astype(np.float64))
            assert err <= 2e-4 * ref + 1e-30, f"{nme}: {err:.3e} vs {ref:.3e}"
    if deterministic:
        dp2, dl2, dm2 = R.raster_backward(dev(g), im, p6, len3, ff, labels, mus, c["msv"], flags=flags)
        torch.cuda.synchronize()
        assert np.array_equal(bits(dp2.cpu().numpy()), bits(dp)) and np.array_equal(bits(dl2.cpu().numpy()), bits(dl))
        assert np.array_equal(bits(dm2.cpu().numpy()), bits(dm))


def test_autograd_function_surface(cuda_lib):
    """Rasterizer.apply keeps the reference signature and gradient pattern (rasterizer.py:13,107)."""
    from shapefromprojections_b200.function.rasterizer import Rasterizer
    c = CASES["two_material_cone"]()
    p6 = dev(c["p6"]).requires_grad_(True); len3 = dev(c["len3"]).requires_grad_(True)
    ff = dev(c["ff"]).requires_grad_(True); mus = dev(c["mus"]).requires_grad_(True)
    im = Rasterizer.apply(p6, len3, ff, dev(c["labels"]), mus, c["H"], c["W"], c["msv"], True)
    assert im.shape == (c["p6"].shape[0], c["H"], c["W"], 1)
    (im * im).sum().backward()
    assert p6.grad is not None and len3.grad is not None and mus.grad is not None and ff.grad is None
    im2 = Rasterizer.apply(p6, len3, ff, dev(c["labels"]), mus, c["H"], c["W"], c["msv"], False)
    assert torch.equal(im2, im.detach())


def test_errors_are_loud(cuda_lib):
    from shapefromprojections_b200.function import rasterizer as R
    c = CASES["soup_one_face"]()
    with pytest.raises(RuntimeError):
        R.raster_forward(torch.as_tensor(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"]), 16, 16, 1.0)
    with pytest.raises(TypeError):
        R.raster_forward(dev(c["p6"]).double(), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"]), 16, 16, 1.0)
    with pytest.raises(RuntimeError):
        R.raster_forward(dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"]), 40000, 16, 1.0)


def test_hoisted_reciprocal_quotient_is_ieee(cuda_lib):
    """The only arithmetic shortcut in the kernels: k1/d from the face's 1/d + one fma correction must
    equal the reference's IEEE fp64 division bit for bit (2^32 random float pairs, incl. corner cases)."""
    from shapefromprojections_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 2):
        _lib.check(cuda_lib.ctdr_selftest_quotient(1 << 31, seed, bad.data_ptr(), _lib.stream_ptr(bad.device)), "selftest")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_empty_inputs(cuda_lib):
    """No faces / no angles: zero sinogram, zero gradients (the reference's loops simply do not run)."""
    from shapefromprojections_b200.function import rasterizer as R
    lab = torch.zeros(0, 2, dtype=torch.int32, device="cuda"); mus = torch.tensor([0.0, 1.0], device="cuda")
    im, cnt, _ = R.raster_forward(torch.zeros(2, 0, 6, device="cuda"), torch.zeros(2, 0, 3, device="cuda"),
                                  torch.zeros(2, 0, 1, device="cuda"), lab, mus, 8, 9, 1.0, want_cnt=True)
    assert im.shape == (2, 8, 9, 1) and not im.any() and not cnt.any()
    dp, dl, dm = R.raster_backward(torch.ones(2, 8, 9, 1, device="cuda"), im, torch.zeros(2, 0, 6, device="cuda"),
                                   torch.zeros(2, 0, 3, device="cuda"), torch.zeros(2, 0, 1, device="cuda"), lab, mus, 1.0)
    assert dp.shape == (2, 0, 6) and not dm.any()
    # a mesh entirely off the detector: every face is rejected whole (diff_fp_for.cu:105-106)
    c = cases.soup_case(8, B=1, F=50, H=32, W=32)
    c["p6"] = c["p6"] + 1000.0
    im, cnt, _ = run_forward(c)
    assert not im.any() and not cnt.any()


@pytest.mark.parametrize("max_box", [4, 64, 4096])
@pytest.mark.parametrize("name", ["soup_small", "soup_ragged", "soup_slivers", "sphere_parallel", "two_material_cone",
                                  "torus_axis_aligned"])
def test_fine_mesh_variant_direct_walk(cuda_lib, monkeypatch, name, max_box):
    """The face-per-lane walk of the fine-mesh kernel variant (normally chosen when F >= 0.4*H*W) forced on
    ordinary cases: max_box 4 mixes direct and general batches inside one launch, 4096 walks everything
    directly.  Forward stays bit-identical to the oracle (the backward walks row spans and is not affected)."""
    from shapefromprojections_b200.function import rasterizer as R
    c = CASES[name]()
    B, H, W = c["p6"].shape[0], c["H"], c["W"]
    ref = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], H, W, c["msv"], False, fast=True)
    g = dev(np.random.default_rng(3).normal(size=(B, H, W, 1)).astype(np.float32))
    args = (dev(c["p6"]), dev(c["len3"]), dev(c["ff"]), dev(c["labels"]), dev(c["mus"]))
    from shapefromprojections_b200 import _lib
    base = R.raster_backward(g, dev(ref["im"]), *args, c["msv"])
    monkeypatch.setenv("CTDR_DIRECT_MAX_BOX", str(max_box))
    _lib.reload_tuning()                                    # the library reads its CTDR_* variables once, not per launch
    try:
        for want_stats in (True, False):                    # both the counting and the lean instantiation
            im, cnt, stats = R.raster_forward(*args, H, W, c["msv"], want_cnt=True, want_stats=want_stats)
            assert np.array_equal(cnt.cpu().numpy(), ref["cnt"])
            assert np.array_equal(bits(im.cpu().numpy()), bits(ref["im"]))
            if want_stats:
                assert stats[1].item() == ref["cnt"].sum() and stats[0].item() == ref["n_nonpositive_len"]
        got = R.raster_backward(g, dev(ref["im"]), *args, c["msv"])     # the backward has no such variant: unchanged
    finally:
        monkeypatch.delenv("CTDR_DIRECT_MAX_BOX")
        _lib.reload_tuning()
    for a, b in zip(got, base):
        assert (a - b).norm().item() <= 2e-5 * b.norm().item() + 1e-30


def _nonfinite_case():
    """A few faces with NaN / Inf coordinates among ordinary ones: torch.min/max propagate NaN into the
    reference's bounding box (rasterizer.py:24-25), every comparison with NaN is false (so such a face is
    neither rejected nor bounded along that axis) and NaN weights pass the inside test (diff_fp_for.cu:149-151)."""
    c = cases.soup_case(11, B=2, F=300, H=48, W=56, reference_safe=True)
    p = c["p6"]
    p[0, 5, 0] = np.nan          # NaN x of vertex a
    p[0, 17, 3] = np.nan         # NaN y of vertex b
    p[1, 40, 4] = np.inf         # +Inf x of vertex c: bbox edge off the detector -> whole face rejected
    p[1, 41, 1] = -np.inf
    p[1, 99, :] = np.nan         # everything NaN
    c["len3"][0, 7, 1] = np.nan  # NaN length on an ordinary face
    return c


def test_nonfinite_coordinates_follow_the_reference(cuda_lib):
    c = _nonfinite_case()
    ref = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"],
                                    False, fast=True)
    lit = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"],
                                    False, fast=False)                # the literal pixel-parallel restatement
    assert np.array_equal(ref["cnt"], lit["cnt"])
    im, cnt, stats = run_forward(c)
    assert np.array_equal(cnt, ref["cnt"])
    nan_ref, nan_got = np.isnan(ref["im"]), np.isnan(im)
    assert np.array_equal(nan_ref, nan_got) and nan_ref.any()
    assert np.array_equal(bits(im)[~nan_ref], bits(ref["im"])[~nan_ref])
