"""GPU tests of the vertex stage and the fused projector (mesh + geometry in, sinogram out)."""
import glob
import os

import numpy as np
import pytest
import torch

import cases
from oracle import oracle, vertex_stage
from test_gpu_raster import bits, dev
from test_oracle_golden import _geom_from_npz

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

GEOMS = {
    "parallel3d": lambda: cases.parallel_geometry(6, 160),
    "parallel3d_vec": lambda: __import__("shapefromprojections_b200.utils.util_sino", fromlist=["x"]).geom_2vec(cases.parallel_geometry(6, 160)),
    "cone_vec": lambda: cases.cone_geometry(6, 160),
}


def setup(geom, mesh):
    from shapefromprojections_b200.function import rasterizer as R
    verts, faces, labels, mus = mesh
    g = R.GeometryDevice(geom, "cuda")
    return g, dev(verts), dev(faces, torch.int32), dev(labels), dev(mus)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "vertex_stage_*_f32.npz"))))
def test_vertex_forward_vs_reference_modules(cuda_lib, path):
    """Device vertex stage vs what the reference's FP modules hand to Rasterizer.apply (golden, torch CPU)."""
    from shapefromprojections_b200.function import rasterizer as R
    g = np.load(path)
    geom = _geom_from_npz(g)
    gd = R.GeometryDevice(geom, "cuda")
    assert gd.max_sino_val == g["msv"]
    p6, len3, ff = R.vertex_forward(gd, dev(g["verts"], torch.float32), dev(g["faces"], torch.int32), g["idx_angles"])
    torch.cuda.synchronize()
    # float32 op-order noise.  Measured on a B200 (round 2): 1.7e-4 px (cone_vec), 1.5e-5 px
    # (parallel3d), 0 (parallel3d_vec: bit-identical) vs the reference's fp32 modules on the torch CPU device —
    # the same distance the reference's own fp32 run has from its fp64 run (1.8e-4 / 1.6e-5 / 6e-5 px).  Bound = 2x.
    tol_px = {"cone_vec": 4e-4, "parallel3d": 4e-5, "parallel3d_vec": 1e-6}[geom["type"]]
    g64 = np.load(path.replace("_f32", "_f64"))                   # the reference modules in float64: the true values
    e_px, e_len = np.abs(p6.cpu().numpy() - g["p6"]).max(), np.abs(len3.cpu().numpy() - g["len3"]).max()
    print(f"[{geom['type']}] vertex stage vs reference fp32 modules: max |dp| = {e_px:.2e} px, max |dlen| = {e_len:.2e}; "
          f"vs reference fp64: {np.abs(p6.cpu().numpy() - g64['p6']).max():.2e} px (reference fp32 vs its own fp64: "
          f"{np.abs(g['p6'] - g64['p6']).max():.2e} px)")
    assert e_px <= tol_px
    assert e_len <= 2e-6                                          # measured: <= 9.5e-7 (one ulp of len ~ 10)
    big = np.abs(g["ff"]) > 1e-6
    assert np.array_equal((ff.cpu().numpy() < 0)[big], (g["ff"] < 0)[big])


@pytest.mark.parametrize("kind", list(GEOMS))
def test_vertex_backward_is_transposed_jacobian(cuda_lib, kind):
    """<J^T g, d> == d/d eps <g, f(v + eps d)> with f evaluated in float64 by the NumPy restatement."""
    from shapefromprojections_b200.function import rasterizer as R
    geom = GEOMS[kind]()
    verts, faces, labels, mus = cases.two_material_mesh(1)
    gd, v, f32, _, _ = setup(geom, (verts, faces, labels, mus))
    idx = np.arange(6)
    rng = np.random.default_rng(0)
    B, F = 6, faces.shape[0]
    gp6 = rng.normal(size=(B, F, 6)).astype(np.float32)
    gl3 = rng.normal(size=(B, F, 3)).astype(np.float32)
    vreq = v.clone().requires_grad_(True)
    p6, len3, ff = R.VertexStage.apply(vreq, f32, idx, gd)
    ((p6 * dev(gp6)).sum() + (len3 * dev(gl3)).sum()).backward()
    gv = vreq.grad.cpu().numpy().astype(np.float64)

    def phi(vv):
        a, b, _, _ = vertex_stage.vertex_stage(geom, vv, faces, idx, np.float64)
        return (a * gp6).sum() + (b * gl3).sum()
    for k in range(3):
        d = rng.normal(size=verts.shape)
        eps = 1e-6
        fd = (phi(verts.astype(np.float64) + eps * d) - phi(verts.astype(np.float64) - eps * d)) / (2 * eps)
        an = (gv * d).sum()
        assert abs(fd - an) <= 2e-4 * max(abs(fd), 1.0), (kind, fd, an)


@pytest.mark.parametrize("kind", list(GEOMS))
def test_fused_forward_equals_staged_and_oracle(cuda_lib, kind):
    from shapefromprojections_b200.function import rasterizer as R
    geom = GEOMS[kind]()
    mesh = cases.two_material_mesh(2)
    gd, v, f32, labels, mus = setup(geom, mesh)
    idx = torch.arange(6, device="cuda")
    im, cnt, stats = R.project_forward(gd, v, f32, labels, mus, idx, want_cnt=True, want_stats=True)
    p6, len3, ff = R.vertex_forward(gd, v, f32, idx)
    im2, cnt2, _ = R.raster_forward(p6, len3, ff, labels, mus, gd.H, gd.W, gd.max_sino_val, want_cnt=True)
    torch.cuda.synchronize()
    assert torch.equal(cnt, cnt2.squeeze(-1)) and torch.equal(im, im2.squeeze(-1))     # same kernels: bit-exact
    # the staged tensors through the CPU oracle: bit-exact again (identical p6 -> identical coverage)
    o = oracle.rasterizer_forward(p6.cpu().numpy(), len3.cpu().numpy(), ff.cpu().numpy(), mesh[2], mesh[3],
                                  gd.H, gd.W, gd.max_sino_val, False, fast=True)
    assert np.array_equal(cnt.cpu().numpy(), o["cnt"][..., 0])
    assert np.array_equal(bits(im.cpu().numpy()), bits(o["im"][..., 0]))
    # against the NumPy vertex stage (different op chain): only knife-edge pixels may differ
    q6, l3, f1, msv = vertex_stage.vertex_stage(geom, mesh[0], mesh[1], np.arange(6), np.float32)
    o2 = oracle.rasterizer_forward(q6, l3, f1, mesh[2], mesh[3], gd.H, gd.W, msv, False, fast=True)
    diff = cnt.cpu().numpy() != o2["cnt"][..., 0]
    assert diff.mean() < 5e-5, f"{diff.sum()} coverage mismatches"       # measured (round 2): 0 of 153600 for all three kinds
    same = ~diff
    a, b = im.cpu().numpy()[same], o2["im"][..., 0][same]
    depth = max(int(o2["cnt"].max()), 1)
    # the stated forward tolerance with no extra allowance for cone (measured: identical bits)
    tol = 1e-5 * np.abs(b) + 4 * depth * np.spacing(np.float32(l3.max())) * np.abs(mesh[3]).max()
    print(f"[{kind}] fused sinogram vs NumPy-vertex-stage oracle: {int(diff.sum())} knife-edge pixels of {diff.size}, "
          f"max |d im| = {np.abs(a - b).max():.2e}, max (|d im| - 1e-5|im|) = {(np.abs(a - b) - 1e-5 * np.abs(b)).max():.2e}, "
          f"base tol term = {4 * depth * np.spacing(np.float32(l3.max())) * np.abs(mesh[3]).max():.2e}")
    assert (np.abs(a - b) <= tol).all(), np.abs(a - b).max()


@pytest.mark.parametrize("kind", list(GEOMS))
def test_fused_backward_equals_staged_autograd(cuda_lib, kind):
    from shapefromprojections_b200.function import rasterizer as R
    geom = GEOMS[kind]()
    mesh = cases.two_material_mesh(2)
    gd, v, f32, labels, mus = setup(geom, mesh)
    idx = torch.arange(6, device="cuda")
    w = torch.randn(6, gd.H, gd.W, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    v1 = v.clone().requires_grad_(True); m1 = mus.clone().requires_grad_(True)
    im = R.ProjectMesh.apply(v1, m1, f32, labels, idx, gd, True)
    (im * w).sum().backward()
    v2 = v.clone().requires_grad_(True); m2 = mus.clone().requires_grad_(True)
    p6, len3, ff = R.VertexStage.apply(v2, f32, idx, gd)
    im2 = R.Rasterizer.apply(p6, len3, ff, labels, m2, gd.H, gd.W, gd.max_sino_val, True)
    (im2.squeeze(-1) * w).sum().backward()
    torch.cuda.synchronize()
    for a, b in ((v1.grad, v2.grad), (m1.grad, m2.grad)):
        assert (a - b).norm() <= 1e-5 * b.norm() + 1e-12, ((a - b).norm().item(), b.norm().item())
    assert v1.grad.abs().max() > 0


def test_residual_step_matches_autograd_of_masked_l2(cuda_lib):
    """ctdr/optimize.py:162-168: loss = mean over valid of (p - phat)^2, through the fused step."""
    from shapefromprojections_b200.function import rasterizer as R
    from shapefromprojections_b200.nn.fp_vecs import get_valid_mask
    geom = GEOMS["parallel3d"]()
    mesh = cases.two_material_mesh(2)
    gd, v, f32, labels, mus = setup(geom, mesh)
    idx = torch.arange(6, device="cuda")
    target = torch.rand(6, gd.H, gd.W, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    phat, gv, gm, out2 = R.project_residual_step(gd, v, f32, labels, mus, idx, target)
    v1 = v.clone().requires_grad_(True); m1 = mus.clone().requires_grad_(True)
    im = R.ProjectMesh.apply(v1, m1, f32, labels, idx, gd, True)
    mask = get_valid_mask(im, gd.max_sino_val)
    loss = (target - im)[mask].pow(2).mean()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.equal(phat, im.detach())
    n_valid = int(mask.sum())
    assert int(out2[1].item()) == n_valid
    assert abs(out2[0].item() / n_valid - loss.item()) <= 1e-5 * loss.item()
    assert (gv / n_valid - v1.grad).norm() <= 2e-5 * v1.grad.norm()
    assert (gm / n_valid - m1.grad).norm() <= 2e-5 * m1.grad.norm()


@pytest.mark.parametrize("kind", list(GEOMS))
def test_fp_modules_surface(cuda_lib, kind):
    """FP(proj_geom, labels, dtype).forward(verts, faces, idx_angles, mus, backprop) -> (phat, mask_valid)."""
    geom = GEOMS[kind]()
    verts, faces, labels, mus = cases.two_material_mesh(1)
    if kind == "parallel3d":
        from shapefromprojections_b200.nn.fp_opengl import FP
    else:
        from shapefromprojections_b200.nn.fp_vecs import FP
    lab = torch.as_tensor(labels)
    fp = FP(geom, lab, torch.float32)
    fp_staged = FP(geom, lab, torch.float32, fused=False)
    v = dev(verts).requires_grad_(True); m = dev(mus).requires_grad_(True)
    faces64 = dev(faces)                                 # int64, as Model registers it (vanilla.py:47)
    idx_cpu = torch.arange(6)                            # the reference passes a CPU LongTensor (optimize.py:125)
    phat, mask = fp.forward(v, faces64, idx_cpu, m)
    phat2, mask2 = fp_staged.forward(v, faces64, idx_cpu, m)
    assert phat.shape == (6, fp.height, fp.width) and mask.dtype == torch.bool
    assert torch.equal(phat, phat2) and torch.equal(mask, mask2)
    assert fp_staged.proj.shape == (6, faces.shape[0], 6) and fp_staged.len.shape == (6, faces.shape[0], 3)
    phat.sum().backward()
    assert v.grad is not None and m.grad is not None
    one, _ = fp.forward(v, faces64, torch.tensor([2]), m, backprop=False)
    assert one.shape == (fp.height, fp.width)            # phat.squeeze() drops B == 1 (fp_vecs.py:203)
    assert torch.equal(one, phat[2].detach())
    with pytest.raises(TypeError):
        FP(geom, lab, torch.float16)                         # float and double only, like AT_DISPATCH_FLOATING_TYPES (diff_fp_for.cu:251)
    fp64 = FP(geom, lab, torch.float64)                      # the float64 instantiation: staged path in double (tests/test_gpu_f64.py)
    assert fp64.fused is False and fp64.dtype == torch.float64
    with pytest.raises(TypeError):
        fp64.forward(v, faces64, idx_cpu, m)                 # float32 operands into a float64 projector


@pytest.mark.parametrize("kind", list(GEOMS))
def test_deterministic_fused_step_is_bit_reproducible_and_agrees(cuda_lib, kind):
    """CTDR_FLAG_DETERMINISTIC through the fused residual step: single-writer kernels and fixed
    summation orders -> identical bits run to run; same gradients as the default path to fp32 noise."""
    from shapefromprojections_b200 import _lib
    from shapefromprojections_b200.function import rasterizer as R
    geom = GEOMS[kind]()
    mesh = cases.two_material_mesh(2)
    gd, v, f32, labels, mus = setup(geom, mesh)
    idx = torch.arange(6, device="cuda")
    target = torch.rand(6, gd.H, gd.W, device="cuda", generator=torch.Generator("cuda").manual_seed(4))
    runs = [R.project_residual_step(gd, v, f32, labels, mus, idx, target, flags=_lib.FLAG_DETERMINISTIC) for _ in range(2)]
    fast = R.project_residual_step(gd, v, f32, labels, mus, idx, target, flags=0)
    torch.cuda.synchronize()
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b)
    phat, gv, gm, out2 = runs[0]
    assert torch.equal(phat, fast[0]) and out2[1].item() == fast[3][1].item()
    assert abs(out2[0].item() - fast[3][0].item()) <= 1e-9 * abs(fast[3][0].item())
    assert (gv - fast[1]).norm() <= 2e-5 * fast[1].norm() and (gm - fast[2]).norm() <= 1e-3 * fast[2].norm()   # grad_mus: a cancelling sum of ~1e5 terms of size ~10 (random target)
    # and through ctdr_project_backward with an explicit upstream gradient
    w = torch.randn(6, gd.H, gd.W, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    d1 = R.project_backward(gd, v, f32, labels, mus, idx, w, phat, flags=_lib.FLAG_DETERMINISTIC)
    d2 = R.project_backward(gd, v, f32, labels, mus, idx, w, phat, flags=_lib.FLAG_DETERMINISTIC)
    ref = R.project_backward(gd, v, f32, labels, mus, idx, w, phat, flags=0)
    torch.cuda.synchronize()
    assert torch.equal(d1[0], d2[0]) and torch.equal(d1[1], d2[1])
    assert (d1[0] - ref[0]).norm() <= 2e-5 * ref[0].norm() and (d1[1] - ref[1]).norm() <= 1e-3 * ref[1].norm()


def test_size_independent_properties_on_a_larger_case(cuda_lib):
    """Properties that need no oracle (usable at any size): coverage is invariant under a permutation
    of the faces (only the summation order changes), the projector is linear in the attenuations,
    and d(sum im)/d mus equals the per-label partial sinogram sums."""
    from shapefromprojections_b200 import workloads
    from shapefromprojections_b200.function import rasterizer as R
    wl = workloads.make_workload("cfg3_small")
    gd = R.GeometryDevice(wl["geom"], "cuda")
    v = dev(wl["verts"]); f32 = dev(wl["faces"], torch.int32); lab = dev(wl["labels"])
    idx = torch.arange(0, 48, 4, device="cuda")
    mu_a, mu_b = torch.tensor([0.0, 1.0], device="cuda"), torch.tensor([0.0, 0.37], device="cuda")
    im_a, cnt_a, st = R.project_forward(gd, v, f32, lab, mu_a, idx, want_cnt=True, want_stats=True)
    perm = torch.randperm(f32.shape[0], generator=torch.Generator().manual_seed(0)).cuda()
    im_p, cnt_p, _ = R.project_forward(gd, v, f32[perm].contiguous(), lab[perm].contiguous(), mu_a, idx, want_cnt=True)
    assert torch.equal(cnt_a, cnt_p)                                   # integer coverage: exact
    assert int(cnt_a.sum()) == int(st[1])                              # fragments == sum of counts (the reference's vf)
    assert (im_a - im_p).abs().max() <= 1e-5 * im_a.abs().max() + 4 * 4 * 1e-6
    im_b, _, _ = R.project_forward(gd, v, f32, lab, mu_b, idx)
    im_ab, _, _ = R.project_forward(gd, v, f32, lab, mu_a + mu_b, idx)
    assert (im_ab - (im_a + im_b)).abs().max() <= 2e-5 * im_ab.abs().max()
    ones = torch.ones_like(im_a)
    _, gmus = R.project_backward(gd, v, f32, lab, mu_a, idx, ones, im_a)
    assert abs(gmus[1].item() - im_a.double().sum().item()) <= 1e-4 * im_a.double().sum().item()   # im is linear in mu_1


def test_full_size_config3_properties(cuda_lib):
    """BASELINE.json config 3 in full (cone 720 x 1024 x 1024, torus of 100k triangles, 7.5e8 pixels) —
    far beyond what the oracle can check, so through properties that hold at any size:
    angles are independent (a strided shard reproduces its slices of the full run bit for bit — what the
    multi-GPU sharding relies on), coverage is invariant under a face permutation, the projector is linear
    in the attenuations, the fused residual step reports exactly the masked-L2 of its own sinogram, and
    gradients add up over angle shards."""
    from shapefromprojections_b200 import workloads
    from shapefromprojections_b200.function import rasterizer as R
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs ~60 GB of device memory")
    wl = workloads.make_workload("cfg3")
    gd = R.GeometryDevice(wl["geom"], "cuda")
    v = dev(wl["verts"]); f32 = dev(wl["faces"], torch.int32); lab = dev(wl["labels"])
    B = wl["nangles"]
    idx = torch.arange(B, device="cuda")
    mu_a, mu_b = torch.tensor([0.0, 1.0], device="cuda"), torch.tensor([0.0, 0.37], device="cuda")
    im_a, cnt_a, st = R.project_forward(gd, v, f32, lab, mu_a, idx, want_cnt=True, want_stats=True)
    assert int(cnt_a.sum(dtype=torch.int64)) == int(st[1]) and int(st[0]) == 0          # vf; no non-positive length
    # chords of the torus: at most 2*sqrt((R+r)^2 - (R-r)^2) = 1.358 (cone magnification adds a little); the
    # few knife-edge pixels where only one of the two surfaces is counted are what get_valid_mask removes
    odd = (im_a < -1e-4) | (im_a > 1.5)
    assert int(odd.sum()) <= 1e-5 * im_a.numel()
    del odd
    # angle independence (strided shard of 8 ranks, rank 3)
    shard = idx[3::8].contiguous()
    im_s, cnt_s, _ = R.project_forward(gd, v, f32, lab, mu_a, shard, want_cnt=True)
    assert torch.equal(im_s, im_a[3::8]) and torch.equal(cnt_s, cnt_a[3::8])
    del im_s, cnt_s
    # face permutation: integer coverage exact, sinogram within the reordering tolerance
    perm = torch.randperm(f32.shape[0], generator=torch.Generator().manual_seed(0)).cuda()
    im_p, cnt_p, _ = R.project_forward(gd, v, f32[perm].contiguous(), lab[perm].contiguous(), mu_a, idx, want_cnt=True)
    assert torch.equal(cnt_a, cnt_p)
    assert float((im_a - im_p).abs().max()) <= 1e-5 * float(im_a.abs().max()) + 4 * 4 * 1e-6
    del im_p, cnt_p, cnt_a
    # linearity in mus
    im_b, _, _ = R.project_forward(gd, v, f32, lab, mu_b, idx)
    im_ab, _, _ = R.project_forward(gd, v, f32, lab, mu_a + mu_b, idx)
    assert float((im_ab - (im_a + im_b)).abs().max()) <= 2e-5 * float(im_ab.abs().max())
    del im_b, im_ab
    # fused residual step: loss / count are those of its own sinogram; gradients add over shards
    vt = v * dev(workloads.TARGET_SCALE)
    target, _, _ = R.project_forward(gd, vt, f32, lab, mu_a, idx)
    phat, gv, gm, out2 = R.project_residual_step(gd, v, f32, lab, mu_a, idx, target)
    assert torch.equal(phat, im_a)
    mask = (phat > -1e-6) & (phat <= gd.max_sino_val - 1e-6)
    assert int(out2[1].item()) == int(mask.sum())
    ref_loss = ((target - phat).double() ** 2)[mask].sum().item()
    assert abs(out2[0].item() - ref_loss) <= 1e-9 * ref_loss
    del mask, phat
    gv_sum, gm_sum, loss_sum = torch.zeros_like(gv), torch.zeros_like(gm), 0.0
    for r in range(2):
        sh = idx[r::2].contiguous()
        _, gvr, gmr, o2 = R.project_residual_step(gd, v, f32, lab, mu_a, sh, target[r::2].contiguous())
        gv_sum += gvr; gm_sum += gmr; loss_sum += o2[0].item()
    assert abs(loss_sum - out2[0].item()) <= 1e-9 * out2[0].item()
    assert float((gv_sum - gv).norm()) <= 1e-4 * float(gv.norm()) and float((gm_sum - gm).norm()) <= 1e-4 * float(gm.norm())
    assert float(gv.norm()) > 0


def nasty_mesh():
    """A sphere plus what the recorded coverage has to survive: zero-area faces (repeated and collinear corners), long
    slivers a fraction of a pixel thick (their rows have no usable span bounds, most candidates of a row fail), faces
    that span many tiles, and a fan of needle faces around one vertex."""
    verts, faces, labels, mus = cases.sphere_mesh(2, 0.45)
    rng = np.random.default_rng(11)
    V0 = verts.shape[0]
    extra_v, extra_f = [], []

    def add(tri):
        base = V0 + len(extra_v)
        extra_v.extend(tri)
        extra_f.append([base, base + 1, base + 2])

    for _ in range(30):                                   # slivers: 0.6 long, 1e-4 .. 3e-3 thick, any orientation
        p = rng.uniform(-0.4, 0.4, 3); d = rng.normal(size=3); d /= np.linalg.norm(d)
        n = np.cross(d, rng.normal(size=3)); n /= np.linalg.norm(n)
        add([p, p + 0.6 * d, p + 0.6 * d + rng.uniform(1e-4, 3e-3) * n])
    for _ in range(10):                                   # collinear corners: exactly zero area in world space
        p = rng.uniform(-0.4, 0.4, 3); d = rng.normal(size=3) * 0.2
        add([p, p + d, p + 2 * d])
    for _ in range(6):                                    # big faces over many tiles
        add(list(rng.uniform(-0.6, 0.6, (3, 3))))
    c = rng.uniform(-0.2, 0.2, 3)                         # needle fan
    for k in range(24):
        a0, a1 = 2 * np.pi * k / 24, 2 * np.pi * (k + 0.02) / 24
        add([c, c + 0.5 * np.array([np.cos(a0), np.sin(a0), 0.3]), c + 0.5 * np.array([np.cos(a1), np.sin(a1), 0.3])])
    nv = np.concatenate([verts, np.asarray(extra_v, np.float32)]).astype(np.float32)
    nf = np.concatenate([faces, np.asarray(extra_f, faces.dtype)])
    for i in range(0, 40, 2):                             # repeated corners on existing vertices
        nf = np.concatenate([nf, np.asarray([[i, i, i + 1], [i, i + 1, i + 1]], faces.dtype)])
    nl = np.tile(np.array([1, 0], labels.dtype), (nf.shape[0], 1))
    return nv, nf, nl, mus


LINK_CASES = {
    # degenerate, sliver and oversized faces: rows re-derived from the hit masks, faces listed for k_step_facelist,
    # zero-area faces in the replay epilogue
    "nasty_parallel": lambda: (cases.parallel_geometry(6, 160), nasty_mesh()),
    "nasty_cone": lambda: (cases.cone_geometry(6, 160), nasty_mesh()),
    # ordinary case: everything is recorded and replayed
    "two_material": lambda: (cases.cone_geometry(5, 160), cases.two_material_mesh(2)),
    # 100k faces on a 128 x 128 detector: region lists take several segments, chains continue across them
    "many_faces_multi_segment": lambda: (cases.parallel_geometry(3, 128), cases.torus_mesh(200, 250)),
    # coarse sphere filling a 512 x 512 detector: more non-empty tiles than record space -> regions fall back
    # to the ordinary residual/backward kernel
    "record_space_overflow": lambda: (cases.parallel_geometry(3, 512), cases.sphere_mesh(2, 0.97)),
}


@pytest.mark.parametrize("name", list(LINK_CASES))
def test_linked_sweeps_equal_unlinked(cuda_lib, monkeypatch, name):
    """Residual step with batch records (forward sweep records, residual/backward sweep replays) vs the same
    step with both sweeps rebuilding their lists and queues: identical sinogram and loss, gradients equal up
    to the order of the atomics."""
    from shapefromprojections_b200.function import rasterizer as R
    geom, mesh = LINK_CASES[name]()
    gd, v, f32, labels, mus = setup(geom, mesh)
    B = geom["Vectors"].shape[0] if "Vectors" in geom else len(geom["ProjectionAngles"])
    idx = torch.arange(B, device="cuda")
    target = torch.rand(B, gd.H, gd.W, device="cuda", generator=torch.Generator("cuda").manual_seed(4))
    from shapefromprojections_b200 import _lib
    monkeypatch.setenv("CTDR_NO_BATCH_RECORDS", "1")
    _lib.reload_tuning()                                   # the library reads its CTDR_* variables once, not per launch
    try:
        ph0, gv0, gm0, o0 = R.project_residual_step(gd, v, f32, labels, mus, idx, target)
    finally:
        monkeypatch.delenv("CTDR_NO_BATCH_RECORDS")
        _lib.reload_tuning()
    ph1, gv1, gm1, o1 = R.project_residual_step(gd, v, f32, labels, mus, idx, target)
    assert torch.equal(ph0, ph1)
    if name == "many_faces_multi_segment":                 # the case does what it is meant to exercise
        _, _, st = R.project_forward(gd, v, f32, labels, mus, idx, want_stats=True)
        assert st[3].item() > 0
    assert o0[1].item() == o1[1].item() and abs(o0[0].item() - o1[0].item()) <= 1e-12 * abs(o0[0].item())
    assert torch.isfinite(gv1).all() and torch.isfinite(gm1).all()
    rel_v = (gv1 - gv0).norm().item() / gv0.norm().item()
    rel_m = (gm1 - gm0).norm().item() / gm0.norm().item()
    print(f"[{name}] linked vs unlinked residual step: grad_verts rel {rel_v:.2e}, grad_mus rel {rel_m:.2e}")
    assert rel_v <= 2e-5 and gv0.norm().item() > 0
    assert rel_m <= 2e-5


@pytest.mark.parametrize("kind", ["cone_vec", "parallel3d_vec"])
def test_facing_sign_of_edge_on_and_degenerate_faces(cuda_lib, kind):
    """The fused vertex stage decides the facing SIGN from R.n without the reference's divisions when that is safe and
    falls back to the reference's own sequence otherwise (csrc/ctdr_vertex.cuh, k_face_assemble<false>).  Faces whose
    plane contains the source of some angle (cone) / the ray direction (parallel) sit exactly in the fallback's
    domain, as do zero-area faces: the fused sinogram and counts must equal the staged path's (exact ``ff``) bit for bit."""
    from shapefromprojections_b200.function import rasterizer as R
    geom = GEOMS[kind]()
    verts, faces, labels, mus = cases.two_material_mesh(2)
    vecs = np.asarray(geom["Vectors"], dtype=np.float64)
    rng = np.random.default_rng(7)
    extra_v, extra_f = [], []
    base = verts.shape[0]
    for k in range(240):
        S = vecs[k % vecs.shape[0], 0:3]                          # source position (cone) / ray direction (parallel)
        p = rng.uniform(-0.3, 0.3, 3)
        a = rng.normal(size=3) * 0.05
        along = (p - S) if kind == "cone_vec" else S / np.linalg.norm(S)
        q = p + 0.04 * along / np.linalg.norm(along) * (1 if kind == "cone_vec" else 1)   # p -> q runs along the ray: the plane contains it
        if k % 8 == 7:
            tri = [p, p + a, p + 2 * a]                             # three collinear corners
        elif k % 8 == 6:
            tri = [p, p, p + a]                                     # two identical corners
        else:
            tri = [p, p + a, q]
        extra_v += tri
        extra_f.append([base + 3 * k, base + 3 * k + 1, base + 3 * k + 2])
    verts2 = np.concatenate([verts, np.asarray(extra_v)]).astype(np.float32)
    faces2 = np.concatenate([faces, np.asarray(extra_f)]).astype(np.int64)
    labels2 = np.concatenate([labels, np.tile([1, 0], (len(extra_f), 1))]).astype(np.int32)
    gd, v, f32, lab, m = setup(geom, (verts2, faces2, labels2, mus))
    idx = torch.arange(6, device="cuda")
    p6, len3, ff = R.vertex_forward(gd, v, f32, idx)
    im_ref, cnt_ref, _ = R.raster_forward(p6, len3, ff, lab, m, gd.H, gd.W, gd.max_sino_val, want_cnt=True)
    near_zero = int((ff[:, faces.shape[0]:].abs() < 1e-7 * ff.abs().max()).sum())
    assert near_zero >= 40, near_zero                               # the construction really produces edge-on faces
    for switch in ("1", "0"):
        os.environ["CTDR_FACE_NORMALS"] = switch
        try:
            from shapefromprojections_b200 import _lib
            _lib.reload_tuning()
            im, cnt, _ = R.project_forward(gd, v, f32, lab, m, idx, want_cnt=True)
            torch.cuda.synchronize()
        finally:
            del os.environ["CTDR_FACE_NORMALS"]
            _lib.reload_tuning()
        assert torch.equal(cnt, cnt_ref.squeeze(-1)), switch
        assert torch.equal(im, im_ref.squeeze(-1)), switch


@pytest.mark.parametrize("name", list(LINK_CASES))
@pytest.mark.parametrize("lanes", [2, 3])
def test_lanes_of_the_residual_step_agree_with_one_lane(cuda_lib, monkeypatch, name, lanes):
    """The residual step cut into lanes (angle parts on several streams, each in its own part of the workspace) vs the
    same step on the caller's stream only: identical sinogram, valid count and (to float64 reassociation) loss, gradients
    equal up to the order of the atomics.  CTDR_LANES is given explicitly, so the lanes run even at these small sizes
    (by default calls below 2^25 pixels stay on one stream)."""
    from shapefromprojections_b200 import _lib
    from shapefromprojections_b200.function import rasterizer as R
    geom, mesh = LINK_CASES[name]()
    gd, v, f32, labels, mus = setup(geom, mesh)
    B = geom["Vectors"].shape[0] if "Vectors" in geom else len(geom["ProjectionAngles"])
    idx = torch.arange(B, device="cuda")
    target = torch.rand(B, gd.H, gd.W, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    out = {}
    for nl in (1, lanes):
        monkeypatch.setenv("CTDR_LANES", str(nl))
        _lib.reload_tuning()
        try:
            out[nl] = R.project_residual_step(gd, v, f32, labels, mus, idx, target)
            torch.cuda.synchronize()
        finally:
            monkeypatch.delenv("CTDR_LANES")
            _lib.reload_tuning()
    (ph0, gv0, gm0, o0), (ph1, gv1, gm1, o1) = out[1], out[lanes]
    assert torch.equal(ph0, ph1)
    assert o0[1].item() == o1[1].item() and abs(o0[0].item() - o1[0].item()) <= 1e-12 * abs(o0[0].item())
    rel_v = (gv1 - gv0).norm().item() / gv0.norm().item()
    rel_m = (gm1 - gm0).norm().item() / gm0.norm().item()
    assert rel_v <= 2e-5 and rel_m <= 2e-5 and gv0.norm().item() > 0, (rel_v, rel_m)
