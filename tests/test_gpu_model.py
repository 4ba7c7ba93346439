"""GPU tests of the callers either side of the hot path: Model.forward and the optimiser iteration."""
import types

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def make_model(geom, subdiv=3, wlap=10.0, wflat=0.01, nmu0=1):
    from ctdr.model.vanilla import Model               # the reference's import path (drop-in package)
    verts, faces, labels, mus = cases.sphere_mesh(subdiv, 0.4)
    labels_v = np.ones(verts.shape[0], dtype=np.int32)
    return Model((verts, faces, labels_v, labels), geom, 2, [0.0, 1.0], nmu0, wlap=wlap, wflat=wflat).cuda()


def target_sinogram(geom, scale=(1.2, 0.9, 1.1)):
    """Config-1 style target: projection of the same sphere scaled anisotropically."""
    m = make_model(geom, wlap=0.0, wflat=0.0)
    with torch.no_grad():
        m.vertices *= torch.tensor(scale, device="cuda")
        phat, _, _, _, _ = m(torch.arange(m.nangles), 0.0, backprop=False)
    return phat.detach()


def test_model_forward_surface(cuda_lib):
    geom = cases.parallel_geometry(8, 96, 2.0 / 96)
    model = make_model(geom)
    phat, mask, edge, lap, flat = model(torch.arange(8), 2.0)
    assert phat.shape == (8, 96, 96) and mask.dtype == torch.bool and mask.all()
    assert edge > 0 and lap >= 0 and 0 <= flat < 0.1
    (phat.sum() + edge + lap + flat).backward()
    assert model.displace.grad.abs().max() > 0 and model.mus.grad is not None
    ph2, _ = model.render()
    assert torch.equal(ph2, phat.detach())
    v, f, lv, lf = model.get_vf_np()
    assert v.shape == (model.vertices.shape[0], 3) and lf.shape == (f.shape[0], 2)
    assert abs(phat.max().item() - 0.8) < 0.02          # chord through the centre of an r = 0.4 sphere


def test_fused_iteration_equals_autograd_iteration(cuda_lib):
    """projection_step (one device sweep, no host sync) gives the loss and gradients of the
    reference's formulation (optimize.py:162-190) evaluated through autograd."""
    from ctdr import optimize
    geom = cases.parallel_geometry(8, 96, 2.0 / 96)
    p = target_sinogram(geom)
    a, b = make_model(geom, wlap=0.0, wflat=0.0), make_model(geom, wlap=0.0, wflat=0.0)
    idx = torch.arange(8)
    phat, mask, _, _, _ = a(idx, 0.0)
    loss = (p - phat)[mask].pow(2).mean()
    loss.backward()
    loss_f, phat_f, _ = optimize.projection_step(b, idx, p)
    assert torch.equal(phat_f, phat.detach())
    assert abs(loss_f.item() - loss.item()) <= 1e-5 * loss.item()
    assert (b.displace.grad - a.displace.grad).norm() <= 2e-5 * a.displace.grad.norm()
    assert (b.mus.grad - a.mus.grad).norm() <= 2e-5 * a.mus.grad.norm() + 1e-9


@pytest.mark.parametrize("fused", [True, False])
def test_optimise_run_reduces_the_data_loss(cuda_lib, tmp_path, fused):
    """Short config-1 style reconstruction (sphere -> anisotropically scaled sphere): the loss drops
    by more than an order of magnitude and log.txt keeps the reference's line format."""
    from ctdr import optimize
    from ctdr.dataset.dataset import SinoDataset
    geom = cases.parallel_geometry(12, 96, 2.0 / 96)
    p = target_sinogram(geom)
    ds = SinoDataset.from_arrays(geom, p.cpu().numpy(), nmaterials=2, eta=0.0)
    model = make_model(geom, wlap=10.0, wflat=0.0)
    model.set_mu0(ds.p.cuda())
    args = types.SimpleNamespace(data="2starA", lr=0.01, b=0, wedge=1.0, wlap=10.0, wflat=0.0, verbose=0,
                                 dresult=str(tmp_path) + "/")
    optimize.run(model, ds, 60, args, fused=fused)
    lines = [l for l in open(tmp_path / "log.txt").read().splitlines() if l.startswith("~ 0")]
    assert len(lines) == 60 and " l2_loss: " in lines[0] and " mus: " in lines[0] and " time: " in lines[0]
    first = float(lines[0].split("l2_loss:")[1].split()[0]); last = float(lines[-1].split("l2_loss:")[1].split()[0])
    assert last < first / 10, (first, last)
    assert model.mus.data[1].item() > 0.9


def test_reference_rasterizer_module_runs_on_new_kernels(cuda_lib):
    """The cuda_diff_fp-compatible module reproduces Rasterizer.forward/backward of the reference flow
    (pass 1, cumsum/firstidx, pass 2, backward) — exercised with this repo's restatement of that flow."""
    import sys
    from shapefromprojections_b200 import cuda_diff_fp as compat
    from oracle import ref_driver
    c = cases.soup_case(5, B=2, F=900, H=64, W=64, reference_safe=True)
    saved = ref_driver._mod
    try:
        ref_driver._mod = compat                      # drive OUR kernels through the reference's host steps
        fwd = ref_driver.forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"])
        g = np.random.default_rng(1).normal(size=(2, 64, 64, 1)).astype(np.float32)
        dp, dl, dm = ref_driver.backward(fwd, g)
    finally:
        ref_driver._mod = saved
    from oracle import oracle
    o = oracle.rasterizer_forward(c["p6"], c["len3"], c["ff"], c["labels"], c["mus"], c["H"], c["W"], c["msv"], True, fast=True)
    assert np.array_equal(fwd["imidx"].cpu().numpy(), o["imidx"]) and np.array_equal(fwd["cnt"].cpu().numpy(), o["cnt"])
    assert np.array_equal(fwd["im"].cpu().numpy().view(np.uint32), o["im"].view(np.uint32))
    od = oracle.rasterizer_backward(g, o, c["p6"], c["len3"], c["ff"], c["labels"], c["mus"])
    for a, b in zip((dp, dl, dm), od):
        assert np.linalg.norm(a.cpu().numpy() - b) <= 2e-4 * np.linalg.norm(b)


def test_cuda_graph_iteration_follows_the_eager_trajectory(cuda_lib, tmp_path):
    """GraphedIteration (whole optimiser iteration replayed from a CUDA graph) == optimize.run eagerly."""
    from ctdr import optimize
    from ctdr.dataset.dataset import SinoDataset
    geom = cases.parallel_geometry(12, 96, 2.0 / 96)
    p = target_sinogram(geom)
    ds = SinoDataset.from_arrays(geom, p.cpu().numpy(), nmaterials=2, eta=0.0)
    args = types.SimpleNamespace(data="2starA", lr=0.01, b=0, wedge=1.0, wlap=10.0, wflat=0.01, verbose=0,
                                 dresult=str(tmp_path) + "/")
    a = make_model(geom, wlap=10.0, wflat=0.01); a.set_mu0(ds.p.cuda())
    optimize.run(a, ds, 25, args, fused=True, log_every=10 ** 9)
    eager = [float(l.split("l2_loss:")[1].split()[0]) for l in open(tmp_path / "log.txt").read().splitlines() if l.startswith("~ 0")]
    b = make_model(geom, wlap=10.0, wflat=0.01); b.set_mu0(ds.p.cuda())
    # the graph constructor runs 3 warm-up iterations on the model (the capture pass itself executes nothing)
    git = optimize.GraphedIteration(b, args, torch.arange(12), ds.p)
    graphed = [git.step().tolist()[0] for _ in range(21)]
    # replay k of the graphed run (0-based) is eager iteration k + 3
    # (float atomics make the two runs drift apart slowly: tight at the start, loose later)
    for k, tol in ((0, 2e-3), (5, 1e-2), (20, 5e-2)):
        assert abs(graphed[k] - eager[k + 3]) <= tol * eager[k + 3] + 1e-7, (k, graphed[k], eager[k + 3])
    assert graphed[-1] < eager[0] / 3


def test_regularisers_at_config3_scale(cuda_lib):
    """SURVEY row f3 at the mesh size of config 3 (V = 50 000, F = 100 000), on the device, values and gradients.
    The reference builds a dense V x V Laplacian (softras_loss.py:31-52: 10 GB here) and scans every face for every edge
    (:64-96: 1.5e10 comparisons); its definitions are restated with scipy.sparse / a per-edge dictionary instead."""
    import time
    import scipy.sparse as sp
    from shapefromprojections_b200 import workloads
    from shapefromprojections_b200.nn.softras_loss import FlattenLoss, LaplacianLoss
    wl = workloads.make_workload("cfg3")
    verts, faces = wl["verts"].astype(np.float64), wl["faces"].astype(np.int64)
    V, F = verts.shape[0], faces.shape[0]
    assert V == 50000 and F == 100000
    rng = np.random.default_rng(0)
    x = verts + 1e-3 * rng.normal(size=verts.shape)

    # Laplacian, softras_loss.py:39-52: -1 on every (undirected) edge, degree on the diagonal, rows / (diag + 1e-10)
    i = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 1], faces[:, 2], faces[:, 2], faces[:, 0]])
    j = np.concatenate([faces[:, 1], faces[:, 0], faces[:, 2], faces[:, 1], faces[:, 0], faces[:, 2]])
    A = sp.coo_matrix((np.ones(i.size), (i, j)), shape=(V, V)).tocsr()
    A.data[:] = 1.0                                             # assignment, not accumulation (:41-46)
    deg = np.asarray(A.sum(1)).ravel()
    L = sp.diags(deg / (deg + 1e-10)) - sp.diags(1.0 / (deg + 1e-10)) @ A
    lap_ref = np.square(L @ x).sum(1)
    glap_ref = 2.0 * (L.T @ (L @ x))                            # d sum(lap) / dx

    t0 = time.time()
    lap = LaplacianLoss(torch.as_tensor(verts), torch.as_tensor(faces)).cuda()
    flat = FlattenLoss(torch.as_tensor(faces)).cuda()
    build_s = time.time() - t0
    xt = torch.as_tensor(x, dtype=torch.float32, device="cuda").requires_grad_(True)
    out = lap(xt)
    out.sum().backward()
    torch.cuda.synchronize()
    assert np.abs(out.detach().cpu().numpy() - lap_ref).max() <= 1e-5 * lap_ref.max() + 1e-9
    g = xt.grad.cpu().numpy()
    assert np.linalg.norm(g - glap_ref) <= 1e-4 * np.linalg.norm(glap_ref)

    # flatten loss, softras_loss.py:64-136: edges = the (v0,v1) and (v1,v2) sides, wings from the first two faces (in
    # face order) that contain the edge
    firsttwo = {}
    for fi, (a, b, c) in enumerate(faces):
        for u, v, w in ((a, b, c), (b, c, a), (c, a, b)):
            firsttwo.setdefault((min(u, v), max(u, v)), []).append(w)
    edges = sorted(set(tuple(sorted(e)) for e in np.concatenate([faces[:, 0:2], faces[:, 1:3]])))
    v0 = np.array([e[0] for e in edges]); v1 = np.array([e[1] for e in edges])
    v2 = np.array([firsttwo[e][0] for e in edges]); v3 = np.array([firsttwo[e][1] for e in edges])
    c10, c20, c30 = x[v1] - x[v0], x[v2] - x[v0], x[v3] - x[v0]
    n0, n1 = np.cross(c10, c20), -np.cross(c10, c30)
    flat_ref = (1.0 - (n0 * n1).sum(1) / (np.linalg.norm(n0, axis=1) * np.linalg.norm(n1, axis=1))).mean()
    xt.grad = None
    fl = flat(xt)
    fl.backward()
    torch.cuda.synchronize()
    assert abs(float(fl) - flat_ref) <= 1e-5 * abs(flat_ref) + 1e-7, (float(fl), flat_ref)
    assert torch.isfinite(xt.grad).all() and float(xt.grad.abs().sum()) > 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        xt.grad = None
        (lap(xt).mean() + flat(xt)).backward()
    e1.record(); torch.cuda.synchronize()
    print(f"regularisers at V={V}, F={F}: tables built in {build_s:.2f} s on the host; Laplacian + flatten forward + backward "
          f"{e0.elapsed_time(e1) / 20:.3f} ms per iteration on the device")
