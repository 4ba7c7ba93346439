"""CPU tests of the host side: C-ABI surface, mesh/geometry I/O, regularisers, angle sharding (gloo)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shared_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports exactly what include/ctdr_b200.h declares."""
    from shapefromprojections_b200 import _lib
    from shapefromprojections_b200.csrc import build
    build.build()
    header = open(os.path.join(ROOT, "include", "ctdr_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ctdr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    # the ctypes signatures carry as many arguments as the header declares for each entry point
    for name, params in re.findall(r"\b(ctdr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        n = 0 if params.strip() in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
    handle.ctdr_abi_version.restype = ctypes.c_int
    assert handle.ctdr_abi_version() == _lib.ABI_VERSION == 5
    # argument validation happens before any CUDA call: usable without a device
    _lib.lib()
    rc = _lib.lib().ctdr_raster_forward(None, None, None, None, None, 1, 1, 1, 8, 8, 1.0, None, None, None, 0, None, 0, None)
    assert rc == -1 and b"null" in _lib.lib().ctdr_last_error()
    rc = _lib.lib().ctdr_raster_forward(None, None, None, None, None, 1, 0, 1, 8, 8, 1.0, None, None, None, 0, None, 0, None)
    assert rc == -2
    assert _lib.lib().ctdr_raster_workspace_bytes(2, 100, 64, 64) >= 2 * 4 * 8


def test_product_has_no_cpu_path():
    from shapefromprojections_b200.function import rasterizer as R
    t = torch.zeros(1, 1, 6)
    with pytest.raises(RuntimeError, match="CUDA"):
        R.raster_forward(t, torch.zeros(1, 1, 3), torch.zeros(1, 1, 1), torch.zeros(1, 2, dtype=torch.int32),
                         torch.zeros(2), 8, 8, 1.0)
    src = "".join(open(os.path.join(dp, f)).read() for dp, _, fs in os.walk(os.path.join(ROOT, "shapefromprojections_b200"))
                  for f in fs if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_regularisers_match_reference_values():
    from shapefromprojections_b200.nn import softras_loss
    from shapefromprojections_b200.utils import util_fun
    g = np.load(os.path.join(HERE, "golden", "regularisers_icosphere2.npz"))   # produced with the reference modules
    v, f = torch.tensor(g["verts"]), torch.tensor(g["faces"])
    assert np.allclose(softras_loss.LaplacianLoss(v, f)(v).numpy(), g["lap"], rtol=1e-5, atol=1e-9)
    assert abs(softras_loss.FlattenLoss(f)(v).item() - g["flat"]) < 1e-6
    assert abs(util_fun.compute_edge_length(v, f).item() - g["edge"]) < 1e-9


def test_obj_template_round_trip_and_init_meshes():
    from shapefromprojections_b200.dataset import init_mesh
    from shapefromprojections_b200.utils import util_mesh
    with tempfile.TemporaryDirectory() as d:
        init_mesh.save_init_mesh(d + "/a.obj", "2starA", 2, 2.0, 4)
        v, f, lv, lf = util_mesh.load_template(d + "/a.obj")
        assert v.shape == (2562, 3) and f.shape == (5120, 3)                    # ctdr_toy_example.ipynb cell 4
        assert (lf == [1, 0]).all() and f.min() == 0
        init_mesh.save_init_mesh(d + "/c.obj", "4x01C", 4, 2.0, 1)
        v, f, lv, lf = util_mesh.load_template(d + "/c.obj")
        assert sorted(set(map(tuple, lf.tolist()))) == [(1, 0), (2, 1), (3, 2)]
        util_mesh.save_mesh(d + "/c2.obj", v, f, lv, lf)
        v2, f2, lv2, lf2 = util_mesh.load_template(d + "/c2.obj")
        assert np.allclose(v, v2) and np.array_equal(f, f2) and np.array_equal(lf, lf2)
        init_mesh.save_init_mesh(d + "/t.obj", "2kitten02A", 2, 2.0, 3)
        assert util_mesh.load_template(d + "/t.obj")[1].shape[0] == 2 * 30 * 30  # torus: 2*sides*rings faces


def test_dataset_noise_and_toml(tmp_path):
    from shapefromprojections_b200.dataset.dataset import SinoDataset
    from shapefromprojections_b200.utils import util_sino
    (tmp_path / "proj_geom.toml").write_text(
        'type = "parallel3d"\nProjectionAngles = [0.0, 0.5, 1.0]\nDetectorRowCount = 4\nDetectorColCount = 5\n'
        "DetectorSpacingX = 0.25\nDetectorSpacingY = 0.25\n")
    p = np.abs(np.random.default_rng(0).normal(size=(3, 4, 5)))
    np.save(tmp_path / "sinogram.npy", p)
    ds = SinoDataset(str(tmp_path), 2, eta=0.3)
    assert ds.proj_geom["ProjectionAngles"].dtype == np.float32 and ds.nangles == 3 and len(ds) == 3
    np.random.seed(1)                                                            # dataset.py:121-130
    e = np.random.normal(size=p.shape)
    want = np.clip(p + 0.3 * e * np.sqrt((p ** 2).sum()) / np.sqrt((e ** 2).sum()), 0, None)
    assert np.allclose(ds.p.numpy(), want.astype(np.float32))
    idx, row = ds[1]
    assert idx == 1 and row.shape == (4, 5)
    vec = util_sino.read_proj_info(str(tmp_path / "proj_geom.toml"), convert_vec=True)
    assert vec["type"] == "parallel3d_vec" and vec["Vectors"].shape == (3, 12)


def test_angle_sharding_is_a_partition():
    from shapefromprojections_b200 import distributed
    for n, w in ((720, 8), (30, 4), (7, 3), (5, 8)):
        blocks = [distributed.shard_angles(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
        strided = [distributed.strided_angles(n, r, w).tolist() for r in range(w)]
        assert sorted(a for part in strided for a in part) == list(range(n))


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from shapefromprojections_b200 import distributed
rank, world, _ = distributed.init_from_env("gloo")
assert world == 2
V, n = 5, 2
xch = distributed.GradientExchange(V, n, "cpu")
b0, b1 = distributed.shard_angles(7, rank, world)
# each rank contributes a gradient that depends only on its own angles
for a in range(b0, b1):
    xch.grad_verts += float(a + 1)
    xch.grad_mus += 0.5 * (a + 1)
    xch.scalars += torch.tensor([2.0 * (a + 1), 10.0], dtype=torch.float64)
xch.all_reduce()
tot = sum(range(1, 8))
assert torch.allclose(xch.grad_verts, torch.full((V, 3), float(tot)))
assert torch.allclose(xch.grad_mus, torch.full((n,), 0.5 * tot))
assert xch.scalars.tolist() == [2.0 * tot, 70.0]
gv, gm, loss = xch.mean_gradients()
assert abs(loss.item() - 2.0 * tot / 70.0) < 1e-12 and torch.allclose(gv, torch.full((V, 3), tot / 70.0))
dist.barrier(); dist.destroy_process_group()
sys.stdout.write(f"rank {{rank}} ok\n"); sys.stdout.flush()
"""


def test_gradient_exchange_two_ranks_gloo(tmp_path):
    """World size 2 on CPU (gloo): sharded sums all-reduce to the single-process result."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout


def test_bench_reference_arm_runs_on_cpu():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1",
                          "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:]
    import json
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_normal_equation_layout_helper():
    """The multi-output pass returns the upper triangle of the Gram matrix and the right-hand side in one
    flat float64 buffer (include/ctdr_b200.h); the binding mirrors it into a symmetric matrix."""
    from shapefromprojections_b200.function.rasterizer import _gram_matrices
    K = 3
    flat = torch.zeros(K * K + K, dtype=torch.float64)
    full = torch.tensor([[4.0, 1.0, 2.0], [1.0, 5.0, 3.0], [2.0, 3.0, 6.0]], dtype=torch.float64)
    for i in range(K):
        for j in range(i, K):
            flat[i * K + j] = full[i, j]           # the kernel leaves the lower triangle at zero
    flat[K * K:] = torch.tensor([7.0, 8.0, 9.0], dtype=torch.float64)
    A, rhs = _gram_matrices(flat, K)
    assert torch.equal(A, full) and torch.equal(rhs, flat[K * K:])
