"""Generate tests/golden/pipeline_*.npz: END-TO-END gradients of the reference's projector modules.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_pipeline_golden.py

The reference's unmodified ``ctdr/nn/fp_opengl.py`` / ``ctdr/nn/fp_vecs.py`` run on the torch CPU device
(``.cuda()`` stubbed, as in make_vertex_golden.py) with their autograd intact; the one piece that cannot run
here — the CUDA-only native module behind ``Rasterizer.apply`` (ctdr/function/rasterizer.py:13-107) — is
replaced by an autograd Function around the C oracle (oracle/ctdr_oracle.c), which is pinned BIT-EXACTLY to the
reference kernels' outputs on a B200 (tests/golden/ref_raster_*.npz, tests/test_oracle_golden.py).  So what is
stored is what ``loss.backward()`` of the reference gives for

  * ``loss_w  = sum(w * phat)``                       (seeded N(0,1) weights ``w``), and
  * ``loss_l2 = mean over mask_valid of (p - phat)^2`` (ctdr/optimize.py:168, seeded target ``p``),

namely ``verts.grad`` and ``mus.grad``, in float32 and float64, for the three geometry kinds.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != REPO]

sys.modules["cuda_diff_fp"] = types.ModuleType("cuda_diff_fp")
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self

from ctdr.nn import fp_opengl as ref_opengl                          # noqa: E402  (reference code)
from ctdr.nn import fp_vecs as ref_vecs                              # noqa: E402
from ctdr.utils import util_sino as ref_sino                         # noqa: E402

assert ref_vecs.__file__.startswith(REF) and ref_opengl.__file__.startswith(REF)
sys.path.insert(1, REPO)
sys.path.insert(1, os.path.join(REPO, "tests"))
import cases                                                         # noqa: E402
from oracle import oracle                                            # noqa: E402

SEED_W, SEED_P = 31, 32


def weights(B, H, W):
    return np.random.default_rng(SEED_W).normal(size=(B, H, W)).astype(np.float32)


def target(B, H, W):
    return (np.random.default_rng(SEED_P).random(size=(B, H, W)) * 1.3).astype(np.float32)


class OracleRasterizer(torch.autograd.Function):
    """Rasterizer.apply (rasterizer.py:13-107) on the CPU oracle: same arguments, same gradient pattern."""

    @staticmethod
    def forward(ctx, p6, len3, ff, labels, mus, H, W, msv, backprop):
        a = lambda t: np.ascontiguousarray(t.detach().numpy())
        fwd = oracle.rasterizer_forward(a(p6), a(len3), a(ff), a(labels), a(mus), H, W, msv, True, fast=True)
        ctx.fwd, ctx.inputs = fwd, (a(p6), a(len3), a(ff), a(labels), a(mus))
        return torch.from_numpy(fwd["im"].copy())

    @staticmethod
    def backward(ctx, g):
        p6, len3, ff, labels, mus = ctx.inputs
        dldp, dldf, dldmu = oracle.rasterizer_backward(np.ascontiguousarray(g.numpy()), ctx.fwd, p6, len3, ff, labels, mus)
        return (torch.from_numpy(dldp), torch.from_numpy(dldf), None, None, torch.from_numpy(dldmu),
                None, None, None, None)


def run(fp_module, proj_geom, mesh, idx, dtype):
    verts, faces, labels, mus = mesh
    fp = fp_module.FP(proj_geom, torch.tensor(labels, dtype=torch.int32), dtype)
    fp.rasterize = OracleRasterizer.apply
    out = {}
    B, H, W = len(idx), proj_geom["DetectorRowCount"], proj_geom["DetectorColCount"]
    for tag in ("w", "l2"):
        v = torch.tensor(verts, dtype=dtype, requires_grad=True)
        m = torch.tensor(mus, dtype=dtype, requires_grad=True)
        phat, mask = fp.forward(v, torch.tensor(faces, dtype=torch.int64), torch.tensor(idx, dtype=torch.int64), m, True)
        if tag == "w":
            loss = (phat * torch.tensor(weights(B, H, W), dtype=dtype)).sum()
        else:
            loss = (torch.tensor(target(B, H, W), dtype=dtype) - phat)[mask].pow(2).mean()      # optimize.py:168
            out["n_valid"] = int(mask.sum())
        loss.backward()
        out[f"loss_{tag}"] = float(loss)
        out[f"grad_verts_{tag}"] = v.grad.numpy().copy()
        out[f"grad_mus_{tag}"] = m.grad.numpy().copy()
    out["phat"] = phat.detach().numpy().astype(np.float32)
    return out


def main():
    geom_par = ref_sino.read_from_toml(os.path.join(REF, "data/2starA/proj_geom.toml"))
    geom_par6 = dict(geom_par, ProjectionAngles=geom_par["ProjectionAngles"][::5].copy())
    geom_cone = dict(geom_par)
    geom_cone.update(type="cone", DistanceOriginSource=10.0, DistanceOriginDetector=4.0)
    geom_cone["ProjectionAngles"] = (np.arange(6) * (2 * np.pi / 6) + 0.1).astype(np.float32)
    mesh = cases.two_material_mesh(2)
    todo = {
        "opengl_parallel3d": (ref_opengl, geom_par, np.arange(0, 30, 5)),
        "vecs_parallel3d_vec": (ref_vecs, ref_sino.geom_2vec(geom_par6), np.arange(6)),
        "vecs_cone_vec": (ref_vecs, ref_sino.geom_2vec(geom_cone), np.arange(6)),
    }
    for name, (mod, geom, idx) in todo.items():
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            got = run(mod, geom, mesh, idx, dtype)
            extra = {"Vectors": geom["Vectors"]} if "Vectors" in geom else {"ProjectionAngles": geom["ProjectionAngles"]}
            path = os.path.join(HERE, f"pipeline_{name}_{tag}.npz")
            np.savez_compressed(path, idx_angles=idx, geom_type=geom["type"], seed_w=SEED_W, seed_p=SEED_P,
                                DetectorRowCount=geom["DetectorRowCount"], DetectorColCount=geom["DetectorColCount"],
                                DetectorSpacingX=geom["DetectorSpacingX"], DetectorSpacingY=geom["DetectorSpacingY"],
                                **extra, **got)
            print(name, tag, "loss_w", got["loss_w"], "loss_l2", got["loss_l2"], "n_valid", got["n_valid"],
                  "|gv_w|", float(np.linalg.norm(got["grad_verts_w"])), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
