"""Generate tests/golden/vertex_stage_*.npz by running the REFERENCE's own projector modules.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_vertex_golden.py

The reference's ``ctdr/nn/fp_opengl.py`` / ``ctdr/nn/fp_vecs.py`` are imported unmodified with
``Tensor.cuda`` / ``Module.cuda`` stubbed to identity (no GPU here) and the native module
``cuda_diff_fp`` replaced by an empty stub; ``FP.rasterize`` is swapped for a recorder so what is
captured is exactly what the reference hands to ``Rasterizer.apply``
(p_bxfx6, len_bxfx3, front_facing_bxfx1, max_sino_val).  The recorder keeps the autograd graph, so
``grad_verts`` = the REFERENCE'S OWN autograd through its ~40 torch ops (fp_opengl.py:87-158,
fp_vecs.py:85-196) for seeded cotangents on p_bxfx6 / len_bxfx3 is stored too (``cotangents(seed, B, F)``
below regenerates them in the tests): the pin for ctdr_vertex_backward and the fused backward.  Geometry dicts come from the reference's
``util_sino.read_from_toml`` / ``geom_2vec`` on the shipped ``data/2starA/proj_geom.toml`` and on
a cone geometry with the numbers of ``run/simulation_cone_beam.py:22-23``.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
# the repository carries an alias package `ctdr/` (a regular package: it would shadow the reference's
# namespace package), so the reference modules are imported with only /root/reference on the path
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") != REPO]

sys.modules["cuda_diff_fp"] = types.ModuleType("cuda_diff_fp")     # native module: not needed here
torch.Tensor.cuda = lambda self, *a, **k: self                       # no GPU in this container
torch.nn.Module.cuda = lambda self, *a, **k: self

from ctdr.nn import fp_opengl as ref_opengl                          # noqa: E402  (reference code)
from ctdr.nn import fp_vecs as ref_vecs                              # noqa: E402
from ctdr.utils import util_sino as ref_sino                         # noqa: E402

assert ref_vecs.__file__.startswith(REF) and ref_opengl.__file__.startswith(REF)
sys.path.insert(1, REPO)
from shapefromprojections_b200.utils import meshgen                  # noqa: E402


COT_SEED = 20


def cotangents(seed, B, F):
    """Seeded float32 cotangents on (p_bxfx6, len_bxfx3); the tests call this with the stored seed."""
    rng = np.random.default_rng(seed)
    return rng.normal(size=(B, F, 6)).astype(np.float32), rng.normal(size=(B, F, 3)).astype(np.float32)


def capture(fp_module, proj_geom, verts, faces, idx, dtype):
    labels = torch.zeros(faces.shape[0], 2, dtype=torch.int32)
    fp = fp_module.FP(proj_geom, labels, dtype)
    got, live = {}, {}

    def recorder(p6, len3, ff, labels_, mus, H, W, msv, backprop):
        got.update(p6=p6.detach().numpy().copy(), len3=len3.detach().numpy().copy(),
                   ff=ff.detach().numpy().copy(), msv=float(msv), H=H, W=W)
        live.update(p6=p6, len3=len3)
        return torch.zeros(p6.shape[0], H, W, 1, dtype=dtype)

    fp.rasterize = recorder
    v = torch.tensor(verts, dtype=dtype, requires_grad=True)
    fp.forward(v, torch.tensor(faces, dtype=torch.int64),
               torch.tensor(idx, dtype=torch.int64), torch.tensor([0.0, 1.0], dtype=dtype), True)
    gp6, gl3 = cotangents(COT_SEED, len(idx), faces.shape[0])
    phi = (live["p6"] * torch.tensor(gp6, dtype=dtype)).sum() + (live["len3"] * torch.tensor(gl3, dtype=dtype)).sum()
    got["grad_verts"] = torch.autograd.grad(phi, v)[0].numpy().copy()     # the reference's autograd (index_backward scatter-adds ...)
    got["cot_seed"] = COT_SEED
    return got


def main():
    toml_path = os.path.join(REF, "data/2starA/proj_geom.toml")
    geom_par = ref_sino.read_from_toml(toml_path)
    # fp_vecs.FP only works for idx_angles == arange(nangles): ``det_pos`` is not indexed by
    # idx_angles (fp_vecs.py:147,161), so the vector cases use 6-angle geometries in full.
    geom_par6 = dict(geom_par, ProjectionAngles=geom_par["ProjectionAngles"][::5].copy())
    geom_par_vec = ref_sino.geom_2vec(geom_par6)
    geom_cone = dict(geom_par)
    geom_cone.update(type="cone", DistanceOriginSource=10.0, DistanceOriginDetector=4.0)
    geom_cone["ProjectionAngles"] = (np.arange(6) * (2 * np.pi / 6) + 0.1).astype(np.float32)
    geom_cone_vec = ref_sino.geom_2vec(geom_cone)

    sv, sf = meshgen.icosphere(2, 0.4, 1e-4 + 1e-6)
    tv, tf = meshgen.torus(0.12, 0.3, 12, 16)
    tv = tv @ meshgen.rigid_rotation(3).T + 1e-4
    verts = np.concatenate([sv, tv + np.array([0.05, -0.1, 0.2])])
    faces = np.concatenate([sf, tf + sv.shape[0]])

    cases = {
        "opengl_parallel3d": (ref_opengl, geom_par, np.arange(0, 30, 5)),
        "vecs_parallel3d_vec": (ref_vecs, geom_par_vec, np.arange(6)),
        "vecs_cone_vec": (ref_vecs, geom_cone_vec, np.arange(6)),
    }
    for name, (mod, geom, idx) in cases.items():
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            got = capture(mod, geom, verts, faces, idx, dtype)
            out = os.path.join(HERE, f"vertex_stage_{name}_{tag}.npz")
            extra = {}
            if "Vectors" in geom:
                extra["Vectors"] = geom["Vectors"]
            else:
                extra["ProjectionAngles"] = geom["ProjectionAngles"]
            for k in ("DistanceOriginSource", "DistanceOriginDetector"):
                if k in geom:
                    extra[k] = geom[k]
            np.savez_compressed(out, verts=verts, faces=faces, idx_angles=idx, geom_type=geom["type"],
                                DetectorRowCount=geom["DetectorRowCount"],
                                DetectorColCount=geom["DetectorColCount"],
                                DetectorSpacingX=geom["DetectorSpacingX"],
                                DetectorSpacingY=geom["DetectorSpacingY"], **extra, **got)
            print(name, tag, got["p6"].shape, "msv", got["msv"], os.path.getsize(out), "bytes")

    # geometry producers themselves (util_sino.py:16-26,127-202; projection.py:65-80)
    from ctdr.utils.projection import compute_P
    np.savez_compressed(os.path.join(HERE, "geometry_2starA.npz"),
                        ProjectionAngles=geom_par["ProjectionAngles"],
                        Vectors_parallel=ref_sino.geom_2vec(geom_par)["Vectors"],
                        cone_angles=geom_cone["ProjectionAngles"],
                        Vectors_cone=geom_cone_vec["Vectors"],
                        P=compute_P(geom_par))
    print("geometry ok")


if __name__ == "__main__":
    main()
