"""Seeded synthetic workloads of BASELINE.json / SURVEY.md section 8d (no dataset ships: every
``sinogram.npy`` of the reference is missing, so targets are synthesised)."""
from __future__ import annotations

import numpy as np

from .utils import meshgen, util_sino

CONFIGS = {
    # name: geometry kind, angles, detector, spacing, mesh
    "cfg1": dict(kind="parallel3d", nangles=30, size=192, spacing=0.01041666666666667,
                 mesh=("icosphere", 4, 0.4), desc="2starA geometry, icosphere(4): 30x192x192, F=5120"),
    "cfg3": dict(kind="cone", nangles=720, size=1024, spacing=3.0 / 1024, dso=10.0, dod=4.0,
                 mesh=("torus", 200, 250), desc="cone 10/4, 720x1024x1024, torus F=100000"),
    "cfg3_small": dict(kind="cone", nangles=48, size=512, spacing=3.0 / 512, dso=10.0, dod=4.0,
                       mesh=("torus", 100, 100), desc="cone 10/4, 48x512x512, torus F=20000 (smoke-sized)"),
    "northstar": dict(kind="cone", nangles=720, size=1024, spacing=3.0 / 1024, dso=10.0, dod=4.0,
                      mesh=("torus", 500, 1000), desc="cone 10/4, 720x1024x1024, torus F=1000000"),
    "cfg4": dict(kind="parallel3d", nangles=1800, size=2048, spacing=2.0 / 2048,
                 mesh=("torus", 500, 1000), desc="parallel, 1800x2048x2048, torus F=1000000"),
    # batched simulation (run/simulation_cone_beam.py style): 64 distinct meshes, forward only; the
    # workload dict carries mesh 0, the others come from cfg5_mesh(i)
    "cfg5": dict(kind="cone", nangles=360, size=512, spacing=3.0 / 512, dso=10.0, dod=4.0,
                 mesh=("cfg5", 0), nmeshes=64, desc="64 x (cone 10/4, 360x512x512, rotated torus F=20000), forward only"),
}
# config 2 (data/2bunnyA: the geometry numbers of 2starA, multi-material ours.py run) shares cfg1's geometry and
# initial mesh; its target is workloads.star_surface + impose_noise (bench.py: optimise_iteration)
CONFIGS["cfg2"] = dict(CONFIGS["cfg1"], desc="2bunnyA geometry 30x192x192, icosphere(4) F=5120, star-shaped noisy target")


def cfg5_mesh(i: int):
    """Mesh ``i`` of config 5: torus(r_i, R_i, 100, 100), r_i in [0.15, 0.3],
    R_i in [0.35, 0.55] and a rigid rotation, all from ``np.random.default_rng(i)``."""
    rng = np.random.default_rng(i)
    v, f = meshgen.torus(rng.uniform(0.15, 0.3), rng.uniform(0.35, 0.55), 100, 100)
    return v @ meshgen.rigid_rotation(i).T + 1e-4, f


def make_mesh(spec):
    if spec[0] == "icosphere":
        v, f = meshgen.icosphere(spec[1], spec[2], 1e-4 + 1e-6)          # init_mesh.py:9-12,71-72
    elif spec[0] == "cfg5":
        v, f = cfg5_mesh(spec[1])
    else:
        v, f = meshgen.torus(0.24, 0.48, spec[1], spec[2])               # util_mesh.py:284-335
        v = v + 1e-4
    labels = np.tile(np.array([1, 0], dtype=np.int32), (f.shape[0], 1))
    mus = np.array([0.0, 1.0], dtype=np.float32)                         # ours.py:46
    return v.astype(np.float32), f.astype(np.int64), labels, mus


def make_workload(name: str):
    cfg = CONFIGS[name]
    geom = util_sino.make_geometry(cfg["kind"], cfg["nangles"], cfg["size"], cfg["size"], cfg["spacing"],
                                   dso=cfg.get("dso"), dod=cfg.get("dod"))
    if cfg["kind"] == "cone":
        geom = util_sino.geom_2vec(geom)                                 # Model rejects plain "cone" (vanilla.py:62)
    verts, faces, labels, mus = make_mesh(cfg["mesh"])
    return dict(name=name, desc=cfg["desc"], geom=geom, verts=verts, faces=faces, labels=labels, mus=mus,
                nangles=cfg["nangles"], H=cfg["size"], W=cfg["size"], nmeshes=cfg.get("nmeshes", 1))


TARGET_SCALE = np.array([1.04, 0.97, 1.02], dtype=np.float32)            # ground-truth shape = scaled mesh


def algorithmic_bytes(B, H, W, V, F, n_gpus=1):
    """8 B per detector pixel (4 B sinogram store + 4 B target/gradient load)
    plus the replicated mesh/geometry term per GPU."""
    return 8 * B * H * W + n_gpus * (24 * V + 20 * F) + 48 * B


def star_surface(subdiv=4):
    """Ground-truth shape of config 2: a star-shaped radial perturbation of the
    icosphere, r(theta, phi) = 0.45 * (1 + 0.25 * sin(3 theta) * sin(4 phi))."""
    v, f = meshgen.icosphere(subdiv, 1.0, 0.0)
    theta = np.arccos(np.clip(v[:, 2], -1.0, 1.0))
    phi = np.arctan2(v[:, 1], v[:, 0])
    r = 0.45 * (1.0 + 0.25 * np.sin(3 * theta) * np.sin(4 * phi))
    return (v * r[:, None] + 1e-4).astype(np.float32), f.astype(np.int64)
