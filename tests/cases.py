"""Seeded inputs shared by the CPU (oracle) and GPU (parity) tests."""
from __future__ import annotations

import numpy as np

from shapefromprojections_b200.utils import meshgen, util_sino


def soup_case(seed=0, B=3, F=1500, H=96, W=96, n_mus=3, reference_safe=False, max_size=30.0):
    """Random triangle soup straight at the Rasterizer boundary (p6/len3/ff/labels/mus).

    Mixes tiny and large faces, faces partly off the detector (whole-face reject,
    diff_fp_for.cu:105-106), vertices exactly on pixel centres (weights exactly 0: the inclusive
    edge rule :149-151), degenerate (zero-area) faces, label_in == label_out, label_in < label_out,
    negative facing values and a face with a NaN-free but huge extent.
    """
    rng = np.random.default_rng(seed)
    centre = rng.uniform(-6, [W + 6, H + 6], size=(B, F, 1, 2))
    size = np.exp(rng.uniform(np.log(0.3), np.log(max_size), size=(B, F, 1, 1)))
    p = centre + rng.normal(size=(B, F, 3, 2)) * size
    snap = rng.random((B, F)) < 0.25                       # integer coordinates: knife-edge pixels
    p[snap] = np.round(p[snap])
    degenerate = rng.random((B, F)) < 0.03
    p[degenerate, 2] = p[degenerate, 1]
    p6 = p.reshape(B, F, 6).astype(np.float32)
    len3 = rng.uniform(0.5, 6.0, size=(B, F, 3)).astype(np.float32)
    ff = rng.normal(size=(B, F, 1)).astype(np.float32)
    ff[rng.random((B, F, 1)) < 0.02] = 0.0
    labels = np.zeros((F, 2), dtype=np.int32)
    labels[:, 0] = rng.integers(0, n_mus, size=F)
    labels[:, 1] = rng.integers(0, n_mus, size=F)
    # mostly the shipped convention [k, smaller]; reference_safe: always (the reference's pass 2
    # overruns its CSR slots for label_in < label_out faces)
    keep = (rng.random(F) < 0.8) | reference_safe
    lo = np.minimum(labels[:, 0], labels[:, 1]); hi = np.maximum(labels[:, 0], labels[:, 1])
    labels[keep, 0] = hi[keep]; labels[keep, 1] = lo[keep]
    mus = np.concatenate([[0.0], rng.uniform(0.2, 2.0, size=n_mus - 1)]).astype(np.float32)
    return dict(p6=p6, len3=len3, ff=ff, labels=labels, mus=mus, H=H, W=W, msv=float(W * 0.02 * 2))


def parallel_geometry(nangles=30, size=192, spacing=None):
    spacing = 2.0 / size if spacing is None else spacing
    return util_sino.make_geometry("parallel3d", nangles, size, size, spacing)


def cone_geometry(nangles=24, size=128, width=3.0, dso=10.0, dod=4.0):
    return util_sino.geom_2vec(util_sino.make_geometry("cone", nangles, size, size, width / size, dso=dso, dod=dod))


def two_material_mesh(subdiv=2):
    """Sphere (label 1 in background 0) with a small rotated torus inside it (label 2 in 1)."""
    sv, sf = meshgen.icosphere(subdiv, 0.55, 1e-4 + 1e-6)
    tv, tf = meshgen.torus(0.08, 0.25, 10, 14)
    tv = tv @ meshgen.rigid_rotation(5).T + np.array([0.03, -0.04, 0.05]) + 1e-4
    verts = np.concatenate([sv, tv]).astype(np.float32)
    faces = np.concatenate([sf, tf + sv.shape[0]]).astype(np.int64)
    labels = np.concatenate([np.tile([1, 0], (sf.shape[0], 1)), np.tile([2, 1], (tf.shape[0], 1))]).astype(np.int32)
    mus = np.array([0.0, 1.0, 1.7], dtype=np.float32)
    return verts, faces, labels, mus


def sphere_mesh(subdiv=4, radius=0.4):
    v, f = meshgen.icosphere(subdiv, radius, 1e-4 + 1e-6)     # init_mesh.py:9-12,71-72
    labels = np.tile([1, 0], (f.shape[0], 1)).astype(np.int32)
    return v.astype(np.float32), f.astype(np.int64), labels, np.array([0.0, 1.0], dtype=np.float32)


def torus_mesh(sides, rings, r=0.24, R=0.48, rotate_seed=None):
    v, f = meshgen.torus(r, R, sides, rings)
    if rotate_seed is not None:
        v = v @ meshgen.rigid_rotation(rotate_seed).T
    v = v + 1e-4
    labels = np.tile([1, 0], (f.shape[0], 1)).astype(np.int32)
    return v.astype(np.float32), f.astype(np.int64), labels, np.array([0.0, 1.0], dtype=np.float32)
