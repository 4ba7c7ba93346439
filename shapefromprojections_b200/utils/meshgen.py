"""Dependency-free seed meshes (the reference needs ``trimesh`` for these).

* :func:`icosphere` — what ``ctdr/dataset/init_mesh.py:9-12`` gets from ``trimesh.creation.icosphere``
  (subdivision 4 -> V=2562, F=5120), vertices offset by ``+1e-4`` like ``make_icosphere``.
* :func:`torus` — same parametrisation and face winding as ``ctdr/utils/util_mesh.py:284-335``
  (exactly ``2*sides*rings`` faces), vectorised, 0-based faces.
"""
from __future__ import annotations

import numpy as np


def icosphere(subdivisions: int = 4, radius: float = 1.0, offset: float = 0.0):
    """Unit icosahedron, each triangle split in 4 per level, re-projected to the sphere.

    The 4 children of a face stay consecutive in the face list, so consecutive faces are
    spatially coherent (the projector's chunk culling benefits; correctness does not depend on it).
    """
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                      [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                      [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    verts /= np.linalg.norm(verts, axis=1, keepdims=True)
    faces = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                      [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                      [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                      [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        edges = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
        uniq, inverse = np.unique(edges, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
        mid = verts[uniq].mean(axis=1)
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        nf = faces.shape[0]
        m01 = inverse[:nf] + verts.shape[0]
        m12 = inverse[nf:2 * nf] + verts.shape[0]
        m20 = inverse[2 * nf:] + verts.shape[0]
        verts = np.concatenate([verts, mid])
        faces = np.stack([np.stack([faces[:, 0], m01, m20], 1), np.stack([m01, faces[:, 1], m12], 1),
                          np.stack([m20, m12, faces[:, 2]], 1), np.stack([m01, m12, m20], 1)],
                         axis=1).reshape(-1, 3)
    return verts * radius + offset, faces


def torus(r: float, R: float, sides: int, rings: int, rotate: bool = False):
    """Torus with tube radius ``r``, centre-line radius ``R``; vertex ``i*sides + j``; 0-based faces."""
    if sides <= 0 or rings <= 0:
        raise ValueError("sides and rings must be > 0.")
    phi = 2 * np.pi * np.arange(rings)[:, None] / rings
    theta = 2 * np.pi * np.arange(sides)[None, :] / sides
    x = ((R + r * np.cos(theta)) * np.cos(phi)).reshape(-1)
    y = ((R + r * np.cos(theta)) * np.sin(phi)).reshape(-1)
    z = np.broadcast_to(r * np.sin(theta), (rings, sides)).reshape(-1)
    verts = np.stack([x, y, z], 1) if rotate else np.stack([y, z, x], 1)
    # the reference walks ring pairs (-1,0),(0,1),... and side pairs likewise (util_mesh.py:324-333)
    i0 = (np.arange(-1, rings - 1) % rings)[:, None] * sides
    i1 = (np.arange(0, rings) % rings)[:, None] * sides
    j0 = (np.arange(-1, sides - 1) % sides)[None, :]
    j1 = (np.arange(0, sides) % sides)[None, :]
    v00, v01, v10, v11 = i0 + j0, i0 + j1, i1 + j0, i1 + j1
    faces = np.stack([np.stack([v00, v10, v11], -1), np.stack([v11, v01, v00], -1)], axis=2)
    return verts, faces.reshape(-1, 3).astype(np.int64)


def rigid_rotation(seed: int) -> np.ndarray:
    """Random rotation matrix from ``np.random.default_rng(seed)`` (QR of a Gaussian, det +1)."""
    q, rmat = np.linalg.qr(np.random.default_rng(seed).normal(size=(3, 3)))
    q = q * np.sign(np.diag(rmat))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q
