"""OBJ templates with material labels + seed-mesh generators, free of trimesh.

File format kept from ``ctdr/utils/util_mesh.py``: one ``o object<m> <label_in> <label_out>`` header
per closed surface (``:62``), parsed by :func:`load_template` (``:135-201``); faces are 1-based in
the file and 0-based in memory.
"""
from __future__ import annotations

import numpy as np

from .meshgen import icosphere, torus as _torus0  # noqa: F401


def torus(r, R, sides, rings, rotate=False):
    """``util_mesh.py:284-335`` (faces 1-based like the reference returns them)."""
    v, f = _torus0(r, R, sides, rings, rotate)
    return v, f + 1


def save_mesh(fname, verts, faces, labels_v=None, labels_f=None):
    """``util_mesh.py:37-70``: plain OBJ, or one labelled object per material."""
    faces = np.asarray(faces)
    if faces.min() == 0:
        faces = faces + 1
    lines = []
    if labels_v is None:
        lines += [f"v {v[0]} {v[1]} {v[2]}" for v in verts]
        lines += [f"f {t[0]} {t[1]} {t[2]}" for t in faces]
    else:
        labels_v, labels_f = np.asarray(labels_v), np.asarray(labels_f)
        for m in range(int(labels_v.max())):
            sel_f = labels_f[:, 0] == m + 1
            lin, lout = np.unique(labels_f[sel_f, 0])[0], np.unique(labels_f[sel_f, 1])[0]
            lines.append(f"o object{m} {lin} {lout}")
            lines += [f"v {v[0]} {v[1]} {v[2]}" for v in verts[labels_v == m + 1]]
            lines += [f"f {t[0]} {t[1]} {t[2]}" for t in faces[sel_f]]
    with open(fname, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def _taubin(verts, faces, lamb=0.5, nu=0.5, iterations=10):
    """Taubin smoothing with the uniform (equal-weight) graph Laplacian, alternating a shrinking step (lamb) and an
    inflating step (nu) — the defaults and update rule of ``trimesh.smoothing.filter_taubin`` used at
    ``util_mesh.py:95-96``."""
    v = np.array(verts, dtype=np.float64)
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    e = np.unique(np.sort(e, axis=1), axis=0)
    nbr_sum = np.zeros_like(v)
    deg = np.zeros(v.shape[0])
    np.add.at(deg, e[:, 0], 1.0); np.add.at(deg, e[:, 1], 1.0)
    for it in range(iterations):
        nbr_sum[:] = 0.0
        np.add.at(nbr_sum, e[:, 0], v[e[:, 1]]); np.add.at(nbr_sum, e[:, 1], v[e[:, 0]])
        dot = nbr_sum / np.maximum(deg, 1.0)[:, None] - v
        v = v + lamb * dot if it % 2 == 0 else v - nu * dot
    return v


def _subdivide(verts, faces):
    """Midpoint subdivision, every triangle into four, shared edge midpoints merged
    (``trimesh.remesh.subdivide`` at ``util_mesh.py:100``)."""
    edges = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
    uniq, inverse = np.unique(edges, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    nf, nv = faces.shape[0], verts.shape[0]
    m01, m12, m20 = inverse[:nf] + nv, inverse[nf:2 * nf] + nv, inverse[2 * nf:] + nv
    new_v = np.concatenate([verts, verts[uniq].mean(axis=1)])
    new_f = np.stack([np.stack([faces[:, 0], m01, m20], 1), np.stack([m01, faces[:, 1], m12], 1),
                      np.stack([m20, m12, faces[:, 2]], 1), np.stack([m01, m12, m20], 1)], axis=1).reshape(-1, 3)
    return new_v, new_f


def subdivide_save(fname, verts, faces, lv, lf, max_edge=0.1):
    """``util_mesh.py:72-133``: per material, Taubin-smooth the closed surface, split every triangle into four and
    write the labelled OBJ (used between the coarse-to-fine stages of ``run/ours.py:94``).  trimesh-free."""
    verts, faces, lv, lf = np.asarray(verts), np.asarray(faces), np.asarray(lv), np.asarray(lf)
    v_list, f_list, labels_v, labels_f = [], [], [], []
    nmaterials = int(lv.max()) + 1
    print("before refine: ", verts.shape)
    offset = 0
    for i in range(1, nmaterials):
        vi = verts[lv == i, :]
        fi = faces[lf[:, 0] == i, :]
        fi = fi - fi.min()
        print("filter by taubin")
        vi = _taubin(vi, fi)
        v_new, f_new = _subdivide(vi, fi)
        v_list.append(v_new)
        f_list.append(f_new + offset + 1)                  # 1-based, continuing after the previous materials
        offset += v_new.shape[0]
        labels_v.append(np.full(v_new.shape[0], i, dtype=np.int32))
        lft = np.zeros([f_new.shape[0], 2], dtype=np.int32)
        lft[:, 0] = i
        lft[:, 1] = lf[lf[:, 0] == i, 1][0]
        labels_f.append(lft)
    print("after refine: ", v_list[-1].shape)
    labels_v_out, labels_f_out = np.concatenate(labels_v), np.concatenate(labels_f)
    assert labels_v_out.max() == labels_f_out.max()
    save_mesh(fname, np.concatenate(v_list), np.concatenate(f_list), labels_v_out, labels_f_out)


def load_template(fname):
    """``util_mesh.py:135-201`` -> (vertices [V,3] f64, faces [F,3] int 0-based, labels_v [V], labels_f [F,2])."""
    verts, faces, objects = [], [], []           # objects: (label_in, label_out, first vertex, first face)
    with open(fname) as fh:
        for line in fh:
            tok = line.split()
            if len(tok) < 4:
                continue
            if line.startswith("v "):
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif line.startswith("f "):
                faces.append([int(t.split("/")[0]) for t in tok[1:4]])
            elif line.startswith("o "):
                objects.append((int(tok[2]), int(tok[3]), len(verts), len(faces)))
    vertices = np.array(verts)
    faces = np.array(faces)
    labels_f = np.zeros((faces.shape[0], 2), dtype=np.int32)
    labels_v = np.ones(vertices.shape[0], dtype=np.int32)
    if not objects:
        labels_f[:, 0] = 1
    else:
        # like the reference, the first object starts at 0 whatever precedes its header
        starts_v = [0] + [o[2] for o in objects[1:]] + [vertices.shape[0]]
        starts_f = [0] + [o[3] for o in objects[1:]] + [faces.shape[0]]
        for i, (lin, lout, _, _) in enumerate(objects):
            labels_v[starts_v[i]:starts_v[i + 1]] = i + 1
            labels_f[starts_f[i]:starts_f[i + 1]] = (lin, lout)
    faces = faces - 1
    print(f"@statistics of mesh: # of v: {vertices.shape[0]}, f: {faces.shape[0]}")
    return vertices, faces, labels_v, labels_f
