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
