"""Initial meshes (``ctdr/dataset/init_mesh.py``) without trimesh: icospheres and tori from
``utils/meshgen`` written as labelled OBJ templates.  Topology types: ``A`` (one closed surface:
sphere, or torus for the two/four-component data sets), ``B`` (independent spheres), ``C`` (nested
spheres, one material inside the other), ``D`` (one outer sphere holding several inner ones)."""
from __future__ import annotations

import numpy as np

from ..utils import util_mesh
from ..utils.meshgen import icosphere

_CENTRES = 0.5 * np.array([[-1, -1, 0], [1, 1, 0], [0, 0, 1], [0, 0, -1], [0, 0, 0], [1, -1, 1], [1, 1, 1],
                           [-1, 1.0, 1], [1, 1, -1], [1, -1, -1]], dtype=np.float32)     # init_mesh.py:60-66


def make_icosphere(subdivisions, radius):
    """(vertices, faces) of an icosphere offset by +1e-4 "to avoid n.r = 0" (``init_mesh.py:9-12``)."""
    return icosphere(subdivisions, radius, 1e-4)


def save_obj_meshes_labels(fname, meshes, lf_list):
    """Concatenate closed surfaces with their (label_in, label_out) pairs (``init_mesh.py:14-41``)."""
    verts = np.concatenate([m[0] for m in meshes])
    offs = np.cumsum([0] + [m[0].shape[0] for m in meshes])
    faces = np.concatenate([m[1] + offs[i] for i, m in enumerate(meshes)])
    labels_v = np.concatenate([np.full(m[0].shape[0], lf[0], np.int32) for m, lf in zip(meshes, lf_list)])
    labels_f = np.concatenate([np.tile(np.array(lf, np.int32), (m[1].shape[0], 1)) for m, lf in zip(meshes, lf_list)])
    util_mesh.save_mesh(fname, verts, faces, labels_v, labels_f)


def save_init_mesh(fname, data, nmaterials, width_physical, subdivision=3, radius=0.4):
    """``init_mesh.py:44-153``: the data-set name encodes the topology (last letter) and the number
    of components (two digits before it)."""
    rad = width_physical * radius
    kind = data[-1]
    if not ("A" <= kind < "Z"):
        raise NotImplementedError
    try:
        ncomp = int(data[-3:-1])
    except ValueError:
        ncomp = 1
    if kind == "A":
        if ncomp == 1:
            v, f = make_icosphere(subdivision, rad * 0.5)
            util_mesh.save_mesh(fname, v + 1e-6, f)
        elif ncomp == 2:
            v, f = util_mesh.torus(rad * 0.3, rad * 0.6, 10 * subdivision, 10 * subdivision, data == "2kitten02A")
            util_mesh.save_mesh(fname, v + 1e-6, f)
        elif ncomp == 4:
            v, f = util_mesh.torus(rad * 0.4, rad * 0.8, 20, 40)
            util_mesh.save_mesh(fname, v + 1e-6, f)
    elif kind == "B":
        meshes = []
        for i in range(ncomp):
            v, f = make_icosphere(subdivision, rad * 0.4)
            meshes.append((v - _CENTRES[i], f))
        save_obj_meshes_labels(fname, meshes, [[1, 0]] * ncomp)
    elif kind == "C":
        meshes = [make_icosphere(subdivision, rad * 0.6), make_icosphere(subdivision, rad * 0.3)]
        labels = [[1, 0], [2, 1]]
        if nmaterials == 4:
            meshes.append(make_icosphere(subdivision, rad * 0.2)); labels.append([3, 2])
        elif nmaterials != 3:
            raise NotImplementedError
        save_obj_meshes_labels(fname, meshes, labels)
    elif kind == "D":
        v, f = make_icosphere(subdivision, rad * 0.99)
        meshes, labels = [(v + 1e-5, f)], [[1, 0]]
        for i in range(ncomp - 1):
            if nmaterials <= 3:
                v, f = make_icosphere(subdivision - 1, rad * 0.2); lab = [2, 1]
            else:
                v, f = make_icosphere(subdivision, rad * 0.1); lab = [i + 2, 1]
            meshes.append((v - _CENTRES[i], f)); labels.append(lab)
        save_obj_meshes_labels(fname, meshes, labels)
