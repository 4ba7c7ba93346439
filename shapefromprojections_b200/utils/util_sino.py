"""Projection-geometry readers: the ``proj_geom.toml`` schema and the ASTRA-style 12-vectors.

Host-side mirror of the geometry producers on the hot path:
``read_from_toml`` / ``read_proj_info`` / ``geom_2vec`` keep the names, arguments and results of
``ctdr/utils/util_sino.py:16-46,127-202`` in jakeoung/ShapeFromProjections, including the quirk
that ``ProjectionAngles`` is stored as **float32** (``:20``) so every ``sin``/``cos`` downstream
carries float32 rounding — the projector must consume these arrays to keep coverage parity.
"""
from __future__ import annotations

import pickle

import numpy as np

try:  # the reference uses the third-party ``toml`` package; the stdlib parser reads the same files
    import tomllib as _toml_reader

    def _load_toml(path):
        with open(path, "rb") as fh:
            return _toml_reader.load(fh)
except ModuleNotFoundError:  # pragma: no cover - python < 3.11
    import toml as _toml_reader

    def _load_toml(path):
        with open(path, "r") as fh:
            return _toml_reader.load(fh)

_GEOM_KEYS = ("DetectorRowCount", "DetectorColCount", "DetectorSpacingX", "DetectorSpacingY")


def read_from_toml(fname_toml, nangles=None):
    """``util_sino.py:16-26``: dict with float32 ``ProjectionAngles`` (float64 linspace if ``nangles``)."""
    geom = dict(_load_toml(fname_toml))
    geom["ProjectionAngles"] = np.asarray(geom["ProjectionAngles"], dtype=np.float32)
    if nangles is not None:
        geom["ProjectionAngles"] = np.linspace(0, np.pi, nangles, False)
        geom["nangles"] = nangles
    return geom


def geom_2vec(proj_geom):
    """``util_sino.py:127-202``: ``cone`` -> ``cone_vec``, ``parallel3d`` -> ``parallel3d_vec``.

    ``Vectors[B,12]`` (float64): 0-2 source position (cone) or ray direction (parallel),
    3-5 detector centre, 6-8 ``u`` = pixel (0,0)->(0,1), 9-11 ``v`` = pixel (0,0)->(1,0).
    Trigonometry is evaluated in the dtype of ``ProjectionAngles`` (float32 from the TOML reader).
    """
    kind = proj_geom["type"]
    if kind not in ("cone", "parallel3d"):
        raise ValueError("No suitable vector geometry found for type: " + kind)
    theta = proj_geom["ProjectionAngles"]
    sin_t = np.array([np.sin(a) for a in theta])
    cos_t = np.array([np.cos(a) for a in theta])
    vectors = np.zeros((len(theta), 12))
    if kind == "cone":
        vectors[:, 0] = sin_t * proj_geom["DistanceOriginSource"]
        vectors[:, 1] = -cos_t * proj_geom["DistanceOriginSource"]
        vectors[:, 3] = -sin_t * proj_geom["DistanceOriginDetector"]
        vectors[:, 4] = cos_t * proj_geom["DistanceOriginDetector"]
    else:
        vectors[:, 0] = sin_t
        vectors[:, 1] = -cos_t
    vectors[:, 6] = cos_t * proj_geom["DetectorSpacingX"]
    vectors[:, 7] = sin_t * proj_geom["DetectorSpacingX"]
    vectors[:, 11] = proj_geom["DetectorSpacingY"]
    out = {"type": kind + "_vec", "Vectors": vectors}
    out.update({k: proj_geom[k] for k in _GEOM_KEYS})
    return out


def read_proj_info(fname, nangles=None, convert_vec=False):
    """``util_sino.py:29-46``: TOML (optionally converted to vectors) or pickled vector geometry."""
    proj_geom = None
    if fname[-4:] == "toml":
        proj_geom = read_from_toml(fname, nangles=nangles)
        if proj_geom["type"] == "cone" or convert_vec:
            proj_geom = geom_2vec(proj_geom)
    elif fname[-3:] == "pkl":
        with open(fname, "rb") as fh:
            proj_geom = pickle.load(fh)
        proj_geom["nangles"] = proj_geom["Vectors"].shape[0]
    return proj_geom


def make_geometry(kind, nangles, rows, cols, spacing_x, spacing_y=None, *, angle_span=None,
                  dso=None, dod=None):
    """Synthetic geometry dict in the TOML schema (not in the reference; used by bench/tests).

    ``kind``: ``parallel3d`` or ``cone``.  Angles are ``k*span/nangles`` stored as float32 exactly
    like a TOML file read through :func:`read_from_toml`.
    """
    if angle_span is None:
        angle_span = np.pi if kind == "parallel3d" else 2 * np.pi
    geom = {"type": kind,
            "ProjectionAngles": (np.arange(nangles) * (angle_span / nangles)).astype(np.float32),
            "DetectorRowCount": int(rows), "DetectorColCount": int(cols),
            "DetectorSpacingX": float(spacing_x),
            "DetectorSpacingY": float(spacing_x if spacing_y is None else spacing_y)}
    if kind == "cone":
        geom["DistanceOriginSource"] = float(dso)
        geom["DistanceOriginDetector"] = float(dod)
    return geom
