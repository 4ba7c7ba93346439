"""Orthographic projection matrix of the ``parallel3d`` projector (``fp_opengl``).

Mirror of ``ctdr/utils/projection.py:16-30,65-80`` (``ortho_Mat`` + the ``parallel3d`` branch of
``compute_P``).  The perspective branches of the reference are unreachable (``Model`` rejects
plain ``cone``: ``ctdr/model/vanilla.py:62-63``) and are not reproduced.
"""
import numpy as np


def ortho_Mat(left, right, bottom, top, near, far):
    """Row-major OpenGL orthographic matrix as float32 (``projection.py:16-30``)."""
    mat = np.eye(4, dtype=np.float64)
    mat[0, 0], mat[0, 3] = 2.0 / (right - left), -(right + left) / (right - left)
    mat[1, 1], mat[1, 3] = 2.0 / (top - bottom), -(top + bottom) / (top - bottom)
    mat[2, 2], mat[2, 3] = 2.0 / (near - far), -(far + near) / (far - near)
    return mat.astype(np.float32)


def compute_P(proj_geom):
    """``projection.py:65-80``: detector window shifted by half a pixel in v, depth range 4x half-width."""
    if proj_geom["type"] != "parallel3d":
        raise NotImplementedError("compute_P: only the parallel3d branch is on the projector path "
                                  "(cone data goes through the *_vec projector)")
    left = -proj_geom["DetectorSpacingX"] * proj_geom["DetectorColCount"] / 2.0
    bottom = -proj_geom["DetectorSpacingY"] * proj_geom["DetectorRowCount"] / 2.0
    half_u = 0.0 * proj_geom["DetectorSpacingX"]
    half_v = 0.5 * proj_geom["DetectorSpacingY"]
    depth = -4.0 * left
    return ortho_Mat(left - half_u, -left - half_u, bottom + half_v, -bottom + half_v,
                     -0.5 * depth, 0.5 * depth)
