"""oracle/vertex_stage.py — TEST INFRASTRUCTURE, NOT THE PRODUCT.

NumPy restatement (op by op, in the working dtype) of the reference's per-angle vertex stage:

* ``ctdr/nn/fp_opengl.py:33-63,87-158`` + ``ctdr/utils/projection.py:16-30,65-80``
  (geometry type ``parallel3d``; what ``run/ours.py`` exercises),
* ``ctdr/nn/fp_vecs.py:53-76,85-196`` + ``ctdr/utils/util_sino.py:127-202``
  (``parallel3d_vec`` / ``cone_vec``; what ``run/simulation_cone_beam.py`` exercises).

Outputs are exactly the three tensors the reference hands to ``Rasterizer.apply``:
``p_bxfx6``, ``len_bxfx3``, ``front_facing_bxfx1`` and ``max_sino_val``.

``vertex_stage_vjp`` is the hand-written reverse mode of the same chains (what torch autograd does through
the reference modules: index_backward scatter-adds of the gathers at fp_opengl.py:126-135 /
fp_vecs.py:176-186, then the transposed per-vertex transform), in float64.

Parity pin: ``tests/golden/vertex_stage_*.npz`` were produced by importing the reference's own
``FP`` modules in the build container with ``.cuda()`` stubbed (``tests/golden/make_vertex_golden.py``);
they hold the forward tensors and ``grad_verts`` from the reference's autograd for seeded cotangents.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- geometry
def geom_2vec(proj_geom: dict) -> dict:
    """util_sino.py:127-202 — ASTRA-style 12-vectors, float64 array, from float32 angles."""
    ang = proj_geom["ProjectionAngles"]
    n = len(ang)
    vec = np.zeros((n, 12))
    s, c = np.sin(ang), np.cos(ang)         # dtype of `ang` (float32 after read_from_toml :20)
    if proj_geom["type"] == "cone":
        dso, dod = proj_geom["DistanceOriginSource"], proj_geom["DistanceOriginDetector"]
        vec[:, 0], vec[:, 1] = s * dso, -c * dso                    # :137-138 source
        vec[:, 3], vec[:, 4] = -s * dod, c * dod                    # :142-143 detector centre
        kind = "cone_vec"
    elif proj_geom["type"] == "parallel3d":
        vec[:, 0], vec[:, 1] = s, -c                                # :171-172 ray direction
        kind = "parallel3d_vec"
    else:
        raise ValueError("No suitable vector geometry found for type: " + proj_geom["type"])
    vec[:, 6] = c * proj_geom["DetectorSpacingX"]                   # u
    vec[:, 7] = s * proj_geom["DetectorSpacingX"]
    vec[:, 11] = proj_geom["DetectorSpacingY"]                      # v
    return {"type": kind, "Vectors": vec,
            "DetectorRowCount": proj_geom["DetectorRowCount"],
            "DetectorColCount": proj_geom["DetectorColCount"],
            "DetectorSpacingX": proj_geom["DetectorSpacingX"],
            "DetectorSpacingY": proj_geom["DetectorSpacingY"]}


def ortho_P(proj_geom: dict) -> np.ndarray:
    """projection.py:65-80 (parallel3d branch) with ortho_Mat (:16-30); float32 4x4."""
    left = -proj_geom["DetectorSpacingX"] * proj_geom["DetectorColCount"] / 2.0
    bottom = -proj_geom["DetectorSpacingY"] * proj_geom["DetectorRowCount"] / 2.0
    half_u = 0.0 * proj_geom["DetectorSpacingX"]
    half_v = 0.5 * proj_geom["DetectorSpacingY"]
    span = -4.0 * left
    l, r = left - half_u, -left - half_u
    b, t = bottom + half_v, -bottom + half_v
    near, far = -0.5 * span, 0.5 * span
    return np.array([[2.0 / (r - l), 0, 0, -(r + l) / (r - l)],
                     [0, 2.0 / (t - b), 0, -(t + b) / (t - b)],
                     [0, 0, 2.0 / (near - far), -(far + near) / (far - near)],
                     [0, 0, 0, 1]], dtype=np.float32)


# ----------------------------------------------------------------------------- fp_opengl
def vertex_stage_opengl(proj_geom, verts, faces, idx_angles, dtype=np.float32):
    """fp_opengl.py:87-158.  verts [V,3], faces [F,3] int, idx_angles [B] int."""
    dt = np.dtype(dtype).type
    W, H = proj_geom["DetectorColCount"], proj_geom["DetectorRowCount"]
    msv = W * proj_geom["DetectorSpacingX"] * 2                      # :42
    P = ortho_P(proj_geom).astype(dtype)                            # :44-45
    angles = np.asarray(proj_geom["ProjectionAngles"]).astype(dtype)  # :63
    v = np.asarray(verts, dtype=dtype)
    th = angles[np.asarray(idx_angles)]
    ct, st = np.cos(th)[:, None], np.sin(th)[:, None]               # :89-90
    x, y, z = v[None, :, 0], v[None, :, 1], v[None, :, 2]
    xr = x * ct + y * st                                            # :91
    yr = -x * st + y * ct                                           # :92
    # M_left = [[1,0,0],[0,0,-1],[0,1,0]] (:50,103); V_t = (0,0,-msv) (:53,104)
    cam = np.stack([xr, np.broadcast_to(-z, xr.shape), yr + dt(-msv)], axis=2).astype(dtype)  # [B,V,3]
    length = np.sqrt((cam * cam).sum(axis=2, dtype=dtype))          # :107 torch.norm
    cam4 = np.concatenate([cam, np.ones_like(cam[..., :1])], axis=2)  # :110
    pv = cam4 @ P.T                                                 # :111
    ndc = pv[..., :3] / (pv[..., 3:4] + dt(1e-8))                   # :113
    f = np.asarray(faces)
    c0, c1, c2 = cam[:, f[:, 0]], cam[:, f[:, 1]], cam[:, f[:, 2]]  # :120-122
    len3 = np.stack([length[:, f[:, 0]], length[:, f[:, 1]], length[:, f[:, 2]]], axis=2)  # :126-129
    p6 = np.concatenate([ndc[:, f[:, 0], :2], ndc[:, f[:, 1], :2], ndc[:, f[:, 2], :2]], axis=2)  # :132-135
    normal = np.cross(c1 - c0, c2 - c0)                             # :137-142
    ff = -normal[..., 2:3]                                          # :148
    p6 = p6.copy()
    p6[..., 0::2] = (p6[..., 0::2] + dt(1.0)) / dt(2.0) * dt(W)     # :157
    p6[..., 1::2] = dt(H) * (dt(1) - p6[..., 1::2]) / dt(2.0) - dt(0.5)  # :158
    return p6.astype(dtype), len3.astype(dtype), ff.astype(dtype), float(msv)


# ----------------------------------------------------------------------------- fp_vecs
def vertex_stage_vecs(proj_geom, verts, faces, idx_angles, dtype=np.float32):
    """fp_vecs.py:85-196 for ``parallel3d_vec`` and ``cone_vec``."""
    dt = np.dtype(dtype).type
    kind = proj_geom["type"]
    W, H = proj_geom["DetectorColCount"], proj_geom["DetectorRowCount"]
    msv = W * proj_geom["DetectorSpacingX"] * (3 if kind == "parallel3d_vec" else 100)  # :59-62
    vecs_all = np.asarray(proj_geom["Vectors"]).astype(dtype)       # :64
    det_pos_all = vecs_all[:, 3:6]                                  # :72
    if kind == "parallel3d_vec":
        det_pos_all = -dt(msv) * vecs_all[:, 0:3]                   # :73-74
    v = np.asarray(verts, dtype=dtype)
    f = np.asarray(faces)
    normal = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]).astype(dtype)  # :85-91
    idx = np.asarray(idx_angles)
    vecs = vecs_all[idx]                                            # :101
    det_pos = det_pos_all[idx]
    B, V = vecs.shape[0], v.shape[0]
    vb = np.broadcast_to(v[None], (B, V, 3))                        # :102
    S = np.broadcast_to(vecs[:, None, 0:3], (B, V, 3))              # :114
    N = np.zeros((B, 2), dtype=dtype)                               # :120-122
    N[:, 0] = vecs[:, 7] * vecs[:, 11]
    N[:, 1] = -vecs[:, 6] * vecs[:, 11]
    if kind == "parallel3d_vec":
        N = N / np.sqrt(N[:, 0:1] ** 2 + N[:, 1:2] ** 2)            # :128
        Rn = -S                                                     # :130
        S = vb                                                      # :131
        NS = (N[:, None, :] * S[:, :, :2]).sum(axis=2, dtype=dtype)  # :132
        t = NS + dt(msv)                                            # :133
        NR = (N[:, None, :] * Rn[:, :, :2]).sum(axis=2, dtype=dtype)  # :135
        t = -t / NR                                                 # :136
        length = t                                                  # :137
    else:
        R = vb - S                                                  # :141
        length = np.sqrt((R * R).sum(axis=2, dtype=dtype))          # :142
        Rn = R / length[..., None]                                  # :145
        NS = (N * vecs[:, 0:2]).sum(axis=1, dtype=dtype)            # :146
        D = -(N * det_pos[:, 0:2]).sum(axis=1, dtype=dtype)         # :147
        coeff = -(NS + D)                                           # :148
        NR = (N[:, None, :] * Rn[:, :, :2]).sum(axis=2, dtype=dtype)  # :149
        t = coeff[:, None] / (NR + dt(1e-10))                       # :150
    Pt = S + t[..., None] * Rn                                      # :153
    DetS = det_pos - dt(0.5) * dt(H) * vecs[:, 9:12] - dt(0.5) * dt(W) * vecs[:, 6:9]  # :161
    PD = Pt - DetS[:, None, :]                                      # :162
    proj = np.zeros((B, V, 2), dtype=dtype)                         # :165
    use_x = np.abs(vecs[:, 6]) > np.abs(vecs[:, 7])                 # :168-169
    proj[use_x, :, 0] = PD[use_x, :, 0] / vecs[use_x, 6][:, None]   # :171
    proj[~use_x, :, 0] = PD[~use_x, :, 1] / vecs[~use_x, 7][:, None]  # :172
    proj[:, :, 1] = PD[:, :, 2] / vecs[:, 11][:, None]              # :173
    p6 = np.concatenate([proj[:, f[:, 0]], proj[:, f[:, 1]], proj[:, f[:, 2]]], axis=2)  # :176-179
    len3 = np.stack([length[:, f[:, 0]], length[:, f[:, 1]], length[:, f[:, 2]]], axis=2)  # :183-186
    if kind == "cone_vec":
        rb = Rn[:, f[:, 0], :]                                      # :192
    else:
        rb = np.broadcast_to(vecs[:, None, 0:3], (B, f.shape[0], 3))  # :194
    ff = (rb * normal[None]).sum(axis=2, dtype=dtype)[..., None]    # :196
    return p6.astype(dtype), len3.astype(dtype), ff.astype(dtype), float(msv)


def vertex_stage(proj_geom, verts, faces, idx_angles, dtype=np.float32):
    """Selection rule of ctdr/model/vanilla.py:66-73."""
    if proj_geom["type"][-3:] == "vec":
        return vertex_stage_vecs(proj_geom, verts, faces, idx_angles, dtype)
    return vertex_stage_opengl(proj_geom, verts, faces, idx_angles, dtype)


# ----------------------------------------------------------------------------- reverse mode
def _scatter_corners(faces, gp6, glen3, V):
    """index_backward of the per-face gathers: [B,F,6]/[B,F,3] -> per-vertex (gpx, gpy, glen) [B,V]."""
    f = np.asarray(faces)
    B = gp6.shape[0]
    gpx, gpy, gl = (np.zeros((B, V)) for _ in range(3))
    for k in range(3):
        for b in range(B):
            np.add.at(gpx[b], f[:, k], gp6[b, :, 2 * k])
            np.add.at(gpy[b], f[:, k], gp6[b, :, 2 * k + 1])
            np.add.at(gl[b], f[:, k], glen3[b, :, k])
    return gpx, gpy, gl


def vertex_stage_vjp(proj_geom, verts, faces, idx_angles, gp6, glen3):
    """grad_verts [V,3] (float64) = J^T (gp6, glen3) of ``vertex_stage`` (front_facing carries no
    gradient, rasterizer.py:107), summed over the angles in ``idx_angles``."""
    d = np.float64
    v = np.asarray(verts, dtype=d)
    V = v.shape[0]
    gp6, glen3 = np.asarray(gp6, dtype=d), np.asarray(glen3, dtype=d)
    gpx, gpy, gl = _scatter_corners(faces, gp6, glen3, V)
    W, H = proj_geom["DetectorColCount"], proj_geom["DetectorRowCount"]
    idx = np.asarray(idx_angles)
    x, y, z = v[None, :, 0], v[None, :, 1], v[None, :, 2]
    if proj_geom["type"][-3:] != "vec":                              # fp_opengl.py:87-158 reversed
        msv = W * proj_geom["DetectorSpacingX"] * 2
        P = ortho_P(proj_geom).astype(d)
        th = np.asarray(proj_geom["ProjectionAngles"]).astype(d)[idx]
        ct, st = np.cos(th)[:, None], np.sin(th)[:, None]
        xr, yr = x * ct + y * st, -x * st + y * ct
        cam = np.stack([xr, np.broadcast_to(-z, xr.shape), yr - msv], axis=2)
        length = np.sqrt((cam * cam).sum(axis=2))
        g_ndc0 = gpx * (W / 2.0) / (1.0 + 1e-8)                      # :157 then :113
        g_ndc1 = -gpy * (H / 2.0) / (1.0 + 1e-8)                     # :158 then :113
        g_cam = gl[..., None] * cam / length[..., None]              # :107
        g_cam[..., 0] += P[0, 0] * g_ndc0                            # :111
        g_cam[..., 1] += P[1, 1] * g_ndc1
        gx = ct * g_cam[..., 0] - st * g_cam[..., 2]                 # :91-92
        gy = st * g_cam[..., 0] + ct * g_cam[..., 2]
        gz = -g_cam[..., 1]                                          # M_left: cam y = -z
        return np.stack([gx.sum(0), gy.sum(0), gz.sum(0)], axis=1)
    kind = proj_geom["type"]
    msv = W * proj_geom["DetectorSpacingX"] * (3 if kind == "parallel3d_vec" else 100)
    vecs_all = np.asarray(proj_geom["Vectors"]).astype(d)
    det_pos_all = vecs_all[:, 3:6] if kind == "cone_vec" else -msv * vecs_all[:, 0:3]
    vecs, det_pos = vecs_all[idx], det_pos_all[idx]
    B = vecs.shape[0]
    N = np.stack([vecs[:, 7] * vecs[:, 11], -vecs[:, 6] * vecs[:, 11]], axis=1)
    use_x = np.abs(vecs[:, 6]) > np.abs(vecs[:, 7])
    gP = np.zeros((B, V, 3))                                         # :171-173 reversed
    gP[use_x, :, 0] = gpx[use_x] / vecs[use_x, 6][:, None]
    gP[~use_x, :, 1] = gpx[~use_x] / vecs[~use_x, 7][:, None]
    gP[:, :, 2] = gpy / vecs[:, 11][:, None]
    if kind == "parallel3d_vec":
        N = N / np.sqrt(N[:, 0:1] ** 2 + N[:, 1:2] ** 2)
        Rn = -vecs[:, None, 0:3]                                     # constant ray
        NR = (N * Rn[:, 0, :2]).sum(axis=1)[:, None]
        g_t = (gP * Rn).sum(axis=2) + gl                             # :153, :137
        g_v = gP.copy()                                              # Pt = v + t*Rn
        g_v[..., 0] += -g_t * N[:, 0:1] / NR                         # t = -(N.v_xy + msv)/NR
        g_v[..., 1] += -g_t * N[:, 1:2] / NR
        return g_v.sum(axis=0)
    S = vecs[:, None, 0:3]
    R = np.broadcast_to(v[None], (B, V, 3)) - S
    length = np.sqrt((R * R).sum(axis=2))
    Rn = R / length[..., None]
    coeff = -((N * vecs[:, 0:2]).sum(axis=1) - (N * det_pos[:, 0:2]).sum(axis=1))
    NRe = (N[:, None, :] * Rn[:, :, :2]).sum(axis=2) + 1e-10
    t = coeff[:, None] / NRe
    g_t = (gP * Rn).sum(axis=2)                                      # Pt = S + t*Rn
    g_Rn = t[..., None] * gP
    g_NR = -g_t * t / NRe                                            # t = coeff / NRe
    g_Rn[..., 0] += g_NR * N[:, 0:1]
    g_Rn[..., 1] += g_NR * N[:, 1:2]
    g_R = (g_Rn - Rn * (Rn * g_Rn).sum(axis=2, keepdims=True)) / length[..., None] + gl[..., None] * Rn
    return g_R.sum(axis=0)
