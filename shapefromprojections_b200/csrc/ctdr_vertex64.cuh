// ctdr_vertex64.cuh — the per-angle vertex stage in float64 (FP(..., dtype=torch.float64), vanilla.py:27,31).
// Same op order as ctdr_vertex.cuh (ctdr/nn/fp_opengl.py:87-158, ctdr/nn/fp_vecs.py:85-196) with double operands;
// geometry tables arrive as float64 (the reference builds them with torch.tensor(..., dtype=dtype): the float32
// ProjectionAngles widened, the float64 Vectors as they are).  Staged path only: p6 / len3 / ff are materialised
// for the float64 rasteriser (ctdr_raster64.cuh).
#pragma once
#include "ctdr_common.cuh"

namespace ctdr {

struct Vertex64Params {
    int kind;
    const double* geom;           // angles[nangles] or Vectors[nangles,12]
    const long long* idx_angles;  // [B]
    int nangles;
    int B, V, F, H, W;
    const double* verts;          // [V,3]
    const int* faces;             // [F,3]
    double msv;
    double P00, P11, P13;         // orthographic matrix entries: float32 values widened (fp_opengl.py:44-45)
};

struct Angle64 {
    double ct, st;
    double Sx, Sy, Sz, N0, N1, ux, uy, vz, DSx, DSy, DSz, coeff;
    bool use_x;
};

__device__ __forceinline__ Angle64 angle_const64(const Vertex64Params& P, int b) {
    Angle64 A = {};
    const long long row = P.idx_angles[b];
    if ((unsigned long long)row >= (unsigned long long)P.nangles) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        A.ct = A.st = A.Sx = A.Sy = A.Sz = A.N0 = A.N1 = A.ux = A.uy = A.vz = A.DSx = A.DSy = A.DSz = A.coeff = nan;
        return A;
    }
    if (P.kind == 0) {
        const double th = __ldg(P.geom + row);
        A.ct = cos(th); A.st = sin(th);                                   // fp_opengl.py:89-90
        return A;
    }
    const double* g = P.geom + row * 12;
    A.Sx = __ldg(g + 0); A.Sy = __ldg(g + 1); A.Sz = __ldg(g + 2);
    const double Dx = __ldg(g + 3), Dy = __ldg(g + 4), Dz = __ldg(g + 5);
    A.ux = __ldg(g + 6); A.uy = __ldg(g + 7);
    const double uz = __ldg(g + 8), vx = __ldg(g + 9), vy = __ldg(g + 10);
    A.vz = __ldg(g + 11);
    A.N0 = __dmul_rn(A.uy, A.vz);                                         // fp_vecs.py:120-122
    A.N1 = __dmul_rn(-A.ux, A.vz);
    double px, py, pz;
    if (P.kind == 1) {
        const double nn = __dsqrt_rn(__dadd_rn(__dmul_rn(A.N0, A.N0), __dmul_rn(A.N1, A.N1)));
        A.N0 = __ddiv_rn(A.N0, nn); A.N1 = __ddiv_rn(A.N1, nn);           // :128
        px = __dmul_rn(-P.msv, A.Sx); py = __dmul_rn(-P.msv, A.Sy); pz = __dmul_rn(-P.msv, A.Sz);  // :73-74
        A.coeff = 0.0;
    } else {
        px = Dx; py = Dy; pz = Dz;                                        // :72
        const double NS = __dadd_rn(__dmul_rn(A.N0, A.Sx), __dmul_rn(A.N1, A.Sy));   // :146
        const double D = -__dadd_rn(__dmul_rn(A.N0, px), __dmul_rn(A.N1, py));       // :147
        A.coeff = -__dadd_rn(NS, D);                                      // :148
    }
    const double hh = 0.5 * (double)P.H, hw = 0.5 * (double)P.W;          // :161
    A.DSx = __dsub_rn(__dsub_rn(px, __dmul_rn(hh, vx)), __dmul_rn(hw, A.ux));
    A.DSy = __dsub_rn(__dsub_rn(py, __dmul_rn(hh, vy)), __dmul_rn(hw, A.uy));
    A.DSz = __dsub_rn(__dsub_rn(pz, __dmul_rn(hh, A.vz)), __dmul_rn(hw, uz));
    A.use_x = fabs(A.ux) > fabs(A.uy);                                    // :168
    return A;
}

struct Vertex64Out {
    double px, py, len, cx, cy, cz;
};

__device__ __forceinline__ Vertex64Out project_vertex64(const Vertex64Params& P, const Angle64& A, double x, double y, double z) {
    Vertex64Out o = {};
    if (P.kind == 0) {
        const double xr = __dadd_rn(__dmul_rn(x, A.ct), __dmul_rn(y, A.st));      // fp_opengl.py:91
        const double yr = __dadd_rn(__dmul_rn(-x, A.st), __dmul_rn(y, A.ct));     // :92
        o.cx = xr; o.cy = -z; o.cz = __dadd_rn(yr, -P.msv);                       // :103-104
        o.len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(o.cx, o.cx), __dmul_rn(o.cy, o.cy)), __dmul_rn(o.cz, o.cz)));   // :107
        const double w = 1.0 + 1e-8;                                              // :113 (a no-op in float32, not in float64)
        const double nx = __ddiv_rn(__dmul_rn(P.P00, o.cx), w);                   // :111,113
        const double ny = __ddiv_rn(__dadd_rn(__dmul_rn(P.P11, o.cy), P.P13), w);
        o.px = __dmul_rn(__ddiv_rn(__dadd_rn(nx, 1.0), 2.0), (double)P.W);        // :157
        o.py = __dsub_rn(__ddiv_rn(__dmul_rn((double)P.H, __dsub_rn(1.0, ny)), 2.0), 0.5);   // :158
        return o;
    }
    double bx, by, bz, rx, ry, rz, t;
    if (P.kind == 1) {
        rx = -A.Sx; ry = -A.Sy; rz = -A.Sz;                                       // fp_vecs.py:130
        bx = x; by = y; bz = z;                                                   // :131
        const double NS = __dadd_rn(__dmul_rn(A.N0, x), __dmul_rn(A.N1, y));      // :132
        const double NR = __dadd_rn(__dmul_rn(A.N0, rx), __dmul_rn(A.N1, ry));    // :135
        t = __ddiv_rn(-__dadd_rn(NS, P.msv), NR);                                 // :133,136
        o.len = t;                                                                // :137
    } else {
        const double Rx = __dsub_rn(x, A.Sx), Ry = __dsub_rn(y, A.Sy), Rz = __dsub_rn(z, A.Sz);   // :141
        o.len = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(Rx, Rx), __dmul_rn(Ry, Ry)), __dmul_rn(Rz, Rz)));   // :142
        rx = __ddiv_rn(Rx, o.len); ry = __ddiv_rn(Ry, o.len); rz = __ddiv_rn(Rz, o.len);          // :145
        bx = A.Sx; by = A.Sy; bz = A.Sz;
        const double NR = __dadd_rn(__dmul_rn(A.N0, rx), __dmul_rn(A.N1, ry));    // :149
        t = __ddiv_rn(A.coeff, __dadd_rn(NR, 1e-10));                             // :150
    }
    o.cx = rx; o.cy = ry; o.cz = rz;
    const double Px = __dadd_rn(bx, __dmul_rn(t, rx));                            // :153
    const double Py = __dadd_rn(by, __dmul_rn(t, ry));
    const double Pz = __dadd_rn(bz, __dmul_rn(t, rz));
    const double dx = __dsub_rn(Px, A.DSx), dy = __dsub_rn(Py, A.DSy), dz = __dsub_rn(Pz, A.DSz);   // :162
    o.px = A.use_x ? __ddiv_rn(dx, A.ux) : __ddiv_rn(dy, A.uy);                   // :171-172
    o.py = __ddiv_rn(dz, A.vz);                                                   // :173
    return o;
}

// thread per (angle, face)
__global__ void __launch_bounds__(256)
k_vertex_forward64(const Vertex64Params P, double* __restrict__ p6, double* __restrict__ len3, double* __restrict__ ff) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.B * P.F) return;
    const int b = (int)(gid / P.F), f = (int)(gid % P.F);
    const Angle64 A = angle_const64(P, b);
    const int i0 = __ldg(P.faces + 3 * f), i1 = __ldg(P.faces + 3 * f + 1), i2 = __ldg(P.faces + 3 * f + 2);
    const double x0 = __ldg(P.verts + 3 * i0), y0 = __ldg(P.verts + 3 * i0 + 1), z0 = __ldg(P.verts + 3 * i0 + 2);
    const double x1 = __ldg(P.verts + 3 * i1), y1 = __ldg(P.verts + 3 * i1 + 1), z1 = __ldg(P.verts + 3 * i1 + 2);
    const double x2 = __ldg(P.verts + 3 * i2), y2 = __ldg(P.verts + 3 * i2 + 1), z2 = __ldg(P.verts + 3 * i2 + 2);
    const Vertex64Out a = project_vertex64(P, A, x0, y0, z0);
    const Vertex64Out bb = project_vertex64(P, A, x1, y1, z1);
    const Vertex64Out c = project_vertex64(P, A, x2, y2, z2);
    double facing;
    if (P.kind == 0) {      // camera-space normal, ray fixed at (0,0,-1): front_facing = -n_z (fp_opengl.py:137-148)
        const double e1x = __dsub_rn(bb.cx, a.cx), e1y = __dsub_rn(bb.cy, a.cy);
        const double e2x = __dsub_rn(c.cx, a.cx), e2y = __dsub_rn(c.cy, a.cy);
        facing = -__dsub_rn(__dmul_rn(e1x, e2y), __dmul_rn(e1y, e2x));
    } else {                // world-space normal (fp_vecs.py:85-91) . ray (parallel) or R^(v0) (cone) (:191-196)
        const double e1x = __dsub_rn(x1, x0), e1y = __dsub_rn(y1, y0), e1z = __dsub_rn(z1, z0);
        const double e2x = __dsub_rn(x2, x0), e2y = __dsub_rn(y2, y0), e2z = __dsub_rn(z2, z0);
        const double nx = __dsub_rn(__dmul_rn(e1y, e2z), __dmul_rn(e1z, e2y));
        const double ny = __dsub_rn(__dmul_rn(e1z, e2x), __dmul_rn(e1x, e2z));
        const double nz = __dsub_rn(__dmul_rn(e1x, e2y), __dmul_rn(e1y, e2x));
        const double rx = (P.kind == 2) ? a.cx : A.Sx, ry = (P.kind == 2) ? a.cy : A.Sy, rz = (P.kind == 2) ? a.cz : A.Sz;
        facing = __dadd_rn(__dadd_rn(__dmul_rn(rx, nx), __dmul_rn(ry, ny)), __dmul_rn(rz, nz));
    }
    double* o = p6 + gid * 6;
    o[0] = a.px; o[1] = a.py; o[2] = bb.px; o[3] = bb.py; o[4] = c.px; o[5] = c.py;
    len3[gid * 3 + 0] = a.len; len3[gid * 3 + 1] = bb.len; len3[gid * 3 + 2] = c.len;
    ff[gid] = facing;
}

// reverse mode of project_vertex64 (== torch autograd through fp_opengl.py:87-158 / fp_vecs.py:101-173)
__device__ __forceinline__ void vertex_vjp64(const Vertex64Params& P, const Angle64& A, double x, double y, double z,
                                             double gpx, double gpy, double glen, double& gx, double& gy, double& gz) {
    if (P.kind == 0) {
        const double xr = x * A.ct + y * A.st, yr = -x * A.st + y * A.ct;
        const double cz = yr - P.msv;
        const double il = 1.0 / sqrt(xr * xr + z * z + cz * cz);
        const double w = 1.0 + 1e-8;
        const double gxr = gpx * (0.5 * (double)P.W * P.P00 / w) + glen * xr * il;
        const double gyr = glen * cz * il;
        gx = gxr * A.ct - gyr * A.st;
        gy = gxr * A.st + gyr * A.ct;
        gz = gpy * (0.5 * (double)P.H * P.P11 / w) + glen * z * il;
        return;
    }
    double gPx = 0.0, gPy = 0.0;
    if (A.use_x) gPx = gpx / A.ux; else gPy = gpx / A.uy;
    const double gPz = gpy / A.vz;
    if (P.kind == 1) {
        const double rx = -A.Sx, ry = -A.Sy, rz = -A.Sz;
        const double NR = A.N0 * rx + A.N1 * ry;
        const double gt = gPx * rx + gPy * ry + gPz * rz + glen;
        const double gNS = -gt / NR;
        gx = gPx + gNS * A.N0; gy = gPy + gNS * A.N1; gz = gPz;
        return;
    }
    const double Rx = x - A.Sx, Ry = y - A.Sy, Rz = z - A.Sz;
    const double il = 1.0 / sqrt(Rx * Rx + Ry * Ry + Rz * Rz);
    const double rx = Rx * il, ry = Ry * il, rz = Rz * il;
    const double iNR = 1.0 / (A.N0 * rx + A.N1 * ry + 1e-10);
    const double t = A.coeff * iNR;
    const double gt = gPx * rx + gPy * ry + gPz * rz;
    double grx = t * gPx, gry = t * gPy, grz = t * gPz;
    const double gNR = -gt * t * iNR;
    grx += gNR * A.N0; gry += gNR * A.N1;
    const double rdot = rx * grx + ry * gry + rz * grz;
    gx = (grx - rx * rdot) * il + glen * rx;
    gy = (gry - ry * rdot) * il + glen * ry;
    gz = (grz - rz * rdot) * il + glen * rz;
}

// thread per face, angle slices along grid.y (as k_vertex_backward)
__global__ void __launch_bounds__(256)
k_vertex_backward64(const Vertex64Params P, const double* __restrict__ grad_p6, const double* __restrict__ grad_len3,
                    double* __restrict__ grad_verts, int angles_per_slice) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.F) return;
    const int b0 = blockIdx.y * angles_per_slice, b1 = min(b0 + angles_per_slice, P.B);
    const int vi[3] = {__ldg(P.faces + 3 * f), __ldg(P.faces + 3 * f + 1), __ldg(P.faces + 3 * f + 2)};
    double vx[3], vy[3], vz[3], ax[3] = {0, 0, 0}, ay[3] = {0, 0, 0}, az[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vx[k] = __ldg(P.verts + 3 * vi[k]); vy[k] = __ldg(P.verts + 3 * vi[k] + 1); vz[k] = __ldg(P.verts + 3 * vi[k] + 2);
    }
    bool any = false;
    for (int b = b0; b < b1; ++b) {
        const long long bf = (long long)b * P.F + f;
        double gp[6], gl[3];
        bool nz = false;
#pragma unroll
        for (int k = 0; k < 6; ++k) { gp[k] = __ldg(grad_p6 + bf * 6 + k); nz |= gp[k] != 0.0; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { gl[k] = __ldg(grad_len3 + bf * 3 + k); nz |= gl[k] != 0.0; }
        if (!nz) continue;
        any = true;
        const Angle64 A = angle_const64(P, b);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double gx, gy, gz;
            vertex_vjp64(P, A, vx[k], vy[k], vz[k], gp[2 * k], gp[2 * k + 1], gl[k], gx, gy, gz);
            ax[k] += gx; ay[k] += gy; az[k] += gz;
        }
    }
    if (!any) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        atomicAdd(grad_verts + 3 * vi[k] + 0, ax[k]);
        atomicAdd(grad_verts + 3 * vi[k] + 1, ay[k]);
        atomicAdd(grad_verts + 3 * vi[k] + 2, az[k]);
    }
}

}  // namespace ctdr
