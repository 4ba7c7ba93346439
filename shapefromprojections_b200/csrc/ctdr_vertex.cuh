// ctdr_vertex.cuh — per-angle vertex stage (world vertices -> detector pixels, ray lengths,
// facing sign) and its transposed Jacobian.  Replaces the ~40 torch ops of
//   ctdr/nn/fp_opengl.py:87-158  (geometry "parallel3d",      kind CTDR_GEOM_PARALLEL3D)
//   ctdr/nn/fp_vecs.py:85-196    ("parallel3d_vec"/"cone_vec", kinds *_VEC)
// and their autograd.  fp32 operations follow the reference's op order with explicit
// round-to-nearest intrinsics (torch evaluates every op separately, so nothing is fused there);
// results agree with the torch chain to a few ulp (exact agreement is not defined across
// cuBLAS/torch versions — SURVEY.md section 7, hard part 1).
#pragma once
#include "ctdr_common.cuh"

namespace ctdr {

struct VertexParams {
    int kind;
    const float* geom;            // angles[nangles] or Vectors[nangles,12]
    const long long* idx_angles;  // [B]
    int nangles;                  // rows of `geom`; an index outside [0, nangles) poisons its angle with NaN
    int B, V, F, H, W;
    const float* verts;           // [V,3]
    const int* faces;             // [F,3]
    float msv;                    // max_sino_val as float32
    float P00, P11, P13;          // orthographic matrix entries (projection.py:16-30), PARALLEL3D only
    const struct AngleConst* atab;  // [B] per-angle constants formed once by k_angle_table, or null (computed per thread)
};

struct __align__(16) AngleConst {
    // PARALLEL3D
    float ct, st;
    // *_VEC
    float Sx, Sy, Sz, N0, N1, ux, uy, vz, DSx, DSy, DSz, coeff;
    int use_x;
    int pad_;
};
static_assert(sizeof(AngleConst) == 64, "one angle = four 16-byte loads from the angle table");

__device__ __forceinline__ AngleConst angle_const(const VertexParams& P, int b) {
    if (P.atab) {
        // fused projector: the angle's constants were formed once (k_angle_table, same arithmetic) instead of by every
        // thread of the three vertex-stage kernels
        const float4* t = reinterpret_cast<const float4*>(P.atab + b);
        const float4 a = __ldg(t), c = __ldg(t + 1), d = __ldg(t + 2), e = __ldg(t + 3);
        AngleConst A;
        A.ct = a.x; A.st = a.y; A.Sx = a.z; A.Sy = a.w; A.Sz = c.x; A.N0 = c.y; A.N1 = c.z; A.ux = c.w;
        A.uy = d.x; A.vz = d.y; A.DSx = d.z; A.DSy = d.w; A.DSz = e.x; A.coeff = e.y; A.use_x = __float_as_int(e.z); A.pad_ = 0;
        return A;
    }
    AngleConst A = {};
    const long long row = P.idx_angles[b];
    if ((unsigned long long)row >= (unsigned long long)P.nangles) {
        // the reference raises IndexError (self.vecs[idx_angles], fp_vecs.py:101); a kernel cannot, so the
        // angle's output becomes NaN instead of reading outside the geometry table (hosts validate beforehand)
        const float nan = __int_as_float(0x7fc00000);
        A.ct = A.st = A.Sx = A.Sy = A.Sz = A.N0 = A.N1 = A.ux = A.uy = A.vz = A.DSx = A.DSy = A.DSz = A.coeff = nan;
        return A;
    }
    if (P.kind == 0) {
        const float th = __ldg(P.geom + row);
        A.ct = cosf(th); A.st = sinf(th);                               // fp_opengl.py:89-90
        return A;
    }
    const float* g = P.geom + row * 12;
    A.Sx = __ldg(g + 0); A.Sy = __ldg(g + 1); A.Sz = __ldg(g + 2);
    const float Dx = __ldg(g + 3), Dy = __ldg(g + 4), Dz = __ldg(g + 5);
    A.ux = __ldg(g + 6); A.uy = __ldg(g + 7);
    const float uz = __ldg(g + 8), vx = __ldg(g + 9), vy = __ldg(g + 10);
    A.vz = __ldg(g + 11);
    A.N0 = __fmul_rn(A.uy, A.vz);                                       // fp_vecs.py:120-122
    A.N1 = __fmul_rn(-A.ux, A.vz);
    float px, py, pz;                                                   // det_pos
    if (P.kind == 1) {
        const float nn = __fsqrt_rn(__fadd_rn(__fmul_rn(A.N0, A.N0), __fmul_rn(A.N1, A.N1)));
        A.N0 = __fdiv_rn(A.N0, nn); A.N1 = __fdiv_rn(A.N1, nn);         // :128
        px = __fmul_rn(-P.msv, A.Sx); py = __fmul_rn(-P.msv, A.Sy); pz = __fmul_rn(-P.msv, A.Sz);  // :73-74
        A.coeff = 0.0f;
    } else {
        px = Dx; py = Dy; pz = Dz;                                      // :72
        const float NS = __fadd_rn(__fmul_rn(A.N0, A.Sx), __fmul_rn(A.N1, A.Sy));   // :146
        const float D = -__fadd_rn(__fmul_rn(A.N0, px), __fmul_rn(A.N1, py));       // :147
        A.coeff = -__fadd_rn(NS, D);                                    // :148
    }
    const float hh = 0.5f * (float)P.H, hw = 0.5f * (float)P.W;        // :161
    A.DSx = __fsub_rn(__fsub_rn(px, __fmul_rn(hh, vx)), __fmul_rn(hw, A.ux));
    A.DSy = __fsub_rn(__fsub_rn(py, __fmul_rn(hh, vy)), __fmul_rn(hw, A.uy));
    A.DSz = __fsub_rn(__fsub_rn(pz, __fmul_rn(hh, A.vz)), __fmul_rn(hw, uz));
    A.use_x = fabsf(A.ux) > fabsf(A.uy) ? 1 : 0;                        // :168
    return A;
}

struct VertexOut {
    float px, py, len;
    float cx, cy, cz;   // PARALLEL3D: camera-space position; CONE: unit ray R/|R|; PARALLEL_VEC: unused
    float t;            // *_VEC: ray parameter
};

__device__ __forceinline__ VertexOut project_vertex(const VertexParams& P, const AngleConst& A,
                                                    float x, float y, float z) {
    VertexOut o = {};
    if (P.kind == 0) {
        const float xr = __fadd_rn(__fmul_rn(x, A.ct), __fmul_rn(y, A.st));      // fp_opengl.py:91
        const float yr = __fadd_rn(__fmul_rn(-x, A.st), __fmul_rn(y, A.ct));     // :92
        o.cx = xr; o.cy = -z; o.cz = __fadd_rn(yr, -P.msv);                      // :103-104
        o.len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(o.cx, o.cx), __fmul_rn(o.cy, o.cy)),
                                     __fmul_rn(o.cz, o.cz)));                    // :107
        const float nx = __fmul_rn(P.P00, o.cx);                                 // :111 (P03 == -0)
        const float ny = __fadd_rn(__fmul_rn(P.P11, o.cy), P.P13);
        o.px = __fmul_rn(__fdiv_rn(__fadd_rn(nx, 1.0f), 2.0f), (float)P.W);      // :157
        o.py = __fsub_rn(__fdiv_rn(__fmul_rn((float)P.H, __fsub_rn(1.0f, ny)), 2.0f), 0.5f);  // :158
        return o;
    }
    float bx, by, bz, rx, ry, rz, t;
    if (P.kind == 1) {
        rx = -A.Sx; ry = -A.Sy; rz = -A.Sz;                                      // fp_vecs.py:130
        bx = x; by = y; bz = z;                                                  // :131
        const float NS = __fadd_rn(__fmul_rn(A.N0, x), __fmul_rn(A.N1, y));      // :132
        const float NR = __fadd_rn(__fmul_rn(A.N0, rx), __fmul_rn(A.N1, ry));    // :135
        t = __fdiv_rn(-__fadd_rn(NS, P.msv), NR);                                // :133,136
        o.len = t;                                                               // :137
    } else {
        const float Rx = __fsub_rn(x, A.Sx), Ry = __fsub_rn(y, A.Sy), Rz = __fsub_rn(z, A.Sz);  // :141
        o.len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(Rx, Rx), __fmul_rn(Ry, Ry)), __fmul_rn(Rz, Rz)));  // :142
        rx = __fdiv_rn(Rx, o.len); ry = __fdiv_rn(Ry, o.len); rz = __fdiv_rn(Rz, o.len);      // :145
        bx = A.Sx; by = A.Sy; bz = A.Sz;
        const float NR = __fadd_rn(__fmul_rn(A.N0, rx), __fmul_rn(A.N1, ry));    // :149
        t = __fdiv_rn(A.coeff, __fadd_rn(NR, 1e-10f));                           // :150
    }
    o.cx = rx; o.cy = ry; o.cz = rz; o.t = t;
    const float Px = __fadd_rn(bx, __fmul_rn(t, rx));                            // :153
    const float Py = __fadd_rn(by, __fmul_rn(t, ry));
    const float Pz = __fadd_rn(bz, __fmul_rn(t, rz));
    const float dx = __fsub_rn(Px, A.DSx), dy = __fsub_rn(Py, A.DSy), dz = __fsub_rn(Pz, A.DSz);  // :162
    o.px = A.use_x ? __fdiv_rn(dx, A.ux) : __fdiv_rn(dy, A.uy);                  // :171-172
    o.py = __fdiv_rn(dz, A.vz);                                                  // :173
    return o;
}

// one thread per window angle: the per-angle constants of the vertex stage, once (VertexParams::atab must be null here)
__global__ void __launch_bounds__(128)
k_angle_table(const VertexParams P, AngleConst* __restrict__ tab) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    tab[b] = angle_const(P, b);
}

// thread per (angle, face): p6 / len3 / ff exactly as handed to Rasterizer.apply
__global__ void __launch_bounds__(256)
k_vertex_forward(const VertexParams P, float* __restrict__ p6, float* __restrict__ len3,
                 float* __restrict__ ff) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.B * P.F) return;
    const int b = (int)(gid / P.F), f = (int)(gid % P.F);
    const AngleConst A = angle_const(P, b);
    const int i0 = __ldg(P.faces + 3 * f), i1 = __ldg(P.faces + 3 * f + 1), i2 = __ldg(P.faces + 3 * f + 2);
    const float x0 = __ldg(P.verts + 3 * i0), y0 = __ldg(P.verts + 3 * i0 + 1), z0 = __ldg(P.verts + 3 * i0 + 2);
    const float x1 = __ldg(P.verts + 3 * i1), y1 = __ldg(P.verts + 3 * i1 + 1), z1 = __ldg(P.verts + 3 * i1 + 2);
    const float x2 = __ldg(P.verts + 3 * i2), y2 = __ldg(P.verts + 3 * i2 + 1), z2 = __ldg(P.verts + 3 * i2 + 2);
    const VertexOut a = project_vertex(P, A, x0, y0, z0);
    const VertexOut bb = project_vertex(P, A, x1, y1, z1);
    const VertexOut c = project_vertex(P, A, x2, y2, z2);
    float facing;
    if (P.kind == 0) {
        // camera-space normal, ray fixed at (0,0,-1): front_facing = -n_z (fp_opengl.py:137-148)
        const float e1x = __fsub_rn(bb.cx, a.cx), e1y = __fsub_rn(bb.cy, a.cy);
        const float e2x = __fsub_rn(c.cx, a.cx), e2y = __fsub_rn(c.cy, a.cy);
        facing = -__fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
    } else {
        // world-space normal (fp_vecs.py:85-91) dotted with the ray (parallel) or R^(v0) (cone) (:191-196)
        const float e1x = __fsub_rn(x1, x0), e1y = __fsub_rn(y1, y0), e1z = __fsub_rn(z1, z0);
        const float e2x = __fsub_rn(x2, x0), e2y = __fsub_rn(y2, y0), e2z = __fsub_rn(z2, z0);
        const float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
        const float ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z));
        const float nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
        const float rx = (P.kind == 2) ? a.cx : A.Sx, ry = (P.kind == 2) ? a.cy : A.Sy, rz = (P.kind == 2) ? a.cz : A.Sz;
        facing = __fadd_rn(__fadd_rn(__fmul_rn(rx, nx), __fmul_rn(ry, ny)), __fmul_rn(rz, nz));
    }
    float2* o = reinterpret_cast<float2*>(p6 + gid * 6);
    o[0] = make_float2(a.px, a.py); o[1] = make_float2(bb.px, bb.py); o[2] = make_float2(c.px, c.py);
    len3[gid * 3 + 0] = a.len; len3[gid * 3 + 1] = bb.len; len3[gid * 3 + 2] = c.len;
    ff[gid] = facing;
}

// Two-stage variant used by the fused projector: every vertex is transformed once per angle
// (thread per (angle, vertex) -> float4 {px, py, len, camera x}) and the faces gather three records
// each, instead of transforming each vertex once per incident face (~6x).  Same arithmetic, same
// results as k_vertex_forward.
__global__ void __launch_bounds__(256)
k_vertex_transform(const VertexParams P, float4* __restrict__ vt) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.B * P.V) return;
    const int b = (int)(gid / P.V), v = (int)(gid % P.V);
    const AngleConst A = angle_const(P, b);
    const VertexOut o = project_vertex(P, A, __ldg(P.verts + 3 * v), __ldg(P.verts + 3 * v + 1), __ldg(P.verts + 3 * v + 2));
    vt[gid] = make_float4(o.px, o.py, o.len, o.cx);
}

// grid = (ceil(F / 256), B): a warp holds 32 consecutive faces of ONE angle, i.e. exactly one chunk of the
// raster kernels, so the per-face pixel boxes and the chunk box (k_chunk_bbox's job at the Rasterizer
// boundary) come out of the same pass while the projected vertices are still in registers.
// FULL = false (gather mode of the raster kernels): only the boxes are written — the pixel box of every face, which
// also carries the facing sign (pack_box), and the chunk boxes; p6 / len3 / ff are not materialised at all.
// World-space face normals of the *_VEC geometries (fp_vecs.py:85-91), angle-independent: formed once per call
// instead of once per (angle, face) by k_face_assemble<false>.
__global__ void __launch_bounds__(256)
k_face_normals(const float* __restrict__ verts, const int* __restrict__ faces, int F, float4* __restrict__ fnorm) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int i0 = __ldg(faces + 3 * f), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
    const float x0 = __ldg(verts + 3 * i0), y0 = __ldg(verts + 3 * i0 + 1), z0 = __ldg(verts + 3 * i0 + 2);
    const float x1 = __ldg(verts + 3 * i1), y1 = __ldg(verts + 3 * i1 + 1), z1 = __ldg(verts + 3 * i1 + 2);
    const float x2 = __ldg(verts + 3 * i2), y2 = __ldg(verts + 3 * i2 + 1), z2 = __ldg(verts + 3 * i2 + 2);
    const float e1x = __fsub_rn(x1, x0), e1y = __fsub_rn(y1, y0), e1z = __fsub_rn(z1, z0);
    const float e2x = __fsub_rn(x2, x0), e2y = __fsub_rn(y2, y0), e2z = __fsub_rn(z2, z0);
    fnorm[f] = make_float4(__fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y)),
                           __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z)),
                           __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x)), 0.0f);
}

// fnorm (FULL = false, *_VEC geometries; may be null): the face normals of k_face_normals.  Only the SIGN of
// front_facing leaves this kernel then, and for the cone geometry it is decided without the three IEEE divisions of
// the unit ray (fp_vecs.py:145) whenever that is safe:  facing_ref = fl-sum of fl(fl(R_i/len) * n_i) differs from
// (R . n)/len by at most 4.1 u * sum|R_i n_i| / len (u = 2^-24: one rounding each for the quotient and the product,
// two for the sums), and E = fma-chain of R . n from sum R_i n_i by at most 3 u * sum|R_i n_i|; so when
// |E| > 1e-6 * sum|R_i n_i| (16.8 u, twice the 7.1 u needed) + 1e-30, facing_ref is non-zero and has the sign of E
// (len > 0).  Faces seen edge-on within that margin — and anything with a NaN — take the reference's own sequence.
template <bool FULL>
__global__ void __launch_bounds__(256)
k_face_assemble(const VertexParams P, const float4* __restrict__ vt, const int* __restrict__ labels2,
                float* __restrict__ p6, float* __restrict__ len3, float* __restrict__ ff,
                short4* __restrict__ face_box, short4* __restrict__ chunk_box, const float4* __restrict__ fnorm = nullptr) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    int x0 = BOX_EMPTY_LO, y0 = BOX_EMPTY_LO, x1 = -1, y1 = -1;
    if (f < P.F) {
    const long long gid = (long long)b * P.F + f;
    const int i0 = __ldg(P.faces + 3 * f), i1 = __ldg(P.faces + 3 * f + 1), i2 = __ldg(P.faces + 3 * f + 2);
    const float4* row = vt + (long long)b * P.V;
    const float4 a = __ldg(row + i0), bb = __ldg(row + i1), c = __ldg(row + i2);
    float facing;
    if (P.kind == 0) {
        // camera-space normal z from (camera x, camera y = -z_world) (fp_opengl.py:137-148)
        const float ya = -__ldg(P.verts + 3 * i0 + 2), yb = -__ldg(P.verts + 3 * i1 + 2), yc = -__ldg(P.verts + 3 * i2 + 2);
        const float e1x = __fsub_rn(bb.w, a.w), e1y = __fsub_rn(yb, ya);
        const float e2x = __fsub_rn(c.w, a.w), e2y = __fsub_rn(yc, ya);
        facing = -__fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
    } else if (!FULL && fnorm) {
        const float4 nrm = __ldg(fnorm + f);
        const AngleConst A = angle_const(P, b);
        if (P.kind == 1) {
            facing = __fadd_rn(__fadd_rn(__fmul_rn(A.Sx, nrm.x), __fmul_rn(A.Sy, nrm.y)), __fmul_rn(A.Sz, nrm.z));
        } else {
            const float Rx = __fsub_rn(__ldg(P.verts + 3 * i0), A.Sx), Ry = __fsub_rn(__ldg(P.verts + 3 * i0 + 1), A.Sy),
                        Rz = __fsub_rn(__ldg(P.verts + 3 * i0 + 2), A.Sz);                       // fp_vecs.py:141
            const float E = fmaf(Rz, nrm.z, fmaf(Ry, nrm.y, Rx * nrm.x));
            const float mag = fmaf(fabsf(Rz), fabsf(nrm.z), fmaf(fabsf(Ry), fabsf(nrm.y), fabsf(Rx * nrm.x)));
            if (fabsf(E) > fmaf(1e-6f, mag, 1e-30f)) {
                facing = E;                                                                  // sign only
            } else {                                                                         // edge-on within rounding, or NaN
                const float rx = __fdiv_rn(Rx, a.z), ry = __fdiv_rn(Ry, a.z), rz = __fdiv_rn(Rz, a.z);   // :145
                facing = __fadd_rn(__fadd_rn(__fmul_rn(rx, nrm.x), __fmul_rn(ry, nrm.y)), __fmul_rn(rz, nrm.z));
            }
        }
    } else {
        const float x0 = __ldg(P.verts + 3 * i0), y0 = __ldg(P.verts + 3 * i0 + 1), z0 = __ldg(P.verts + 3 * i0 + 2);
        const float x1 = __ldg(P.verts + 3 * i1), y1 = __ldg(P.verts + 3 * i1 + 1), z1 = __ldg(P.verts + 3 * i1 + 2);
        const float x2 = __ldg(P.verts + 3 * i2), y2 = __ldg(P.verts + 3 * i2 + 1), z2 = __ldg(P.verts + 3 * i2 + 2);
        const float e1x = __fsub_rn(x1, x0), e1y = __fsub_rn(y1, y0), e1z = __fsub_rn(z1, z0);
        const float e2x = __fsub_rn(x2, x0), e2y = __fsub_rn(y2, y0), e2z = __fsub_rn(z2, z0);
        const float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
        const float ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z));
        const float nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
        const AngleConst A = angle_const(P, b);
        float rx = A.Sx, ry = A.Sy, rz = A.Sz;
        if (P.kind == 2) {      // unit ray to the face's first vertex, as project_vertex forms it (fp_vecs.py:141-145)
            const float Rx = __fsub_rn(x0, A.Sx), Ry = __fsub_rn(y0, A.Sy), Rz = __fsub_rn(z0, A.Sz);
            rx = __fdiv_rn(Rx, a.z); ry = __fdiv_rn(Ry, a.z); rz = __fdiv_rn(Rz, a.z);
        }
        facing = __fadd_rn(__fadd_rn(__fmul_rn(rx, nx), __fmul_rn(ry, ny)), __fmul_rn(rz, nz));
    }
    if (FULL) {
        float2* o = reinterpret_cast<float2*>(p6 + gid * 6);
        o[0] = make_float2(a.x, a.y); o[1] = make_float2(bb.x, bb.y); o[2] = make_float2(c.x, c.y);
        len3[gid * 3 + 0] = a.z; len3[gid * 3 + 1] = bb.z; len3[gid * 3 + 2] = c.z;
        ff[gid] = facing;
    }
    const int2 lab = __ldg(reinterpret_cast<const int2*>(labels2) + f);
    int ix0, iy0, ix1, iy1;
    if (lab.x >= lab.y && face_int_bbox(a.x, a.y, bb.x, bb.y, c.x, c.y, P.H, P.W, ix0, iy0, ix1, iy1)) {   // as k_chunk_bbox
        x0 = ix0; y0 = iy0; x1 = ix1; y1 = iy1;
    }
    face_box[gid] = pack_box(x0, y0, x1, y1, facing < 0.0f);
    }
    // blockDim.x is a multiple of 32 and chunks start at multiples of 32: one warp == one chunk
    x0 = __reduce_min_sync(0xffffffffu, x0); y0 = __reduce_min_sync(0xffffffffu, y0);
    x1 = __reduce_max_sync(0xffffffffu, x1); y1 = __reduce_max_sync(0xffffffffu, y1);
    const int chunk = f >> 5, nchunks = (P.F + 31) >> 5;
    if ((threadIdx.x & 31) == 0 && chunk < nchunks)
        chunk_box[(long long)b * nchunks + chunk] = make_short4((short)x0, (short)y0, (short)x1, (short)y1);
}

// cotangent of one vertex: (gpx, gpy, glen) -> d/d(x,y,z), hand-written reverse mode of
// project_vertex (== what torch autograd does through fp_opengl.py:87-158 / fp_vecs.py:101-173).
// Only the intermediates the transposed Jacobian needs are recomputed (no px/py).
__device__ __forceinline__ void vertex_vjp(const VertexParams& P, const AngleConst& A,
                                           float x, float y, float z, float gpx, float gpy, float glen,
                                           float& gx, float& gy, float& gz) {
    if (P.kind == 0) {
        const float xr = x * A.ct + y * A.st, yr = -x * A.st + y * A.ct;
        const float cz = yr - P.msv;
        const float il = rsqrtf(xr * xr + z * z + cz * cz);
        const float gxr = gpx * (0.5f * (float)P.W * P.P00) + glen * xr * il;
        const float gyr = glen * cz * il;
        gx = gxr * A.ct - gyr * A.st;
        gy = gxr * A.st + gyr * A.ct;
        gz = gpy * (0.5f * (float)P.H * P.P11) + glen * z * il;      // cam y = -z: d(py)/dz = +H*P11/2
        return;
    }
    // P = base + t*r ; px = (P - DetS).{x|y} / u.{x|y} ; py = (P - DetS).z / vz
    float gPx = 0.0f, gPy = 0.0f;
    if (A.use_x) gPx = __fdividef(gpx, A.ux); else gPy = __fdividef(gpx, A.uy);
    const float gPz = __fdividef(gpy, A.vz);
    if (P.kind == 1) {
        // base = v, r = -S (constant), t = -(N.v + msv)/NR, len = t
        const float rx = -A.Sx, ry = -A.Sy, rz = -A.Sz;
        const float NR = A.N0 * rx + A.N1 * ry;
        const float gt = gPx * rx + gPy * ry + gPz * rz + glen;
        const float gNS = -__fdividef(gt, NR);
        gx = gPx + gNS * A.N0;
        gy = gPy + gNS * A.N1;
        gz = gPz;
        return;
    }
    // cone: R = v - S, len = |R|, r = R/len, t = coeff/(N.r_xy + 1e-10), P = S + t*r
    const float Rx = x - A.Sx, Ry = y - A.Sy, Rz = z - A.Sz;
    const float il = rsqrtf(Rx * Rx + Ry * Ry + Rz * Rz);
    const float rx = Rx * il, ry = Ry * il, rz = Rz * il;
    const float iNR = __fdividef(1.0f, A.N0 * rx + A.N1 * ry + 1e-10f);
    const float t = A.coeff * iNR;
    const float gt = gPx * rx + gPy * ry + gPz * rz;
    float grx = t * gPx, gry = t * gPy, grz = t * gPz;
    const float gNR = -gt * t * iNR;
    grx += gNR * A.N0; gry += gNR * A.N1;
    const float rdot = rx * grx + ry * gry + rz * grz;
    gx = (grx - rx * rdot) * il + glen * rx;
    gy = (gry - ry * rdot) * il + glen * ry;
    gz = (grz - rz * rdot) * il + glen * rz;
}

// thread per face, angle slices along grid.y: sums the cotangents of the face's three corners over
// the slice's angles in registers, then one atomic per (slice, corner, axis)
__global__ void __launch_bounds__(256)
k_vertex_backward(const VertexParams P, const float* __restrict__ grad_p6,
                  const float* __restrict__ grad_len3, float* __restrict__ grad_verts, int angles_per_slice) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.F) return;
    const int b0 = blockIdx.y * angles_per_slice, b1 = min(b0 + angles_per_slice, P.B);
    const int vi[3] = {__ldg(P.faces + 3 * f), __ldg(P.faces + 3 * f + 1), __ldg(P.faces + 3 * f + 2)};
    float vx[3], vy[3], vz[3], ax[3] = {0, 0, 0}, ay[3] = {0, 0, 0}, az[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vx[k] = __ldg(P.verts + 3 * vi[k]); vy[k] = __ldg(P.verts + 3 * vi[k] + 1); vz[k] = __ldg(P.verts + 3 * vi[k] + 2);
    }
    bool any = false;
    for (int b = b0; b < b1; ++b) {
        const long long bf = (long long)b * P.F + f;
        const float2* gp = reinterpret_cast<const float2*>(grad_p6 + bf * 6);
        const float2 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
        const float l0 = __ldg(grad_len3 + bf * 3), l1 = __ldg(grad_len3 + bf * 3 + 1), l2 = __ldg(grad_len3 + bf * 3 + 2);
        if (g0.x == 0.f && g0.y == 0.f && g1.x == 0.f && g1.y == 0.f && g2.x == 0.f && g2.y == 0.f &&
            l0 == 0.f && l1 == 0.f && l2 == 0.f) continue;
        any = true;
        const AngleConst A = angle_const(P, b);
        const float gpx[3] = {g0.x, g1.x, g2.x}, gpy[3] = {g0.y, g1.y, g2.y}, gl[3] = {l0, l1, l2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float gx, gy, gz;
            vertex_vjp(P, A, vx[k], vy[k], vz[k], gpx[k], gpy[k], gl[k], gx, gy, gz);
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

// Fused projector: the raster backward has already summed the corner cotangents per (angle, vertex)
// (RasterParams::grad_vt), so the transposed vertex stage is thread per vertex, angle slices along grid.y: no face
// gather, coalesced float4 reads, three atomics per (vertex, slice).
__global__ void __launch_bounds__(256)
k_vertex_backward_pervertex(const VertexParams P, const float4* __restrict__ grad_vt, float* __restrict__ grad_verts,
                            int angles_per_slice) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.V) return;
    const int b0 = blockIdx.y * angles_per_slice, b1 = min(b0 + angles_per_slice, P.B);
    const float x = __ldg(P.verts + 3 * v), y = __ldg(P.verts + 3 * v + 1), z = __ldg(P.verts + 3 * v + 2);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    bool any = false;
    for (int b = b0; b < b1; ++b) {
        const float4 g = __ldg(grad_vt + (long long)b * P.V + v);
        if (g.x == 0.0f && g.y == 0.0f && g.z == 0.0f) continue;
        any = true;
        const AngleConst A = angle_const(P, b);
        float gx, gy, gz;
        vertex_vjp(P, A, x, y, z, g.x, g.y, g.z, gx, gy, gz);
        ax += gx; ay += gy; az += gz;
    }
    if (!any) return;
    atomicAdd(grad_verts + 3 * v + 0, ax);
    atomicAdd(grad_verts + 3 * v + 1, ay);
    atomicAdd(grad_verts + 3 * v + 2, az);
}

}  // namespace ctdr
