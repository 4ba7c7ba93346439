// ctdr_common.cuh — shared device helpers of the B200 mesh projector (sm_100a).
//
// The arithmetic in cover_exact()/face_int_bbox() restates the reference kernel
// ctdr/cuda/diff_fp_for.cu:100-151 (jakeoung/ShapeFromProjections) AS COMPILED for sm_100a:
// products rounded before the fused multiply-add (FMUL + FFMA), IEEE fp64 divide narrowed to
// fp32.  Round-to-nearest intrinsics pin that pattern so nvcc cannot re-associate it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctdr {

constexpr int      TILE       = 16;    // tile edge in detector pixels; one warp owns one tile
constexpr int      TILE_PIX   = TILE * TILE;
constexpr int      WARPS      = 4;     // warps per CTA
constexpr int      THREADS    = WARPS * 32;
constexpr int      CHUNK      = 32;    // consecutive faces culled/binned together (one lane each)
constexpr int      LIST_CAP   = 320;   // chunk-list capacity per CTA segment (>= THREADS)
constexpr unsigned FULL       = 0xffffffffu;
constexpr int      BOX_EMPTY_LO = 32767;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// fp32 add on a shared-memory address (explicit state space: a 4-instruction CAS loop instead of
// the generic-address atomicAdd expansion)
__device__ __forceinline__ void shared_add(float* p, float v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// Integer pixel range covered by the bounding box of one projected face, with the reference's
// whole-face reject (diff_fp_for.cu:105-106: any bbox edge off the detector drops the face) and
// candidate rule (:109-111: x0 >= xmin && x0 <= xmax ...).  torch.min/max propagate NaN
// (rasterizer.py:24-25): a NaN coordinate makes every comparison false, i.e. the face is neither
// rejected nor bounded along that axis.  Returns false when the face contributes nowhere.
__device__ __forceinline__ bool face_int_bbox(float ax, float ay, float bx, float by, float cx, float cy,
                                              int H, int W, int& ix0, int& iy0, int& ix1, int& iy1) {
    const bool nanx = (ax != ax) | (bx != bx) | (cx != cx);
    const bool nany = (ay != ay) | (by != by) | (cy != cy);
    const float xmin = fminf(fminf(ax, bx), cx), xmax = fmaxf(fmaxf(ax, bx), cx);
    const float ymin = fminf(fminf(ay, by), cy), ymax = fmaxf(fmaxf(ay, by), cy);
    bool reject = false;
    if (!nanx) reject |= (xmin < 0.0f) | (xmax > (float)(W - 1));
    if (!nany) reject |= (ymin < 0.0f) | (ymax > (float)(H - 1));
    if (reject) return false;
    if (nanx) { ix0 = 0; ix1 = W - 1; } else { ix0 = (int)ceilf(xmin); ix1 = (int)floorf(xmax); }
    if (nany) { iy0 = 0; iy1 = H - 1; } else { iy0 = (int)ceilf(ymin); iy1 = (int)floorf(ymax); }
    return (ix0 <= ix1) & (iy0 <= iy1);
}

// The per-face pixel box also carries the face's facing sign (the only thing the rasteriser needs of
// front_facing_bxfx1, diff_fp_for.cu:94-96) in bit 15 of .x: coordinates are <= 32767, so the bit is free.
__device__ __forceinline__ short4 pack_box(int x0, int y0, int x1, int y1, bool facing_neg) {
    return make_short4((short)(x0 | (facing_neg ? 0x8000 : 0)), (short)y0, (short)x1, (short)y1);
}
__device__ __forceinline__ short4 unpack_box(short4 fb, bool& facing_neg) {
    facing_neg = fb.x < 0;
    fb.x = (short)(fb.x & 0x7fff);
    return fb;
}
__device__ __forceinline__ short4 unpack_box(short4 fb) { fb.x = (short)(fb.x & 0x7fff); return fb; }

// Per-face constants of the barycentric solve (diff_fp_for.cu:126-130,142).
struct EdgeSetup {
    float ax, ay, m, p, n, q, k3;
};

__device__ __forceinline__ EdgeSetup edge_setup(float ax, float ay, float bx, float by, float cx, float cy) {
    EdgeSetup e;
    e.ax = ax; e.ay = ay;
    e.m = __fsub_rn(bx, ax); e.p = __fsub_rn(by, ay);
    e.n = __fsub_rn(cx, ax); e.q = __fsub_rn(cy, ay);
    e.k3 = __fmaf_rn(e.m, e.q, -__fmul_rn(e.n, e.p));
    return e;
}

// One (pixel, face) coverage evaluation, bit-identical to diff_fp_for.cu:132-151.
// PREFILTER: reject candidates that are provably outside without the two fp64 divides:
//   * w1 = fl32(k1/d) is negative  <=>  k1 and d have opposite signs, as long as the quotient
//     cannot round to -0 (|k1| > 1e-30, |d| < 1e10) — and d has the sign of k3 when |k3| > 1e-6;
//     same for w2;
//   * w0 = (1 - w1) - w2: an fp32 estimate that is below -1e-4*(1+|w1|+|w2|) cannot be >= 0
//     (the estimate is within ~5e-7*(1+|w1|+|w2|) of the reference value).
// Survivors (all true fragments and knife-edge candidates) take the exact path, so results are
// identical with and without the prefilter.
template <bool PREFILTER>
__device__ __forceinline__ bool cover_exact(const EdgeSetup& e, float x0, float y0,
                                            float& s, float& t, float& k1, float& k2,
                                            float& w0, float& w1, float& w2) {
    s = __fsub_rn(x0, e.ax);
    t = __fsub_rn(y0, e.ay);
    k1 = __fmaf_rn(s, e.q, -__fmul_rn(e.n, t));
    k2 = __fmaf_rn(e.m, t, -__fmul_rn(s, e.p));
    const double d = (double)e.k3 + 1e-15;
    if (PREFILTER) {
        // |k3| in (1e-6, 1e10): d = k3 + 1e-15 has the sign of k3 and 1/k3 approximates 1/d to 1e-9.
        const float ak3 = fabsf(e.k3);
        if ((ak3 > 1e-6f) & (ak3 < 1e10f)) {
            const bool dneg = e.k3 < 0.0f;
            if (fabsf(k1) > 1e-30f && ((k1 < 0.0f) != dneg)) return false;
            if (fabsf(k2) > 1e-30f && ((k2 < 0.0f) != dneg)) return false;
            const float rd = __frcp_rn(e.k3);
            const float a1 = k1 * rd, a2 = k2 * rd;
            const float a0 = 1.0f - a1 - a2;
            if (a0 < -1e-4f * (1.0f + fabsf(a1) + fabsf(a2))) return false;
        }
    }
    w1 = (float)((double)k1 / d);
    w2 = (float)((double)k2 / d);
    w0 = __fsub_rn(__fsub_rn(1.0f, w1), w2);
    return !(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f);
}

}  // namespace ctdr
