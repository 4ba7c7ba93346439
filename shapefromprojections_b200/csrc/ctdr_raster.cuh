// ctdr_raster.cuh — tile-binned mesh rasteriser: forward (sinogram), fragment list (pass 2),
// backward (moment-reduced gradients), fused forward+residual+backward.  sm_100a, fp32 + the
// reference's fp64 barycentric divide.
//
// Work decomposition:
//   angle b  ->  regions of RT x RT tiles (one CTA each)  ->  16x16-pixel tiles (one warp each)
//   faces    ->  chunks of 32 consecutive faces (one lane each); per (angle, chunk) a pixel
//                bounding box is precomputed by k_chunk_bbox.
//   A CTA compacts, IN FACE ORDER, the chunks whose box overlaps its region into a shared-memory
//   list (ballot + popc ranking); each warp then walks that list for its tile (32 boxes per
//   ballot), stages the chunk's 32 faces in shared memory and spreads the covered bounding-box
//   pixels ("candidates") evenly over its lanes through a prefix sum, so lane utilisation does not
//   depend on triangle size.  Because lists, chunks and candidates are all visited in ascending
//   face order and a tile is owned by one warp, every pixel accumulates its faces in exactly the
//   order of the reference's per-pixel loop (diff_fp_for.cu:83) -> bit-identical sinograms, with no
//   global atomics and one coalesced store per pixel.
#pragma once
#include "ctdr_common.cuh"

namespace ctdr {

enum RasterMode { MODE_FWD = 0, MODE_FRAG = 1, MODE_BWD = 2, MODE_STEP = 3 };

struct RasterParams {
    // inputs at the Rasterizer boundary (rasterizer.py:13)
    const float* p6;        // [B,F,6]
    const float* ff;        // [B,F]
    const float* len3;      // [B,F,3]
    const int*   labels2;   // [F,2]
    const float* mus;       // [n_mus]
    int n_mus;
    int B, F, H, W;
    float msv;
    // decomposition
    int NC;                 // chunks per angle = ceil(F/32)
    int RT;                 // tiles per region edge
    int regions_x, regions_y;
    const short4* chunk_box;  // [B,NC] pixel bbox of each chunk (x0,y0,x1,y1), empty: x0 > x1
    // MODE_FWD outputs / MODE_STEP phat
    float* im;              // [B,H,W]
    int*   cnt;             // [B,H,W] or null
    // MODE_FRAG
    const int* firstidx;    // [B,H,W]
    int*   imidx;           // [vf]
    float* imwei;           // [3 vf]
    // MODE_BWD inputs
    const float* grad_im;   // [B,H,W]
    const float* im_in;     // [B,H,W]
    // MODE_BWD / MODE_STEP outputs (accumulated)
    float* grad_p6;         // [B,F,6]
    float* grad_len3;       // [B,F,3]
    float* grad_mus;        // [n_mus]
    // MODE_STEP
    const float* target;    // [B,H,W]
    float mask_hi;          // float32(max_sino_val - 1e-6), get_valid_mask upper bound (fp_vecs.py:13)
    double* out2;           // {loss_sum, n_valid}
    unsigned long long* stats;  // may be null
};

// ------------------------------------------------------------------------------------------
// per-(angle, chunk) pixel bounding boxes; one warp per chunk
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
k_chunk_bbox(const float* __restrict__ p6, const int* __restrict__ labels2,
             int B, int F, int H, int W, int NC, short4* __restrict__ chunk_box) {
    const long long gw = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * NC) return;
    const int b = (int)(gw / NC), c = (int)(gw % NC);
    const int f = c * CHUNK + lane;
    int x0 = BOX_EMPTY_LO, y0 = BOX_EMPTY_LO, x1 = -1, y1 = -1;
    if (f < F) {
        const int2 lab = __ldg(reinterpret_cast<const int2*>(labels2) + f);
        if (lab.x >= lab.y) {   // diff_fp_for.cu:161-163 (applied in every pass)
            const float2* pp = reinterpret_cast<const float2*>(p6 + ((long long)b * F + f) * 6);
            const float2 a = __ldg(pp), bb = __ldg(pp + 1), cc = __ldg(pp + 2);
            int ix0, iy0, ix1, iy1;
            if (face_int_bbox(a.x, a.y, bb.x, bb.y, cc.x, cc.y, H, W, ix0, iy0, ix1, iy1)) {
                x0 = ix0; y0 = iy0; x1 = ix1; y1 = iy1;
            }
        }
    }
    x0 = __reduce_min_sync(FULL, x0); y0 = __reduce_min_sync(FULL, y0);
    x1 = __reduce_max_sync(FULL, x1); y1 = __reduce_max_sync(FULL, y1);
    if (lane == 0) chunk_box[gw] = make_short4((short)x0, (short)y0, (short)x1, (short)y1);
}

// ------------------------------------------------------------------------------------------
// shared memory
// ------------------------------------------------------------------------------------------
struct __align__(16) WarpSmem {
    // staged chunk: one slot per lane/face, structure of arrays (bank = slot index)
    float ax[32], ay[32], m[32], p[32], n[32], q[32], k3[32];
    float r0[32], r1[32], r2[32];
    float cA[32], cB[32];        // forward: +sign*mu_in, -(sign*mu_out')   (diff_fp_for.cu:187,190)
    int   box[32];               // lx0 | ly0<<4 | w<<8 | ceil(2^13/w)<<13
    int   excl[32];              // exclusive candidate offset of the slot
    int   offs[32];              // inclusive candidate offsets (binary-searched)
    float mom[5][32];            // backward: sum g, g*s, g*t, g*k1, g*k2 per slot
    // tile accumulators
    float acc[TILE_PIX];         // forward: running sinogram; backward: g (0 where no gradient)
    int   cntv[TILE_PIX];        // forward: fragment count; MODE_FRAG: CSR cursor, -1 = invalid pixel
};

struct CtaSmem {
    int    list_id[LIST_CAP];
    short4 list_box[LIST_CAP];
    int    wsum[WARPS];
    int    list_n;
    int    next_tile;
    WarpSmem warp[WARPS];
    // followed by float gmus[n_mus] (dynamic)
};

__device__ __forceinline__ bool box_overlap(const short4 bx, int x0, int y0, int x1, int y1) {
    return (bx.x <= x1) & (bx.z >= x0) & (bx.y <= y1) & (bx.w >= y0);
}

// segmented (by contiguous key) reduction toward the lowest lane of each run
__device__ __forceinline__ void seg_reduce5(int key, int lane, float (&v)[5]) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int kk = __shfl_down_sync(FULL, key, d);
        float u[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) u[k] = __shfl_down_sync(FULL, v[k], d);
        if (lane + d < 32 && kk == key) {
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] += u[k];
        }
    }
}


// Gradients of one face from five moments of the upstream gradient g over the face's fragments:
// M0 = sum g, Ms = sum g*s, Mt = sum g*t, A1 = sum g*k1, A2 = sum g*k2 (s = x0-ax, t = y0-ay, k1/k2
// exactly as the forward computes them).  Every per-fragment term of diff_fp_back.cu:126-136
// (grad_len), :143-147 (interpolated length * g, for grad_mus) and :206-292 (grad_points2d) is a
// per-face constant times one of {1, s, t, k1, k2}, and s,t only ever appear multiplied by k3 — so
// a zero-area face (k1 = k2 = k3 = 0) yields exact zeros like the reference, before the 1/(k3^2+eps)
// factor.  cgeo = sign*(mu_in - mu_out): the two label iterations of back.cu:111 merged.
__device__ __forceinline__ void face_grads_from_moments(const EdgeSetup& es, float r0, float r1, float r2,
                                                        float cgeo, float M0, float Ms, float Mt,
                                                        float A1, float A2,
                                                        float gp[6], float gl[3], float& bsum) {
    const float k3 = es.k3;
    const double dd = (double)k3 + 1e-15;
    const float W1 = (float)((double)A1 / dd), W2 = (float)((double)A2 / dd);
    const float W0 = M0 - W1 - W2;                            // sum g*w_k
    gl[0] = cgeo * W0; gl[1] = cgeo * W1; gl[2] = cgeo * W2;
    bsum = r0 * W0 + r1 * W1 + r2 * W2;
    const float d1m = -es.q * A1,               d2m = k3 * Mt - es.q * A2;      // back.cu:231-243
    const float d1n = -k3 * Mt + es.p * A1,     d2n = es.p * A2;
    const float d1p = es.n * A1,                d2p = -k3 * Ms + es.n * A2;
    const float d1q = k3 * Ms - es.m * A1,      d2q = -es.m * A2;
    const float d1s = es.q * k3 * M0,           d2s = -es.p * k3 * M0;
    const float d1t = -es.n * k3 * M0,          d2t = es.m * k3 * M0;
    const float c10 = r1 - r0, c20 = r2 - r0;                 // back.cu:271-280
    const float G0 = (float)((double)cgeo / ((double)(k3 * k3) + 1e-15));       // back.cu:282-283
    gp[0] = -G0 * (c10 * (d1m + d1n + d1s) + c20 * (d2m + d2n + d2s));          // back.cu:246-258
    gp[1] = -G0 * (c10 * (d1p + d1q + d1t) + c20 * (d2p + d2q + d2t));
    gp[2] = G0 * (c10 * d1m + c20 * d2m); gp[3] = G0 * (c10 * d1p + c20 * d2p);
    gp[4] = G0 * (c10 * d1n + c20 * d2n); gp[5] = G0 * (c10 * d1q + c20 * d2q);
}

struct ThreadStats {
    unsigned int cand, frag, nonpos;
};

// ------------------------------------------------------------------------------------------
// one (tile, chunk): stage the chunk's faces, spread candidates over lanes, evaluate coverage
// ------------------------------------------------------------------------------------------
template <int MODE, bool PREFILTER>
__device__ __forceinline__ void process_chunk(const RasterParams& P, WarpSmem& S, float* gmus,
                                              int b, int chunk, int tx0, int ty0, int tx1, int ty1,
                                              int lane, ThreadStats& st) {
    constexpr bool IS_FWD = (MODE == MODE_FWD);
    constexpr bool IS_FRAG = (MODE == MODE_FRAG);
    constexpr bool IS_BWD = (MODE == MODE_BWD || MODE == MODE_STEP);
    const int f = chunk * CHUNK + lane;
    const long long bf = (long long)b * P.F + f;
    int ncand = 0;
    int2 lab = make_int2(0, 0);
    float ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;
    int cx0 = 0, cy0 = 0, bw = 0;
    if (f < P.F) {
        lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
        if (lab.x >= lab.y) {
            const float2* pp = reinterpret_cast<const float2*>(P.p6 + bf * 6);
            const float2 va = __ldg(pp), vb = __ldg(pp + 1), vc = __ldg(pp + 2);
            ax = va.x; ay = va.y; bx = vb.x; by = vb.y; cx = vc.x; cy = vc.y;
            int ix0, iy0, ix1, iy1;
            if (face_int_bbox(ax, ay, bx, by, cx, cy, P.H, P.W, ix0, iy0, ix1, iy1)) {
                cx0 = max(ix0, tx0); cy0 = max(iy0, ty0);
                const int cx1 = min(ix1, tx1), cy1 = min(iy1, ty1);
                bw = cx1 - cx0 + 1;
                const int bh = cy1 - cy0 + 1;
                if (bw > 0 && bh > 0) ncand = bw * bh;
            }
        }
    }
    if (__ballot_sync(FULL, ncand > 0) == 0u) return;

    int incl = ncand;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int T = __shfl_sync(FULL, incl, 31);

    // per-face gradient constants kept in registers by the owning lane (backward only)
    float sgn = 1.0f, r0 = 0, r1 = 0, r2 = 0;
    EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
    S.offs[lane] = incl;
    if (ncand > 0) {
        S.ax[lane] = es.ax; S.ay[lane] = es.ay; S.m[lane] = es.m; S.p[lane] = es.p;
        S.n[lane] = es.n; S.q[lane] = es.q; S.k3[lane] = es.k3;
        S.box[lane] = (cx0 - tx0) | ((cy0 - ty0) << 4) | (bw << 8) | (((8192 + bw - 1) / bw) << 13);
        S.excl[lane] = incl - ncand;
        if (!IS_FRAG) {
            const float* pl = P.len3 + bf * 3;
            r0 = __ldg(pl); r1 = __ldg(pl + 1); r2 = __ldg(pl + 2);
            if (__ldg(P.ff + bf) < 0.0f) sgn = -1.0f;                 // diff_fp_for.cu:94-96
            if (lab.x == lab.y) sgn = -sgn;                           // :167-169
        }
        if (IS_FWD) {
            const float mu_in = __ldg(P.mus + lab.x);
            float mu_out = __ldg(P.mus + lab.y);
            if (lab.y == 0) mu_out = -mu_out;                         // :174-175
            S.r0[lane] = r0; S.r1[lane] = r1; S.r2[lane] = r2;
            S.cA[lane] = __fmul_rn(sgn, mu_in);                       // :187
            S.cB[lane] = -__fmul_rn(sgn, mu_out);                     // :190
        }
        if (IS_BWD) {
#pragma unroll
            for (int k = 0; k < 5; ++k) S.mom[k][lane] = 0.0f;
        }
    }
    __syncwarp();

    for (int base = 0; base < T; base += 32) {
        const int j = base + lane;
        const bool act = j < T;
        const int jj = act ? j : T - 1;
        int i = 0;                                   // smallest slot with offs[i] > jj
        if (S.offs[i + 15] <= jj) i += 16;
        if (S.offs[i + 7] <= jj) i += 8;
        if (S.offs[i + 3] <= jj) i += 4;
        if (S.offs[i + 1] <= jj) i += 2;
        if (S.offs[i] <= jj) i += 1;
        const int e = jj - S.excl[i];
        const int box = S.box[i];
        const int w = (box >> 8) & 31;
        const int ly = (e * (box >> 13)) >> 13;
        const int lx = e - ly * w;
        const int px = (box & 15) + lx, py = ((box >> 4) & 15) + ly;
        const int pl = py * TILE + px;
        EdgeSetup c;
        c.ax = S.ax[i]; c.ay = S.ay[i]; c.m = S.m[i]; c.p = S.p[i]; c.n = S.n[i]; c.q = S.q[i]; c.k3 = S.k3[i];
        float s, t, k1, k2, w0, w1, w2;
        bool in = cover_exact<PREFILTER>(c, (float)(tx0 + px), (float)(ty0 + py), s, t, k1, k2, w0, w1, w2);
        in = in && act;
        if (act) st.cand++;

        if (IS_FWD) {
            float bary = 0.0f, cA = 0.0f, cB = 0.0f;
            if (in) {
                bary = __fmaf_rn(w2, S.r2[i], __fmaf_rn(w0, S.r0[i], __fmul_rn(w1, S.r1[i])));  // :182
                cA = S.cA[i]; cB = S.cB[i];
                if (!(bary > 0.0f)) st.nonpos++;                      // :183 assert -> counter
                st.frag++;
            }
            const unsigned inm = __ballot_sync(FULL, in);
            if (inm) {
                int rank = -1, np = 0;
                if (in) {
                    const unsigned peers = __match_any_sync(inm, pl);  // same pixel, different faces
                    rank = __popc(peers & lanemask_lt());
                    np = __popc(peers);
                }
                const int rounds = __reduce_max_sync(FULL, np);
                for (int r = 0; r < rounds; ++r) {                   // ascending lane == ascending face id
                    if (rank == r) {
                        S.acc[pl] = __fmaf_rn(bary, cB, __fmaf_rn(bary, cA, S.acc[pl]));
                        S.cntv[pl] += 1;                              // :192
                    }
                    __syncwarp();
                }
            }
        } else if (IS_FRAG) {
            if (in && S.cntv[pl] < 0) in = false;                     // invalid pixel (:69-73)
            const unsigned inm = __ballot_sync(FULL, in);
            if (inm) {
                int rank = -1, np = 0;
                if (in) {
                    const unsigned peers = __match_any_sync(inm, pl);
                    rank = __popc(peers & lanemask_lt());
                    np = __popc(peers);
                }
                const int rounds = __reduce_max_sync(FULL, np);
                for (int r = 0; r < rounds; ++r) {
                    if (rank == r) {
                        const int k = S.cntv[pl];
                        S.cntv[pl] = k + 1;
                        const long long pix = ((long long)b * P.H + (ty0 + py)) * P.W + (tx0 + px);
                        const long long slot = (long long)__ldg(P.firstidx + pix) + k;  // :208-209
                        P.imidx[slot] = chunk * CHUNK + i + 1;        // :212
                        P.imwei[3 * slot + 0] = w0;
                        P.imwei[3 * slot + 1] = w1;
                        P.imwei[3 * slot + 2] = w2;                   // :214-216
                    }
                    __syncwarp();
                }
            }
        } else {  // backward: per-face moments of the upstream gradient
            float g = 0.0f;
            if (in) { g = S.acc[pl]; if (g != 0.0f) st.frag++; }
            float v[5] = {g, g * s, g * t, g * k1, g * k2};
            if (__ballot_sync(FULL, g != 0.0f)) {
                seg_reduce5(i, lane, v);
                const int iprev = __shfl_up_sync(FULL, i, 1);
                if (lane == 0 || iprev != i) {        // run head: distinct slots, no conflicts
#pragma unroll
                    for (int k = 0; k < 5; ++k) S.mom[k][i] += v[k];
                }
                __syncwarp();
            }
        }
    }

    if (IS_BWD && ncand > 0) {
        // per-face gradients from the five moments of this tile's fragments (face_grads_from_moments)
        const float M0 = S.mom[0][lane], Ms = S.mom[1][lane], Mt = S.mom[2][lane];
        const float A1 = S.mom[3][lane], A2 = S.mom[4][lane];
        if (M0 != 0.0f || Ms != 0.0f || Mt != 0.0f || A1 != 0.0f || A2 != 0.0f) {
            const float mu_in = __ldg(P.mus + lab.x), mu_out = __ldg(P.mus + lab.y);
            const float cgeo = sgn * (mu_in - mu_out);                // raw mu, no label-0 flip (back.cu:115)
            float gp[6], gl[3], bsum;
            face_grads_from_moments(es, r0, r1, r2, cgeo, M0, Ms, Mt, A1, A2, gp, gl, bsum);
            float* dl = P.grad_len3 + bf * 3;                         // back.cu:126-136
            atomicAdd(dl + 0, gl[0]); atomicAdd(dl + 1, gl[1]); atomicAdd(dl + 2, gl[2]);
            // grad_mus (back.cu:143-147): sign*(+-)bary*g for the inward label, opposite for the outward
            atomicAdd(gmus + lab.x, (lab.x == 0 ? -sgn : sgn) * bsum);
            atomicAdd(gmus + lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
            float* dp = P.grad_p6 + bf * 6;                           // back.cu:285-292
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(dp + k, gp[k]);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// tile load / store helpers (one warp, 16x16 pixels, rows of 64 B)
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void tile_store(T* __restrict__ dst, const T* src_smem, int b, int H, int W,
                                           int tx0, int ty0, int tx1, int ty1, int lane) {
    const long long base = (long long)b * H * W;
    if (tx1 - tx0 == TILE - 1 && (W & 3) == 0) {
        // 4 lanes x 16 B per row, 8 rows per pass
        const int c4 = (lane & 3) * 4, r = lane >> 2;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int row = r + pass * 8;
            if (ty0 + row <= ty1) {
                const int4 v = *reinterpret_cast<const int4*>(src_smem + row * TILE + c4);
                *reinterpret_cast<int4*>(dst + base + (long long)(ty0 + row) * W + tx0 + c4) = v;
            }
        }
    } else {
        const int col = lane & 15, r = lane >> 4;
        for (int row = r; row < TILE; row += 2)
            if (ty0 + row <= ty1 && tx0 + col <= tx1)
                dst[base + (long long)(ty0 + row) * W + tx0 + col] = src_smem[row * TILE + col];
    }
}

template <typename T>
__device__ __forceinline__ void tile_load(T* dst_smem, const T* __restrict__ src, T fill, int b, int H, int W,
                                          int tx0, int ty0, int tx1, int ty1, int lane) {
    const long long base = (long long)b * H * W;
    const int col = lane & 15, r = lane >> 4;
    for (int row = r; row < TILE; row += 2) {
        T v = fill;
        if (ty0 + row <= ty1 && tx0 + col <= tx1) v = src[base + (long long)(ty0 + row) * W + tx0 + col];
        dst_smem[row * TILE + col] = v;
    }
}

// ------------------------------------------------------------------------------------------
// the raster kernel: one CTA per (angle, region)
// ------------------------------------------------------------------------------------------
template <int MODE, bool PREFILTER>
__global__ void __launch_bounds__(THREADS)
k_raster(const RasterParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaSmem& C = *reinterpret_cast<CtaSmem*>(smem_raw);
    float* gmus = reinterpret_cast<float*>(smem_raw + sizeof(CtaSmem));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int regions = P.regions_x * P.regions_y;
    const int b = (int)(blockIdx.x / (unsigned)regions);
    const int rr = (int)(blockIdx.x % (unsigned)regions);
    const int rx0 = (rr % P.regions_x) * P.RT * TILE, ry0 = (rr / P.regions_x) * P.RT * TILE;
    const int rx1 = min(rx0 + P.RT * TILE, P.W) - 1, ry1 = min(ry0 + P.RT * TILE, P.H) - 1;
    const short4* boxes = P.chunk_box + (long long)b * P.NC;
    WarpSmem& S = C.warp[warp];
    ThreadStats st = {0u, 0u, 0u};

    constexpr bool K_BWD = (MODE == MODE_BWD || MODE == MODE_STEP);
    double loss_acc = 0.0;
    unsigned nvalid_acc = 0u;
    if (K_BWD) {
        for (int k = tid; k < P.n_mus; k += THREADS) gmus[k] = 0.0f;
    }

    int c = 0;
    int seg = 0;
    do {
        // ---- phase A: ordered compaction of the chunks overlapping this region -------------
        if (tid == 0) { C.list_n = 0; C.next_tile = 0; }
        __syncthreads();
        while (c < P.NC) {
            const int ci = c + tid;
            bool hit = false;
            short4 bx = make_short4(0, 0, 0, 0);
            if (ci < P.NC) { bx = __ldg(boxes + ci); hit = box_overlap(bx, rx0, ry0, rx1, ry1); }
            const unsigned m = __ballot_sync(FULL, hit);
            if (lane == 0) C.wsum[warp] = __popc(m);
            __syncthreads();
            int total = 0, mybase = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const int v = C.wsum[w]; if (w < warp) mybase += v; total += v; }
            const int n0 = C.list_n;
            if (n0 + total > LIST_CAP) break;              // uniform: segment is full, resume at c later
            if (hit) {
                const int pos = n0 + mybase + __popc(m & lanemask_lt());
                C.list_id[pos] = ci; C.list_box[pos] = bx;
            }
            __syncthreads();
            if (tid == 0) C.list_n = n0 + total;
            c += THREADS;
        }
        __syncthreads();
        const int nlist = C.list_n;
        const bool first = (seg == 0), last = (c >= P.NC);

        if (nlist == 0 && first && last) {
            // nothing projects into this region: zero-fill (forward); the residual step still owes
            // the data term of its (all-zero, hence valid) pixels; nothing to do otherwise
            const int rw = rx1 - rx0 + 1, rh = ry1 - ry0 + 1;
            const long long base = (long long)b * P.H * P.W;
            if (MODE == MODE_FWD) {
                if ((P.W & 3) == 0 && (rw & 3) == 0) {
                    const int rw4 = rw >> 2;
                    for (int k = tid; k < rw4 * rh; k += THREADS) {
                        const int y = k / rw4, x = (k - y * rw4) * 4;
                        const long long o = base + (long long)(ry0 + y) * P.W + rx0 + x;
                        *reinterpret_cast<float4*>(P.im + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (P.cnt) *reinterpret_cast<int4*>(P.cnt + o) = make_int4(0, 0, 0, 0);
                    }
                } else {
                    for (int k = tid; k < rw * rh; k += THREADS) {
                        const int y = k / rw, x = k - y * rw;
                        const long long o = base + (long long)(ry0 + y) * P.W + rx0 + x;
                        P.im[o] = 0.0f;
                        if (P.cnt) P.cnt[o] = 0;
                    }
                }
            } else if (MODE == MODE_STEP) {
                if (0.0f <= P.mask_hi) {
                    for (int k = tid; k < rw * rh; k += THREADS) {
                        const int y = k / rw, x = k - y * rw;
                        const float t = __ldg(P.target + base + (long long)(ry0 + y) * P.W + rx0 + x);
                        loss_acc += (double)t * (double)t;
                        nvalid_acc++;
                    }
                }
            }
            break;
        }

        // ---- phase B: warps pull tiles of the region ---------------------------------------
        const int ntiles = P.RT * P.RT;
        while (true) {
            int tix = 0;
            if (lane == 0) tix = atomicAdd(&C.next_tile, 1);
            tix = __shfl_sync(FULL, tix, 0);
            if (tix >= ntiles) break;
            const int tx0 = rx0 + (tix % P.RT) * TILE, ty0 = ry0 + (tix / P.RT) * TILE;
            if (tx0 > rx1 || ty0 > ry1) continue;
            const int tx1 = min(tx0 + TILE - 1, rx1), ty1 = min(ty0 + TILE - 1, ry1);

            // tile prologue
            if (MODE == MODE_FWD) {
                if (first) {
#pragma unroll
                    for (int k = 0; k < TILE_PIX / 32; ++k) { S.acc[lane + 32 * k] = 0.0f; S.cntv[lane + 32 * k] = 0; }
                } else {
                    tile_load<float>(S.acc, P.im, 0.0f, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                    if (P.cnt) tile_load<int>(S.cntv, P.cnt, 0, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                }
            } else if (MODE == MODE_FRAG) {
                // cursor = fragments already emitted for the pixel; -1 marks pixels the reference
                // skips in pass 2 (diff_fp_for.cu:69-73)
                tile_load<float>(S.acc, P.im_in, 0.0f, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                if (!first) tile_load<int>(S.cntv, P.cnt, 0, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < TILE_PIX / 32; ++k) {
                    const float v = S.acc[lane + 32 * k];
                    const bool bad = ((double)v < -1e-5) || (v > P.msv);
                    if (bad) S.cntv[lane + 32 * k] = -1;
                    else if (first) S.cntv[lane + 32 * k] = 0;
                }
            } else if (MODE == MODE_BWD) {
                tile_load<float>(S.acc, P.grad_im, 0.0f, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                // pixels the reference leaves without fragments (pass-2 skip) carry no gradient
                const long long base = (long long)b * P.H * P.W;
                const int col = lane & 15, r = lane >> 4;
                for (int row = r; row < TILE; row += 2) {
                    if (ty0 + row <= ty1 && tx0 + col <= tx1) {
                        const float v = __ldg(P.im_in + base + (long long)(ty0 + row) * P.W + tx0 + col);
                        if (((double)v < -1e-5) || (v > P.msv)) S.acc[row * TILE + col] = 0.0f;
                    }
                }
            } else if (MODE == MODE_STEP) {
                // residual prologue: g = d/dphat of sum_valid (target - phat)^2 = 2*(phat - target) on
                // pixels passing get_valid_mask (fp_vecs.py:10-13; a subset of the kernel's own
                // gradient predicate, diff_fp_for.cu:69-73); loss and count once per tile
                const long long base = (long long)b * P.H * P.W;
                const int col = lane & 15, r = lane >> 4;
                for (int row = r; row < TILE; row += 2) {
                    float g = 0.0f;
                    if (ty0 + row <= ty1 && tx0 + col <= tx1) {
                        const long long o = base + (long long)(ty0 + row) * P.W + tx0 + col;
                        const float v = __ldg(P.im_in + o);
                        if (v > -1e-6f && v <= P.mask_hi) {
                            const float diff = v - __ldg(P.target + o);
                            g = 2.0f * diff;
                            if (first) { loss_acc += (double)diff * (double)diff; nvalid_acc++; }
                        }
                    }
                    S.acc[row * TILE + col] = g;
                }
            }
            __syncwarp();

            // walk the region's chunk list, 32 boxes per ballot, in face order
            for (int lb = 0; lb < nlist; lb += 32) {
                const int e = lb + lane;
                bool hit = false;
                int cid = 0;
                if (e < nlist) { hit = box_overlap(C.list_box[e], tx0, ty0, tx1, ty1); cid = C.list_id[e]; }
                unsigned m = __ballot_sync(FULL, hit);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int chunk = __shfl_sync(FULL, cid, src);
                    process_chunk<MODE, PREFILTER>(P, S, gmus, b, chunk, tx0, ty0, tx1, ty1, lane, st);
                }
            }

            // tile epilogue
            if (MODE == MODE_FWD) {
                tile_store<float>(P.im, S.acc, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                if (P.cnt) tile_store<int>(P.cnt, S.cntv, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
            } else if (MODE == MODE_FRAG) {
                if (!last) {   // persist cursors between list segments (cnt doubles as scratch)
#pragma unroll
                    for (int k = 0; k < TILE_PIX / 32; ++k) if (S.cntv[lane + 32 * k] < 0) S.cntv[lane + 32 * k] = 0;
                    __syncwarp();
                    tile_store<int>(const_cast<int*>(P.cnt), S.cntv, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                }
            }
            __syncwarp();
        }
        __syncthreads();
        ++seg;
    } while (c < P.NC);

    if (K_BWD) {
        __syncthreads();
        for (int k = tid; k < P.n_mus; k += THREADS) {
            const float v = gmus[k];
            if (v != 0.0f) atomicAdd(P.grad_mus + k, v);
        }
    }
    if (MODE == MODE_STEP) {
        // masked L2 data term of ctdr/optimize.py:168: sum over valid pixels and their count
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) loss_acc += __shfl_down_sync(FULL, loss_acc, d);
        nvalid_acc = __reduce_add_sync(FULL, nvalid_acc);
        if (lane == 0 && nvalid_acc) {
            atomicAdd(P.out2 + 0, loss_acc);
            atomicAdd(P.out2 + 1, (double)nvalid_acc);
        }
    }
    if (P.stats) {
        const unsigned cand = __reduce_add_sync(FULL, st.cand);
        const unsigned frag = __reduce_add_sync(FULL, st.frag);
        const unsigned nonpos = __reduce_add_sync(FULL, st.nonpos);
        if (lane == 0) {
            if (nonpos) atomicAdd(P.stats + 0, (unsigned long long)nonpos);
            if (frag) atomicAdd(P.stats + 1, (unsigned long long)frag);
            if (cand) atomicAdd(P.stats + 2, (unsigned long long)cand);
            if (warp == 0 && seg > 1) atomicAdd(P.stats + 3, (unsigned long long)(seg - 1));
        }
    }
}

}  // namespace ctdr
