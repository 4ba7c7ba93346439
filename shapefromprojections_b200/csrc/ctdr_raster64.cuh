// ctdr_raster64.cuh — the float64 instantiation of the rasteriser (forward pass 1, backward).
//
// The reference dispatches its kernels over float AND double (AT_DISPATCH_FLOATING_TYPES,
// diff_fp_for.cu:251, diff_fp_back.cu:326; Model(..., dtype) vanilla.py:27,31).  The double instantiation is the
// high-precision referee of the gradient tolerance, not the production path, so this file
// trades the fp32 kernels' machinery (row spans, candidate redistribution, hit masks, batch records) for a short
// kernel that is still tile-binned and still accumulates every pixel in ascending face order:
//
//   same decomposition and binning as k_raster (regions of RT x RT tiles, ordered chunk lists, per-tile face
//   queue drained 32 faces at a time), boxes computed from the double coordinates;
//   a batch is staged in shared memory (12 doubles per face) and then walked PIXEL-PARALLEL: a lane owns eight
//   pixels of the 16x16 tile and keeps their running sums in registers; the batch's faces are visited one after
//   the other (warp-uniform loop), each lane tests its pixels against the face's box and evaluates the
//   reference's coverage arithmetic in double (diff_fp_for.cu:126-151, IEEE division) for those inside.
//   Forward: acc = fma(bary, cB, fma(bary, cA, acc)) per covering face in ascending order — the reference's own
//   per-pixel loop (:83) restricted to the tile's faces; one coalesced store per pixel, no atomics.
//   Backward: per face the three moments (sum g, sum g*k1, sum g*k2) are reduced over the warp, the per-face
//   gradient formulas of face_grads_from_moments are evaluated in double with exact divisions, and nine
//   atomicAdd(double) per (tile, face) go to grad_p6 / grad_len3.
#pragma once
#include "ctdr_raster.cuh"

namespace ctdr {

struct Raster64Params {
    const double* p6;       // [B,F,6]
    const double* ff;       // [B,F]
    const double* len3;     // [B,F,3]
    const int*    labels2;  // [F,2]
    const double* mus;      // [n_mus]
    int n_mus;
    int B, F, H, W;
    double msv;
    int NC, RT, regions_x, regions_y;
    const short4* chunk_box;
    const short4* face_box;
    const int* slist;
    const int* scount;
    int nsr_x, nsr_y, sr_rx, sr_ry;
    int region_major;
    double* im;             // forward [B,H,W]
    int*    cnt;            // forward [B,H,W] or null
    const double* grad_im;  // backward
    const double* im_in;
    double* grad_p6;
    double* grad_len3;
    double* grad_mus;
    unsigned long long* stats;   // [0] non-positive lengths, [1] fragments, [2] candidates
};

// pixel boxes from double coordinates: the reference's whole-face reject and candidate rule in double
// (diff_fp_for.cu:105-111), NaN handling as face_int_bbox
__device__ __forceinline__ bool face_int_bbox64(double ax, double ay, double bx, double by, double cx, double cy,
                                                int H, int W, int& ix0, int& iy0, int& ix1, int& iy1) {
    const bool nanx = (ax != ax) | (bx != bx) | (cx != cx);
    const bool nany = (ay != ay) | (by != by) | (cy != cy);
    const double xmin = fmin(fmin(ax, bx), cx), xmax = fmax(fmax(ax, bx), cx);
    const double ymin = fmin(fmin(ay, by), cy), ymax = fmax(fmax(ay, by), cy);
    bool reject = false;
    if (!nanx) reject |= (xmin < 0.0) | (xmax > (double)W - 1.0);
    if (!nany) reject |= (ymin < 0.0) | (ymax > (double)H - 1.0);
    if (reject) return false;
    if (nanx) { ix0 = 0; ix1 = W - 1; } else { ix0 = (int)ceil(xmin); ix1 = (int)floor(xmax); }
    if (nany) { iy0 = 0; iy1 = H - 1; } else { iy0 = (int)ceil(ymin); iy1 = (int)floor(ymax); }
    return (ix0 <= ix1) & (iy0 <= iy1);
}

__global__ void __launch_bounds__(THREADS)
k_chunk_bbox64(const double* __restrict__ p6, const int* __restrict__ labels2,
               int B, int F, int H, int W, int NC, short4* __restrict__ chunk_box, short4* __restrict__ face_box) {
    const long long gw = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * NC) return;
    const int b = (int)(gw / NC), c = (int)(gw % NC);
    const int f = c * CHUNK + lane;
    int x0 = BOX_EMPTY_LO, y0 = BOX_EMPTY_LO, x1 = -1, y1 = -1;
    if (f < F) {
        const int2 lab = __ldg(reinterpret_cast<const int2*>(labels2) + f);
        if (lab.x >= lab.y) {
            const double* pp = p6 + ((long long)b * F + f) * 6;
            int ix0, iy0, ix1, iy1;
            if (face_int_bbox64(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2), __ldg(pp + 3), __ldg(pp + 4), __ldg(pp + 5), H, W, ix0, iy0, ix1, iy1)) {
                x0 = ix0; y0 = iy0; x1 = ix1; y1 = iy1;
            }
        }
        face_box[(long long)b * F + f] = make_short4((short)x0, (short)y0, (short)x1, (short)y1);
    }
    x0 = __reduce_min_sync(FULL, x0); y0 = __reduce_min_sync(FULL, y0);
    x1 = __reduce_max_sync(FULL, x1); y1 = __reduce_max_sync(FULL, y1);
    if (lane == 0) chunk_box[gw] = make_short4((short)x0, (short)y0, (short)x1, (short)y1);
}

constexpr int SLOT64 = 12;      // doubles per staged face: ax ay m p n q d r0 r1 r2 cA cB  (backward: cA = k3, cB unused)
static_assert(32 * SLOT64 * sizeof(double) <= 32 * SLOT_VEC * sizeof(float4) + 32 * TILE, "staged double faces overlay slot + span");

// one batch of <= 32 queued faces on the warp's tile; acc / cntv / g: the lane's eight pixels (pixel k of the lane =
// tile row 2k + (lane >> 4), column lane & 15)
template <bool BWD>
__device__ __forceinline__ void process_batch64(const Raster64Params& P, WarpSmem& S, double* gmus, int b, int nb,
                                                int tx0, int ty0, int tx1, int ty1, int lane,
                                                double (&acc)[8], int (&cntv)[8], unsigned& nonpos, unsigned& nfrag, unsigned& ncand) {
    double* slot = reinterpret_cast<double*>(S.slot);
    int* boxes = S.cntv;                     // per staged face: in-tile box, x0 | y0 << 4 | x1 << 8 | y1 << 12
    const bool queued = lane < nb;
    const int f = queued ? S.qface[lane] : 0;
    const long long bf = (long long)b * P.F + f;
    double sgn = 1.0, r0 = 0, r1 = 0, r2 = 0, m = 0, p = 0, n = 0, q = 0, k3 = 0;
    int2 lab = make_int2(0, 0);
    if (queued) {
        const short4 fb = __ldg(P.face_box + bf);
        lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
        const double* pp = P.p6 + bf * 6;
        const double ax = __ldg(pp), ay = __ldg(pp + 1), bx = __ldg(pp + 2), by = __ldg(pp + 3), cx = __ldg(pp + 4), cy = __ldg(pp + 5);
        m = bx - ax; p = by - ay; n = cx - ax; q = cy - ay;                   // diff_fp_for.cu:126-130
        k3 = __fma_rn(m, q, -__dmul_rn(n, p));                                // :142 as compiled (mul + fma)
        r0 = __ldg(P.len3 + bf * 3); r1 = __ldg(P.len3 + bf * 3 + 1); r2 = __ldg(P.len3 + bf * 3 + 2);
        if (__ldg(P.ff + bf) < 0.0) sgn = -1.0;                               // :94-96
        if (lab.x == lab.y) sgn = -sgn;                                       // :167-169
        double* sl = slot + lane * SLOT64;
        sl[0] = ax; sl[1] = ay; sl[2] = m; sl[3] = p; sl[4] = n; sl[5] = q;
        sl[6] = k3 + 1e-15;                                                   // :144 denominator
        sl[7] = r0; sl[8] = r1; sl[9] = r2;
        if (!BWD) {
            const double mu_in = __ldg(P.mus + lab.x);
            double mu_out = __ldg(P.mus + lab.y);
            if (lab.y == 0) mu_out = -mu_out;                                 // :174-175
            sl[10] = __dmul_rn(sgn, mu_in);                                   // :187
            sl[11] = -__dmul_rn(sgn, mu_out);                                 // :190
        }
        const int cx0 = max((int)fb.x, tx0) - tx0, cy0 = max((int)fb.y, ty0) - ty0;
        const int cx1 = min((int)fb.z, tx1) - tx0, cy1 = min((int)fb.w, ty1) - ty0;
        boxes[lane] = cx0 | (cy0 << 4) | (cx1 << 8) | (cy1 << 12);
    }
    __syncwarp();
    const int col = lane & 15, rbase = lane >> 4;
    const double x0 = (double)(tx0 + col);
    double fm0 = 0.0, fa1 = 0.0, fa2 = 0.0;     // backward: the moments of this lane's own face (slot == lane)
#pragma unroll 1
    for (int i = 0; i < nb; ++i) {
        const int bx = boxes[i];
        const int bx0 = bx & 15, by0 = (bx >> 4) & 15, bx1 = (bx >> 8) & 15, by1 = (bx >> 12) & 15;
        double M0 = 0.0, A1 = 0.0, A2 = 0.0;
        if (col >= bx0 && col <= bx1) {
            const double* sl = slot + i * SLOT64;
            const double ax = sl[0], ay = sl[1], fm = sl[2], fp = sl[3], fn = sl[4], fq = sl[5], d = sl[6];
            const double s = x0 - ax;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = 2 * k + rbase;
                if (row < by0 || row > by1) continue;
                ++ncand;
                const double t = (double)(ty0 + row) - ay;
                const double k1 = __fma_rn(s, fq, -__dmul_rn(fn, t));         // :140-141 as compiled
                const double k2 = __fma_rn(fm, t, -__dmul_rn(s, fp));
                const double w1 = k1 / d, w2 = k2 / d;                        // :144-145
                const double w0 = (1.0 - w1) - w2;                            // :146
                if (w0 < 0.0 || w1 < 0.0 || w2 < 0.0) continue;               // :149-151 (NaN counts as inside)
                ++nfrag;
                if (!BWD) {
                    const double bary = __fma_rn(w2, sl[9], __fma_rn(w0, sl[7], __dmul_rn(w1, sl[8])));   // :182
                    if (!(bary > 0.0)) ++nonpos;                              // :183 assert -> counter
                    acc[k] = __fma_rn(bary, sl[11], __fma_rn(bary, sl[10], acc[k]));                     // :187,190
                    cntv[k] += 1;                                             // :192
                } else {
                    const double g = acc[k];
                    M0 += g; A1 = fma(g, k1, A1); A2 = fma(g, k2, A2);
                }
            }
        }
        if (BWD) {
            if (__ballot_sync(FULL, M0 != 0.0 || A1 != 0.0 || A2 != 0.0)) {
#pragma unroll
                for (int dlt = 16; dlt > 0; dlt >>= 1) {
                    M0 += __shfl_xor_sync(FULL, M0, dlt); A1 += __shfl_xor_sync(FULL, A1, dlt); A2 += __shfl_xor_sync(FULL, A2, dlt);
                }
                if (lane == i) { fm0 = M0; fa1 = A1; fa2 = A2; }
            }
        }
    }
    if (BWD && queued && (fm0 != 0.0 || fa1 != 0.0 || fa2 != 0.0)) {
        // per-face gradients from the moments, all in double with IEEE divisions (see face_grads_from_moments for
        // the algebra; back.cu:126-147,206-292)
        const double mu_in = __ldg(P.mus + lab.x), mu_out = __ldg(P.mus + lab.y);
        const double cgeo = sgn * (mu_in - mu_out);                           // raw mu, no label-0 flip (back.cu:115)
        const double dd = k3 + 1e-15;
        const double W1 = fa1 / dd, W2 = fa2 / dd, W0 = fm0 - W1 - W2;        // sum g*w_k
        const double bsum = r0 * W0 + r1 * W1 + r2 * W2;
        const double c10 = r1 - r0, c20 = r2 - r0;
        const double G0 = cgeo / (k3 * k3 + 1e-15);                           // back.cu:282-283
        const double d1m = -q * fa1, d2m = p * fa1, d1n = -q * fa2, d2n = p * fa2;
        const double d1p = n * fa1, d2p = -m * fa1, d1q = n * fa2, d2q = -m * fa2;
        const double d1s = q * k3 * fm0, d2s = -p * k3 * fm0, d1t = -n * k3 * fm0, d2t = m * k3 * fm0;
        double* dl = P.grad_len3 + bf * 3;
        atomicAdd(dl + 0, cgeo * W0); atomicAdd(dl + 1, cgeo * W1); atomicAdd(dl + 2, cgeo * W2);
        double* dp = P.grad_p6 + bf * 6;
        atomicAdd(dp + 0, -G0 * (c10 * (d1m + d1n + d1s) + c20 * (d2m + d2n + d2s)));
        atomicAdd(dp + 1, -G0 * (c10 * (d1p + d1q + d1t) + c20 * (d2p + d2q + d2t)));
        atomicAdd(dp + 2, G0 * (c10 * d1m + c20 * d2m)); atomicAdd(dp + 3, G0 * (c10 * d1p + c20 * d2p));
        atomicAdd(dp + 4, G0 * (c10 * d1n + c20 * d2n)); atomicAdd(dp + 5, G0 * (c10 * d1q + c20 * d2q));
        // grad_mus (back.cu:143-147): shared-memory accumulators per CTA
        atomicAdd(gmus + lab.x, (lab.x == 0 ? -sgn : sgn) * bsum);
        atomicAdd(gmus + lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
    }
    __syncwarp();
}

template <bool BWD>
__global__ void __launch_bounds__(THREADS)
k_raster64(const Raster64Params P) {
    __shared__ CtaSmem C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* gmus = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int regions = P.regions_x * P.regions_y;
    const int b = P.region_major ? (int)(blockIdx.x % (unsigned)P.B) : (int)(blockIdx.x / (unsigned)regions);
    const int rr = P.region_major ? (int)(blockIdx.x / (unsigned)P.B) : (int)(blockIdx.x % (unsigned)regions);
    const int rx0 = (rr % P.regions_x) * P.RT * TILE, ry0 = (rr / P.regions_x) * P.RT * TILE;
    const int rx1 = min(rx0 + P.RT * TILE, P.W) - 1, ry1 = min(ry0 + P.RT * TILE, P.H) - 1;
    const short4* boxes = P.chunk_box + (long long)b * P.NC;
    const int srx = (rr % P.regions_x) / P.sr_rx, sry = (rr / P.regions_x) / P.sr_ry;
    const long long sidx = (long long)b * (P.nsr_x * P.nsr_y) + sry * P.nsr_x + srx;
    const int* super = P.slist + sidx * P.NC;
    const int nsuper = __ldg(P.scount + sidx);
    WarpSmem& S = C.warp[warp];
    unsigned nonpos = 0u, nfrag = 0u, ncand = 0u;
    if (BWD) for (int k = tid; k < P.n_mus; k += THREADS) gmus[k] = 0.0;
    const long long base = (long long)b * P.H * P.W;
    const int col = lane & 15, rbase = lane >> 4;
    int c = 0, seg = 0;
    do {
        // ---- phase A: ordered compaction of the chunks overlapping this region (as k_raster) ----
        if (tid == 0) { C.list_n = 0; C.next_tile = 0; }
        __syncthreads();
        while (c < nsuper) {
            const int cpos = c + tid;
            bool hit = false;
            int ci = 0;
            if (cpos < nsuper) {
                ci = __ldg(super + cpos);
                const short4 bx = __ldg(boxes + ci);
                hit = box_overlap(bx, rx0, ry0, rx1, ry1);
                const int c0 = max((int)bx.x - rx0, 0) >> 4, c1 = min(((int)bx.z - rx0) >> 4, 7);
                const int w0 = max((int)bx.y - ry0, 0) >> 4, w1 = min(((int)bx.w - ry0) >> 4, 7);
                ci |= (c0 << 20) | (c1 << 23) | (w0 << 26) | (w1 << 29);
            }
            const unsigned m = __ballot_sync(FULL, hit);
            if (lane == 0) C.wsum[warp] = __popc(m);
            __syncthreads();
            int total = 0, mybase = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) { const int v = C.wsum[w]; if (w < warp) mybase += v; total += v; }
            const int n0 = C.list_n;
            if (n0 + total > LIST_CAP) break;
            if (hit) C.list_id[n0 + mybase + __popc(m & lanemask_lt())] = ci;
            __syncthreads();
            if (tid == 0) C.list_n = n0 + total;
            c += THREADS;
        }
        __syncthreads();
        const int nlist = C.list_n;
        const bool first = (seg == 0), last = (c >= nsuper);
        if (nlist == 0 && first && last) {
            if (!BWD) {             // nothing projects into this region: zero-fill
                const int rw = rx1 - rx0 + 1, rh = ry1 - ry0 + 1;
                for (int k = tid; k < rw * rh; k += THREADS) {
                    const int y = k / rw, x = k - y * rw;
                    const long long o = base + (long long)(ry0 + y) * P.W + rx0 + x;
                    P.im[o] = 0.0;
                    if (P.cnt) P.cnt[o] = 0;
                }
            }
            break;
        }
        // ---- phase B: warps pull tiles of the region ----
        const int ntiles = P.RT * P.RT;
        while (true) {
            int tix = 0;
            if (lane == 0) tix = atomicAdd(&C.next_tile, 1);
            tix = __shfl_sync(FULL, tix, 0);
            if (tix >= ntiles) break;
            const unsigned tcol = (unsigned)(tix % P.RT), trow = (unsigned)(tix / P.RT);
            const int tx0 = rx0 + (int)tcol * TILE, ty0 = ry0 + (int)trow * TILE;
            if (tx0 > rx1 || ty0 > ry1) continue;
            const int tx1 = min(tx0 + TILE - 1, rx1), ty1 = min(ty0 + TILE - 1, ry1);
            // the lane's eight pixels: running sums (forward) / upstream gradient, 0 where none (backward)
            double acc[8];
            int cntv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = 2 * k + rbase;
                const bool ok = (ty0 + row <= ty1) && (tx0 + col <= tx1);
                const long long o = base + (long long)(ty0 + row) * P.W + tx0 + col;
                acc[k] = 0.0; cntv[k] = 0;
                if (!BWD) {
                    if (!first && ok) { acc[k] = P.im[o]; if (P.cnt) cntv[k] = P.cnt[o]; }
                } else if (ok) {
                    const double v = __ldg(P.im_in + o);
                    if (!(v < -1e-5 || v > P.msv)) acc[k] = __ldg(P.grad_im + o);     // diff_fp_for.cu:69-73
                }
            }
            int qn = 0, lb = 0, cid = 0;
            unsigned m = 0u;
            bool exhausted = false;
            const short4* fbrow = P.face_box + (long long)b * P.F;
            while (true) {
                while (qn < 32 && !exhausted) {
                    if (m == 0u) {
                        if (lb >= nlist) { exhausted = true; break; }
                        const int e = lb + lane;
                        bool hit = false;
                        if (e < nlist) {
                            const unsigned w = (unsigned)C.list_id[e];
                            cid = (int)(w & 0xfffffu);
                            hit = (tcol >= ((w >> 20) & 7u)) & (tcol <= ((w >> 23) & 7u)) & (trow >= ((w >> 26) & 7u)) & (trow <= (w >> 29));
                        }
                        m = __ballot_sync(FULL, hit);
                        lb += 32;
                        continue;
                    }
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int f = __shfl_sync(FULL, cid, src) * CHUNK + lane;
                    short4 fb = make_short4(BOX_EMPTY_LO, BOX_EMPTY_LO, -1, -1);
                    if (f < P.F) fb = __ldg(fbrow + f);
                    const bool fh = box_overlap(fb, tx0, ty0, tx1, ty1);
                    const unsigned hm = __ballot_sync(FULL, fh);
                    if (fh) S.qface[qn + __popc(hm & lanemask_lt())] = f;
                    qn += __popc(hm);
                    __syncwarp();
                }
                if (qn == 0) break;
                const int nbatch = min(qn, 32);
                process_batch64<BWD>(P, S, gmus, b, nbatch, tx0, ty0, tx1, ty1, lane, acc, cntv, nonpos, nfrag, ncand);
                const int rest = qn - nbatch;
                int tf = 0;
                if (lane < rest) tf = S.qface[32 + lane];
                __syncwarp();
                if (lane < rest) S.qface[lane] = tf;
                qn = rest;
                __syncwarp();
            }
            if (!BWD) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int row = 2 * k + rbase;
                    if ((ty0 + row <= ty1) && (tx0 + col <= tx1)) {
                        const long long o = base + (long long)(ty0 + row) * P.W + tx0 + col;
                        P.im[o] = acc[k];
                        if (P.cnt) P.cnt[o] = cntv[k];
                    }
                }
            }
        }
        __syncthreads();
        ++seg;
    } while (c < nsuper);
    if (BWD) {
        __syncthreads();
        for (int k = tid; k < P.n_mus; k += THREADS) {
            const double v = gmus[k];
            if (v != 0.0) atomicAdd(P.grad_mus + k, v);
        }
    }
    if (P.stats) {
        nonpos = __reduce_add_sync(FULL, nonpos); nfrag = __reduce_add_sync(FULL, nfrag); ncand = __reduce_add_sync(FULL, ncand);
        if (lane == 0) {
            if (nonpos) atomicAdd(P.stats + 0, (unsigned long long)nonpos);
            if (nfrag) atomicAdd(P.stats + 1, (unsigned long long)nfrag);
            if (ncand) atomicAdd(P.stats + 2, (unsigned long long)ncand);
        }
    }
}

}  // namespace ctdr
