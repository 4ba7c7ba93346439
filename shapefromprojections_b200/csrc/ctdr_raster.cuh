// ctdr_raster.cuh — tile-binned mesh rasteriser: forward (sinogram), fragment list (pass 2),
// backward (moment-reduced gradients), fused forward+residual+backward.  sm_100a, fp32 + the
// reference's fp64 barycentric divide.
//
// Work decomposition:
//   angle b  ->  super-regions (<= 4x4) -> regions of RT x RT tiles (one CTA each) -> 16x16-pixel
//                tiles (one warp each, pulled from a CTA-local counter)
//   faces    ->  chunks of 32 consecutive faces; per (angle, chunk) and per (angle, face) a pixel
//                bounding box is precomputed by k_chunk_bbox; k_super_lists bins chunks coarsely.
//                (the fused projector gets both box arrays from k_face_assemble, ctdr_vertex.cuh)
//   A CTA compacts, IN FACE ORDER, the chunks whose box overlaps its region into a shared-memory
//   list (ballot + popc ranking; each entry = chunk id + the chunk's tile range inside the region);
//   each warp then walks that list for its tile (32 entries per ballot, bit tests only); for every
//   overlapping chunk the lanes test their face's box against the tile and append the hits, still
//   in face order, to a per-warp queue.  Whenever 32 faces are queued the
//   warp stages them in shared memory (edge constants, 1/d in fp64, lengths, coefficients, a
//   conservative pixel span per detector row) and splits the batch's candidate pixels into 32
//   contiguous ranges, one per lane.  Forward: fragments are applied optimistically while a
//   per-pixel hit mask records which faces of the batch touched each pixel; pixels touched by more
//   than one face are restored and recomputed in ascending face order.  Batches, queue and lists are
//   all in face order and a tile is owned by one warp, so every pixel accumulates its faces in
//   exactly the order of the reference's per-pixel loop (diff_fp_for.cu:83) -> bit-identical
//   sinograms, with no global atomics and one coalesced store per pixel.  Backward: per-face
//   moments in registers, merged across lanes by a segmented shuffle scan, nine RED per (tile, face).
//   Variants (template parameters of k_raster): EXACT_DIV (literal fp64 divide, debug), DIRECT
//   (fine meshes: batches of small boxes are walked face-per-lane), STATS (counters compiled in
//   only when the caller passes a stats buffer).
#pragma once
#include "ctdr_common.cuh"

namespace ctdr {

#ifdef CTDR_PROBE_ROWSTATS
__device__ unsigned long long g_rowstats[4];     // diagnostic build only (tools/probe_rowstats.py)
#endif
enum RasterMode { MODE_FWD = 0, MODE_FRAG = 1, MODE_BWD = 2, MODE_STEP = 3, MODE_MULTI = 4 };
constexpr int MAX_OUT = 4;       // sinograms per MODE_MULTI pass
// Batch record of the linked sweeps: what the forward sweep found for one batch of <= 32 faces of one tile,
//   [0] next record of the tile or -1, [1] face count | flags << 8, [2 .. 33] face id | (facing < 0) << 31,
//   [36 .. 43] one byte per face: first row of the face in the tile | (rows - 1) << 4,
//   [44 .. 171] 16 bytes per face: the pixels of each TILE row the face covers, first column << 4 | last column
//               (0x10 = none) — the exact coverage decided by the forward's own test, so the residual/backward sweep
//               needs no edge functions, margins or exact re-tests: its fragments are the forward's by construction.
constexpr int BREC_INTS = 172;   // ints per batch record (688 B, a multiple of 16)
constexpr int BREC_INFO = 36, BREC_SPANS = 44;
constexpr int BREC_FLAG_EMPTY = 0x100;   // no face of the batch covers a pixel of the tile: nothing to replay
constexpr int XLIST_CAP = 8192;          // (angle, face, tile) triples per angle window whose coverage is not one interval per row
constexpr unsigned SPAN_NONE4 = 0x10101010u;

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
    const short4* face_box;   // [B,F]  pixel bbox of each face, empty when the face contributes nowhere; bit 15 of .x
                              // carries the facing sign (front_facing < 0, diff_fp_for.cu:94-96): see unpack_box
    // coarse level: per (angle, super-region) ordered list of the chunks overlapping it
    const int* slist;         // [B, nsr, NC]
    const int* scount;        // [B, nsr]
    int nsr_x, nsr_y;         // super-regions per detector edge (<= 4 x 4)
    int sr_rx, sr_ry;         // regions per super-region edge
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
    // fused projector: cotangents go straight to the projected VERTICES instead of the face corners — per angle a
    // [V] array of float4 {d/dpx, d/dpy, d/dlen, -} (six times smaller than grad_p6 + grad_len3 to zero, to RED
    // into and to read back; the scatter over `faces` that autograd's index_backward does, fp_vecs.py:176-186)
    float4* grad_vt;        // [B,V] or null (then grad_p6 / grad_len3 are written)
    const int* faces;       // [F,3] vertex ids, needed with grad_vt / vt
    int V;
    // fused projector, gather mode: no [B,F] corner arrays at all — a face's projected corners and lengths are
    // gathered from the per-(angle, vertex) records {px, py, len, -} through `faces` (p6 / len3 / ff are null then;
    // the gathers fp_vecs.py:176-186 materialises as [B,F,3,*] tensors)
    const float4* vt;       // [B,V] or null
    // MODE_STEP
    const float* target;    // [B,H,W]
    float mask_hi;          // float32(max_sino_val - 1e-6), get_valid_mask upper bound (fp_vecs.py:13)
    double* out2;           // {loss_sum, n_valid}
    // MODE_MULTI: n_out sinograms in one pass; mus is [n_out, n_mus], im is n_out slabs out_stride apart
    int n_out;
    long long out_stride;   // elements between consecutive output sinograms
    double* gram;           // [n_out*n_out + n_out]: sum p_i*p_j (i <= j) and sum p_i*target, or null
    unsigned long long* stats;  // may be null
    // batch records: the forward sweep of the residual step writes, per tile, the chain of face batches it
    // processed; the residual/backward sweep replays them instead of rebuilding region lists and queues
    int    bmode;           // 0 off, 1 record (MODE_FWD), 2 replay (MODE_STEP)
    int*   brec;            // [B][bcap][BREC_INTS]: {next record or -1, face count, 32 face ids}
    int*   bhead;           // [B, tiles]: first record of the tile, -1 when no face reached it
    int*   btail;           // [B, tiles]: last record so far (only regions whose list takes several segments)
    unsigned char* bflag;   // [B, regions]: 0 = not recorded (fall back to the walk), 1 = recorded, 2 = empty region
    int*   bcursor;         // [B] next free record of each angle's slice
    // Faces whose coverage on some row of a tile is not one interval (faces without a usable area whose pixel centres
    // fall erratically on their line: a few per million batches) cannot be recorded as spans: the forward sweep lists
    // them here and leaves their rows empty in the record; k_step_facelist handles them pixel by pixel.
    int4*  xlist;           // [XLIST_CAP] {angle, face, tile x0, tile y0}
    int*   xcount;          // entries appended (may exceed XLIST_CAP: the region is then left to the walking kernel)
    int    bcap;            // records per angle
    int rts;                // k_step_replay: tiles per edge of ITS regions, a multiple of RT (its CTAs amortise their wind-down over
                            // more tiles than the forward's do; the records are per tile, only the region flags are per forward region)
    int region_major;       // CTA order, see k_raster
    int direct_max_box;     // batches whose largest in-tile face box has at most this many pixels take the direct walk
};

// Projected corners (and ray lengths) of face f on window angle b: from the Rasterizer-boundary arrays, or —
// fused projector, P.vt != null — gathered from the per-(angle, vertex) records through the face's vertex ids.
template <bool GATHER>
__device__ __forceinline__ void load_face(const RasterParams& P, int b, int f, long long bf,
                                          float& ax, float& ay, float& bx, float& by, float& cx, float& cy,
                                          float& r0, float& r1, float& r2) {
    if (GATHER) {
        const int* fi = P.faces + 3 * (long long)f;
        const int i0 = __ldg(fi), i1 = __ldg(fi + 1), i2 = __ldg(fi + 2);
        const float4* row = P.vt + (long long)b * P.V;
        const float4 a = __ldg(row + i0), bb = __ldg(row + i1), c = __ldg(row + i2);
        ax = a.x; ay = a.y; bx = bb.x; by = bb.y; cx = c.x; cy = c.y; r0 = a.z; r1 = bb.z; r2 = c.z;
    } else {
        const float2* pp = reinterpret_cast<const float2*>(P.p6 + bf * 6);
        const float2 va = __ldg(pp), vb = __ldg(pp + 1), vc = __ldg(pp + 2);
        const float* pl = P.len3 + bf * 3;
        ax = va.x; ay = va.y; bx = vb.x; by = vb.y; cx = vc.x; cy = vc.y;
        r0 = __ldg(pl); r1 = __ldg(pl + 1); r2 = __ldg(pl + 2);
    }
}
template <bool GATHER>
__device__ __forceinline__ void load_face_xy(const RasterParams& P, int b, int f, long long bf,
                                             float& ax, float& ay, float& bx, float& by, float& cx, float& cy) {
    if (GATHER) {
        const int* fi = P.faces + 3 * (long long)f;
        const int i0 = __ldg(fi), i1 = __ldg(fi + 1), i2 = __ldg(fi + 2);
        const float2* row = reinterpret_cast<const float2*>(P.vt + (long long)b * P.V);
        const float2 a = __ldg(row + 2 * i0), bb = __ldg(row + 2 * i1), c = __ldg(row + 2 * i2);
        ax = a.x; ay = a.y; bx = bb.x; by = bb.y; cx = c.x; cy = c.y;
    } else {
        const float2* pp = reinterpret_cast<const float2*>(P.p6 + bf * 6);
        const float2 va = __ldg(pp), vb = __ldg(pp + 1), vc = __ldg(pp + 2);
        ax = va.x; ay = va.y; bx = vb.x; by = vb.y; cx = vc.x; cy = vc.y;
    }
}

// ------------------------------------------------------------------------------------------
// per-(angle, chunk) pixel bounding boxes; one warp per chunk
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
k_chunk_bbox(const float* __restrict__ p6, const float* __restrict__ ff, const int* __restrict__ labels2,
             int B, int F, int H, int W, int NC, short4* __restrict__ chunk_box, short4* __restrict__ face_box) {
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
    if (f < F) face_box[(long long)b * F + f] = pack_box(x0, y0, x1, y1, __ldg(ff + (long long)b * F + f) < 0.0f);
    x0 = __reduce_min_sync(FULL, x0); y0 = __reduce_min_sync(FULL, y0);
    x1 = __reduce_max_sync(FULL, x1); y1 = __reduce_max_sync(FULL, y1);
    if (lane == 0) chunk_box[gw] = make_short4((short)x0, (short)y0, (short)x1, (short)y1);
}

// ------------------------------------------------------------------------------------------
// coarse binning: one CTA per (angle, super-region) compacts, in face order, the chunks whose box
// overlaps the super-region (<= 4 x 4 super-regions per detector, so every chunk box is read 16
// times instead of once per region)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
k_super_lists(const short4* __restrict__ chunk_box, int NC, int H, int W, int nsr_x, int nsr_y,
              int sr_w, int sr_h, int* __restrict__ slist, int* __restrict__ scount) {
    __shared__ int wsum[WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsr = nsr_x * nsr_y;
    const int b = blockIdx.x / nsr, sr = blockIdx.x % nsr;
    const int x0 = (sr % nsr_x) * sr_w, y0 = (sr / nsr_x) * sr_h;
    const int x1 = min(x0 + sr_w, W) - 1, y1 = min(y0 + sr_h, H) - 1;
    const short4* boxes = chunk_box + (long long)b * NC;
    int* out = slist + ((long long)b * nsr + sr) * NC;
    int n = 0;
    for (int c = 0; c < NC; c += THREADS) {
        const int ci = c + tid;
        bool hit = false;
        if (ci < NC) {
            const short4 bx = __ldg(boxes + ci);
            hit = (bx.x <= x1) & (bx.z >= x0) & (bx.y <= y1) & (bx.w >= y0);
        }
        const unsigned m = __ballot_sync(FULL, hit);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int total = 0, mybase = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { const int v = wsum[w]; if (w < warp) mybase += v; total += v; }
        if (hit) out[n + mybase + __popc(m & lanemask_lt())] = ci;
        n += total;
        __syncthreads();
    }
    if (tid == 0) scount[(long long)b * nsr + sr] = n;
}

// CTA order over the regions of an angle.  region_major == 2: centre-out — rows from the detector's middle row outwards,
// and within a row the columns from the middle outwards — so a launch ends on the detector's border regions (few
// faces, short CTAs) wherever the object sits: the tail of a small grid (an angle shard) is then made of light CTAs
// (measured: -0.5 % on 720 angles of config 3, -1.3 % on a 90-angle shard and on the 1M-triangle mesh).
__device__ __forceinline__ int centre_out(int k, int n) {
    const int c = n >> 1;
    return (k & 1) ? c - ((k + 1) >> 1) : c + (k >> 1);
}
__device__ __forceinline__ int region_of_rank(const RasterParams& P, int rank, int regions_x, int regions_y) {
    if (P.region_major != 2) return rank;
    const int ky = rank / regions_x, kx = rank - ky * regions_x;
    return centre_out(ky, regions_y) * regions_x + centre_out(kx, regions_x);
}
__device__ __forceinline__ int region_of_rank(const RasterParams& P, int rank) { return region_of_rank(P, rank, P.regions_x, P.regions_y); }

// ------------------------------------------------------------------------------------------
// shared memory
// ------------------------------------------------------------------------------------------
constexpr int SLOT_VEC = 5;       // float4 per staged face slot (stride 80 B: conflict-free LDS.128)

// CNT = false: no fragment-count tile (the fused projector's forward never asks for counts) — with it the CTA fits
// eight times beside 60 KB of L1 instead of 28 KB (see k_raster)
template <bool CNT>
struct __align__(16) WarpSmemT {
    static constexpr bool HAS_CNT = CNT;
    // Staged chunk, array of structures, nonempty faces compacted to the low slots:
    //   v0 = (ax, ay, m, p)   v1 = (n, q, k3, box)   v2 = (d, 1/d) as two doubles
    //   v3 = (r0, r1, r2, cA) v4 = (cB, excl, labels/sign word, face id)
    // box = queue lane | row mask << 5 | (candidates - 1) << 21; excl = exclusive candidate offset of the
    // slot; cA/cB: forward coefficients +sign*mu_in, -(sign*mu_out') (diff_fp_for.cu:187,190);
    // labels/sign word (MODE_MULTI): label_in | label_out << 12 | (sign < 0) << 24.
    float4 slot[32 * SLOT_VEC];
    // Per-row pixel spans of the staged faces inside the tile: one byte per (slot, row) =
    // first column << 4 | last column (tile-relative); first > last marks an empty row.  Once the batch's candidates
    // are walked the same bytes hold the list of pixels hit by more than one face of the batch (cpix).
    union {
        unsigned char span[32][TILE];
        unsigned char cpix[TILE_PIX];
    };
    union {
        float mom[3][32];        // backward: sum g, sum g*k1, sum g*k2 per slot
        struct {
            unsigned hitm[TILE_PIX];         // forward: which slots of the current batch hit each pixel
        } f;
    };
    // forward: fragment count; MODE_FRAG: CSR cursor, -1 = invalid pixel.  The backward modes of k_raster use its
    // first 256 bytes as the tail of the prefix-sum overlay (prefix_e0) and the next 144 as the record landing zone.
    int   cntv[CNT ? TILE_PIX : 4];
    // queue of faces whose box overlaps the tile, in face order (drained 32 at a time)
    int    qface[64];
    // tile accumulator
    float acc[TILE_PIX];                     // forward: running sinogram; backward: g (0 where no gradient)
    // recording forward sweep: per face (queue lane) the tile rows whose coverage must be re-derived from the hit masks
    // (a candidate inside the conservative span failed, or its lane had no room left to remember it); 16 bits per face
    unsigned susp[16];
};
typedef WarpSmemT<true> WarpSmem;

template <bool CNT>
struct CtaSmemT {
    // chunks overlapping the region, ascending: chunk id (20 bits) | first/last tile column and first/last
    // tile row of its box inside the region (3 bits each), so a tile tests a chunk without touching memory
    int    list_id[LIST_CAP];
    int    wsum[WARPS];
    int    list_n;
    int    next_tile;
    int    rec_fail;              // batch recording ran out of record space somewhere in this region
    WarpSmemT<CNT> warp[WARPS];
    // the dynamic shared memory holds float gmus[n_mus] (backward modes) or the MODE_MULTI extras
};
typedef CtaSmemT<true> CtaSmem;

__device__ __forceinline__ bool box_overlap(const short4 bx, int x0, int y0, int x1, int y1) {
    return (bx.x <= x1) & (bx.z >= x0) & (bx.y <= y1) & (bx.w >= y0);
}

// segmented (by contiguous key) reduction toward the lowest lane of each run
__device__ __forceinline__ void seg_reduce3(int key, int lane, float (&v)[3]) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int kk = __shfl_down_sync(FULL, key, d);
        float u[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) u[k] = __shfl_down_sync(FULL, v[k], d);
        if (lane + d < 32 && kk == key) {
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] += u[k];
        }
    }
}

// Gradients of one face from three moments of the upstream gradient g over the face's fragments:
// M0 = sum g, A1 = sum g*k1, A2 = sum g*k2 (k1, k2 exactly as the forward computes them).
// The per-fragment terms of diff_fp_back.cu:231-243 are  dw1/dn = -t*k3 + p*k1,  dw1/dq = s*k3 - m*k1,
// dw2/dm = t*k3 - q*k2,  dw2/dp = -s*k3 + n*k2; with the identities  s*k3 = m*k1 + n*k2  and
// t*k3 = p*k1 + q*k2  (k1 = s*q - n*t, k2 = m*t - s*p, k3 = m*q - n*p) they reduce to -q*k2, n*k2,
// p*k1, -m*k1: every term is a per-face constant times one of {1, k1, k2}.  A zero-area face
// (k1 = k2 = k3 = 0) therefore yields exact zeros like the reference, before the 1/(k3^2+eps) factor,
// and the cancellation the reference suffers in fp32 (-t*k3 + p*k1) is avoided.
// grad_len (:126-136) and the interpolated length (:143-147) use sum g*w_k = A_k/d.
// cgeo = sign*(mu_in - mu_out): the two label iterations of back.cu:111 merged.
// fp64 reciprocal to ~2^-40 (hardware seed + one Newton step): the backward is specified to a relative
// tolerance, so its three per-face fp64 divisions need not be IEEE-rounded
__device__ __forceinline__ double fast_rcp64(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y * fma(-d, y, 2.0);
}

__device__ __forceinline__ void face_grads_from_moments(const EdgeSetup& es, float r0, float r1, float r2,
                                                        float cgeo, float M0, float A1, float A2,
                                                        float gp[6], float gl[3], float& bsum) {
    const float k3 = es.k3;
    const double dd = (double)k3 + 1e-15;
    const double ydd = fast_rcp64(dd);
    const float W1 = (float)((double)A1 * ydd), W2 = (float)((double)A2 * ydd);
    const float W0 = M0 - W1 - W2;                            // sum g*w_k
    gl[0] = cgeo * W0; gl[1] = cgeo * W1; gl[2] = cgeo * W2;
    bsum = r0 * W0 + r1 * W1 + r2 * W2;
    const float d1m = -es.q * A1,        d2m = es.p * A1;
    const float d1n = -es.q * A2,        d2n = es.p * A2;
    const float d1p = es.n * A1,         d2p = -es.m * A1;
    const float d1q = es.n * A2,         d2q = -es.m * A2;
    const float d1s = es.q * k3 * M0,    d2s = -es.p * k3 * M0;
    const float d1t = -es.n * k3 * M0,   d2t = es.m * k3 * M0;
    const float c10 = r1 - r0, c20 = r2 - r0;                 // back.cu:271-280
    const float G0 = (float)((double)cgeo * fast_rcp64((double)(k3 * k3) + 1e-15));   // back.cu:282-283
    gp[0] = -G0 * (c10 * (d1m + d1n + d1s) + c20 * (d2m + d2n + d2s));          // back.cu:246-258
    gp[1] = -G0 * (c10 * (d1p + d1q + d1t) + c20 * (d2p + d2q + d2t));
    gp[2] = G0 * (c10 * d1m + c20 * d2m); gp[3] = G0 * (c10 * d1p + c20 * d2p);
    gp[4] = G0 * (c10 * d1n + c20 * d2n); gp[5] = G0 * (c10 * d1q + c20 * d2q);
}

// counter that compiles to nothing: the statistics are optional (`stats` may be NULL) and five live
// registers per thread are worth more to the hot loops than to a caller who does not read them
struct NoCount {
    __device__ __forceinline__ NoCount& operator++() { return *this; }
    __device__ __forceinline__ NoCount operator++(int) { return *this; }
    template <typename T> __device__ __forceinline__ NoCount& operator+=(T) { return *this; }
    __device__ __forceinline__ operator unsigned() const { return 0u; }
};
template <bool ON> struct CountType { typedef unsigned int type; };
template <> struct CountType<false> { typedef NoCount type; };

template <bool STATS>
struct ThreadStatsT {
    typedef typename CountType<STATS>::type count_t;
    count_t cand{}, frag{}, nonpos{}, visits{}, hits{};   // visits/hits: batches staged / (tile, chunk) pairs with an overlapping face
    // Attenuation gradient: each lane keeps a running sum per (label_in, label_out) pair and only
    // touches the CTA's shared accumulators when its face's labels change (meshes carry long runs
    // of equal labels), instead of two contended shared-memory atomics per (tile, face).
    int lab_in = 0, lab_out = 0;
    float sum_in = 0.0f, sum_out = 0.0f;
    __device__ __forceinline__ void mus_flush(float* gmus) {
        if (sum_in != 0.0f) shared_add(gmus + lab_in, sum_in);
        if (sum_out != 0.0f) shared_add(gmus + lab_out, sum_out);
        sum_in = 0.0f; sum_out = 0.0f;
    }
    __device__ __forceinline__ void mus_add(float* gmus, int li, float vi, int lo, float vo) {
        if (li != lab_in || lo != lab_out) { mus_flush(gmus); lab_in = li; lab_out = lo; }
        sum_in += vi; sum_out += vo;
    }
};

// ------------------------------------------------------------------------------------------
// one (tile, chunk): stage the chunk's faces, spread candidates over lanes, evaluate coverage
// ------------------------------------------------------------------------------------------
template <int MODE, bool EXACT_DIV, class ThreadStats, class WS>
__device__ __forceinline__ void process_batch(const RasterParams& P, WS& S, float* gmus,
                                              int b, int nb, int tx0, int ty0, int tx1, int ty1,
                                              int lane, ThreadStats& st) {
    constexpr bool IS_FWD = (MODE == MODE_FWD);
    constexpr bool IS_FRAG = (MODE == MODE_FRAG);
    constexpr bool IS_BWD = (MODE == MODE_BWD || MODE == MODE_STEP);
    // lane < nb owns queued face qface[lane]; its box overlaps the tile, so it has >= 1 candidate
    const bool mine = lane < nb;
    const int f = mine ? S.qface[lane] : 0;
    const long long bf = (long long)b * P.F + f;
    int ncand = 0;
    int2 lab = make_int2(0, 0);
    float ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;
    int cx0 = 0, cy0 = 0, bw = 1;
    bool fneg = false;
    float l0 = 0, l1 = 0, l2 = 0;
    if (mine) {
        const short4 fb = unpack_box(__ldg(P.face_box + bf), fneg);
        lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
        load_face<false>(P, b, f, bf, ax, ay, bx, by, cx, cy, l0, l1, l2);
        cx0 = max((int)fb.x, tx0); cy0 = max((int)fb.y, ty0);
        bw = min((int)fb.z, tx1) - cx0 + 1;
        ncand = bw * (min((int)fb.w, ty1) - cy0 + 1);
    }
    __syncwarp();                                       // queue entries consumed: the caller may shift it

    int incl = ncand;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int T = __shfl_sync(FULL, incl, 31);
    const int nslots = nb;
    const int rank = lane;                              // queued faces are dense: slot == lane
    if (lane == 0) { st.visits++; }

    // per-face constants kept in registers by the owning lane (backward epilogue)
    float sgn = 1.0f, r0 = 0, r1 = 0, r2 = 0;
    const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
    if (mine) {
        float4* sl = S.slot + rank * SLOT_VEC;
        const int box = (cx0 - tx0) | ((cy0 - ty0) << 4) | (bw << 8) |
                        (__float2int_ru(8192.0f * __frcp_rn((float)bw)) << 13);
        sl[0] = make_float4(es.ax, es.ay, es.m, es.p);
        sl[1] = make_float4(es.n, es.q, es.k3, __int_as_float(box));
        const double d = (double)es.k3 + 1e-15;                        // diff_fp_for.cu:144 denominator
        reinterpret_cast<double2*>(sl)[2] = make_double2(d, 1.0 / d);
        float cA = 0.0f, cB = 0.0f;
        if (!IS_FRAG) {
            r0 = l0; r1 = l1; r2 = l2;
            if (fneg) sgn = -1.0f;                                    // diff_fp_for.cu:94-96
            if (lab.x == lab.y) sgn = -sgn;                           // :167-169
        }
        if (IS_FWD) {
            const float mu_in = __ldg(P.mus + lab.x);
            float mu_out = __ldg(P.mus + lab.y);
            if (lab.y == 0) mu_out = -mu_out;                         // :174-175
            cA = __fmul_rn(sgn, mu_in);                               // :187
            cB = -__fmul_rn(sgn, mu_out);                             // :190
            sl[3] = make_float4(r0, r1, r2, cA);
        }
        const float ak3 = fabsf(es.k3);
        const float rd = ((ak3 > 1e-6f) & (ak3 < 1e10f)) ? __frcp_rn(es.k3) : 0.0f;
        sl[4] = make_float4(cB, __int_as_float(incl - ncand), rd, __int_as_float(f));
        if (IS_BWD) { S.mom[0][rank] = 0.0f; S.mom[1][rank] = 0.0f; S.mom[2][rank] = 0.0f; }
    }
    __syncwarp();
    // lane r < nslots tracks the first candidate of compacted slot r ("segment heads")
    const int myexcl = (lane < nslots) ? __float_as_int(S.slot[lane * SLOT_VEC + 4].y) : 0x3fffffff;
    int nbefore = 0;                                    // slots starting before `base`

    for (int base = 0; base < T; base += 32) {
        // which slot owns candidate base+lane: count the segment heads at or before this lane
        const unsigned hp = (unsigned)(myexcl - base);
        const unsigned heads = __reduce_or_sync(FULL, hp < 32u ? (1u << hp) : 0u);
        const int i = nbefore + __popc(heads & (lanemask_lt() | (1u << lane))) - 1;
        nbefore += __popc(heads);
        const bool act = base + lane < T;
        const float4* sl = S.slot + i * SLOT_VEC;
        const float4 v0 = sl[0], v1 = sl[1], v4 = sl[4];
        const int e = base + lane - __float_as_int(v4.y);
        const int box = __float_as_int(v1.w);
        const int w = (box >> 8) & 31;
        const int ly = (e * (box >> 13)) >> 13;
        const int lx = e - ly * w;
        const int px = (box & 15) + lx, py = ((box >> 4) & 15) + ly;
        const int pl = (py * TILE + px) & (TILE_PIX - 1);
        // coverage: diff_fp_for.cu:132-151, bit-identical (see cover_exact for the prefilter argument)
        const float s = __fsub_rn((float)(tx0 + px), v0.x), t = __fsub_rn((float)(ty0 + py), v0.y);
        const float k1 = __fmaf_rn(s, v1.y, -__fmul_rn(v1.x, t));
        const float k2 = __fmaf_rn(v0.z, t, -__fmul_rn(s, v0.w));
        bool in = act;
        if (!EXACT_DIV) {
            // Prefilter (see cover_exact in ctdr_common.cuh for the argument): v4.z = 1/k3, or 0 when
            // |k3| is outside (1e-6, 1e10) — then a1 = a2 = 0 and every test below passes.
            // a < -1e-30 => the fp64 quotient is strictly negative and far from underflow => w < 0;
            // the fp32 estimate of w0 = 1 - w1 - w2 is within ~5e-7*(1+|w1|+|w2|) of the reference value.
            const float a1 = k1 * v4.z, a2 = k2 * v4.z;
            if (fminf(a1, a2) < -1e-30f) in = false;
            if ((1.0f - a1) - a2 < -1e-4f * ((1.0f + fabsf(a1)) + fabsf(a2))) in = false;
        }
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
        if (in) {
            const double2 dy = reinterpret_cast<const double2*>(sl)[2];
            if (EXACT_DIV) {
                w1 = (float)((double)k1 / dy.x);
                w2 = (float)((double)k2 / dy.x);
            } else {
                // IEEE quotient from the face's reciprocal: q0 = a*y, r = a - q0*d (exact), q = q0 + r*y
                // (the correction step of the compiler's own fp64 division, with 1/d hoisted per face)
                // A zero numerator keeps q0 = (+-0)*y: the correction step would turn -0 into +0.
                const double a1 = (double)k1, a2 = (double)k2;
                const double q1 = __dmul_rn(a1, dy.y), q2 = __dmul_rn(a2, dy.y);
                w1 = (float)(k1 == 0.0f ? q1 : __fma_rn(__fma_rn(-q1, dy.x, a1), dy.y, q1));
                w2 = (float)(k2 == 0.0f ? q2 : __fma_rn(__fma_rn(-q2, dy.x, a2), dy.y, q2));
            }
            w0 = __fsub_rn(__fsub_rn(1.0f, w1), w2);
            in = !(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f);
        }
        if (act) st.cand++;

        if (IS_FWD) {
            float bary = 0.0f, cA = 0.0f;
            if (in) {
                const float4 v3 = sl[3];
                bary = __fmaf_rn(w2, v3.z, __fmaf_rn(w0, v3.x, __fmul_rn(w1, v3.y)));  // :182
                cA = v3.w;
                if (!(bary > 0.0f)) st.nonpos++;                      // :183 assert -> counter
                st.frag++;
            }
            const unsigned inm = __ballot_sync(FULL, in);
            if (inm) {
                int rk = -1, np = 0;
                if (in) {
                    const unsigned peers = __match_any_sync(inm, pl);  // same pixel, different faces
                    rk = __popc(peers & lanemask_lt());
                    np = __popc(peers);
                }
                const int rounds = __reduce_max_sync(FULL, np);
                for (int r = 0; r < rounds; ++r) {                   // ascending lane == ascending face id
                    if (rk == r) {
                        S.acc[pl] = __fmaf_rn(bary, v4.x, __fmaf_rn(bary, cA, S.acc[pl]));
                        if (P.cnt) S.cntv[pl] += 1;                   // :192
                    }
                    __syncwarp();
                }
            }
        } else if (IS_FRAG) {
            if (in && S.cntv[pl] < 0) in = false;                     // invalid pixel (:69-73)
            const unsigned inm = __ballot_sync(FULL, in);
            if (inm) {
                int rk = -1, np = 0;
                if (in) {
                    const unsigned peers = __match_any_sync(inm, pl);
                    rk = __popc(peers & lanemask_lt());
                    np = __popc(peers);
                }
                const int rounds = __reduce_max_sync(FULL, np);
                for (int r = 0; r < rounds; ++r) {
                    if (rk == r) {
                        const int k = S.cntv[pl];
                        S.cntv[pl] = k + 1;
                        const long long pix = ((long long)b * P.H + (ty0 + py)) * P.W + (tx0 + px);
                        const long long slot = (long long)__ldg(P.firstidx + pix) + k;  // :208-209
                        P.imidx[slot] = __float_as_int(v4.w) + 1;                      // :212
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
            float v[3] = {g, g * k1, g * k2};
            if (__ballot_sync(FULL, g != 0.0f)) {
                seg_reduce3(i, lane, v);
                const int iprev = __shfl_up_sync(FULL, i, 1);
                if (lane == 0 || iprev != i) {        // run head: distinct slots, no conflicts
                    S.mom[0][i] += v[0]; S.mom[1][i] += v[1]; S.mom[2][i] += v[2];
                }
                __syncwarp();
            }
        }
    }

    if (IS_BWD && mine) {
        // per-face gradients from the moments of this tile's fragments (face_grads_from_moments)
        const float M0 = S.mom[0][rank], A1 = S.mom[1][rank], A2 = S.mom[2][rank];
        if (M0 != 0.0f || A1 != 0.0f || A2 != 0.0f) {
            const float mu_in = __ldg(P.mus + lab.x), mu_out = __ldg(P.mus + lab.y);
            const float cgeo = sgn * (mu_in - mu_out);                // raw mu, no label-0 flip (back.cu:115)
            float gp[6], gl[3], bsum;
            face_grads_from_moments(es, r0, r1, r2, cgeo, M0, A1, A2, gp, gl, bsum);
            float* dl = P.grad_len3 + bf * 3;                         // back.cu:126-136
            atomicAdd(dl + 0, gl[0]); atomicAdd(dl + 1, gl[1]); atomicAdd(dl + 2, gl[2]);
            // grad_mus (back.cu:143-147): sign*(+-)bary*g for the inward label, opposite for the outward
            st.mus_add(gmus, lab.x, (lab.x == 0 ? -sgn : sgn) * bsum, lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
            float* dp = P.grad_p6 + bf * 6;                           // back.cu:285-292
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(dp + k, gp[k]);
        }
    }
    __syncwarp();
}


// ------------------------------------------------------------------------------------------
// One batch of <= 32 queued faces, forward and backward modes (the ordered candidate walk above
// is kept for the fragment-list mode).
//
// Candidates (bbox-in-tile pixels of the batch, face-major) are split into 32 CONTIGUOUS ranges,
// one per lane, so every lane walks T/32 (+-1) candidates with a plain x/y increment: perfectly
// balanced whatever the triangle sizes, no per-step collectives, face constants re-read from the
// staged slot (LDS.128, mostly broadcast).
//
// Backward: a lane sums (g, g*k1, g*k2) in registers while it stays on one face and adds them to
// the face's shared-memory moments when it moves on.
//
// Forward: fragments are applied to the tile optimistically, in whatever order the lanes reach
// them, while an 8-bit counter per pixel records how many faces of THIS batch hit it.  A pixel hit
// by exactly one face of the batch has received exactly the reference's update
// (acc = fma(bary, cB, fma(bary, cA, acc)), batches being processed in face order).  Pixels hit by
// several faces of the batch (shared edges under the inclusive rule, silhouette folds) are restored
// from a register snapshot and recomputed pixel-parallel by looping over the batch's faces in
// ascending order — the reference's own per-pixel loop (diff_fp_for.cu:83) restricted to 32 faces.
// The sinogram stays bit-identical.
// ------------------------------------------------------------------------------------------
template <bool EXACT_DIV>
__device__ __forceinline__ bool cover_slot(const float4 v0, const float4 v1, const double2 dy,
                                           float x0, float y0, float& k1, float& k2,
                                           float& w0, float& w1, float& w2) {
    // diff_fp_for.cu:132-151 bit for bit (hoisted quotient as argued in DESIGN.md section 5); no prefilter
    // here: the row spans already leave almost only covered pixels
    const float s = __fsub_rn(x0, v0.x), t = __fsub_rn(y0, v0.y);
    k1 = __fmaf_rn(s, v1.y, -__fmul_rn(v1.x, t));
    k2 = __fmaf_rn(v0.z, t, -__fmul_rn(s, v0.w));
    if (!EXACT_DIV) {
        const double d1 = (double)k1, d2 = (double)k2;
        const double q1 = __dmul_rn(d1, dy.y), q2 = __dmul_rn(d2, dy.y);
        w1 = (float)(k1 == 0.0f ? q1 : __fma_rn(__fma_rn(-q1, dy.x, d1), dy.y, q1));
        w2 = (float)(k2 == 0.0f ? q2 : __fma_rn(__fma_rn(-q2, dy.x, d2), dy.y, q2));
    } else {
        w1 = (float)((double)k1 / dy.x);
        w2 = (float)((double)k2 / dy.x);
    }
    w0 = __fsub_rn(__fsub_rn(1.0f, w1), w2);
    return !(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f);
}

// Pixels hit by several faces of the current batch: recompute them over exactly the faces in their
// hit mask, in ascending face order (kept out of line: rare, and it keeps the hot loop's code small).
template <bool EXACT_DIV, class WS>
__device__ __noinline__ void redo_conflicted(WS& S, int ncp, int tx0, int ty0, int lane) {
#pragma unroll 1
    for (int base = 0; base < ncp; base += 32) {
        const bool have = base + lane < ncp;
        const int pix = have ? (int)S.cpix[base + lane] : 0;
        const int px = pix & 15, py = pix >> 4;
        float acc = S.acc[pix];
        unsigned fm = have ? S.f.hitm[pix] : 0u;
#pragma unroll 1
        while (fm) {                                          // ascending face id (diff_fp_for.cu:83)
            const int fi = __ffs(fm) - 1;
            fm &= fm - 1u;
            const float4* sl = S.slot + fi * SLOT_VEC;
            const float4 v0 = sl[0], v1 = sl[1], v3 = sl[3], v4 = sl[4];
            float k1, k2, w0, w1, w2;
            // the face is known to cover the pixel: only the weights are needed again
            cover_slot<EXACT_DIV>(v0, v1, reinterpret_cast<const double2*>(sl)[2],
                                  (float)(tx0 + px), (float)(ty0 + py), k1, k2, w0, w1, w2);
            const float bary = __fmaf_rn(w2, v3.z, __fmaf_rn(w0, v3.x, __fmul_rn(w1, v3.y)));
            acc = __fmaf_rn(bary, v4.x, __fmaf_rn(bary, v3.w, acc));
        }
        if (have) { S.acc[pix] = acc; S.f.hitm[pix] = 0u; }
    }
}

// 4-byte asynchronous global -> shared copy (LDGSTS): no register holds the value while it is in flight
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Replay of recorded batches (residual/backward sweep): the next record is fetched into a landing zone in
// the part of the forward/backward union the backward does not use (behind the three moment rows), while
// the current batch is being processed.
template <class WS>
__device__ __forceinline__ int* record_landing(WS& S) { return S.cntv + 64; }   // the count tile is unused by the backward modes (its first 256 B: prefix overlay)
template <bool CNT>
__device__ __forceinline__ void prefetch_record(const RasterParams& P, WarpSmemT<CNT>& S, int off, int lane) {
    const int* r = P.brec + (long long)off * BREC_INTS;
    int* L = record_landing(S);
    cp_async4(L + 2 + lane, r + 2 + lane);
    if (lane < 2) cp_async4(L + lane, r + lane);
}

// MODE_MULTI dynamic shared memory after CtaSmem: per-warp normal-equation sums (NGRAM doubles each),
// then per-warp tile accumulators of outputs 2 and 3
constexpr int NGRAM = MAX_OUT * (MAX_OUT + 1) / 2 + MAX_OUT;
__device__ __forceinline__ float* multi_xacc(float* dyn, int warp) {
    return reinterpret_cast<float*>(reinterpret_cast<double*>(dyn) + WARPS * NGRAM) + warp * (MAX_OUT - 2) * TILE_PIX;
}

// MODE_MULTI: accumulator of output k for this warp's tile (k = 0: acc, k = 1: the count array's
// storage, k >= 2: extra dynamic shared memory)
template <class WS>
__device__ __forceinline__ float* multi_acc(WS& S, float* xacc, int k) {
    return k == 0 ? S.acc : (k == 1 ? reinterpret_cast<float*>(S.cntv) : xacc + (k - 2) * TILE_PIX);
}

// MODE_MULTI: every pixel hit by the batch, pixel-parallel, faces in ascending order; the n_out
// sinograms share the coverage test and the interpolated length and differ only in the
// attenuation coefficients, so each of them receives exactly the update sequence of a
// single-output pass with its own mus (diff_fp_for.cu:174-190).  Returns the non-positive-length count.
template <bool EXACT_DIV, class WS>
__device__ __noinline__ unsigned redo_multi(WS& S, float* xacc, const float* __restrict__ mus, int n_mus,
                                            int n_out, int ncp, int tx0, int ty0, int lane) {
    unsigned nonpos = 0u;
#pragma unroll 1
    for (int base = 0; base < ncp; base += 32) {
        const bool have = base + lane < ncp;
        const int pix = have ? (int)S.cpix[base + lane] : 0;
        const int px = pix & 15, py = pix >> 4;
        float a[MAX_OUT];
#pragma unroll
        for (int k = 0; k < MAX_OUT; ++k) a[k] = (k < n_out) ? multi_acc(S, xacc, k)[pix] : 0.0f;
        unsigned fm = have ? S.f.hitm[pix] : 0u;
#pragma unroll 1
        while (fm) {                                          // ascending face id (diff_fp_for.cu:83)
            const int fi = __ffs(fm) - 1;
            fm &= fm - 1u;
            const float4* sl = S.slot + fi * SLOT_VEC;
            const float4 v0 = sl[0], v1 = sl[1], v3 = sl[3], v4 = sl[4];
            float k1, k2, w0, w1, w2;
            cover_slot<EXACT_DIV>(v0, v1, reinterpret_cast<const double2*>(sl)[2],
                                  (float)(tx0 + px), (float)(ty0 + py), k1, k2, w0, w1, w2);
            const float bary = __fmaf_rn(w2, v3.z, __fmaf_rn(w0, v3.x, __fmul_rn(w1, v3.y)));  // :182
            if (!(bary > 0.0f)) nonpos++;                     // :183 assert -> counter
            const int word = __float_as_int(v4.z);
            const int li = word & 0xfff, lo = (word >> 12) & 0xfff;
            const float sgn = (word >> 24) ? -1.0f : 1.0f;
#pragma unroll
            for (int k = 0; k < MAX_OUT; ++k) {
                if (k < n_out) {
                    const float mu_in = __ldg(mus + k * n_mus + li);
                    float mu_out = __ldg(mus + k * n_mus + lo);
                    if (lo == 0) mu_out = -mu_out;                                     // :174-175
                    a[k] = __fmaf_rn(bary, -__fmul_rn(sgn, mu_out), __fmaf_rn(bary, __fmul_rn(sgn, mu_in), a[k]));  // :187,190
                }
            }
        }
        if (have) {
#pragma unroll
            for (int k = 0; k < MAX_OUT; ++k) if (k < n_out) multi_acc(S, xacc, k)[pix] = a[k];
            S.f.hitm[pix] = 0u;
        }
    }
    return nonpos;
}

// Conservative column range of one face on one detector row (absolute pixels, clipped to [x_lo, x_hi]).
// Inside <=> sg*k1 >= 0, sg*k2 >= 0, sg*(k1+k2) <= |d| with sg = sign(d) (diff_fp_for.cu:140-151);
// for a fixed row each is a linear inequality a*s + b(t) >= 0 in s = x - ax with b linear in
// t = y - ay, so the root -b/a is R*t + C with per-face R, C.  Bounds are widened by a margin that
// dominates the rounding of the reference's own k1/k2/w0 evaluation (and of this estimate), so
// every pixel the reference counts lies inside the range; the exact test still decides.
// Near-zero slopes give no bound.
struct SpanSetup {
    // x >= RL_k*t + CL_k and x <= RU_k*t + CU_k for k = 0, 1   (t = y - ay, x absolute).  The three edge
    // coefficients a_j sum to zero, so at most two edges bound the row from the left and at most two from
    // the right; unused slots are (0, -3e38) / (0, +3e38); margins and ax are folded into CL / CU.
    float RL0, CL0, RL1, CL1, RU0, CU0, RU1, CU1;
};

__device__ __forceinline__ SpanSetup span_setup(const EdgeSetup& es, float tmax) {
    SpanSetup sp = {0.0f, -3.0e38f, 0.0f, -3.0e38f, 0.0f, 3.0e38f, 0.0f, 3.0e38f};
    const float sg = es.k3 < 0.0f ? -1.0f : 1.0f;
    // a_j*s + beta_j*t + gamma_j >= 0
    const float a[3] = {sg * es.q, -sg * es.p, -sg * (es.q - es.p)};
    const float beta[3] = {-sg * es.n, sg * es.m, -sg * (es.m - es.n)};
    const float gamma[3] = {0.0f, 0.0f, fabsf(es.k3)};
    bool have_l = false, have_u = false;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool use = fabsf(a[j]) > 1e-4f;
        const float ia = use ? __fdividef(1.0f, a[j]) : 0.0f;       // MUFU.RCP: ~1 ulp, far inside the margin below
        const float R = -beta[j] * ia, C = -gamma[j] * ia;       // root of constraint j on row t: R*t + C
        // rounding of the reference's k1/k2/w0 (~2^-23 of the products involved) and of this estimate
        // (a few ulp of |root| <= |R||t| + |C|), times 10, plus 0.02 px; worst case over the rows used
        const float margin = 0.02f + 8e-6f * (fabsf(C) + fabsf(R) * tmax) + 1e-6f * fabsf(es.ax);
        const bool lower = use && a[j] > 0.0f, upper = use && a[j] < 0.0f;
        const float cl = es.ax + C - margin, cu = es.ax + C + margin;
        if (lower && !have_l) { sp.RL0 = R; sp.CL0 = cl; } else if (lower) { sp.RL1 = R; sp.CL1 = cl; }
        if (upper && !have_u) { sp.RU0 = R; sp.CU0 = cu; } else if (upper) { sp.RU1 = R; sp.CU1 = cu; }
        have_l |= lower; have_u |= upper;
    }
    return sp;
}

__device__ __forceinline__ bool row_span(const SpanSetup& sp, float t, int x_lo, int x_hi, int& xs, int& xe) {
    const float lo = fmaxf(fmaf(sp.RL0, t, sp.CL0), fmaf(sp.RL1, t, sp.CL1));
    const float hi = fminf(fmaf(sp.RU0, t, sp.CU0), fmaf(sp.RU1, t, sp.CU1));
    // float -> int conversions saturate, so far-away bounds clip correctly
    xs = max(x_lo, __float2int_ru(lo));
    xe = min(x_hi, __float2int_rd(hi));
    return xs <= xe;
}

// RECSPAN (recording forward sweep): `rec` = this batch's record (negative: none) receives the exact coverage; a face
// whose coverage on some row is not one interval goes to RasterParams::xlist instead (its rows stay empty in the record)
template <int MODE, bool EXACT_DIV, bool DIRECT, bool GATHER, bool RECSPAN, class ThreadStats, class WS>
__device__ __forceinline__ void process_batch_blocked(const RasterParams& P, WS& S, float* gmus,
                                                      int b, int nb, int tx0, int ty0, int tx1, int ty1,
                                                      int lane, ThreadStats& st, int rec, int* rec_fail) {
    constexpr bool IS_FWD = (MODE == MODE_FWD);
    constexpr bool IS_MULTI = (MODE == MODE_MULTI);          // gmus doubles as the extra accumulators
    const bool queued = lane < nb;
    const int f = queued ? S.qface[lane] : 0;
    const long long bf = (long long)b * P.F + f;
    int ncand = 0;
    int2 lab = make_int2(0, 0);
    float ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;
    int cx0 = 0, cy0 = 0, bw = 1, bh = 1;
    float sgn = 1.0f, r0 = 0, r1 = 0, r2 = 0;
    bool fneg = false;
    if (queued) {
        const short4 fb = unpack_box(__ldg(P.face_box + bf), fneg);
        lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
        load_face<GATHER>(P, b, f, bf, ax, ay, bx, by, cx, cy, r0, r1, r2);   // the lengths are in flight while the spans are computed
        cx0 = max((int)fb.x, tx0); cy0 = max((int)fb.y, ty0);
        bw = min((int)fb.z, tx1) - cx0 + 1;
        bh = min((int)fb.w, ty1) - cy0 + 1;
    }
    __syncwarp();                                       // queue entries consumed
    const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
    if (RECSPAN) {
        *reinterpret_cast<uint4*>(S.span[lane]) = make_uint4(SPAN_NONE4, SPAN_NONE4, SPAN_NONE4, SPAN_NONE4);
        if (lane < 16) S.susp[lane] = 0u;
    }
    bool irregular = false;
    // recording: the (at most two) candidates this lane saw fail the exact test at an END of their row's conservative
    // span, 16 bits each: 1 | queue lane << 1 | tile row << 6 | (right end) << 10; applied to the span bytes after the walk
    // (other lanes still read the spans during it)
    unsigned myfail = 0u;

    if (lane == 0) st.visits++;
    if (fneg) sgn = -1.0f;                                            // diff_fp_for.cu:94-96
    if (lab.x == lab.y) sgn = -sgn;                                   // :167-169
    // degenerate / NaN faces keep their bounding-box rows (measured: spans pay off even on 2x2 boxes)
    const bool tight = (fabsf(es.k3) > 1e-6f) && (fabsf(es.k3) < 1e10f);
    const SpanSetup sps = span_setup(es, fmaxf(fabsf((float)cy0 - es.ay), fabsf((float)(cy0 + bh - 1) - es.ay)));
    float4 old0 = make_float4(0, 0, 0, 0), old1 = old0;   // forward: snapshot of the 8 pixels this lane checks afterwards
    float fm0 = 0.0f, fa1 = 0.0f, fa2 = 0.0f;             // backward: moments of this lane's face
    bool mine = false;
    int rank = lane;                                      // slot (and hit-mask bit) of this lane's face

    // ---- direct walk: every face of the batch is small, so each lane walks the row spans of its OWN face
    // with the face constants in registers — no candidate redistribution (prefix sums, slot search, span
    // bytes in shared memory), no merging of partial moments.  The walk takes max-over-lanes steps, which
    // for small boxes is cheaper than that machinery (config 3: 8 steps on average, fine meshes: 2).
    const bool direct = DIRECT && !IS_MULTI && __reduce_max_sync(FULL, queued ? bw * bh : 0) <= P.direct_max_box;
    if (direct) {
        mine = queued;
        const double d = (double)es.k3 + 1e-15;                       // diff_fp_for.cu:144 denominator
        const double2 dy = make_double2(d, 1.0 / d);
        const float4 v0 = make_float4(es.ax, es.ay, es.m, es.p), v1 = make_float4(es.n, es.q, es.k3, 0.0f);
        float cA = 0.0f, cB = 0.0f;
        if (IS_FWD) {
            if (queued) {
                const float mu_in = __ldg(P.mus + lab.x);
                float mu_out = __ldg(P.mus + lab.y);
                if (lab.y == 0) mu_out = -mu_out;                     // :174-175
                cA = __fmul_rn(sgn, mu_in);                           // :187
                cB = -__fmul_rn(sgn, mu_out);                         // :190
                // the fix-up of pixels hit by several faces re-reads the faces from their slots (slot = lane)
                float4* sl = S.slot + lane * SLOT_VEC;
                sl[0] = v0; sl[1] = v1;
                reinterpret_cast<double2*>(sl)[2] = dy;
                sl[3] = make_float4(r0, r1, r2, cA);
                sl[4] = make_float4(cB, 0.0f, 0.0f, __int_as_float(f));
            }
            old0 = *reinterpret_cast<const float4*>(S.acc + 4 * lane);
            old1 = *reinterpret_cast<const float4*>(S.acc + 128 + 4 * lane);
            __syncwarp();
        }
        int r = -1, x = 0, xe = -1;
        int cfirst = TILE, clast = -1, ccount = 0;                    // RECSPAN: covered columns of the current row
        bool active = queued;
#pragma unroll 1
        while (active) {                                              // first non-empty row
            if (++r >= bh) { active = false; break; }
            int xs = cx0; xe = cx0 + bw - 1;
            if (!tight || row_span(sps, (float)(cy0 + r) - es.ay, cx0, cx0 + bw - 1, xs, xe)) { x = xs; break; }
        }
#pragma unroll 1
        while (__any_sync(FULL, active)) {
            if (active) {
                float k1, k2, w0, w1, w2;
                const bool in = cover_slot<EXACT_DIV>(v0, v1, dy, (float)x, (float)(cy0 + r), k1, k2, w0, w1, w2);
                st.cand++;
                if (RECSPAN && in) { cfirst = min(cfirst, x - tx0); clast = x - tx0; ++ccount; }
                if (in) {
                    const int pl = (cy0 + r - ty0) * TILE + (x - tx0);
                    if (IS_FWD) {
                        const float bary = __fmaf_rn(w2, r2, __fmaf_rn(w0, r0, __fmul_rn(w1, r1)));       // :182
                        if (!(bary > 0.0f)) st.nonpos++;              // :183 assert -> counter
                        st.frag++;
                        atomicOr(&S.f.hitm[pl], 1u << lane);
                        S.acc[pl] = __fmaf_rn(bary, cB, __fmaf_rn(bary, cA, S.acc[pl]));                  // :187,190
                    } else {
                        const float g = S.acc[pl];
                        if (g != 0.0f) { st.frag++; fm0 += g; fa1 = fmaf(g, k1, fa1); fa2 = fmaf(g, k2, fa2); }
                    }
                }
                if (++x > xe) {                                       // next non-empty row of this face
                    if (RECSPAN && ccount) {
                        S.span[lane][cy0 + r - ty0] = (unsigned char)((cfirst << 4) | clast);
                        irregular |= (clast - cfirst + 1 != ccount);
                        cfirst = TILE; clast = -1; ccount = 0;
                    }
#pragma unroll 1
                    while (true) {
                        if (++r >= bh) { active = false; break; }
                        int xs = cx0; xe = cx0 + bw - 1;
                        if (!tight || row_span(sps, (float)(cy0 + r) - es.ay, cx0, cx0 + bw - 1, xs, xe)) { x = xs; break; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
    // ---- general path: candidates of the whole batch redistributed evenly over the lanes ------------
    // candidate rows: conservative per-row spans (bounding-box rows when the face is degenerate or NaN)
    unsigned rowmask = 0u;
    if (queued) {
        unsigned char* sprow = &S.span[lane][cy0 - ty0];
        const int pack0 = -(tx0 << 4) - tx0;                           // ((xs - tx0) << 4) | (xe - tx0) == (xs << 4) + xe + pack0
        unsigned bit = 1u << (cy0 - ty0);
#pragma unroll 1
        for (int r = 0; r < bh; ++r, bit <<= 1) {
            int xs = cx0, xe = cx0 + bw - 1;
            bool ok = true;
            if (tight) ok = row_span(sps, (float)(cy0 + r) - es.ay, cx0, cx0 + bw - 1, xs, xe);
            if (ok) {
                rowmask |= bit;
                ncand += xe - xs + 1;
                sprow[r] = (unsigned char)((xs << 4) + xe + pack0);
            }
        }
    }
    const unsigned nz = __ballot_sync(FULL, ncand > 0);
    if (nz == 0u) {
        if (RECSPAN && rec >= 0 && lane == 0) P.brec[(long long)rec * BREC_INTS + 1] = nb | BREC_FLAG_EMPTY;
        return;
    }
    mine = ncand > 0;
    const int nslots = __popc(nz);
    rank = __popc(nz & lanemask_lt());

    int incl = ncand;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int T = __shfl_sync(FULL, incl, 31);

    if (mine) {
        float4* sl = S.slot + rank * SLOT_VEC;
        // box word: queue lane (5 bits) | row mask (16 bits) << 5 | (ncand - 1) << 21   (ncand <= 256)
        const int box = lane | (int)(rowmask << 5) | ((ncand - 1) << 21);
        sl[0] = make_float4(es.ax, es.ay, es.m, es.p);
        sl[1] = make_float4(es.n, es.q, es.k3, __int_as_float(box));
        const double d = (double)es.k3 + 1e-15;                        // diff_fp_for.cu:144 denominator
        reinterpret_cast<double2*>(sl)[2] = make_double2(d, 1.0 / d);
        float cA = 0.0f, cB = 0.0f;
        if (IS_FWD) {
            const float mu_in = __ldg(P.mus + lab.x);
            float mu_out = __ldg(P.mus + lab.y);
            if (lab.y == 0) mu_out = -mu_out;                         // :174-175
            cA = __fmul_rn(sgn, mu_in);                               // :187
            cB = -__fmul_rn(sgn, mu_out);                             // :190
            sl[3] = make_float4(r0, r1, r2, cA);
        } else if (IS_MULTI) {
            sl[3] = make_float4(r0, r1, r2, 0.0f);
        } else {
            S.mom[0][rank] = 0.0f; S.mom[1][rank] = 0.0f; S.mom[2][rank] = 0.0f;
        }
        const int word = lab.x | (lab.y << 12) | (sgn < 0.0f ? 1 << 24 : 0);
        sl[4] = make_float4(cB, __int_as_float(incl - ncand), __int_as_float(word), __int_as_float(f));
    }
    if (IS_FWD) {
        old0 = *reinterpret_cast<const float4*>(S.acc + 4 * lane);
        old1 = *reinterpret_cast<const float4*>(S.acc + 128 + 4 * lane);
    }
    __syncwarp();

    // ---- contiguous candidate range of this lane --------------------------------------------
    const int q = T >> 5, rem = T & 31;
    const int mycount = q + (lane < rem ? 1 : 0);
    const int j0 = lane * q + min(lane, rem);
    const int iters = q + (rem ? 1 : 0);
    int i = 0;                                           // slot owning candidate j0: last slot with excl <= j0
    if (mycount > 0) {
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int c = i + step;
            if (c < nslots && __float_as_int(S.slot[c * SLOT_VEC + 4].y) <= j0) i = c;
        }
    }
    // position inside the face: row r, column offset lx within the row's span
    int e = j0 - __float_as_int(S.slot[i * SLOT_VEC + 4].y);
    int box = __float_as_int(S.slot[i * SLOT_VEC + 1].w);
    int n_i = (box >> 21) + 1, ql = box & 31;
    unsigned rows = ((unsigned)box >> 5) & 0xffffu;
    int r = __ffs(rows) - 1, rx0 = 0, rlen = 1, lx = 0;
    {
        int left = e;
        while (true) {
            const unsigned sp = S.span[ql][r];
            rx0 = sp >> 4; rlen = (int)(sp & 15u) - rx0 + 1;
            if (left < rlen || (rows & (rows - 1u)) == 0u) break;
            left -= rlen; rows &= rows - 1u; r = __ffs(rows) - 1;
        }
        lx = min(left, rlen - 1);
    }
    float M0 = 0.0f, A1 = 0.0f, A2 = 0.0f;
    float H0 = 0.0f, H1 = 0.0f, H2 = 0.0f;
    int hkey = -1;                                       // slot of the face this lane finished for earlier lanes
    bool whole = (e == 0);                               // current face entered at its first candidate

#pragma unroll 1
    for (int t = 0; t < iters; ++t) {
        const bool act = t < mycount;
        const float4* sl = S.slot + i * SLOT_VEC;
        const float4 v0 = sl[0], v1 = sl[1];
        const int px = rx0 + lx, py = r;
        const int pl = (py * TILE + px) & (TILE_PIX - 1);
        bool in = false;
        float k1 = 0.0f, k2 = 0.0f, w0, w1, w2;
        if (act) {
            in = cover_slot<EXACT_DIV>(v0, v1, reinterpret_cast<const double2*>(sl)[2],
                                       (float)(tx0 + px), (float)(ty0 + py), k1, k2, w0, w1, w2);
            st.cand++;
            if (RECSPAN && !in) {                        // a candidate of the conservative span that is not covered (~3 %)
                // normally an end of the span (a one-pixel row is trimmed from the left, or from the right in column 15, so
                // that no nibble overflows): remembered and applied after the walk; anything else marks the row
                const bool left = (lx == 0) && (rlen > 1 || px < 15), right = (lx == rlen - 1);
                if ((left || right) && (myfail >> 16) == 0u) myfail = (myfail << 16) | 1u | (ql << 1) | (py << 6) | (left ? 0u : 1u << 10);
                else atomicOr(&S.susp[ql >> 1], (1u << py) << (16 * (ql & 1)));
            }
        }
        if (IS_FWD) {
            if (in) {
                const float4 v3 = sl[3];
                const float cB = sl[4].x;
                const float bary = __fmaf_rn(w2, v3.z, __fmaf_rn(w0, v3.x, __fmul_rn(w1, v3.y)));  // :182
                if (!(bary > 0.0f)) st.nonpos++;                      // :183 assert -> counter
                st.frag++;
                atomicOr(&S.f.hitm[pl], 1u << i);
                S.acc[pl] = __fmaf_rn(bary, cB, __fmaf_rn(bary, v3.w, S.acc[pl]));               // :187,190
            }
        } else if (IS_MULTI) {
            if (in) { st.frag++; atomicOr(&S.f.hitm[pl], 1u << i); }
        } else {
            if (in) {
                const float g = S.acc[pl];
                if (g != 0.0f) { st.frag++; M0 += g; A1 = fmaf(g, k1, A1); A2 = fmaf(g, k2, A2); }
            }
        }
        if (act) {                                       // next column, next non-empty row, or next face
            ++e; ++lx;
            if (lx == rlen) {
                lx = 0;
                if (e == n_i) {
                    if (!IS_FWD && !IS_MULTI) {
                        if (whole) {           // entered at its first candidate and finished here: sole contributor
                            S.mom[0][i] = M0; S.mom[1][i] = A1; S.mom[2][i] = A2;
                        } else {               // this lane finishes a face that earlier lanes started
                            hkey = i; H0 = M0; H1 = A1; H2 = A2;
                        }
                        M0 = 0.0f; A1 = 0.0f; A2 = 0.0f;
                        whole = true;
                    }
                    i = min(i + 1, nslots - 1); e = 0;
                    box = __float_as_int(S.slot[i * SLOT_VEC + 1].w);
                    n_i = (box >> 21) + 1; ql = box & 31;
                    rows = ((unsigned)box >> 5) & 0xffffu;
                } else {
                    rows &= rows - 1u;
                }
                r = __ffs(rows) - 1;
                const unsigned sp = S.span[ql][r & 15];
                rx0 = sp >> 4; rlen = (int)(sp & 15u) - rx0 + 1;
            }
        }
    }
    if (!IS_FWD && !IS_MULTI) {
        // Faces split over several lanes: the lanes that leave a face unfinished hold "tail" partials
        // (consecutive lanes, same slot); the lane that finishes it holds the "head" partial.  A
        // segmented shuffle scan sums each run of tails into its last lane; the finishing lane adds
        // that to its own part and is the face's only writer — no shared-memory atomics.
        const int tkey = (mycount > 0 && e > 0) ? i : -1;
        float t0 = M0, t1 = A1, t2 = A2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int kk = __shfl_up_sync(FULL, tkey, d);
            const float u0 = __shfl_up_sync(FULL, t0, d), u1 = __shfl_up_sync(FULL, t1, d), u2 = __shfl_up_sync(FULL, t2, d);
            if (lane >= d && kk == tkey && tkey >= 0) { t0 += u0; t1 += u1; t2 += u2; }
        }
        const int pk = __shfl_up_sync(FULL, tkey, 1);
        const float p0 = __shfl_up_sync(FULL, t0, 1), p1 = __shfl_up_sync(FULL, t1, 1), p2 = __shfl_up_sync(FULL, t2, 1);
        if (hkey >= 0) {
            const bool chain = lane > 0 && pk == hkey;
            S.mom[0][hkey] = H0 + (chain ? p0 : 0.0f);
            S.mom[1][hkey] = H1 + (chain ? p1 : 0.0f);
            S.mom[2][hkey] = H2 + (chain ? p2 : 0.0f);
        }
    }
    __syncwarp();
    if (!IS_FWD && !IS_MULTI && mine) { fm0 = S.mom[0][rank]; fa1 = S.mom[1][rank]; fa2 = S.mom[2][rank]; }
    if (RECSPAN) {
        // trim the conservative spans to the exact coverage (the walk is over: nobody reads the spans any more): one
        // shared-memory add per remembered failure on the word that holds the row's byte — first column + 1 or last
        // column - 1, neither can carry into the neighbouring byte
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned e = (myfail >> (16 * h)) & 0xffffu;
            if (e) {
                const unsigned q = (e >> 1) & 31u, rr = (e >> 6) & 15u;
                atomicAdd(reinterpret_cast<unsigned*>(S.span[q]) + (rr >> 2), ((e >> 10) ? 0xffffffffu : 0x10u) << (8u * (rr & 3u)));
            }
        }
        __syncwarp();
        unsigned rows = queued ? (S.susp[lane >> 1] >> (16 * (lane & 1))) & 0xffffu : 0u;
        // rare: rows re-derived from the hit masks, which hold the coverage of every pixel of the tile by this batch
        while (rows) {
            const int rr = __ffs(rows) - 1;
            rows &= rows - 1u;
            int first = 0, last = -1, cnt = 0;
            for (int x = 0; x < TILE; ++x)
                if ((S.f.hitm[rr * TILE + x] >> rank) & 1u) { if (!cnt) first = x; last = x; ++cnt; }
            irregular |= cnt > 0 && (last - first + 1 != cnt);
            S.span[lane][rr] = (unsigned char)(cnt > 0 ? ((first << 4) | last) : 0x10);
        }
    }
    }   // general path

    if (RECSPAN) {
        // the batch's exact coverage goes into its record (the spans are overwritten by the conflicted-pixel list below)
        if (__any_sync(FULL, irregular)) {              // rare (see RasterParams::xlist)
#ifdef CTDR_PROBE_ROWSTATS
            if (lane == 0) atomicAdd(&g_rowstats[3], 1ull);
#endif
            if (irregular && queued) {
                const int k = atomicAdd(P.xcount, 1);
                if (k < XLIST_CAP) P.xlist[k] = make_int4(b, f, tx0, ty0);
                else *rec_fail = 1;
                *reinterpret_cast<uint4*>(S.span[lane]) = make_uint4(SPAN_NONE4, SPAN_NONE4, SPAN_NONE4, SPAN_NONE4);
            }
        }
        if (rec >= 0 && queued) {
            int* R = P.brec + (long long)rec * BREC_INTS;
            R[2 + lane] = f | (fneg ? (int)0x80000000 : 0);
            reinterpret_cast<unsigned char*>(R + BREC_INFO)[lane] = (unsigned char)((cy0 - ty0) | ((bh - 1) << 4));
            reinterpret_cast<uint4*>(R + BREC_SPANS)[lane] = *reinterpret_cast<const uint4*>(S.span[lane]);
        }
        __syncwarp();
    }

    if constexpr (IS_MULTI) {
        // ---- every pixel hit by the batch is (re)computed pixel-parallel for all outputs -----------
        const uint4 h0 = *reinterpret_cast<const uint4*>(S.f.hitm + 4 * lane);
        const uint4 h1 = *reinterpret_cast<const uint4*>(S.f.hitm + 128 + 4 * lane);
        const unsigned hm[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        int ncp = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool hit = hm[k] != 0u;
            const unsigned m = __ballot_sync(FULL, hit);
            if (hit) S.cpix[ncp + __popc(m & lanemask_lt())] = (unsigned char)((k < 4 ? 0 : 128) + 4 * lane + (k & 3));
            ncp += __popc(m);
        }
        __syncwarp();
        if (ncp) st.nonpos += redo_multi<EXACT_DIV, WS>(S, multi_xacc(gmus, threadIdx.x >> 5), P.mus, P.n_mus,
                                                    P.n_out, ncp, tx0, ty0, lane);
        __syncwarp();
        return;
    }

    if (IS_FWD) {
        // ---- pixels hit by several faces of this batch: restore and redo them in face order --------
        const uint4 h0 = *reinterpret_cast<const uint4*>(S.f.hitm + 4 * lane);
        const uint4 h1 = *reinterpret_cast<const uint4*>(S.f.hitm + 128 + 4 * lane);
        *reinterpret_cast<uint4*>(S.f.hitm + 4 * lane) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(S.f.hitm + 128 + 4 * lane) = make_uint4(0u, 0u, 0u, 0u);
        int ncp = 0;
        const unsigned hm[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float ov[8] = {old0.x, old0.y, old0.z, old0.w, old1.x, old1.y, old1.z, old1.w};
        unsigned mymask[8];
        bool anymulti = false;
        if (WS::HAS_CNT && P.cnt) {                                   // coverage counts are optional
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (hm[k]) S.cntv[(k < 4 ? 0 : 128) + 4 * lane + (k & 3)] += __popc(hm[k]);   // :192
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool multi = (hm[k] & (hm[k] - 1u)) != 0u;          // more than one face of the batch
            mymask[k] = multi ? hm[k] : 0u;
            anymulti |= multi;
        }
        if (__ballot_sync(FULL, anymulti)) {                          // rare: most batches have no such pixel
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int pix = (k < 4 ? 0 : 128) + 4 * lane + (k & 3);
                const bool multi = mymask[k] != 0u;
                const unsigned m = __ballot_sync(FULL, multi);
                if (m) {
                    if (multi) {
                        S.acc[pix] = ov[k];
                        S.cpix[ncp + __popc(m & lanemask_lt())] = (unsigned char)pix;
                    }
                    ncp += __popc(m);
                }
            }
        }
        if (ncp) {
            // the hit masks of the conflicted pixels go back to shared memory for their new owners
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (mymask[k]) S.f.hitm[(k < 4 ? 0 : 128) + 4 * lane + (k & 3)] = mymask[k];
            __syncwarp();
            redo_conflicted<EXACT_DIV, WS>(S, ncp, tx0, ty0, lane);
        }
        __syncwarp();
        return;
    }

    if (mine) {
        // per-face gradients from the moments of this tile's fragments (face_grads_from_moments)
        const float m0 = fm0, a1 = fa1, a2 = fa2;
        if (m0 != 0.0f || a1 != 0.0f || a2 != 0.0f) {
            const float mu_in = __ldg(P.mus + lab.x), mu_out = __ldg(P.mus + lab.y);
            const float cgeo = sgn * (mu_in - mu_out);                // raw mu, no label-0 flip (back.cu:115)
            float gp[6], gl[3], bsum;
            face_grads_from_moments(es, r0, r1, r2, cgeo, m0, a1, a2, gp, gl, bsum);
            float* dl = P.grad_len3 + bf * 3;                         // back.cu:126-136
            atomicAdd(dl + 0, gl[0]); atomicAdd(dl + 1, gl[1]); atomicAdd(dl + 2, gl[2]);
            st.mus_add(gmus, lab.x, (lab.x == 0 ? -sgn : sgn) * bsum, lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
            float* dp = P.grad_p6 + bf * 6;                           // back.cu:285-292
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(dp + k, gp[k]);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Backward batch (MODE_BWD / MODE_STEP): lane per face, row spans + per-row prefix sums.
//
// The gradients of a face need three sums over its covered pixels only (face_grads_from_moments):
// sum g, sum g*k1, sum g*k2, and k1 = q*(x-ax) - n*(y-ay), k2 = m*(y-ay) - p*(x-ax) are affine in the pixel
// position, so sum g, sum g*x, sum g*y suffice.  On one detector row a triangle covers an interval, so with
// the exclusive prefix sums E0 = sum g and E1 = sum g*x along every row of the tile (built once per tile)
// a (face, row) pair costs four shared-memory loads instead of one exact coverage test per pixel.
// The interval comes from the same linear inequalities as the forward's conservative spans, with a margin
// (1e-4 px + the rounding of the estimate and of the reference's own k1/k2/w0, times ~10) on BOTH sides:
// pixels farther inside than the margin are certainly counted by the reference, pixels farther outside are
// certainly not; only a pixel centre within the margin of an edge (probability ~1e-3 per row end) takes
// the reference's exact test (cover_exact, bit-identical decision), so the set of fragments is the forward's.
// Faces without usable bounds (zero area, NaN, huge, two near-horizontal edges) and knife-edge rows on a
// near-horizontal edge test every bounding-box pixel exactly.
// ------------------------------------------------------------------------------------------
constexpr int PREFIX_STRIDE = TILE + 1;
// The prefix sums are kept in float64: a span sum is a DIFFERENCE of two prefixes, and with fp32 prefixes its
// error would be an ulp of the whole row's running sum (measured: 2e-4 relative on grad_p6, the moments then
// cancel against each other) instead of an ulp of the span sum.  2 x 16 x 17 doubles = 4352 B: the forward-only
// part of WarpSmem (slot + span + hit masks = 4096 B) and the head of the count tile, which the backward modes do not use.
template <bool CNT>
__device__ __forceinline__ double* prefix_e0(WarpSmemT<CNT>& S) { return reinterpret_cast<double*>(S.slot); }
static_assert(offsetof(WarpSmem, cntv) + 64 * sizeof(int) >= 2 * TILE * PREFIX_STRIDE * sizeof(double), "prefix arrays overlay slot + span + hit masks + the head of the count tile");

// per-warp shared memory of the replay-only residual/backward kernel (k_step_replay): nothing of the forward's state
struct __align__(16) ReplayWarpSmem {
    double pre[2 * TILE * PREFIX_STRIDE];   // E0 | E1 (the g tile itself is never stored: g = differences of E0)
    int    landing[2][BREC_INTS];           // the current batch record and the next one in flight (cp.async, 16 B pieces)
};
__device__ __forceinline__ double* prefix_e0(ReplayWarpSmem& S) { return S.pre; }

// exclusive prefix sums of g (S.acc) and g*x' (x' = tile-relative column) along the 16 rows of the tile;
// lane = (row, half row)
template <class WS>
__device__ __forceinline__ void build_row_prefix(WS& S, int lane) {
    const int row = lane >> 1, half = lane & 1;
    const float4 a = *reinterpret_cast<const float4*>(S.acc + row * TILE + half * 8);
    const float4 b = *reinterpret_cast<const float4*>(S.acc + row * TILE + half * 8 + 4);
    const float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    double e0[8], e1[8], p0 = 0.0, p1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        e0[k] = p0; e1[k] = p1;
        p0 += (double)g[k];
        p1 = fma((double)g[k], (double)(half * 8 + k), p1);
    }
    const double t0 = __shfl_sync(FULL, p0, lane & ~1), t1 = __shfl_sync(FULL, p1, lane & ~1);
    const double o0 = half ? t0 : 0.0, o1 = half ? t1 : 0.0;
    __syncwarp();                                       // the overlay may still be read as forward state by a slower lane
    double* E0 = prefix_e0(S) + row * PREFIX_STRIDE + half * 8;
    double* E1 = E0 + TILE * PREFIX_STRIDE;
#pragma unroll
    for (int k = 0; k < 8; ++k) { E0[k] = e0[k] + o0; E1[k] = e1[k] + o1; }
    if (half) { E0[8] = p0 + o0; E1[8] = p1 + o1; }
    __syncwarp();
}

// the same sums straight from registers: lane = (row, half row) holds its eight g values (replay kernel: the tile of g
// is never staged in shared memory)
__device__ __forceinline__ void build_row_prefix_regs(ReplayWarpSmem& S, const float (&g)[8], int lane) {
    const int row = lane >> 1, half = lane & 1;
    double e0[8], e1[8], p0 = 0.0, p1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        e0[k] = p0; e1[k] = p1;
        p0 += (double)g[k];
        p1 = fma((double)g[k], (double)(half * 8 + k), p1);
    }
    const double t0 = __shfl_sync(FULL, p0, lane & ~1), t1 = __shfl_sync(FULL, p1, lane & ~1);
    const double o0 = half ? t0 : 0.0, o1 = half ? t1 : 0.0;
    double* E0 = S.pre + row * PREFIX_STRIDE + half * 8;
    double* E1 = E0 + TILE * PREFIX_STRIDE;
#pragma unroll
    for (int k = 0; k < 8; ++k) { E0[k] = e0[k] + o0; E1[k] = e1[k] + o1; }
    if (half) { E0[8] = p0 + o0; E1[8] = p1 + o1; }
    __syncwarp();
}

struct SpanSetupB {
    SpanSetup sp;        // bounds widened outward by the face's margin
    float din;           // lo + din <= x <= hi - din: certainly covered
    unsigned doubt;      // bit r: row r of the face's rows in this tile takes the exact test on its whole bounding-box
                         // row — all rows of a face without usable bounds (zero area, NaN, huge, two near-horizontal
                         // edges); the first / last row when it lies within rounding of a near-horizontal edge
};

// t_first / t_last: y - ay of the first and last row of the face in this tile (nrows rows); smax: max |x - ax| over the
// face's columns in this tile; axr = ax - tile origin: the bounds come out TILE-RELATIVE, so their own rounding is that
// of numbers of tile size, not of detector size.
//
// Margin of edge j (in pixels, both sides of the estimated root).  The reference decides on COMPUTED values:
// sign(k1), sign(k2) (edges 0, 1) and w0 = fl(fl(1 - w1) - w2) >= 0 (edge 2), with k1 = fma(s, q, -fl(n*t)),
// k2 = fma(m, t, -fl(s*p)), s = fl(x - ax), t = fl(y - ay), w = fl32(k/d).  Against the exact line a_j*(x - root), with
// u = 2^-24 = 6e-8 (s and t carry ONE rounding of their own magnitude — the difference is formed exactly, then rounded):
//   edge 0 (a = q):    |k1 - exact| <= u*(|q|*|s| + 2*|n|*|t|)                        -> u*(smax + 2*|R_0|*tmax) pixels
//   edge 1 (a = p):    |k2 - exact| <= u*(2*|p|*|s| + |m|*|t|)                        -> u*(2*smax + |R_1|*tmax) pixels
//   edge 2 (a = q-p):  both of the above, over |q - p|, plus the float32 roundings of k1, k2 (now of size |d|), of w1,
//                      w2 and of the two subtractions, 6u in barycentric units = 6u*|C_2| pixels (C_2 = |k3| / a_2)
//                                                        -> u*((|q|+2|p|)*smax + (2|n|+|m|)*tmax)/|q-p| + 6u*|C_2| pixels
//   this estimate:     1/a_j from MUFU.RCP and two products (4e-7 relative on R_j*t and C_j), and the roundings of
//                      axr + C -/+ margin and of the fma (1.2e-7 each, of |axr| + |C_j| + |R_j|*tmax)
//                                                        -> 5.2e-7*(|R_j|*tmax + |C_j|) + 2.4e-7*|axr| pixels
// (|R_2|*tmax <= (2|n|+|m|)*tmax/|q-p|.)  The margins used are at least twice these bounds, plus a constant floor:
//   margin_j = 2e-5 + 5e-7*|axr| + 4.8e-7*smax + (1.3e-6*T_j + 4.8e-7*S_j)/|a_j| + 3e-6*|C_j|
//   T_0 = |n|*tmax, T_1 = |m|*tmax, T_2 = (2|n|+|m|)*tmax;  S_0 = S_1 = 0, S_2 = (|q|+2|p|)*smax.
// (First version: one margin 1e-4 + 8e-6*(|C| + |R|*tmax) + 1e-6*|ax| for all edges of a face: 0.34 % of the (face, row)
// pairs of config 3 took the exact path, 0.24 % with these; separate inner margins for the lower and the upper bound
// bring it to 0.13 % but cost a register in the row loop and were measured slower.)
__device__ __forceinline__ SpanSetupB span_setup_bwd(const EdgeSetup& es, float axr, float t_first, float t_last, int nrows, float smax) {
    SpanSetupB B;
    B.sp = {0.0f, -3.0e38f, 0.0f, -3.0e38f, 0.0f, 3.0e38f, 0.0f, 3.0e38f};
    const float tmax = fmaxf(fabsf(t_first), fabsf(t_last));
    float fb = 0.0f, fg = 1.0f, fthr = 0.0f;     // near-horizontal edge: value fb*t + fg on row t, decided when |value| > fthr
    const float sg = es.k3 < 0.0f ? -1.0f : 1.0f;
    const float a[3] = {sg * es.q, -sg * es.p, -sg * (es.q - es.p)};
    const float beta[3] = {-sg * es.n, sg * es.m, -sg * (es.m - es.n)};
    const float gamma[3] = {0.0f, 0.0f, fabsf(es.k3)};
    const float an = fabsf(es.n), am = fabsf(es.m);
    const float TS[3] = {1.3e-6f * an * tmax, 1.3e-6f * am * tmax,
                         1.3e-6f * (2.0f * an + am) * tmax + 4.8e-7f * (fabsf(es.q) + 2.0f * fabsf(es.p)) * smax};
    const float mbase = 2e-5f + 5e-7f * fabsf(axr) + 4.8e-7f * smax;
    bool have_l = false, have_u = false;
    float mmax_l = 0.0f, mmax_u = 0.0f;
    int nflat = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool use = fabsf(a[j]) > 1e-4f;
        const float ia = use ? __fdividef(1.0f, a[j]) : 0.0f;
        const float R = -beta[j] * ia, C = -gamma[j] * ia;
        const float margin = mbase + TS[j] * fabsf(ia) + 3e-6f * fabsf(C);
        const bool lower = use && a[j] > 0.0f, upper = use && a[j] < 0.0f;
        const float cl = axr + C - margin, cu = axr + C + margin;
        if (lower && !have_l) { B.sp.RL0 = R; B.sp.CL0 = cl; } else if (lower) { B.sp.RL1 = R; B.sp.CL1 = cl; }
        if (upper && !have_u) { B.sp.RU0 = R; B.sp.CU0 = cu; } else if (upper) { B.sp.RU1 = R; B.sp.CU1 = cu; }
        have_l |= lower; have_u |= upper;
        if (lower) mmax_l = fmaxf(mmax_l, margin);
        if (upper) mmax_u = fmaxf(mmax_u, margin);
        if (!use) {
            ++nflat;
            fb = beta[j]; fg = gamma[j];
            fthr = fabsf(a[j]) * smax + 1e-5f * (fabsf(beta[j]) * tmax + fabsf(gamma[j])) + 1e-30f;
        }
    }
    B.din = 2.0f * fmaxf(mmax_l, mmax_u);     // one value for both sides: a second live register in the row loop costs more than it saves (measured)
    // A near-horizontal edge is the top or the bottom of the triangle, so the rows of the bounding box lie on its
    // inner side; only the first or the last row can come within rounding of it (the value is linear in t).
    const float ak3 = fabsf(es.k3);
    const bool exactall = !((ak3 > 1e-6f) && (ak3 < 1e10f)) || nflat > 1;   // NaN coordinates: every `use` is false
    B.doubt = exactall ? 0xffffu : 0u;
    if (!(fmaf(fb, t_first, fg) > fthr)) B.doubt |= 1u;
    if (!(fmaf(fb, t_last, fg) > fthr)) B.doubt |= 1u << ((nrows - 1) & 15);
    return B;
}

// rare path of one (face, row): columns outside [xsi, xei] (all of them when that range is empty) take the
// reference's exact test.  Few scalars in, one float4 out: sum g, sum g*k1, sum g*k2 with k1, k2 evaluated per pixel
// exactly as the forward does (a zero-area face has k1 = k2 = 0 on every pixel it covers, so its moments — and with
// them its gradients, before the 1/(k3^2 + eps) factor — are exact zeros like the reference's), and the fragment
// count.  The face is re-read (L1 hits); nothing of the caller's state has its address taken.
// grow = the tile row of g (16 floats); xs/xe/xsi/xei tile-relative columns; y absolute.
// src: the face's p6 row, or (vtrow != null, gather mode) its three vertex ids into the angle's vertex records.
__device__ __noinline__ float4 bwd_row_exact(const void* __restrict__ src, const float4* __restrict__ vtrow, const float* grow,
                                             int xs, int xe, int xsi, int xei, int tx0, int y) {
    float ax, ay, bx, by, cx, cy;
    if (vtrow) {
        const int* fi = static_cast<const int*>(src);
        const float2* row = reinterpret_cast<const float2*>(vtrow);
        const float2 a = __ldg(row + 2 * __ldg(fi)), bb = __ldg(row + 2 * __ldg(fi + 1)), c = __ldg(row + 2 * __ldg(fi + 2));
        ax = a.x; ay = a.y; bx = bb.x; by = bb.y; cx = c.x; cy = c.y;
    } else {
        const float* p6row = static_cast<const float*>(src);
        ax = __ldg(p6row); ay = __ldg(p6row + 1); bx = __ldg(p6row + 2); by = __ldg(p6row + 3); cx = __ldg(p6row + 4); cy = __ldg(p6row + 5);
    }
    const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
    float M0 = 0.0f, A1 = 0.0f, A2 = 0.0f;
    int cnt = 0;
    for (int x = xs; x <= xe; ++x) {
        float s, t, k1, k2, w0, w1, w2;
        bool in = (x >= xsi) & (x <= xei);
        if (!in) {
            in = cover_exact<false>(es, (float)(tx0 + x), (float)y, s, t, k1, k2, w0, w1, w2);
        } else {
            s = __fsub_rn((float)(tx0 + x), es.ax); t = __fsub_rn((float)y, es.ay);
            k1 = __fmaf_rn(s, es.q, -__fmul_rn(es.n, t));
            k2 = __fmaf_rn(es.m, t, -__fmul_rn(s, es.p));
        }
        if (in) {
            const float g = grow[x];
            M0 += g; A1 = fmaf(g, k1, A1); A2 = fmaf(g, k2, A2);
            ++cnt;
        }
    }
    return make_float4(M0, A1, A2, __int_as_float(cnt));
}

template <bool GATHER, class ThreadStats, class WS>
__device__ __forceinline__ void process_batch_rows(const RasterParams& P, WS& S, float* gmus,
                                                   int b, int nb, int tx0, int ty0, int tx1, int ty1,
                                                   int lane, ThreadStats& st) {
    const bool queued = lane < nb;
    const int f = queued ? S.qface[lane] : 0;
    const long long bf = (long long)b * P.F + f;
    float M0 = 0.0f, Ss = 0.0f, St = 0.0f;              // sum g, sum g*(x - ax), sum g*(y - ay) over the face's fragments
    float A1x = 0.0f, A2x = 0.0f;                       // sum g*k1, sum g*k2 of the rows that took the exact path
    typename ThreadStats::count_t nfrag{};
    {
        // Only what the row loop needs stays live in it (the kernel runs at the 64-register cap): the edge
        // constants, lengths, labels and facing sign are (re)loaded afterwards — L1 hits — for the epilogue.
        int cx0 = 0, cy0 = 0, bw = 1, bh = 0;
        float ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;
        if (queued) {
            const short4 fb = unpack_box(__ldg(P.face_box + bf));
            load_face_xy<GATHER>(P, b, f, bf, ax, ay, bx, by, cx, cy);
            cx0 = max((int)fb.x, tx0); cy0 = max((int)fb.y, ty0);
            bw = min((int)fb.z, tx1) - cx0 + 1;
            bh = min((int)fb.w, ty1) - cy0 + 1;
            if (!GATHER) prefetch_l1(P.len3 + bf * 3);    // the epilogue's operands travel while the rows are walked
        }
        __syncwarp();                                   // queue entries consumed
        if (lane == 0) st.visits++;
        const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
        const int x_hi = cx0 + bw - 1;
        // everything below in tile-relative coordinates (columns 0..15, rows 0..15)
        const float axr = es.ax - (float)tx0;
        const SpanSetupB sb = span_setup_bwd(es, axr, (float)cy0 - es.ay, (float)(cy0 + bh - 1) - es.ay, bh,
                                             fmaxf(fabsf((float)cx0 - es.ax), fabsf((float)x_hi - es.ax)));
        const double axd = (double)axr;
        const float xlo = (float)(cx0 - tx0), xhi = (float)(x_hi - tx0);
        const float cl0 = sb.sp.CL0, cl1 = sb.sp.CL1, cu0 = sb.sp.CU0, cu1 = sb.sp.CU1;
        const double* E0 = prefix_e0(S);
        const int bhmax = __reduce_max_sync(FULL, bh);
        const float t0 = (float)cy0 - es.ay;            // y - ay of the face's first row in this tile
        const int rowbase0 = (cy0 - ty0) * PREFIX_STRIDE;
        unsigned exact_rows = 0u;                       // rows left to the exact path below (kept out of the hot loop:
                                                        // a call inside it would force spills around every iteration)
        {
            float rf = 0.0f;
            int rowbase = rowbase0;
#pragma unroll 1
            for (int r = 0; r < bhmax; ++r, rf += 1.0f, rowbase += PREFIX_STRIDE) {
                const float t = t0 + rf;                // one rounding, like the reference's own t = fl(y - ay): no drift over the rows
                if (r < bh) {
                    const float lo = fmaxf(fmaf(sb.sp.RL0, t, cl0), fmaf(sb.sp.RL1, t, cl1));
                    const float hi = fminf(fmaf(sb.sp.RU0, t, cu0), fmaf(sb.sp.RU1, t, cu1));
                    const float fs = ceilf(fmaxf(lo, xlo)), fe = floorf(fminf(hi, xhi));   // candidate columns of this row
                    // the reference's exact test must decide: the row ends when a pixel centre lies within the margin
                    // of an edge; the whole bounding-box row when the row is in doubt (see SpanSetupB::doubt)
                    if (((sb.doubt >> r) & 1u) | (fs - lo < sb.din) | (hi - fe < sb.din)) {
                        exact_rows |= 1u << r;
                    } else if (fs <= fe) {
                        const int xs = (int)fs, xe = (int)fe;
                        const double* q0 = E0 + rowbase + xs;
                        const double* q1 = E0 + rowbase + xe + 1;
                        const double d0 = q1[0] - q0[0];
                        const double d1 = (q1[TILE * PREFIX_STRIDE] - q0[TILE * PREFIX_STRIDE]) - axd * d0;     // sum g*(x' - ax')
                        const float s0 = (float)d0;
                        M0 += s0; Ss += (float)d1; St = fmaf(t, s0, St);
                        nfrag += (unsigned)(xe - xs + 1);
                    }
                }
            }
        }
        if (__any_sync(FULL, exact_rows != 0u)) {       // about one batch in five has such a row
#pragma unroll 1
            while (exact_rows) {
                const int r = __ffs(exact_rows) - 1;
                exact_rows &= exact_rows - 1u;
                const float t = t0 + (float)r;
                const bool doubt = (sb.doubt >> r) & 1u;
                const float lo = fmaxf(fmaf(sb.sp.RL0, t, cl0), fmaf(sb.sp.RL1, t, cl1));
                const float hi = fminf(fmaf(sb.sp.RU0, t, cu0), fmaf(sb.sp.RU1, t, cu1));
                const float fs = doubt ? xlo : ceilf(fmaxf(lo, xlo)), fe = doubt ? xhi : floorf(fminf(hi, xhi));
                if (fs <= fe) {
                    // certain columns: [ceil(lo + din), floor(hi - din)], none when the row is in doubt
                    const int xsi = doubt ? 32 : __float2int_ru(lo + sb.din), xei = doubt ? -1 : __float2int_rd(hi - sb.din);
                    const int ry = cy0 - ty0 + r;
                    const float4* vtrow = GATHER ? P.vt + (long long)b * P.V : nullptr;
                    const void* src = GATHER ? static_cast<const void*>(P.faces + 3 * (long long)f) : static_cast<const void*>(P.p6 + bf * 6);
                    const float4 e = bwd_row_exact(src, vtrow, S.acc + ry * TILE, (int)fs, (int)fe, xsi, xei, tx0, ty0 + ry);
                    M0 += e.x; A1x += e.y; A2x += e.z; nfrag += (unsigned)__float_as_int(e.w);
                }
            }
        }
    }
    st.cand += nfrag; st.frag += nfrag;
    if (queued && (M0 != 0.0f || Ss != 0.0f || St != 0.0f || A1x != 0.0f || A2x != 0.0f)) {
        float ax, ay, bx, by, cx, cy, r0, r1, r2;
        load_face<GATHER>(P, b, f, bf, ax, ay, bx, by, cx, cy, r0, r1, r2);
        const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
        const int2 lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
        float sgn = (__ldg(reinterpret_cast<const short*>(P.face_box + bf)) < 0) ? -1.0f : 1.0f;   // diff_fp_for.cu:94-96 (unpack_box)
        if (lab.x == lab.y) sgn = -sgn;                               // :167-169
        // the k1 / k2 moments from sum g*(x - ax), sum g*(y - ay)  (k1 = q*s - n*t, k2 = m*t - p*s)
        const float A1 = fmaf(es.q, Ss, -(es.n * St)) + A1x, A2 = fmaf(es.m, St, -(es.p * Ss)) + A2x;
        const float mu_in = __ldg(P.mus + lab.x), mu_out = __ldg(P.mus + lab.y);
        const float cgeo = sgn * (mu_in - mu_out);                    // raw mu, no label-0 flip (back.cu:115)
        float gp[6], gl[3], bsum;
        face_grads_from_moments(es, r0, r1, r2, cgeo, M0, A1, A2, gp, gl, bsum);
        // grad_mus (back.cu:143-147): sign*(+-)bary*g for the inward label, opposite for the outward
        st.mus_add(gmus, lab.x, (lab.x == 0 ? -sgn : sgn) * bsum, lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
        if (P.grad_vt) {
            // one vector RED per corner onto the corner's vertex (back.cu:126-136 and :285-292 merged)
            const int i0 = __ldg(P.faces + 3 * f), i1 = __ldg(P.faces + 3 * f + 1), i2 = __ldg(P.faces + 3 * f + 2);
            float4* gv = P.grad_vt + (long long)b * P.V;
            atomicAdd(gv + i0, make_float4(gp[0], gp[1], gl[0], 0.0f));
            atomicAdd(gv + i1, make_float4(gp[2], gp[3], gl[1], 0.0f));
            atomicAdd(gv + i2, make_float4(gp[4], gp[5], gl[2], 0.0f));
        } else {
            float* dl = P.grad_len3 + bf * 3;                         // back.cu:126-136
            atomicAdd(dl + 0, gl[0]); atomicAdd(dl + 1, gl[1]); atomicAdd(dl + 2, gl[2]);
            float* dp = P.grad_p6 + bf * 6;                           // back.cu:285-292
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(dp + k, gp[k]);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Backward batch of the replay kernel: lane per face, coverage READ from the batch record.
//
// The forward sweep recorded, per face and tile row, the interval of pixels its own exact test accepted (BREC_SPANS).
// With the row prefix sums E0 = sum g, E1 = sum g*x' of the tile, a (face, row) pair is one byte, four shared-memory
// loads and five float64 operations: no edge functions, no margins, no exact re-tests — the set of fragments is the
// forward's by construction.  The moments are taken about the TILE origin in float64 (sum g, sum g*x', sum g*y') and
// moved to the face's first corner afterwards, so nothing of the face is needed before the rows are done.
// ------------------------------------------------------------------------------------------
// rare: a face without a usable area (|k3| outside (1e-6, 1e10), e.g. three collinear corners): its moments are formed
// pixel by pixel with k1, k2 exactly as the forward evaluates them — they are exact zeros where the reference's are,
// before the 1/(k3^2 + eps) factor gets to amplify rounding
__device__ __noinline__ float2 bwd_face_exact_moments(const double* E0, const unsigned char* spans, int r0, int nr,
                                                      float ax, float ay, float m, float p, float n, float q, int tx0, int ty0) {
    float A1 = 0.0f, A2 = 0.0f;
    for (int k = 0; k < nr; ++k) {
        const unsigned byte = spans[r0 + k];
        const int xs = (int)(byte >> 4), xe = (int)(byte & 15u);
        const double* row = E0 + (r0 + k) * PREFIX_STRIDE;
        const float t = __fsub_rn((float)(ty0 + r0 + k), ay);
        for (int x = xs; x <= xe; ++x) {
            const float g = (float)(row[x + 1] - row[x]);
            const float sx = __fsub_rn((float)(tx0 + x), ax);
            A1 = fmaf(g, __fmaf_rn(sx, q, -__fmul_rn(n, t)), A1);
            A2 = fmaf(g, __fmaf_rn(m, t, -__fmul_rn(sx, p)), A2);
        }
    }
    return make_float2(A1, A2);
}

// L: the batch record in shared memory; nb faces
template <bool GATHER, class ThreadStats>
__device__ __forceinline__ void process_batch_spans(const RasterParams& P, ReplayWarpSmem& S, const int* L, float* gmus,
                                                    int b, int nb, int tx0, int ty0, int lane, ThreadStats& st) {
    const bool queued = lane < nb;
    const unsigned char* spans = reinterpret_cast<const unsigned char*>(L + BREC_SPANS) + lane * TILE;
    int r0 = 0, nr = 0;
    if (queued) {
        const unsigned info = reinterpret_cast<const unsigned char*>(L + BREC_INFO)[lane];
        r0 = (int)(info & 15u); nr = (int)(info >> 4) + 1;
    }
    double M0 = 0.0, Mx = 0.0, My = 0.0;                // sum g, sum g*x', sum g*y' over the face's fragments (tile-relative)
    typename ThreadStats::count_t nfrag{};
    {
        const double* E0 = prefix_e0(S);
        const int nrmax = __reduce_max_sync(FULL, nr);
        const unsigned char* sp = spans + r0;
        int rowbase = r0 * PREFIX_STRIDE;
        double yd = (double)r0;
#pragma unroll 1
        for (int k = 0; k < nrmax; ++k, rowbase += PREFIX_STRIDE, yd += 1.0) {
            if (k < nr) {
                const unsigned byte = sp[k];
                const int xs = (int)(byte >> 4), xe = (int)(byte & 15u);
                if (xs <= xe) {
                    const double* q0 = E0 + rowbase + xs;
                    const double* q1 = E0 + rowbase + xe + 1;
                    const double d0 = q1[0] - q0[0];
                    const double d1 = q1[TILE * PREFIX_STRIDE] - q0[TILE * PREFIX_STRIDE];
                    M0 += d0; Mx += d1; My = fma(yd, d0, My);
                    nfrag += (unsigned)(xe - xs + 1);
                }
            }
        }
    }
    st.cand += nfrag; st.frag += nfrag;
    if (queued && (M0 != 0.0 || Mx != 0.0 || My != 0.0)) {
        const int idw = L[2 + lane];
        const int f = idw & 0x7fffffff;
        const long long bf = (long long)b * P.F + f;
        float ax, ay, bx, by, cx, cy, r0l, r1l, r2l;
        load_face<GATHER>(P, b, f, bf, ax, ay, bx, by, cx, cy, r0l, r1l, r2l);
        const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
        const int2 lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
        float sgn = (idw < 0) ? -1.0f : 1.0f;                         // diff_fp_for.cu:94-96 (facing bit of the record)
        if (lab.x == lab.y) sgn = -sgn;                               // :167-169
        // moments about the face's first corner: sum g*(x - ax), sum g*(y - ay)
        const float m0 = (float)M0;
        const float Ss = (float)(Mx - ((double)es.ax - (double)tx0) * M0);
        const float St = (float)(My - ((double)es.ay - (double)ty0) * M0);
        // the k1 / k2 moments (k1 = q*s - n*t, k2 = m*t - p*s)
        float A1 = fmaf(es.q, Ss, -(es.n * St)), A2 = fmaf(es.m, St, -(es.p * Ss));
        const float ak3 = fabsf(es.k3);
        if (!((ak3 > 1e-6f) && (ak3 < 1e10f))) {
            const float2 a = bwd_face_exact_moments(prefix_e0(S), spans, r0, nr, es.ax, es.ay, es.m, es.p, es.n, es.q, tx0, ty0);
            A1 = a.x; A2 = a.y;
        }
        const float mu_in = __ldg(P.mus + lab.x), mu_out = __ldg(P.mus + lab.y);
        const float cgeo = sgn * (mu_in - mu_out);                    // raw mu, no label-0 flip (back.cu:115)
        float gp[6], gl[3], bsum;
        face_grads_from_moments(es, r0l, r1l, r2l, cgeo, m0, A1, A2, gp, gl, bsum);
        // grad_mus (back.cu:143-147): sign*(+-)bary*g for the inward label, opposite for the outward
        st.mus_add(gmus, lab.x, (lab.x == 0 ? -sgn : sgn) * bsum, lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
        if (P.grad_vt) {
            // one vector RED per corner onto the corner's vertex (back.cu:126-136 and :285-292 merged)
            const int i0 = __ldg(P.faces + 3 * f), i1 = __ldg(P.faces + 3 * f + 1), i2 = __ldg(P.faces + 3 * f + 2);
            float4* gv = P.grad_vt + (long long)b * P.V;
            atomicAdd(gv + i0, make_float4(gp[0], gp[1], gl[0], 0.0f));
            atomicAdd(gv + i1, make_float4(gp[2], gp[3], gl[1], 0.0f));
            atomicAdd(gv + i2, make_float4(gp[4], gp[5], gl[2], 0.0f));
        } else {
            float* dl = P.grad_len3 + bf * 3;                         // back.cu:126-136
            atomicAdd(dl + 0, gl[0]); atomicAdd(dl + 1, gl[1]); atomicAdd(dl + 2, gl[2]);
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
// Residual/backward sweep of the residual step, replay only: one CTA per (angle, region) like k_raster, but all
// it needs is what the forward sweep recorded — per tile the chain of face batches (RasterParams::brec).  No region
// lists, no chunk walk, no face queue, none of the forward's shared memory: a small kernel whose registers go to
// the row loop of process_batch_rows.  Regions the forward could not record (bflag 0) are left to
// k_raster<MODE_STEP> launched right after it, which skips the recorded ones.
// ------------------------------------------------------------------------------------------
struct ReplayCtaSmem {
    int next_tile;
    ReplayWarpSmem warp[WARPS];
};

__device__ __forceinline__ void prefetch_record(const RasterParams& P, int* L, int off, int lane) {
    const int4* r = reinterpret_cast<const int4*>(P.brec + (long long)off * BREC_INTS);
    int4* l = reinterpret_cast<int4*>(L);
    static_assert(BREC_INTS % 4 == 0 && BREC_INTS / 4 <= 64, "record = at most two 16-byte pieces per lane");
    cp_async16(l + lane, r + lane);
    if (lane + 32 < BREC_INTS / 4) cp_async16(l + lane + 32, r + lane + 32);
}

// 8 resident CTAs (64 registers); 9 CTAs at 56 registers was measured slower (5.06 -> 6.19 ms on config 3).  Also measured
// and rejected: four single-warp CTAs per region (constant shared-memory addresses, no barrier, no tile counter:
// 5.02 -> 5.08 ms); persistent warps pulling tile rows of any region from a device-wide queue (5.13 -> 5.67 ms: the four
// warps of a CTA then work on four angles and thrash the 28 KB of L1); persistent CTAs taking (angle, region) items from
// a device-wide queue with a rolling CTA-local tile ticket, so that no warp waits for the slowest warp of its CTA
// (5.11 -> 5.79 ms, and 0.75 -> 0.82 ms on a 90-angle shard) — the hardware's CTA scheduler beats all three, although
// regions of 4 x 4 instead of 8 x 8 tiles cost this kernel 14 % (per-CTA wind-down) and cheaper CTAs were the aim.
template <bool STATS, bool GATHER>
__global__ void __launch_bounds__(THREADS, 8)
k_step_replay(const RasterParams P) {
    __shared__ ReplayCtaSmem C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* gmus = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    asm volatile("mov.b32 %0, %0;" : "+r"(lane));            // pinned, see k_raster
    // this kernel's regions: rts x rts tiles = sub x sub regions of the forward sweep (sub = 1 or 2)
    const int RTS = P.rts, sub = RTS / P.RT;
    const int sregions_x = (P.W + RTS * TILE - 1) / (RTS * TILE), sregions_y = (P.H + RTS * TILE - 1) / (RTS * TILE);
    const int regions = sregions_x * sregions_y;
    const int b = P.region_major ? (int)(blockIdx.x % (unsigned)P.B) : (int)(blockIdx.x / (unsigned)regions);
    const int rr = region_of_rank(P, P.region_major ? (int)(blockIdx.x / (unsigned)P.B) : (int)(blockIdx.x % (unsigned)regions), sregions_x, sregions_y);
    // what the forward sweep left for its (up to four) regions in here, one byte each: 0 = not recorded (k_raster<MODE_STEP>
    // walks it), 1 = recorded, 2 = nothing projects into it; 0xff = outside the detector
    unsigned flags4 = 0xffffffffu;
    {
        const unsigned char* fl = P.bflag + (long long)b * (P.regions_x * P.regions_y);
        const int fx0 = (rr % sregions_x) * sub, fy0 = (rr / sregions_x) * sub;
        for (int j = 0; j < sub; ++j)
            for (int i = 0; i < sub; ++i)
                if (fx0 + i < P.regions_x && fy0 + j < P.regions_y) {
                    const unsigned v = fl[(fy0 + j) * P.regions_x + fx0 + i];
                    flags4 = (flags4 & ~(0xffu << (8 * (2 * j + i)))) | (v << (8 * (2 * j + i)));
                }
    }
    // over the regions present: none recorded or empty -> nothing to do here; all empty -> one stream over the target;
    // all recorded -> tiles; a mixture -> tiles that look at their own region's flag
    const bool has0 = __vcmpeq4(flags4, 0u) != 0u, has1 = __vcmpeq4(flags4, 0x01010101u) != 0u, has2 = __vcmpeq4(flags4, 0x02020202u) != 0u;
    if (!has1 && !has2) return;
    const bool stream_all = !has0 && !has1;
    const bool mixed = (int)has0 + (int)has1 + (int)has2 > 1;
    const int rx0 = (rr % sregions_x) * RTS * TILE, ry0 = (rr / sregions_x) * RTS * TILE;
    const int rx1 = min(rx0 + RTS * TILE, P.W) - 1, ry1 = min(ry0 + RTS * TILE, P.H) - 1;
    unsigned sbase = (unsigned)__cvta_generic_to_shared(&C.warp[warp]);     // pinned in a register, see k_raster
    asm volatile("mov.b32 %0, %0;" : "+r"(sbase));
    ReplayWarpSmem& S = *reinterpret_cast<ReplayWarpSmem*>(__cvta_shared_to_generic(sbase));
    ThreadStatsT<STATS> st;
    double loss_acc = 0.0;
    unsigned nvalid_acc = 0u;
    for (int k = tid; k < P.n_mus; k += THREADS) gmus[k] = 0.0f;
    if (tid == 0) C.next_tile = 0;
    __syncthreads();
    const long long base = (long long)b * P.H * P.W;
    if (stream_all) {
        // nothing projects into this region: phat == 0 everywhere, residual = -target, no gradient
        if (0.0f <= P.mask_hi) {
            const int rw = rx1 - rx0 + 1, rh = ry1 - ry0 + 1;
            const bool vec = ((P.W & 3) == 0) && ((rw & 3) == 0);
            if (vec) {
                // a pure stream of the target: several rows per lane in flight (float4 = four pixels)
                const int rw4 = rw >> 2, n4 = rw4 * rh;
                const float* org = P.target + base + (long long)ry0 * P.W + rx0;
#pragma unroll 4
                for (int k = tid; k < n4; k += THREADS) {
                    const int y = k / rw4, x = (k - y * rw4) * 4;
                    const float4 t = __ldg(reinterpret_cast<const float4*>(org + (long long)y * P.W + x));
                    loss_acc += (double)(t.x * t.x + t.y * t.y) + (double)(t.z * t.z + t.w * t.w);
                }
            } else {
                for (int y = warp; y < rh; y += WARPS) {
                    const float* row = P.target + base + (long long)(ry0 + y) * P.W + rx0;
                    for (int x = lane; x < rw; x += 32) { const float t = __ldg(row + x); loss_acc += (double)t * (double)t; }
                }
            }
            if (tid == 0) nvalid_acc += (unsigned)(rw * rh);
        }
    } else {
        const int tiles_x = (P.W + TILE - 1) / TILE, tiles_y = (P.H + TILE - 1) / TILE;
        const int ntiles = RTS * RTS;
        while (true) {
            int tix = 0;
            if (lane == 0) tix = atomicAdd(&C.next_tile, 1);
            tix = __shfl_sync(FULL, tix, 0);
            if (tix >= ntiles) break;
            const int tcol = tix % RTS, trow = tix / RTS;
            const int tx0 = rx0 + tcol * TILE, ty0 = ry0 + trow * TILE;
            if (tx0 > rx1 || ty0 > ry1) continue;
            const int tx1 = min(tx0 + TILE - 1, rx1), ty1 = min(ty0 + TILE - 1, ry1);
            int roff0 = -1;
            if (mixed) {
                // the forward region of this tile: unrecorded -> the walking kernel's; empty -> no head was written
                const unsigned f = (flags4 >> (8 * (2 * (trow / P.RT) + tcol / P.RT))) & 0xffu;
                if (f == 0u) continue;
                if (f == 1u) roff0 = __ldg(P.bhead + ((long long)b * tiles_y + (ty0 >> 4)) * tiles_x + (tx0 >> 4));
            } else {
                roff0 = __ldg(P.bhead + ((long long)b * tiles_y + (ty0 >> 4)) * tiles_x + (tx0 >> 4));
            }
            if (roff0 >= 0) prefetch_record(P, S.landing[0], roff0, lane);  // in flight during the tile prologue
            // residual prologue: g = d/dphat of sum_valid (target - phat)^2 = 2*(phat - target) on pixels passing
            // get_valid_mask (fp_vecs.py:10-13; a subset of the kernel's own gradient predicate,
            // diff_fp_for.cu:69-73); loss and count once per tile
            // lane = (row, half row): eight consecutive pixels, as two 16-byte loads per array on whole tiles — the layout
            // the row prefix sums are built in, so g never goes through shared memory
            const int prow = lane >> 1, pcol = (lane & 1) * 8;
            float vv[8], tt[8];
            {
                const bool rowok = ty0 + prow <= ty1;
                const long long o = base + (long long)(ty0 + prow) * P.W + tx0 + pcol;
                if (tx1 - tx0 == TILE - 1 && (P.W & 3) == 0) {
                    float4 va = make_float4(-1.f, -1.f, -1.f, -1.f), vb = va, ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
                    if (rowok) {
                        va = __ldg(reinterpret_cast<const float4*>(P.im_in + o)); vb = __ldg(reinterpret_cast<const float4*>(P.im_in + o) + 1);
                        ta = __ldg(reinterpret_cast<const float4*>(P.target + o)); tb = __ldg(reinterpret_cast<const float4*>(P.target + o) + 1);
                    }
                    vv[0] = va.x; vv[1] = va.y; vv[2] = va.z; vv[3] = va.w; vv[4] = vb.x; vv[5] = vb.y; vv[6] = vb.z; vv[7] = vb.w;
                    tt[0] = ta.x; tt[1] = ta.y; tt[2] = ta.z; tt[3] = ta.w; tt[4] = tb.x; tt[5] = tb.y; tt[6] = tb.z; tt[7] = tb.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool ok = rowok && (tx0 + pcol + k <= tx1);
                        vv[k] = ok ? __ldg(P.im_in + o + k) : -1.0f;
                        tt[k] = ok ? __ldg(P.target + o + k) : 0.0f;
                    }
                }
            }
            // Pipeline: while batch k is processed out of one landing zone, record k+1 is copied into the other by
            // cp.async (16-byte pieces, no register holds the data in flight).  A record carries everything the row
            // phase needs — face ids, rows and the exact coverage — so the only global loads left in the batch are the
            // per-face ones of the gradient epilogue.
            int slot = 0;
            float g[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                g[k] = 0.0f;
                if (vv[k] > -1e-6f && vv[k] <= P.mask_hi) {
                    const float diff = vv[k] - tt[k];
                    g[k] = 2.0f * diff;
                    loss_acc += (double)diff * (double)diff; nvalid_acc++;
                }
            }
            if (roff0 < 0) continue;                                  // no face reaches this tile
            build_row_prefix_regs(S, g, lane);
            while (true) {
                cp_async_wait_all();                                  // record k (issued during the prologue / a whole batch ago)
                __syncwarp();
                const int* L = S.landing[slot];
                const int next = L[0], word = L[1];
                slot ^= 1;
                if (next >= 0) prefetch_record(P, S.landing[slot], next, lane);
                if (!(word & BREC_FLAG_EMPTY)) process_batch_spans<GATHER>(P, S, L, gmus, b, word & 0xff, tx0, ty0, lane, st);
                if (next < 0) break;
            }
        }
    }
    st.mus_flush(gmus);
    __syncthreads();
    for (int k = tid; k < P.n_mus; k += THREADS) {
        const float v = gmus[k];
        if (v != 0.0f) atomicAdd(P.grad_mus + k, v);
    }
    // masked L2 data term of ctdr/optimize.py:168: sum over valid pixels and their count
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) loss_acc += __shfl_down_sync(FULL, loss_acc, d);
    nvalid_acc = __reduce_add_sync(FULL, nvalid_acc);
    if (lane == 0) {
        if (loss_acc != 0.0) atomicAdd(P.out2 + 0, loss_acc);
        if (nvalid_acc) atomicAdd(P.out2 + 1, (double)nvalid_acc);
    }
}

// ------------------------------------------------------------------------------------------
// Residual step, leftovers of the recorded coverage: one thread per listed (angle, face, tile) (RasterParams::xlist) walks
// the face's bounding box inside the tile with the reference's exact test and the residual gradient formed on the fly —
// the single-writer kernel's walk (k_backward_faceowner) restricted to one tile, adding with atomics.  The tile's loss
// and valid count are the replay kernel's business.
// ------------------------------------------------------------------------------------------
template <bool GATHER>
__global__ void __launch_bounds__(128)
k_step_facelist(const RasterParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(*P.xcount, XLIST_CAP);
    if (i >= n) return;
    const int4 e = P.xlist[i];
    const int b = e.x, f = e.y, tx0 = e.z, ty0 = e.w;
    // a region that ran out of record space after this entry was listed is walked as a whole by k_raster<MODE_STEP>
    const int rr = (ty0 / (P.RT * TILE)) * P.regions_x + tx0 / (P.RT * TILE);
    if (P.bflag[(long long)b * (P.regions_x * P.regions_y) + rr] == 0) return;
    const long long bf = (long long)b * P.F + f;
    bool fneg;
    const short4 fb = unpack_box(__ldg(P.face_box + bf), fneg);
    const int x0 = max((int)fb.x, tx0), x1 = min((int)fb.z, min(tx0 + TILE, P.W) - 1);
    const int y0 = max((int)fb.y, ty0), y1 = min((int)fb.w, min(ty0 + TILE, P.H) - 1);
    float ax, ay, bx, by, cx, cy, r0, r1, r2;
    load_face<GATHER>(P, b, f, bf, ax, ay, bx, by, cx, cy, r0, r1, r2);
    const EdgeSetup es = edge_setup(ax, ay, bx, by, cx, cy);
    float M0 = 0.0f, A1 = 0.0f, A2 = 0.0f;
    const long long base = (long long)b * P.H * P.W;
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            const long long o = base + (long long)y * P.W + x;
            const float v = __ldg(P.im_in + o);
            if (!(v > -1e-6f && v <= P.mask_hi)) continue;                       // get_valid_mask (fp_vecs.py:10-13)
            const float g = 2.0f * (v - __ldg(P.target + o));
            if (g == 0.0f) continue;
            float s, t, k1, k2, w0, w1, w2;
            if (!cover_exact<true>(es, (float)x, (float)y, s, t, k1, k2, w0, w1, w2)) continue;
            M0 += g; A1 = fmaf(g, k1, A1); A2 = fmaf(g, k2, A2);
        }
    if (M0 == 0.0f && A1 == 0.0f && A2 == 0.0f) return;
    const int2 lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
    float sgn = fneg ? -1.0f : 1.0f;                                              // diff_fp_for.cu:94-96
    if (lab.x == lab.y) sgn = -sgn;                                               // :167-169
    const float cgeo = sgn * (__ldg(P.mus + lab.x) - __ldg(P.mus + lab.y));       // raw mu, no label-0 flip (back.cu:115)
    float gp[6], gl[3], bsum;
    face_grads_from_moments(es, r0, r1, r2, cgeo, M0, A1, A2, gp, gl, bsum);
    atomicAdd(P.grad_mus + lab.x, (lab.x == 0 ? -sgn : sgn) * bsum);              // back.cu:143-147
    atomicAdd(P.grad_mus + lab.y, (lab.y == 0 ? sgn : -sgn) * bsum);
    if (P.grad_vt) {
        const int i0 = __ldg(P.faces + 3 * f), i1 = __ldg(P.faces + 3 * f + 1), i2 = __ldg(P.faces + 3 * f + 2);
        float4* gv = P.grad_vt + (long long)b * P.V;
        atomicAdd(gv + i0, make_float4(gp[0], gp[1], gl[0], 0.0f));
        atomicAdd(gv + i1, make_float4(gp[2], gp[3], gl[1], 0.0f));
        atomicAdd(gv + i2, make_float4(gp[4], gp[5], gl[2], 0.0f));
    } else {
        float* dl = P.grad_len3 + bf * 3;
        atomicAdd(dl + 0, gl[0]); atomicAdd(dl + 1, gl[1]); atomicAdd(dl + 2, gl[2]);
        float* dp = P.grad_p6 + bf * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) atomicAdd(dp + k, gp[k]);
    }
}

// ------------------------------------------------------------------------------------------
// the raster kernel: one CTA per (angle, region)
// ------------------------------------------------------------------------------------------
// LINKED: the two sweeps of the residual step are linked by batch records.  MODE_FWD + LINKED records the face
// batches of every tile; MODE_STEP + LINKED is a replay-only kernel (no region lists, no queues) that skips
// the regions the forward sweep could not record — those are done by the ordinary MODE_STEP kernel, launched
// right after it with bmode == 2, which in turn skips the recorded ones.
// Residency: the forward sweep of the fused path (NOCNT: 23.6 KB of shared memory per CTA) fits NINE CTAs per SM as long as it
// stays within 56 registers — measured three times in round 2: every edit that cost it a 57th register (short batch records,
// a region weight in the record flags) dropped it to eight CTAs and cost 4-6 % of the sweep.  The launch bound pins that.
template <int MODE, bool EXACT_DIV, bool DIRECT = false, bool STATS = true, bool LINKED = false, bool GATHER = false, bool NOCNT = false>
__global__ void __launch_bounds__(THREADS, (NOCNT && !STATS) ? 9 : 8)
k_raster(const RasterParams P) {
    constexpr bool RECORD = LINKED && MODE == MODE_FWD;
    constexpr bool REPLAY = LINKED && MODE == MODE_STEP;
    // fixed-size state in STATIC shared memory (addresses are link-time constants: no base pointer to keep
    // or rebuild in the hot loops); the dynamic part holds grad_mus accumulators / MODE_MULTI extras
    static_assert(!NOCNT || MODE == MODE_FWD, "only the forward sweep runs without the count tile");
    typedef WarpSmemT<!NOCNT> WS;
    __shared__ CtaSmemT<!NOCNT> C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* gmus = reinterpret_cast<float*>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5;
    int lane = tid & 31;
    asm volatile("mov.b32 %0, %0;" : "+r"(lane));            // pinned like the shared-memory base below (S2R is slow to rebuild)
    const int regions = P.regions_x * P.regions_y;
    // CTA order: angle-major (neighbouring regions of one angle run together and share its face data in
    // L2), or region-major (all angles of a region together: the last CTAs of the launch are then the
    // detector's bottom rows, usually empty — no heavy CTA is left running alone at the end of a small grid)
    const int b = P.region_major ? (int)(blockIdx.x % (unsigned)P.B) : (int)(blockIdx.x / (unsigned)regions);
    const int rr = region_of_rank(P, P.region_major ? (int)(blockIdx.x / (unsigned)P.B) : (int)(blockIdx.x % (unsigned)regions));
    const int rx0 = (rr % P.regions_x) * P.RT * TILE, ry0 = (rr / P.regions_x) * P.RT * TILE;
    const int rx1 = min(rx0 + P.RT * TILE, P.W) - 1, ry1 = min(ry0 + P.RT * TILE, P.H) - 1;
    const short4* boxes = P.chunk_box + (long long)b * P.NC;
    const int srx = (rr % P.regions_x) / P.sr_rx, sry = (rr / P.regions_x) / P.sr_ry;
    const long long sidx = (long long)b * (P.nsr_x * P.nsr_y) + sry * P.nsr_x + srx;
    const int* super = P.slist + sidx * P.NC;
    const int nsuper = __ldg(P.scount + sidx);
    // the warp's shared-memory base is pinned in a register (opaque to the compiler): at the 64-register cap ptxas
    // otherwise rebuilds it from %tid / %cluster_ctarank (two S2R + four ALU instructions) at dozens of places in the hot loops
    unsigned sbase = (unsigned)__cvta_generic_to_shared(&C.warp[warp]);
    asm volatile("mov.b32 %0, %0;" : "+r"(sbase));
    WS& S = *reinterpret_cast<WS*>(__cvta_shared_to_generic(sbase));
    ThreadStatsT<STATS> st;

    constexpr bool K_BWD = (MODE == MODE_BWD || MODE == MODE_STEP);
    double loss_acc = 0.0;
    unsigned nvalid_acc = 0u;
    // MODE_MULTI: after CtaSmem come the per-warp running sums of p_i*p_j (i <= j, row-major upper
    // triangle) and p_i*target, then the tile accumulators of outputs 2 and 3
    double* gwarp = reinterpret_cast<double*>(gmus) + warp * NGRAM;
    float* xacc = multi_xacc(gmus, warp);
    if (MODE == MODE_MULTI && lane < NGRAM) gwarp[lane] = 0.0;
    if (K_BWD) {
        for (int k = tid; k < P.n_mus; k += THREADS) gmus[k] = 0.0f;
    }

    int c = 0;
    int seg = 0;
    // residual/backward sweep: what the forward sweep recorded for this region (0: nothing, walk as usual)
    int bflag = 0;
    if (MODE == MODE_STEP && (REPLAY || P.bmode == 2)) {
        bflag = (int)P.bflag[(long long)b * regions + rr];
        if (REPLAY ? bflag == 0 : bflag != 0) return;         // the other kernel of the pair owns this region
    }
    constexpr bool replay = REPLAY;
    const int tiles_x = (P.W + TILE - 1) / TILE, tiles_y = (P.H + TILE - 1) / TILE;
    if (RECORD && tid == 0) C.rec_fail = 0;
    if (REPLAY) c = nsuper;                                   // no region list needed
    do {
        // ---- phase A: ordered compaction of the chunks overlapping this region -------------
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
            if (n0 + total > LIST_CAP) break;              // uniform: segment is full, resume at c later
            if (hit) {
                const int pos = n0 + mybase + __popc(m & lanemask_lt());
                C.list_id[pos] = ci;
            }
            __syncthreads();
            if (tid == 0) C.list_n = n0 + total;
            c += THREADS;
        }
        __syncthreads();
        const int nlist = C.list_n;
        const bool first = (seg == 0), last = (c >= nsuper);

        constexpr bool recording = RECORD;
        if (nlist == 0 && first && last && !(REPLAY && bflag == 1)) {
            if (RECORD && tid == 0) P.bflag[(long long)b * regions + rr] = 2;
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
            } else if (MODE == MODE_MULTI) {
                for (int o = 0; o < P.n_out; ++o) {
                    float* dst = P.im + (long long)o * P.out_stride + base;
                    for (int k = tid; k < rw * rh; k += THREADS) {
                        const int y = k / rw, x = k - y * rw;
                        dst[(long long)(ry0 + y) * P.W + rx0 + x] = 0.0f;
                    }
                }
            } else if (MODE == MODE_STEP) {
                if (0.0f <= P.mask_hi) {        // phat == 0 everywhere here: residual = -target
                    const bool vec = ((P.W & 3) == 0) && ((rw & 3) == 0);
                    for (int y = warp; y < rh; y += WARPS) {
                        const float* row = P.target + base + (long long)(ry0 + y) * P.W + rx0;
                        if (vec) {
                            for (int x = lane * 4; x < rw; x += 128) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(row + x));
                                loss_acc += (double)(t.x * t.x + t.y * t.y) + (double)(t.z * t.z + t.w * t.w);
                            }
                        } else {
                            for (int x = lane; x < rw; x += 32) { const float t = __ldg(row + x); loss_acc += (double)t * (double)t; }
                        }
                    }
                    if (tid == 0) nvalid_acc += (unsigned)(rw * rh);
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
            const unsigned tcol = (unsigned)(tix % P.RT), trow = (unsigned)(tix / P.RT);
            const int tx0 = rx0 + (int)tcol * TILE, ty0 = ry0 + (int)trow * TILE;
            if (tx0 > rx1 || ty0 > ry1) continue;
            const int tx1 = min(tx0 + TILE - 1, rx1), ty1 = min(ty0 + TILE - 1, ry1);
            // batch records of this tile: replaying, `roff` is the next record; recording, the previous one
            // (the index is recomputed where needed instead of being held across the batch loop)
#define CTDR_HEAD_INDEX (((long long)b * tiles_y + (ty0 >> 4)) * tiles_x + (tx0 >> 4))
            int roff = -1;
            if (replay) {
                roff = __ldg(P.bhead + CTDR_HEAD_INDEX);
                if (roff >= 0) prefetch_record(P, S, roff, lane);      // in flight during the tile prologue
            }
            if (recording) {
                // first segment: no chain yet (the head is overwritten by the first record); later segments
                // continue from the tail the previous one left (plain load: written by another warp of this
                // CTA before the last barrier)
                if (first) { if (lane == 0) P.bhead[CTDR_HEAD_INDEX] = -1; }
                else roff = *reinterpret_cast<volatile int*>(P.btail + CTDR_HEAD_INDEX);
            }

            // tile prologue
            if (MODE == MODE_FWD) {
#pragma unroll
                for (int k = 0; k < TILE_PIX / 32; ++k) S.f.hitm[lane + 32 * k] = 0u;
                if (first) {
#pragma unroll
                    for (int k = 0; k < TILE_PIX / 32; ++k) { S.acc[lane + 32 * k] = 0.0f; if (WS::HAS_CNT) S.cntv[lane + 32 * k] = 0; }
                } else {
                    tile_load<float>(S.acc, P.im, 0.0f, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                    if (WS::HAS_CNT && P.cnt) tile_load<int>(S.cntv, P.cnt, 0, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                }
            } else if (MODE == MODE_MULTI) {
#pragma unroll
                for (int k = 0; k < TILE_PIX / 32; ++k) S.f.hitm[lane + 32 * k] = 0u;
                for (int o = 0; o < P.n_out; ++o) {
                    float* a = multi_acc(S, xacc, o);
                    if (first) {
#pragma unroll
                        for (int k = 0; k < TILE_PIX / 32; ++k) a[lane + 32 * k] = 0.0f;
                    } else {
                        tile_load<float>(a, P.im + (long long)o * P.out_stride, 0.0f, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                    }
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
                float vv[TILE / 2], tt[TILE / 2];
#pragma unroll
                for (int k = 0; k < TILE / 2; ++k) {          // all 16 loads in flight before any use
                    const int row = r + 2 * k;
                    const bool ok = (ty0 + row <= ty1) && (tx0 + col <= tx1);
                    const long long o = base + (long long)(ty0 + row) * P.W + tx0 + col;
                    vv[k] = ok ? __ldg(P.im_in + o) : -1.0f;
                    tt[k] = ok ? __ldg(P.target + o) : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < TILE / 2; ++k) {
                    float g = 0.0f;
                    if (vv[k] > -1e-6f && vv[k] <= P.mask_hi) {
                        const float diff = vv[k] - tt[k];
                        g = 2.0f * diff;
                        if (first) { loss_acc += (double)diff * (double)diff; nvalid_acc++; }
                    }
                    S.acc[(r + 2 * k) * TILE + col] = g;
                }
            }
            __syncwarp();

            // walk the region's chunk list, 32 boxes per ballot, in face order; queue the faces whose
            // own box overlaps the tile and drain the queue 32 faces at a time (single call site: the
            // batch routine is large and must not be inlined twice — it would overflow the I-cache)
            int qn = 0, lb = 0, cid = 0;
            unsigned m = 0u;
            bool exhausted = false;
            bool prefix_ready = false;                      // backward: row prefix sums of g, built at the tile's first batch
            const short4* fbrow = P.face_box + (long long)b * P.F;
            while (true) {
                int nbatch;
                bool rec_now = false;                          // recording: this batch found record space
                if (replay) {
                    // the forward sweep already found this tile's faces and cut them into batches
                    if (roff < 0) break;
                    cp_async_wait_all();
                    __syncwarp();
                    const int* L = record_landing(S);
                    nbatch = L[1];
                    const int next = L[0];
                    const int id = L[2 + lane];
                    __syncwarp();                               // the landing zone may be overwritten now
                    if (lane < nbatch) S.qface[lane] = id;
                    roff = next;
                    if (roff >= 0) prefetch_record(P, S, roff, lane);   // lands while this batch is processed
                    __syncwarp();
                } else {
                while (qn < 32 && !exhausted) {
                    if (m == 0u) {
                        if (lb >= nlist) { exhausted = true; break; }
                        const int e = lb + lane;
                        bool hit = false;
                        if (e < nlist) {
                            const unsigned w = (unsigned)C.list_id[e];
                            cid = (int)(w & 0xfffffu);
                            hit = (tcol >= ((w >> 20) & 7u)) & (tcol <= ((w >> 23) & 7u)) & (trow >= ((w >> 26) & 7u)) & (trow <= (w >> 29));
                            if (hit) {      // the chunk's 32 face boxes (two lines) travel to L1 while earlier chunks are tested
                                prefetch_l1(fbrow + (long long)cid * CHUNK);
                                prefetch_l1(fbrow + (long long)cid * CHUNK + 16);
                            }
                        }
                        m = __ballot_sync(FULL, hit);
                        lb += 32;
                        continue;
                    }
                    // two hit chunks per turn: both face-box loads are in flight before either is tested
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int src2 = m ? __ffs(m) - 1 : -1;
                    const int f = __shfl_sync(FULL, cid, src) * CHUNK + lane;
                    const int f2 = __shfl_sync(FULL, cid, max(src2, 0)) * CHUNK + lane;
                    short4 fb = make_short4(BOX_EMPTY_LO, BOX_EMPTY_LO, -1, -1), fb2 = fb;
                    if (f < P.F) fb = unpack_box(__ldg(fbrow + f));
                    if (src2 >= 0 && f2 < P.F) fb2 = unpack_box(__ldg(fbrow + f2));
                    const bool fh = box_overlap(fb, tx0, ty0, tx1, ty1);
                    const unsigned hm = __ballot_sync(FULL, fh);
                    if (lane == 0) st.hits += (hm != 0u);
                    if (fh) S.qface[qn + __popc(hm & lanemask_lt())] = f;
                    qn += __popc(hm);
                    if (src2 >= 0 && qn < 32) {                // room for another 32: take the second chunk too
                        m &= m - 1;
                        const bool fh2 = box_overlap(fb2, tx0, ty0, tx1, ty1);
                        const unsigned hm2 = __ballot_sync(FULL, fh2);
                        if (lane == 0) st.hits += (hm2 != 0u);
                        if (fh2) S.qface[qn + __popc(hm2 & lanemask_lt())] = f2;
                        qn += __popc(hm2);
                    }
                    __syncwarp();
                }
                if (qn == 0) break;
                nbatch = min(qn, 32);
                if (recording) {
                    // one cursor and one slice of bcap records per angle: concurrent CTAs (region-major order:
                    // same region, different angles) hit different counters
                    int off = 0;
                    if (lane == 0) off = atomicAdd(P.bcursor + b, 1);
                    off = __shfl_sync(FULL, off, 0);
                    if (off < P.bcap) {
                        rec_now = true;
                        off += b * P.bcap;
                        int* r = P.brec + (long long)off * BREC_INTS;       // face ids, rows and coverage: process_batch_blocked
                        if (lane == 0) {
                            r[0] = -1; r[1] = nbatch;
                            if (roff >= 0) P.brec[(long long)roff * BREC_INTS] = off;     // link behind the previous record
                            else P.bhead[CTDR_HEAD_INDEX] = off;                          // first record of the tile
                        }
                        roff = off;
                    } else if (lane == 0) {
                        C.rec_fail = 1;
                    }
                }
                }
                if (MODE == MODE_FRAG) process_batch<MODE, EXACT_DIV>(P, S, gmus, b, nbatch, tx0, ty0, tx1, ty1, lane, st);
                else if (K_BWD) {
                    if (!prefix_ready) { build_row_prefix(S, lane); prefix_ready = true; }
                    process_batch_rows<GATHER>(P, S, gmus, b, nbatch, tx0, ty0, tx1, ty1, lane, st);
                } else process_batch_blocked<MODE, EXACT_DIV, DIRECT, GATHER, RECORD>(P, S, gmus, b, nbatch, tx0, ty0, tx1, ty1, lane, st,
                                                                                          rec_now ? roff : -1, &C.rec_fail);
                if (!replay) {
                    const int rest = qn - nbatch;              // < 32: shift the tail to the front
                    int tf = 0;
                    if (lane < rest) tf = S.qface[32 + lane];
                    __syncwarp();
                    if (lane < rest) S.qface[lane] = tf;
                    qn = rest;
                    __syncwarp();
                }
            }

            // tile epilogue
            if (MODE == MODE_FWD) {
                if (recording && !last && lane == 0) P.btail[CTDR_HEAD_INDEX] = roff;
                tile_store<float>(P.im, S.acc, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                if (WS::HAS_CNT && P.cnt) tile_store<int>(P.cnt, S.cntv, b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
            } else if (MODE == MODE_MULTI) {
                for (int o = 0; o < P.n_out; ++o)
                    tile_store<float>(P.im + (long long)o * P.out_stride, multi_acc(S, xacc, o), b, P.H, P.W, tx0, ty0, tx1, ty1, lane);
                if (last && P.gram) {
                    // normal-equation sums of Model.estimate_mu (vanilla.py:208-216,245-256) over the finished
                    // tile: lane partials over 8 pixels in float64, one warp reduction per sum
                    const long long base = (long long)b * P.H * P.W;
                    double gs[NGRAM];
#pragma unroll
                    for (int g = 0; g < NGRAM; ++g) gs[g] = 0.0;
#pragma unroll 1
                    for (int k = 0; k < TILE_PIX / 32; ++k) {
                        const int pix = lane + 32 * k, row = pix >> 4, col = pix & 15;
                        float v[MAX_OUT];
#pragma unroll
                        for (int o = 0; o < MAX_OUT; ++o) v[o] = (o < P.n_out) ? multi_acc(S, xacc, o)[pix] : 0.0f;
                        float tg = 0.0f;
                        if (P.target && ty0 + row <= ty1 && tx0 + col <= tx1)
                            tg = __ldg(P.target + base + (long long)(ty0 + row) * P.W + tx0 + col);
                        int g = 0;
#pragma unroll
                        for (int i = 0; i < MAX_OUT; ++i) {
#pragma unroll
                            for (int j = i; j < MAX_OUT; ++j, ++g) gs[g] = fma((double)v[i], (double)v[j], gs[g]);
                        }
#pragma unroll
                        for (int i = 0; i < MAX_OUT; ++i, ++g) gs[g] = fma((double)v[i], (double)tg, gs[g]);
                    }
                    double mine = 0.0;
#pragma unroll
                    for (int g = 0; g < NGRAM; ++g) {
                        double v = gs[g];
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
                        if (lane == g) mine = v;
                    }
                    if (lane < NGRAM) gwarp[lane] += mine;
                }
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
#undef CTDR_HEAD_INDEX
        __syncthreads();
        ++seg;
    } while (c < nsuper);
    // forward sweep with recording: the region is replayable when every batch found record space
    // (seg == 0: empty region, flagged above)
    if (RECORD && tid == 0 && seg > 0)
        P.bflag[(long long)b * regions + rr] = (C.rec_fail == 0) ? 1 : 0;

    if (K_BWD) {
        st.mus_flush(gmus);
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
        if (lane == 0) {
            if (loss_acc != 0.0) atomicAdd(P.out2 + 0, loss_acc);
            if (nvalid_acc) atomicAdd(P.out2 + 1, (double)nvalid_acc);
        }
    }
    if (MODE == MODE_MULTI && P.gram) {
        // upper triangle (i <= j) -> gram[i*n_out + j]; right-hand side -> gram[n_out*n_out + i]
        __syncwarp();
        if (lane < NGRAM) {
            const double v = gwarp[lane];
            int g = 0, dst = -1;
#pragma unroll
            for (int i = 0; i < MAX_OUT; ++i) {
#pragma unroll
                for (int j = i; j < MAX_OUT; ++j, ++g) if (g == lane && j < P.n_out) dst = i * P.n_out + j;
            }
#pragma unroll
            for (int i = 0; i < MAX_OUT; ++i, ++g) if (g == lane && i < P.n_out) dst = P.n_out * P.n_out + i;
            if (dst >= 0 && v != 0.0) atomicAdd(P.gram + dst, v);
        }
    }
    if (STATS && P.stats) {
        const unsigned cand = __reduce_add_sync(FULL, st.cand);
        const unsigned frag = __reduce_add_sync(FULL, st.frag);
        const unsigned nonpos = __reduce_add_sync(FULL, st.nonpos);
        const unsigned visits = __reduce_add_sync(FULL, st.visits);
        const unsigned hits = __reduce_add_sync(FULL, st.hits);
        if (lane == 0) {
            if (nonpos) atomicAdd(P.stats + 0, (unsigned long long)nonpos);
            if (frag) atomicAdd(P.stats + 1, (unsigned long long)frag);
            if (cand) atomicAdd(P.stats + 2, (unsigned long long)cand);
            if (visits) atomicAdd(P.stats + 4, (unsigned long long)visits);
            if (hits) atomicAdd(P.stats + 5, (unsigned long long)hits);
            if (warp == 0 && seg > 1) atomicAdd(P.stats + 3, (unsigned long long)(seg - 1));
        }
    }
}

}  // namespace ctdr
