// ctdr_b200.cu — C ABI (include/ctdr_b200.h) over the sm_100a kernels.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC
// (shapefromprojections_b200/csrc/build.py).  No torch types, no allocation, no synchronisation.
#include "../../include/ctdr_b200.h"
#include "ctdr_raster.cuh"
#include "ctdr_vertex.cuh"
#include "ctdr_raster64.cuh"
#include "ctdr_vertex64.cuh"
#include "ctdr_peer.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

using namespace ctdr;
static_assert(MAX_OUT == CTDR_MAX_OUTPUTS, "kernel and ABI disagree on the output count");

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

#define CTDR_CUDA(call, where)                                   \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return cuda_fail(e__, where);    \
    } while (0)

// every kernel launch of the library is counted (ctdr_launch_count): bench.py reports the launches inside its
// timed region from this counter instead of a constant
std::atomic<unsigned long long> g_launches{0};
#define CTDR_LAUNCHED(n, where)                                  \
    do {                                                         \
        g_launches.fetch_add((n), std::memory_order_relaxed);    \
        CTDR_CUDA(cudaGetLastError(), where);                    \
    } while (0)

// Optional per-stage device timing of the residual step (ctdr_project_residual_step_timed): events recorded on
// the caller's stream between the stages; elapsed times are read after one synchronisation at the end.
struct StageTimer {
    static constexpr int MAX_MARKS = 512;
    cudaEvent_t ev[MAX_MARKS];
    int stage[MAX_MARKS];
    int n = 0;
    bool ok = true;
    void mark(int stage_id, cudaStream_t st) {
        if (n >= MAX_MARKS) { ok = false; return; }
        if (cudaEventCreate(&ev[n]) != cudaSuccess || cudaEventRecord(ev[n], st) != cudaSuccess) { ok = false; return; }
        stage[n++] = stage_id;
    }
    // stage_ms[k] += time between the mark that closes stage k and the mark before it
    void collect(float* stage_ms, int nstages) {
        if (n > 0) cudaEventSynchronize(ev[n - 1]);
        for (int i = 1; i < n; ++i) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ev[i - 1], ev[i]) == cudaSuccess && stage[i] >= 0 && stage[i] < nstages) stage_ms[stage[i]] += ms;
        }
        for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
        n = 0;
    }
};
#define CTDR_MARK(tm, id, st) do { if (tm) (tm)->mark((id), (st)); } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Tuning / debug overrides.  The environment is read ONCE (first use, or ctdr_tuning_reload()), never per launch:
// behaviour does not follow ambient changes of the environment behind the caller's back.
struct Tuning {
    int region_tiles = 0;        // CTDR_REGION_TILES: 2, 4 or 8; 0 = heuristic
    int region_major = 2;        // CTDR_REGION_MAJOR: CTA order of the raster kernels: 0 angle-major, 1 region-major (rows top to bottom), 2 region-major centre-out (measured: 2 wins)
    int direct_max_box = -1;     // CTDR_DIRECT_MAX_BOX: -1 = heuristic, 0 disables the fine-mesh forward walk
    bool no_batch_records = false;   // CTDR_NO_BATCH_RECORDS: residual step without linked sweeps
    int grad_per_vertex = 1;
    int replay_carveout = -1;    // CTDR_REPLAY_CARVEOUT: preferred shared-memory carve-out of the replay kernel in percent; -1 = the driver's choice
    bool lanes_explicit = false; // CTDR_LANES given: used for every call size
    int lanes = 2;               // CTDR_LANES: 1 = the residual step on the caller's stream only; 2 (default) .. 4 = angles in that many parts on as many streams
    int replay_tiles = 0;        // CTDR_REPLAY_TILES: region edge of the replay kernel when the forward uses 4 x 4 tiles: 4 or 8; 0 = heuristic
    int face_normals = 1;        // CTDR_FACE_NORMALS: 0 = k_face_assemble forms the face normal and the unit ray per (angle, face) again
    int gather = 1;              // CTDR_GATHER: 1 = fused raster kernels gather corners from the per-vertex records (no p6/len3/ff), 0 = face arrays     // CTDR_GRAD_PER_VERTEX: 0 = per-corner gradient window (grad_p6 / grad_len3), 1 = per (angle, vertex)
    void load() {
        *this = Tuning();
        if (const char* e = getenv("CTDR_REGION_TILES")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) region_tiles = v; }
        if (const char* e = getenv("CTDR_REGION_MAJOR")) { const int v = atoi(e); region_major = (v == 2) ? 2 : (v != 0); }
        if (const char* e = getenv("CTDR_DIRECT_MAX_BOX")) direct_max_box = atoi(e) < 0 ? -1 : atoi(e);
        no_batch_records = getenv("CTDR_NO_BATCH_RECORDS") != nullptr;
        if (const char* e = getenv("CTDR_GRAD_PER_VERTEX")) grad_per_vertex = atoi(e);
        if (const char* e = getenv("CTDR_GATHER")) gather = atoi(e) != 0;
        if (const char* e = getenv("CTDR_FACE_NORMALS")) face_normals = atoi(e) != 0;
        if (const char* e = getenv("CTDR_REPLAY_TILES")) replay_tiles = atoi(e);
        if (const char* e = getenv("CTDR_REPLAY_CARVEOUT")) replay_carveout = atoi(e);
        if (const char* e = getenv("CTDR_LANES")) { const int v = atoi(e); lanes = v < 1 ? 1 : (v > 4 ? 4 : v); lanes_explicit = true; }
    }
};
Tuning& tuning() {
    static Tuning t = [] { Tuning x; x.load(); return x; }();
    return t;
}

int check_sizes(int B, int F, int H, int W, int n_mus) {
    if (B <= 0 || F <= 0 || H <= 0 || W <= 0 || n_mus <= 0)
        return fail(CTDR_E_SHAPE, "sizes must be positive (B=%d F=%d H=%d W=%d n_mus=%d)", B, F, H, W, n_mus);
    if (H > 32767 || W > 32767) return fail(CTDR_E_SHAPE, "detector larger than 32767 pixels (H=%d W=%d)", H, W);
    if (n_mus > 4096) return fail(CTDR_E_SHAPE, "n_mus=%d > 4096", n_mus);
    if ((F + CHUNK - 1) / CHUNK > (1 << 20)) return fail(CTDR_E_SHAPE, "F=%d > %d faces", F, CHUNK << 20);   // region list entries hold 20-bit chunk ids
    return 0;
}

int pick_region_tiles(int B, int F, int H, int W) {
    if (tuning().region_tiles) return tuning().region_tiles;      // tuning/debug override: 2, 4 or 8
    // One CTA per region of RT x RT tiles.  8x8 amortises the per-CTA list build best; 4x4 gives four
    // times more CTAs, which wins when the grid is small (multi-GPU angle shards: tail effect) or when
    // region lists are long (very fine meshes: fewer list segments).  Measured on config 3 / north star.
    int rt = 8;
    const long long regions8 = (long long)((W + 8 * TILE - 1) / (8 * TILE)) * ((H + 8 * TILE - 1) / (8 * TILE));
    if (regions8 * B < 88LL * 148 || F > 300000) rt = 4;      // config 3: 4x4 wins up to 180 angles per GPU, 8x8 from 240
    const long long regions4 = (long long)((W + 4 * TILE - 1) / (4 * TILE)) * ((H + 4 * TILE - 1) / (4 * TILE));
    if (regions4 * B < 2LL * 148) rt = 2;
    return rt;
}

size_t raster_smem_bytes(int n_mus) { return (size_t)n_mus * sizeof(float); }   // dynamic part; CtaSmem is static

constexpr int MAX_SR = 4;   // super-regions per detector edge

// concurrent_B: angles in flight on the GPU when other launches of the same kind run beside this one (the two lanes of the
// residual step): the grid-size heuristics look at what the machine sees, not at one lane's share
void fill_decomposition(RasterParams& P, const short4* boxes, const short4* face_boxes, const int* slist, const int* scount, int concurrent_B = 0) {
    P.face_box = face_boxes; P.slist = slist; P.scount = scount;
    P.NC = (P.F + CHUNK - 1) / CHUNK;
    const int HB = concurrent_B > P.B ? concurrent_B : P.B;
    P.RT = pick_region_tiles(HB, P.F, P.H, P.W);
    P.regions_x = (P.W + P.RT * TILE - 1) / (P.RT * TILE);
    P.regions_y = (P.H + P.RT * TILE - 1) / (P.RT * TILE);
    P.chunk_box = boxes;
    // the replay kernel of the residual step prefers 8 x 8 tiles per CTA at every grid size (measured, config 3: -14 % on a
    // 180-angle shard) while the forward sweep wants 4 x 4 on small grids: its regions are 2 x 2 forward regions then
    // (neutral at 90 angles; the 1M-triangle mesh loses 9 % with it: heavy tiles, CTA-level imbalance)
    const long long regions8 = (long long)((P.W + 8 * TILE - 1) / (8 * TILE)) * ((P.H + 8 * TILE - 1) / (8 * TILE));
    P.rts = P.RT;
    if (P.RT == 4 && (tuning().replay_tiles == 8 || (tuning().replay_tiles == 0 && P.F <= 300000 && regions8 * HB >= 40LL * 148))) P.rts = 8;
    P.region_major = tuning().region_major;     // region-major measured: -1.5 % on one GPU, -4 % on a 1/8 angle shard (config 3); centre-out another -0.5 % / -1.3 %
    P.sr_rx = (P.regions_x + MAX_SR - 1) / MAX_SR;
    P.sr_ry = (P.regions_y + MAX_SR - 1) / MAX_SR;
    P.nsr_x = (P.regions_x + P.sr_rx - 1) / P.sr_rx;
    P.nsr_y = (P.regions_y + P.sr_ry - 1) / P.sr_ry;
}

int launch_chunk_bbox(const RasterParams& P, short4* boxes, short4* face_boxes, cudaStream_t st, bool have_boxes = false) {
    if (!have_boxes) {      // the fused path gets both box arrays from k_face_assemble
        const long long warps = (long long)P.B * P.NC;
        const long long blocks = (warps + WARPS - 1) / WARPS;
        if (blocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*F too large for one launch");
        k_chunk_bbox<<<(unsigned)blocks, THREADS, 0, st>>>(P.p6, P.ff, P.labels2, P.B, P.F, P.H, P.W, P.NC, boxes, face_boxes);
        CTDR_LAUNCHED(1, "k_chunk_bbox launch");
    }
    if (P.slist) {      // tile-binned modes: coarse (super-region) lists for phase A of k_raster
        const long long sblocks = (long long)P.B * P.nsr_x * P.nsr_y;
        if (sblocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B too large for one launch");
        k_super_lists<<<(unsigned)sblocks, THREADS, 0, st>>>(boxes, P.NC, P.H, P.W, P.nsr_x, P.nsr_y,
                                                             P.sr_rx * P.RT * TILE, P.sr_ry * P.RT * TILE,
                                                             const_cast<int*>(P.slist), const_cast<int*>(P.scount));
        CTDR_LAUNCHED(1, "k_super_lists launch");
    }
    return 0;
}

template <int MODE, bool DIRECT, bool LINKED, bool GATHER>
void launch_raster_variant(const RasterParams& P, unsigned blocks, size_t smem, cudaStream_t st) {
    if constexpr (MODE == MODE_FWD && GATHER) {
        // the fused projector's forward without a count output: no count tile in shared memory, so eight CTAs fit
        // beside 60 KB of L1 instead of 28 KB
        if (!P.cnt) {
            if (P.stats) k_raster<MODE, false, DIRECT, true, LINKED, GATHER, true><<<blocks, THREADS, smem, st>>>(P);
            else k_raster<MODE, false, DIRECT, false, LINKED, GATHER, true><<<blocks, THREADS, smem, st>>>(P);
            return;
        }
    }
    if (P.stats) k_raster<MODE, false, DIRECT, true, LINKED, GATHER><<<blocks, THREADS, smem, st>>>(P);
    else k_raster<MODE, false, DIRECT, false, LINKED, GATHER><<<blocks, THREADS, smem, st>>>(P);
}

// gather (P.vt != null, fused projector only): the kernels instantiated for corner gathers from the vertex records
template <int MODE, bool DIRECT, bool LINKED>
void launch_raster_gather(const RasterParams& P, unsigned blocks, size_t smem, cudaStream_t st) {
    constexpr bool CAN_GATHER = (MODE == MODE_FWD || MODE == MODE_BWD || MODE == MODE_STEP);
    if constexpr (CAN_GATHER) {
        if (P.vt) { launch_raster_variant<MODE, DIRECT, LINKED, true>(P, blocks, smem, st); return; }
    }
    launch_raster_variant<MODE, DIRECT, LINKED, false>(P, blocks, smem, st);
}

// linked: MODE_FWD records its face batches, MODE_STEP is the replay-only kernel (see k_raster)
template <int MODE>
int launch_raster(const RasterParams& P, unsigned flags, cudaStream_t st, bool linked = false) {
    const long long blocks = (long long)P.B * P.regions_x * P.regions_y;
    if (blocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*regions too large for one launch");
    // dynamic shared memory: the CTA's grad_mus accumulators, or (MODE_MULTI) the per-warp normal-equation
    // sums followed by the tile accumulators of outputs 2 and 3; with the static CtaSmem it stays below the
    // 48 KB that need no opt-in (n_mus <= 4096)
    static_assert(sizeof(CtaSmem) + 4096 * sizeof(float) <= 48 * 1024, "static + dynamic shared memory");
    constexpr size_t gram_bytes = (size_t)WARPS * NGRAM * sizeof(double);
    const size_t smem = (MODE == MODE_MULTI)
        ? gram_bytes + (P.n_out > 2 ? (size_t)WARPS * (MAX_OUT - 2) * TILE_PIX * sizeof(float) : 0)
        : raster_smem_bytes(P.n_mus);
    // fine meshes (about a pixel per face): the forward walks batches of small boxes face-per-lane (DIRECT); the
    // backward modes walk row spans face-per-lane in any case (process_batch_rows)
    const bool direct = MODE == MODE_FWD && P.direct_max_box > 0;
    if (P.vt && (MODE == MODE_FRAG || MODE == MODE_MULTI || (flags & CTDR_FLAG_NO_PREFILTER)))
        return fail(CTDR_E_SHAPE, "internal: this raster mode reads the face arrays, not vertex records");
    if (flags & CTDR_FLAG_NO_PREFILTER) {   // debug: literal fp64 divide for every candidate (never linked)
        k_raster<MODE, true><<<(unsigned)blocks, THREADS, smem, st>>>(P);
    } else if (linked && MODE == MODE_STEP) {
        // replay-only residual/backward kernel over the batches the forward sweep recorded; its own region size (P.rts)
        const long long blocks = (long long)P.B * ((P.W + P.rts * TILE - 1) / (P.rts * TILE)) * ((P.H + P.rts * TILE - 1) / (P.rts * TILE));
        if (tuning().replay_carveout >= 0) {       // tuning probe: shared-memory carve-out in percent (L1 = the rest)
            cudaFuncSetAttribute(k_step_replay<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, tuning().replay_carveout);
        }
        if (P.vt) {
            if (P.stats) k_step_replay<true, true><<<(unsigned)blocks, THREADS, smem, st>>>(P);
            else k_step_replay<false, true><<<(unsigned)blocks, THREADS, smem, st>>>(P);
        } else {
            if (P.stats) k_step_replay<true, false><<<(unsigned)blocks, THREADS, smem, st>>>(P);
            else k_step_replay<false, false><<<(unsigned)blocks, THREADS, smem, st>>>(P);
        }
    } else if constexpr (MODE == MODE_FWD) {
        if (linked) {
            if (direct) launch_raster_gather<MODE_FWD, true, true>(P, (unsigned)blocks, smem, st);
            else launch_raster_gather<MODE_FWD, false, true>(P, (unsigned)blocks, smem, st);
        } else if (direct) {
            launch_raster_gather<MODE_FWD, true, false>(P, (unsigned)blocks, smem, st);
        } else {
            launch_raster_gather<MODE_FWD, false, false>(P, (unsigned)blocks, smem, st);
        }
    } else {
        launch_raster_gather<MODE, false, false>(P, (unsigned)blocks, smem, st);
    }
    CTDR_LAUNCHED(1, "k_raster launch");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// deterministic backward: one thread owns one (angle, face), walks its whole bounding box in
// row-major order and is the only writer of its 9 gradient values; the per-face attenuation
// terms go to `fs` and are reduced in a fixed tree by k_mus_reduce.
// ---------------------------------------------------------------------------------------------
template <bool PREFILTER, bool RESID>
__global__ void __launch_bounds__(128)
k_backward_faceowner(const RasterParams P, float* __restrict__ fs) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.B * P.F) return;
    const int b = (int)(gid / P.F), f = (int)(gid % P.F);
    fs[gid] = 0.0f;
    const int2 lab = __ldg(reinterpret_cast<const int2*>(P.labels2) + f);
    if (lab.x < lab.y) return;
    const float2* pp = reinterpret_cast<const float2*>(P.p6 + gid * 6);
    const float2 va = __ldg(pp), vb = __ldg(pp + 1), vc = __ldg(pp + 2);
    int ix0, iy0, ix1, iy1;
    if (!face_int_bbox(va.x, va.y, vb.x, vb.y, vc.x, vc.y, P.H, P.W, ix0, iy0, ix1, iy1)) return;
    const EdgeSetup es = edge_setup(va.x, va.y, vb.x, vb.y, vc.x, vc.y);
    float M0 = 0.0f, A1 = 0.0f, A2 = 0.0f;
    const long long base = (long long)b * P.H * P.W;
    for (int y = iy0; y <= iy1; ++y)
        for (int x = ix0; x <= ix1; ++x) {
            const long long o = base + (long long)y * P.W + x;
            float g;
            if (RESID) {        // gradient of sum_valid (target - phat)^2 formed on the fly (fp_vecs.py:10-13 mask)
                const float v = __ldg(P.im_in + o);
                if (!(v > -1e-6f && v <= P.mask_hi)) continue;
                g = 2.0f * (v - __ldg(P.target + o));
                if (g == 0.0f) continue;
            } else {
                g = __ldg(P.grad_im + o);
                if (g == 0.0f) continue;
                const float v = __ldg(P.im_in + o);
                if (((double)v < -1e-5) || (v > P.msv)) continue;          // diff_fp_for.cu:69-73
            }
            float s, t, k1, k2, w0, w1, w2;
            if (!cover_exact<PREFILTER>(es, (float)x, (float)y, s, t, k1, k2, w0, w1, w2)) continue;
            M0 += g; A1 += g * k1; A2 += g * k2;
        }
    if (M0 == 0.0f && A1 == 0.0f && A2 == 0.0f) return;
    float sgn = (__ldg(P.ff + gid) < 0.0f) ? -1.0f : 1.0f;
    if (lab.x == lab.y) sgn = -sgn;
    const float r0 = __ldg(P.len3 + gid * 3), r1 = __ldg(P.len3 + gid * 3 + 1), r2 = __ldg(P.len3 + gid * 3 + 2);
    const float cgeo = sgn * (__ldg(P.mus + lab.x) - __ldg(P.mus + lab.y));
    float gp[6], gl[3], bsum;
    face_grads_from_moments(es, r0, r1, r2, cgeo, M0, A1, A2, gp, gl, bsum);
#pragma unroll
    for (int k = 0; k < 6; ++k) P.grad_p6[gid * 6 + k] += gp[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) P.grad_len3[gid * 3 + k] += gl[k];
    fs[gid] = sgn * bsum;
}

// grad_mus[l] += sum over (b,f) of the face's attenuation term, fixed summation order:
// each of the 1024 threads sums a strided slice sequentially, then a fixed shared-memory tree.
__global__ void __launch_bounds__(1024)
k_mus_reduce(const float* __restrict__ fs, const int* __restrict__ labels2, int B, int F, int n_mus,
             float* __restrict__ grad_mus) {
    __shared__ float sh[1024];
    const int l = blockIdx.x;
    float acc = 0.0f;
    const long long n = (long long)B * F;
    for (long long i = threadIdx.x; i < n; i += 1024) {
        const int f = (int)(i % F);
        const int2 lab = __ldg(reinterpret_cast<const int2*>(labels2) + f);
        const float v = fs[i];
        if (lab.x == l) acc += (l == 0 ? -v : v);
        if (lab.y == l) acc += (l == 0 ? v : -v);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 512; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) grad_mus[l] += sh[0];
}

// ---------------------------------------------------------------------------------------------
// self-test: the hoisted-reciprocal quotient used by the raster kernels vs the compiler's IEEE
// fp64 division, over pseudo-random (k1, k3) float pairs; counts bitwise mismatches of fl32(k1/d).
// ---------------------------------------------------------------------------------------------
__global__ void k_selftest_quotient(unsigned long long per_thread, unsigned seed, unsigned long long* mismatches) {
    unsigned long long x = ((unsigned long long)seed << 32) ^ (0x9E3779B97F4A7C15ULL * (blockIdx.x * blockDim.x + threadIdx.x + 1));
    unsigned long long bad = 0;
    for (unsigned long long it = 0; it < per_thread; ++it) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;                       // xorshift64
        // exponents spread over 2^-40 .. 2^40, full random mantissas, both signs; every 64th sample
        // is made tiny / zero / equal to exercise the corners
        unsigned ua = (unsigned)x, ub = (unsigned)(x >> 32);
        const unsigned ea = 87u + (ua >> 24) % 81u, eb = 87u + (ub >> 24) % 81u;
        float k1 = __uint_as_float((ua & 0x807fffffu) | (ea << 23));
        float k3 = __uint_as_float((ub & 0x807fffffu) | (eb << 23));
        const unsigned sel = (unsigned)(x >> 17) & 63u;
        if (sel == 0) k1 = (ua & 1u) ? 0.0f : -0.0f;
        if (sel == 1) k3 = 0.0f;
        if (sel == 2) k1 = k3;
        if (sel == 3) k1 = __uint_as_float(ua & 0x800fffffu);           // denormal numerator
        if (sel == 4) k3 = __uint_as_float((ub & 0x807fffffu) | (60u << 23));  // |k3| ~ 1e-20 (near eps)
        const double d = (double)k3 + 1e-15;
        const float ref = (float)((double)k1 / d);
        const double y = 1.0 / d, a = (double)k1;
        const double q0 = __dmul_rn(a, y);
        const float got = (float)(k1 == 0.0f ? q0 : __fma_rn(__fma_rn(-q0, d, a), y, q0));
        if (__float_as_uint(ref) != __float_as_uint(got) && !(ref != ref && got != got)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// deterministic masked-L2 data term: fixed slice per block, fixed shared-memory tree, then one thread
// adds the block partials in order
__global__ void __launch_bounds__(256)
k_residual_partials(const float* __restrict__ phat, const float* __restrict__ target, long long n, float mask_hi,
                    double* __restrict__ partials) {
    __shared__ double sl[256], sc[256];
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long i0 = (long long)blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    double loss = 0.0, cnt = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
        const float v = __ldg(phat + i);
        if (v > -1e-6f && v <= mask_hi) { const double d = (double)(v - __ldg(target + i)); loss += d * d; cnt += 1.0; }
    }
    sl[threadIdx.x] = loss; sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) { sl[threadIdx.x] += sl[threadIdx.x + d]; sc[threadIdx.x] += sc[threadIdx.x + d]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = sl[0]; partials[2 * blockIdx.x + 1] = sc[0]; }
}

__global__ void k_residual_final(const double* __restrict__ partials, int nblocks, double* __restrict__ out2) {
    double loss = 0.0, cnt = 0.0;
    for (int i = 0; i < nblocks; ++i) { loss += partials[2 * i]; cnt += partials[2 * i + 1]; }
    out2[0] += loss; out2[1] += cnt;
}

// deterministic vertex backward: one thread owns one vertex, sums the cotangents of its incident
// face corners (adjacency in fixed order) angle by angle and is the only writer of its gradient
__global__ void __launch_bounds__(128)
k_vertex_backward_gather(const VertexParams P, const int* __restrict__ vstart, const int* __restrict__ vcorner,
                         const float* __restrict__ gp6, const float* __restrict__ glen3, float* __restrict__ grad_verts) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.V) return;
    const float x = __ldg(P.verts + 3 * v), y = __ldg(P.verts + 3 * v + 1), z = __ldg(P.verts + 3 * v + 2);
    const int c0 = __ldg(vstart + v), c1 = __ldg(vstart + v + 1);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    for (int b = 0; b < P.B; ++b) {
        float gpx = 0.0f, gpy = 0.0f, gl = 0.0f;
        for (int c = c0; c < c1; ++c) {
            const int corner = __ldg(vcorner + c);
            const int f = corner / 3, k = corner - 3 * f;
            const long long bf = (long long)b * P.F + f;
            gpx += __ldg(gp6 + bf * 6 + 2 * k); gpy += __ldg(gp6 + bf * 6 + 2 * k + 1); gl += __ldg(glen3 + bf * 3 + k);
        }
        if (gpx == 0.0f && gpy == 0.0f && gl == 0.0f) continue;
        const AngleConst A = angle_const(P, b);
        float gx, gy, gz;
        vertex_vjp(P, A, x, y, z, gpx, gpy, gl, gx, gy, gz);
        ax += gx; ay += gy; az += gz;
    }
    grad_verts[3 * v] += ax; grad_verts[3 * v + 1] += ay; grad_verts[3 * v + 2] += az;
}

struct RasterWs {
    short4* boxes;
    short4* face_boxes;
    int* slist;
    int* scount;
    float* fs;       // deterministic backward only
    int* cursors;    // fragment mode, multi-segment only
};

size_t raster_ws_layout(int B, int F, int H, int W, bool want_fs, bool want_cursors, RasterWs* out, void* base) {
    const int NC = (F + CHUNK - 1) / CHUNK;
    size_t off = 0;
    char* p = static_cast<char*>(base);
    if (out) out->boxes = reinterpret_cast<short4*>(p + off);
    off += align_up((size_t)B * NC * sizeof(short4), 256);
    if (out) out->face_boxes = reinterpret_cast<short4*>(p + off);
    off += align_up((size_t)B * F * sizeof(short4), 256);
    if (out) out->slist = reinterpret_cast<int*>(p + off);
    off += align_up((size_t)B * MAX_SR * MAX_SR * NC * sizeof(int), 256);
    if (out) out->scount = reinterpret_cast<int*>(p + off);
    off += align_up((size_t)B * MAX_SR * MAX_SR * sizeof(int), 256);
    if (out) out->fs = nullptr;
    if (want_fs) {
        if (out) out->fs = reinterpret_cast<float*>(p + off);
        off += align_up((size_t)B * F * sizeof(float), 256);
    }
    if (out) out->cursors = nullptr;
    if (want_cursors && NC > LIST_CAP) {
        if (out) out->cursors = reinterpret_cast<int*>(p + off);
        off += align_up((size_t)B * H * W * sizeof(int), 256);
    }
    return off;
}

RasterParams base_params(const float* p6, const float* ff, const float* len3, const int32_t* labels2,
                         const float* mus, int n_mus, int B, int F, int H, int W, float msv, int64_t* stats) {
    RasterParams P;
    memset(&P, 0, sizeof(P));
    P.p6 = p6; P.ff = ff; P.len3 = len3; P.labels2 = labels2; P.mus = mus; P.n_mus = n_mus;
    P.B = B; P.F = F; P.H = H; P.W = W; P.msv = msv;
    P.stats = reinterpret_cast<unsigned long long*>(stats);
    // Direct (face-per-lane) walk for fine meshes: with about two surface layers over roughly a third of the
    // detector, F >= 0.4*H*W means an average projected face of at most ~1.5 pixels (measured: north-star mesh,
    // 1M faces on 1024^2, -9 %; config 3, 100k faces, +3 % — hence a per-call choice, not a per-batch one).
    P.direct_max_box = ((double)F >= 0.4 * (double)H * (double)W) ? 64 : 0;
    if (tuning().direct_max_box >= 0) P.direct_max_box = tuning().direct_max_box;   // tuning/debug override; 0 disables
    return P;
}

// gather mode of the fused projector: corners come from the per-vertex records, the [B,F] corner arrays do not exist
void use_vertex_records(RasterParams& P, const float4* vt, const int32_t* faces, int V) {
    P.vt = vt; P.faces = faces; P.V = V;
    P.p6 = nullptr; P.len3 = nullptr; P.ff = nullptr;
}

int raster_backward_impl(RasterParams P, RasterWs ws, unsigned flags, cudaStream_t st) {
    if (flags & CTDR_FLAG_DETERMINISTIC) {
        const long long n = (long long)P.B * P.F;
        const long long blocks = (n + 127) / 128;
        if (blocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*F too large for one launch");
        if (flags & CTDR_FLAG_NO_PREFILTER) k_backward_faceowner<false, false><<<(unsigned)blocks, 128, 0, st>>>(P, ws.fs);
        else k_backward_faceowner<true, false><<<(unsigned)blocks, 128, 0, st>>>(P, ws.fs);
        CTDR_LAUNCHED(1, "k_backward_faceowner launch");
        k_mus_reduce<<<P.n_mus, 1024, 0, st>>>(ws.fs, P.labels2, P.B, P.F, P.n_mus, P.grad_mus);
        CTDR_LAUNCHED(1, "k_mus_reduce launch");
        return 0;
    }
    fill_decomposition(P, ws.boxes, ws.face_boxes, ws.slist, ws.scount);
    int rc = launch_chunk_bbox(P, ws.boxes, ws.face_boxes, st);
    if (rc) return rc;
    return launch_raster<MODE_BWD>(P, flags, st);
}

VertexParams vertex_params(int kind, const float* geom, int nangles, const int64_t* idx, int B, const float* verts, int V,
                           const int32_t* faces, int F, int H, int W, double sx, double sy, double msv) {
    VertexParams P;
    memset(&P, 0, sizeof(P));
    P.kind = kind; P.geom = geom; P.nangles = nangles; P.idx_angles = reinterpret_cast<const long long*>(idx);
    P.B = B; P.V = V; P.F = F; P.H = H; P.W = W; P.verts = verts; P.faces = faces;
    P.msv = (float)msv;
    if (kind == CTDR_GEOM_PARALLEL3D) {
        // ortho_Mat / compute_P (ctdr/utils/projection.py:16-30,65-80) in double, stored as float32
        const double left = -sx * W / 2.0, bottom = -sy * H / 2.0;
        const double half_u = 0.0 * sx, half_v = 0.5 * sy;
        const double l = left - half_u, r = -left - half_u, bo = bottom + half_v, t = -bottom + half_v;
        P.P00 = (float)(2.0 / (r - l));
        P.P11 = (float)(2.0 / (t - bo));
        P.P13 = (float)(-(t + bo) / (t - bo));
    }
    return P;
}

int check_geom(int kind) {
    if (kind < 0 || kind > 2) return fail(CTDR_E_GEOMETRY, "unknown geometry kind %d", kind);
    return 0;
}

struct ProjectWs {
    float *p6, *len3, *ff, *gp6, *glen3;
    float4* gvt;         // [window, V] per-vertex cotangents of the raster backward (default backward)
    float4* vt;          // [window, V] per-vertex transform records
    float4* fnorm;       // [F] world-space face normals (k_face_normals), gather mode of the *_VEC geometries
    AngleConst* atab;    // [window] per-angle constants of the vertex stage
    float* fs;           // [window, F] per-face attenuation terms (deterministic backward)
    double* partials;    // [2 * DET_LOSS_BLOCKS] (deterministic residual)
    // batch records of the residual step (forward sweep records, residual/backward sweep replays)
    int* brec; int* bhead; int* btail; unsigned char* bflag; int* bcursor; int bcap;
    int4* xlist;         // faces the forward sweep could not record as spans (RasterParams::xlist); its counter is bcursor[-1]
    RasterWs rws;
    int window;   // angles per window
};

constexpr int DET_LOSS_BLOCKS = 1024;

// batch-record capacity per angle: one record per 32 (face, tile) pairs plus one partly filled record per
// non-empty tile; sized for faces that touch about two tiles (larger ones fall back to the walk, region by region)
size_t batch_records_per_angle(int F, int H, int W) {
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    return (size_t)F / 16 + tiles / 2 + 64;
}

size_t batch_record_bytes_per_angle(int F, int H, int W) {
    const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    const size_t regions2 = (size_t)((W + 2 * TILE - 1) / (2 * TILE)) * ((H + 2 * TILE - 1) / (2 * TILE));   // smallest regions
    return batch_records_per_angle(F, H, W) * BREC_INTS * sizeof(int) + 2 * tiles * sizeof(int) + regions2 + sizeof(int);
}

// face_arrays: the [B,F] corner arrays (p6/len3/ff, and their gradient twins + per-face terms when bwd) are laid out;
// the default (gather) mode of the fused projector needs none of them, so its angle windows hold more angles in the
// same bytes.  The size query always budgets for them (the flags of the later call are not known to it).
size_t project_per_angle_bytes(int V, int F, bool bwd, int H = 0, int W = 0, bool face_arrays = true) {
    const int NC = (F + CHUNK - 1) / CHUNK;
    const size_t per_face = (face_arrays ? 40 + (bwd ? 36 + 4 : 0) : 0) + sizeof(short4);
    return (bwd && H > 0 ? batch_record_bytes_per_angle(F, H, W) : 0) + (size_t)V * sizeof(float4) * (bwd ? 2 : 1) + (size_t)F * per_face + (size_t)NC * (sizeof(short4) + MAX_SR * MAX_SR * sizeof(int)) +
           MAX_SR * MAX_SR * sizeof(int) + sizeof(AngleConst) + 512;
}

// window-independent part of the fused projector's workspace: alignment padding, the irregular-face list, the
// deterministic loss partials and the per-face normals
size_t project_fixed_bytes(int F) {
    return 32 * 256 + XLIST_CAP * sizeof(int4) + 2 * DET_LOSS_BLOCKS * sizeof(double) + align_up((size_t)F * sizeof(float4), 256);
}

int project_ws_layout(int B, int V, int F, int H, int W, bool bwd, void* base, size_t bytes, ProjectWs* ws, bool face_arrays = true) {
    const size_t slack = project_fixed_bytes(F);
    const size_t per = project_per_angle_bytes(V, F, bwd, H, W, face_arrays);
    if (bytes < per + slack) return fail(CTDR_E_WORKSPACE, "workspace %zu B < one angle (%zu B)", bytes, per + slack);
    long long win = (long long)((bytes - slack) / per);
    if (win > B) win = B;
    if (win > 65535) win = 65535;                        // k_face_assemble puts the angle in gridDim.y
    const long long nwin = (B + win - 1) / win;          // equal windows: no small straggler at the end
    win = (B + nwin - 1) / nwin;
    ws->window = (int)win;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    ws->fnorm = reinterpret_cast<float4*>(p + off); off += align_up((size_t)F * sizeof(float4), 256);
    ws->vt = reinterpret_cast<float4*>(p + off);  off += align_up((size_t)win * V * sizeof(float4), 256);
    ws->atab = reinterpret_cast<AngleConst*>(p + off); off += align_up((size_t)win * sizeof(AngleConst), 256);
    ws->p6 = ws->len3 = ws->ff = nullptr;
    if (face_arrays) {
        ws->p6 = reinterpret_cast<float*>(p + off);   off += align_up((size_t)win * F * 24, 256);
        ws->len3 = reinterpret_cast<float*>(p + off); off += align_up((size_t)win * F * 12, 256);
        ws->ff = reinterpret_cast<float*>(p + off);   off += align_up((size_t)win * F * 4, 256);
    }
    ws->gp6 = ws->glen3 = ws->fs = nullptr; ws->gvt = nullptr;
    if (bwd) {
        ws->gvt = reinterpret_cast<float4*>(p + off);  off += align_up((size_t)win * V * sizeof(float4), 256);
        if (face_arrays) {
            ws->gp6 = reinterpret_cast<float*>(p + off);   off += align_up((size_t)win * F * 24, 256);
            ws->glen3 = reinterpret_cast<float*>(p + off); off += align_up((size_t)win * F * 12, 256);
            ws->fs = reinterpret_cast<float*>(p + off);    off += align_up((size_t)win * F * 4, 256);
        }
        ws->partials = reinterpret_cast<double*>(p + off); off += align_up(2 * DET_LOSS_BLOCKS * sizeof(double), 256);
        const size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
        const size_t regions2 = (size_t)((W + 2 * TILE - 1) / (2 * TILE)) * ((H + 2 * TILE - 1) / (2 * TILE));
        size_t cap_a = batch_records_per_angle(F, H, W);          // records per angle; win * cap_a must index as int
        if (cap_a * (size_t)win > 0x7fffffffULL) cap_a = 0x7fffffffULL / (size_t)win;
        ws->bcap = (int)cap_a;
        ws->brec = reinterpret_cast<int*>(p + off);    off += align_up(cap_a * (size_t)win * BREC_INTS * sizeof(int), 256);
        ws->bhead = reinterpret_cast<int*>(p + off);   off += align_up((size_t)win * tiles * sizeof(int), 256);
        ws->btail = reinterpret_cast<int*>(p + off);   off += align_up((size_t)win * tiles * sizeof(int), 256);
        ws->bflag = reinterpret_cast<unsigned char*>(p + off); off += align_up((size_t)win * regions2, 256);
        // [0]: entries of xlist, [1 .. win]: record cursors — one memset zeroes both
        ws->bcursor = reinterpret_cast<int*>(p + off) + 1; off += align_up(((size_t)win + 1) * sizeof(int), 256);
        ws->xlist = reinterpret_cast<int4*>(p + off); off += align_up((size_t)XLIST_CAP * sizeof(int4), 256);
    }
    raster_ws_layout((int)win, F, H, W, false, false, &ws->rws, p + off);
    return 0;
}

int launch_vertex_forward(const VertexParams& P, float* p6, float* len3, float* ff, cudaStream_t st) {
    const long long n = (long long)P.B * P.F;
    const long long blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*F too large for one launch");
    k_vertex_forward<<<(unsigned)blocks, 256, 0, st>>>(P, p6, len3, ff);
    CTDR_LAUNCHED(1, "k_vertex_forward launch");
    return 0;
}

// per-angle constants of one angle window, formed once; the window's vertex-stage kernels then read the table
int launch_angle_table(VertexParams& P, AngleConst* tab, cudaStream_t st) {
    P.atab = nullptr;
    k_angle_table<<<(unsigned)((P.B + 127) / 128), 128, 0, st>>>(P, tab);
    CTDR_LAUNCHED(1, "k_angle_table launch");
    P.atab = tab;
    return 0;
}

// vertex stage of one angle window in the fused path; also produces the face and chunk pixel boxes
// full = false: boxes only — the raster kernels gather the corners from `vt` (RasterParams::vt)
int launch_vertex_forward_staged(const VertexParams& P, float4* vt, const int32_t* labels2, float* p6, float* len3, float* ff,
                                 short4* face_boxes, short4* chunk_boxes, cudaStream_t st, bool full = true, float4* fnorm = nullptr) {
    const long long nv = (long long)P.B * P.V;
    if ((nv + 255) / 256 > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*V too large for one launch");
    if (P.B > 65535) return fail(CTDR_E_SHAPE, "angle window of %d angles > 65535", P.B);
    k_vertex_transform<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(P, vt);
    CTDR_LAUNCHED(1, "k_vertex_transform launch");
    const dim3 grid((unsigned)((P.F + 255) / 256), (unsigned)P.B);
    if (full) k_face_assemble<true><<<grid, 256, 0, st>>>(P, vt, labels2, p6, len3, ff, face_boxes, chunk_boxes);
    else {
        if (P.kind == 0 || !tuning().face_normals) fnorm = nullptr;
        if (fnorm) {
            k_face_normals<<<(unsigned)((P.F + 255) / 256), 256, 0, st>>>(P.verts, P.faces, P.F, fnorm);
            CTDR_LAUNCHED(1, "k_face_normals launch");
        }
        k_face_assemble<false><<<grid, 256, 0, st>>>(P, vt, labels2, nullptr, nullptr, nullptr, face_boxes, chunk_boxes, fnorm);
    }
    CTDR_LAUNCHED(1, "k_face_assemble launch");
    return 0;
}

int launch_vertex_backward(const VertexParams& P, const float* gp6, const float* glen3, float* grad_verts, cudaStream_t st) {
    // enough angle slices to fill the machine, few enough to keep atomics per vertex low
    int slices = 1;
    const long long face_blocks = (P.F + 255) / 256;
    while (slices < P.B && face_blocks * slices < 24 * 148 && slices < 64) slices <<= 1;
    if (slices > P.B) slices = P.B;
    const int per = (P.B + slices - 1) / slices;
    dim3 grid((unsigned)face_blocks, (unsigned)((P.B + per - 1) / per));
    k_vertex_backward<<<grid, 256, 0, st>>>(P, gp6, glen3, grad_verts, per);
    CTDR_LAUNCHED(1, "k_vertex_backward launch");
    return 0;
}

int launch_vertex_backward_pervertex(const VertexParams& P, const float4* gvt, float* grad_verts, cudaStream_t st) {
    // enough angle slices to fill the machine, few enough to keep atomics per vertex low
    int slices = 1;
    const long long vblocks = (P.V + 255) / 256;
    while (slices < P.B && vblocks * slices < 16 * 148 && slices < 128) slices <<= 1;
    if (slices > P.B) slices = P.B;
    const int per = (P.B + slices - 1) / slices;
    dim3 grid((unsigned)vblocks, (unsigned)((P.B + per - 1) / per));
    k_vertex_backward_pervertex<<<grid, 256, 0, st>>>(P, gvt, grad_verts, per);
    CTDR_LAUNCHED(1, "k_vertex_backward_pervertex launch");
    return 0;
}

// deterministic residual/backward sweep of one angle window (CTDR_FLAG_DETERMINISTIC): single-writer
// kernels only, fixed summation orders everywhere
int deterministic_window(RasterParams P, const VertexParams& VP, const ProjectWs& ws, const int* vstart,
                         const int* vcorner, float* grad_verts, bool resid, unsigned flags, cudaStream_t st) {
    const long long n = (long long)P.B * P.F;
    const long long blocks = (n + 127) / 128;
    if (blocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*F too large for one launch");
    const bool exact = (flags & CTDR_FLAG_NO_PREFILTER) != 0;
    if (resid) {
        const long long npix = (long long)P.B * P.H * P.W;
        k_residual_partials<<<DET_LOSS_BLOCKS, 256, 0, st>>>(P.im_in, P.target, npix, P.mask_hi, ws.partials);
        k_residual_final<<<1, 1, 0, st>>>(ws.partials, DET_LOSS_BLOCKS, P.out2);
        if (exact) k_backward_faceowner<false, true><<<(unsigned)blocks, 128, 0, st>>>(P, ws.fs);
        else k_backward_faceowner<true, true><<<(unsigned)blocks, 128, 0, st>>>(P, ws.fs);
    } else {
        if (exact) k_backward_faceowner<false, false><<<(unsigned)blocks, 128, 0, st>>>(P, ws.fs);
        else k_backward_faceowner<true, false><<<(unsigned)blocks, 128, 0, st>>>(P, ws.fs);
    }
    CTDR_LAUNCHED(resid ? 3 : 1, "k_backward_faceowner launch");
    k_mus_reduce<<<P.n_mus, 1024, 0, st>>>(ws.fs, P.labels2, P.B, P.F, P.n_mus, P.grad_mus);
    CTDR_LAUNCHED(1, "k_mus_reduce launch");
    k_vertex_backward_gather<<<(VP.V + 127) / 128, 128, 0, st>>>(VP, vstart, vcorner, ws.gp6, ws.glen3, grad_verts);
    CTDR_LAUNCHED(1, "k_vertex_backward_gather launch");
    return 0;
}

}  // namespace

extern "C" {

int ctdr_abi_version(void) { return CTDR_ABI_VERSION; }
#ifdef CTDR_PROBE_ROWSTATS
int ctdr_debug_rowstats(unsigned long long* out4, int reset) {
    if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; return (int)cudaMemcpyToSymbol(ctdr::g_rowstats, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(out4, ctdr::g_rowstats, 4 * sizeof(unsigned long long));
}
#endif

int ctdr_selftest_quotient(unsigned long long samples, unsigned seed, unsigned long long* mismatches, ctdr_stream_t stream) {
    if (!mismatches) return fail(CTDR_E_NULL, "ctdr_selftest_quotient: null pointer");
    const int blocks = 148 * 8, threads = 256;
    const unsigned long long per = (samples + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
    k_selftest_quotient<<<blocks, threads, 0, stream>>>(per, seed, mismatches);
    CTDR_LAUNCHED(1, "k_selftest_quotient launch");
    return 0;
}
const char* ctdr_last_error(void) { return g_err; }

size_t ctdr_raster_workspace_bytes(int B, int F, int H, int W) {
    if (B <= 0 || F <= 0) return 0;
    // chunk boxes + the per-face attenuation terms of the deterministic backward
    return raster_ws_layout(B, F, H, W, true, false, nullptr, nullptr);
}

size_t ctdr_raster_fragments_workspace_bytes(int B, int F, int H, int W) {
    if (B <= 0 || F <= 0) return 0;
    return raster_ws_layout(B, F, H, W, false, true, nullptr, nullptr);
}

int ctdr_raster_forward(const float* p6, const float* ff, const float* len3, const int32_t* labels2,
                        const float* mus, int n_mus, int B, int F, int H, int W, float max_sino_val,
                        float* im, int32_t* cnt, void* workspace, size_t workspace_bytes,
                        int64_t* stats, unsigned flags, ctdr_stream_t stream) {
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (!p6 || !ff || !len3 || !labels2 || !mus || !im || !workspace) return fail(CTDR_E_NULL, "ctdr_raster_forward: null pointer");
    RasterWs ws;
    const size_t need = raster_ws_layout(B, F, H, W, false, false, &ws, workspace);
    if (workspace_bytes < need) return fail(CTDR_E_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, need);
    RasterParams P = base_params(p6, ff, len3, labels2, mus, n_mus, B, F, H, W, max_sino_val, stats);
    P.im = im; P.cnt = cnt;
    fill_decomposition(P, ws.boxes, ws.face_boxes, ws.slist, ws.scount);
    if (int rc = launch_chunk_bbox(P, ws.boxes, ws.face_boxes, stream)) return rc;
    return launch_raster<MODE_FWD>(P, flags, stream);
}

static int check_multi(int n_out, int n_mus, const float* target, const double* gram) {
    if (n_out < 1 || n_out > MAX_OUT) return fail(CTDR_E_SHAPE, "n_out=%d outside 1..%d", n_out, MAX_OUT);
    if (n_mus > 4096) return fail(CTDR_E_SHAPE, "n_mus=%d > 4096", n_mus);
    if (target && !gram) return fail(CTDR_E_NULL, "target given without gram");
    return 0;
}

int ctdr_raster_forward_multi(const float* p6, const float* ff, const float* len3, const int32_t* labels2,
                              const float* mus_kxn, int n_out, int n_mus, int B, int F, int H, int W,
                              float max_sino_val, float* im_kxbxhxw, const float* target, double* gram,
                              void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                              ctdr_stream_t stream) {
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (int rc = check_multi(n_out, n_mus, target, gram)) return rc;
    if (!p6 || !ff || !len3 || !labels2 || !mus_kxn || !im_kxbxhxw || !workspace) return fail(CTDR_E_NULL, "ctdr_raster_forward_multi: null pointer");
    RasterWs ws;
    const size_t need = raster_ws_layout(B, F, H, W, false, false, &ws, workspace);
    if (workspace_bytes < need) return fail(CTDR_E_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, need);
    RasterParams P = base_params(p6, ff, len3, labels2, mus_kxn, n_mus, B, F, H, W, max_sino_val, stats);
    P.im = im_kxbxhxw; P.n_out = n_out; P.out_stride = (long long)B * H * W; P.target = target; P.gram = gram;
    if (gram) CTDR_CUDA(cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)(n_out * n_out + n_out), stream), "gram memset");
    fill_decomposition(P, ws.boxes, ws.face_boxes, ws.slist, ws.scount);
    if (int rc = launch_chunk_bbox(P, ws.boxes, ws.face_boxes, stream)) return rc;
    return launch_raster<MODE_MULTI>(P, flags, stream);
}

int ctdr_raster_fragments(const float* p6, const float* ff, const float* len3, const int32_t* labels2,
                          const float* mus, int n_mus, int B, int F, int H, int W, float max_sino_val,
                          const float* im, const int32_t* firstidx, int32_t* imidx_vf, float* imwei_3vf,
                          void* workspace, size_t workspace_bytes, unsigned flags, ctdr_stream_t stream) {
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (!p6 || !ff || !len3 || !labels2 || !mus || !im || !firstidx || !imidx_vf || !imwei_3vf || !workspace)
        return fail(CTDR_E_NULL, "ctdr_raster_fragments: null pointer");
    RasterWs ws;
    const size_t need = raster_ws_layout(B, F, H, W, false, true, &ws, workspace);
    if (workspace_bytes < need) return fail(CTDR_E_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, need);
    RasterParams P = base_params(p6, ff, len3, labels2, mus, n_mus, B, F, H, W, max_sino_val, nullptr);
    P.im_in = im; P.firstidx = firstidx; P.imidx = imidx_vf; P.imwei = imwei_3vf;
    P.cnt = ws.cursors;
    fill_decomposition(P, ws.boxes, ws.face_boxes, ws.slist, ws.scount);
    if (int rc = launch_chunk_bbox(P, ws.boxes, ws.face_boxes, stream)) return rc;
    return launch_raster<MODE_FRAG>(P, flags, stream);
}

int ctdr_raster_backward(const float* grad_im, const float* im, const float* p6, const float* ff,
                         const float* len3, const int32_t* labels2, const float* mus, int n_mus,
                         int B, int F, int H, int W, float max_sino_val,
                         float* grad_p6, float* grad_len3, float* grad_mus,
                         void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                         ctdr_stream_t stream) {
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (!grad_im || !im || !p6 || !ff || !len3 || !labels2 || !mus || !grad_p6 || !grad_len3 || !grad_mus || !workspace)
        return fail(CTDR_E_NULL, "ctdr_raster_backward: null pointer");
    RasterWs ws;
    const bool det = (flags & CTDR_FLAG_DETERMINISTIC) != 0;
    const size_t need = raster_ws_layout(B, F, H, W, det, false, &ws, workspace);
    if (workspace_bytes < need) return fail(CTDR_E_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, need);
    RasterParams P = base_params(p6, ff, len3, labels2, mus, n_mus, B, F, H, W, max_sino_val, stats);
    P.grad_im = grad_im; P.im_in = im; P.grad_p6 = grad_p6; P.grad_len3 = grad_len3; P.grad_mus = grad_mus;
    return raster_backward_impl(P, ws, flags, stream);
}

int ctdr_vertex_forward(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                        const float* verts, int V, const int32_t* faces, int F, int H, int W,
                        double spacing_x, double spacing_y, double max_sino_val,
                        float* p6, float* len3, float* ff, ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, 1)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !p6 || !len3 || !ff) return fail(CTDR_E_NULL, "ctdr_vertex_forward: null pointer");
    const VertexParams P = vertex_params(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
    return launch_vertex_forward(P, p6, len3, ff, stream);
}

int ctdr_vertex_backward(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                         const float* verts, int V, const int32_t* faces, int F, int H, int W,
                         double spacing_x, double spacing_y, double max_sino_val,
                         const float* grad_p6, const float* grad_len3, float* grad_verts, ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, 1)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !grad_p6 || !grad_len3 || !grad_verts) return fail(CTDR_E_NULL, "ctdr_vertex_backward: null pointer");
    const VertexParams P = vertex_params(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
    return launch_vertex_backward(P, grad_p6, grad_len3, grad_verts, stream);
}

// ---- float64 instantiation -------------------------------------------------------------------
static int raster64_common(Raster64Params& P, const double* p6, const double* ff, const double* len3, const int32_t* labels2,
                           const double* mus, int n_mus, int B, int F, int H, int W, double msv,
                           void* workspace, size_t workspace_bytes, int64_t* stats, cudaStream_t stream, const char* who) {
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (n_mus > 2048) return fail(CTDR_E_SHAPE, "%s: n_mus=%d > 2048", who, n_mus);
    if (!p6 || !ff || !len3 || !labels2 || !mus || !workspace) return fail(CTDR_E_NULL, "%s: null pointer", who);
    RasterWs ws;
    const size_t need = raster_ws_layout(B, F, H, W, false, false, &ws, workspace);
    if (workspace_bytes < need) return fail(CTDR_E_WORKSPACE, "workspace %zu B < %zu B", workspace_bytes, need);
    RasterParams D;                                          // the decomposition is the float32 kernels'
    memset(&D, 0, sizeof(D));
    D.B = B; D.F = F; D.H = H; D.W = W;
    fill_decomposition(D, ws.boxes, ws.face_boxes, ws.slist, ws.scount);
    memset(&P, 0, sizeof(P));
    P.p6 = p6; P.ff = ff; P.len3 = len3; P.labels2 = labels2; P.mus = mus; P.n_mus = n_mus;
    P.B = B; P.F = F; P.H = H; P.W = W; P.msv = msv;
    P.NC = D.NC; P.RT = D.RT; P.regions_x = D.regions_x; P.regions_y = D.regions_y;
    P.chunk_box = ws.boxes; P.face_box = ws.face_boxes; P.slist = ws.slist; P.scount = ws.scount;
    P.nsr_x = D.nsr_x; P.nsr_y = D.nsr_y; P.sr_rx = D.sr_rx; P.sr_ry = D.sr_ry; P.region_major = D.region_major;
    P.stats = reinterpret_cast<unsigned long long*>(stats);
    const long long warps = (long long)B * P.NC;
    const long long blocks = (warps + WARPS - 1) / WARPS;
    if (blocks > 0x7fffffffLL || (long long)B * P.regions_x * P.regions_y > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*F too large for one launch");
    k_chunk_bbox64<<<(unsigned)blocks, THREADS, 0, stream>>>(p6, labels2, B, F, H, W, P.NC, ws.boxes, ws.face_boxes);
    CTDR_LAUNCHED(1, "k_chunk_bbox64 launch");
    return launch_chunk_bbox(D, ws.boxes, ws.face_boxes, stream, true);      // super-region lists from the chunk boxes
}

int ctdr_raster_forward_f64(const double* p6, const double* ff, const double* len3, const int32_t* labels2,
                            const double* mus, int n_mus, int B, int F, int H, int W, double max_sino_val,
                            double* im, int32_t* cnt, void* workspace, size_t workspace_bytes,
                            int64_t* stats, unsigned flags, ctdr_stream_t stream) {
    (void)flags;
    if (!im) return fail(CTDR_E_NULL, "ctdr_raster_forward_f64: null pointer");
    Raster64Params P;
    if (int rc = raster64_common(P, p6, ff, len3, labels2, mus, n_mus, B, F, H, W, max_sino_val, workspace, workspace_bytes, stats, stream, "ctdr_raster_forward_f64")) return rc;
    P.im = im; P.cnt = cnt;
    k_raster64<false><<<(unsigned)((long long)B * P.regions_x * P.regions_y), THREADS, 0, stream>>>(P);
    CTDR_LAUNCHED(1, "k_raster64 launch");
    return 0;
}

int ctdr_raster_backward_f64(const double* grad_im, const double* im, const double* p6, const double* ff, const double* len3,
                             const int32_t* labels2, const double* mus, int n_mus, int B, int F, int H, int W, double max_sino_val,
                             double* grad_p6, double* grad_len3, double* grad_mus, void* workspace, size_t workspace_bytes,
                             int64_t* stats, unsigned flags, ctdr_stream_t stream) {
    (void)flags;
    if (!grad_im || !im || !grad_p6 || !grad_len3 || !grad_mus) return fail(CTDR_E_NULL, "ctdr_raster_backward_f64: null pointer");
    Raster64Params P;
    if (int rc = raster64_common(P, p6, ff, len3, labels2, mus, n_mus, B, F, H, W, max_sino_val, workspace, workspace_bytes, stats, stream, "ctdr_raster_backward_f64")) return rc;
    P.grad_im = grad_im; P.im_in = im; P.grad_p6 = grad_p6; P.grad_len3 = grad_len3; P.grad_mus = grad_mus;
    k_raster64<true><<<(unsigned)((long long)B * P.regions_x * P.regions_y), THREADS, (size_t)n_mus * sizeof(double), stream>>>(P);
    CTDR_LAUNCHED(1, "k_raster64 launch");
    return 0;
}

static Vertex64Params vertex64_params(int kind, const double* geom, int nangles, const int64_t* idx, int B, const double* verts, int V,
                                      const int32_t* faces, int F, int H, int W, double sx, double sy, double msv) {
    const VertexParams Q = vertex_params(kind, nullptr, nangles, idx, B, nullptr, V, faces, F, H, W, sx, sy, msv);
    Vertex64Params P;
    memset(&P, 0, sizeof(P));
    P.kind = kind; P.geom = geom; P.idx_angles = reinterpret_cast<const long long*>(idx); P.nangles = nangles;
    P.B = B; P.V = V; P.F = F; P.H = H; P.W = W; P.verts = verts; P.faces = faces;
    P.msv = msv;
    P.P00 = (double)Q.P00; P.P11 = (double)Q.P11; P.P13 = (double)Q.P13;     // the float32 matrix widened (fp_opengl.py:44-45)
    return P;
}

int ctdr_vertex_forward_f64(int geom_kind, const double* geom, int nangles, const int64_t* idx_angles, int B,
                            const double* verts, int V, const int32_t* faces, int F, int H, int W,
                            double spacing_x, double spacing_y, double max_sino_val,
                            double* p6, double* len3, double* ff, ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, 1)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !p6 || !len3 || !ff) return fail(CTDR_E_NULL, "ctdr_vertex_forward_f64: null pointer");
    const Vertex64Params P = vertex64_params(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
    const long long blocks = ((long long)B * F + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(CTDR_E_SHAPE, "B*F too large for one launch");
    k_vertex_forward64<<<(unsigned)blocks, 256, 0, stream>>>(P, p6, len3, ff);
    CTDR_LAUNCHED(1, "k_vertex_forward64 launch");
    return 0;
}

int ctdr_vertex_backward_f64(int geom_kind, const double* geom, int nangles, const int64_t* idx_angles, int B,
                             const double* verts, int V, const int32_t* faces, int F, int H, int W,
                             double spacing_x, double spacing_y, double max_sino_val,
                             const double* grad_p6, const double* grad_len3, double* grad_verts, ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, 1)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !grad_p6 || !grad_len3 || !grad_verts) return fail(CTDR_E_NULL, "ctdr_vertex_backward_f64: null pointer");
    const Vertex64Params P = vertex64_params(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
    int slices = 1;
    const long long face_blocks = (F + 255) / 256;
    while (slices < B && face_blocks * slices < 24 * 148 && slices < 64) slices <<= 1;
    if (slices > B) slices = B;
    const int per = (B + slices - 1) / slices;
    k_vertex_backward64<<<dim3((unsigned)face_blocks, (unsigned)((B + per - 1) / per)), 256, 0, stream>>>(P, grad_p6, grad_len3, grad_verts, per);
    CTDR_LAUNCHED(1, "k_vertex_backward64 launch");
    return 0;
}

size_t ctdr_project_workspace_bytes(int B, int V, int F, int H, int W, int need_backward) {
    if (B <= 0 || F <= 0 || V <= 0) return 0;
    const size_t per = project_per_angle_bytes(V, F, need_backward != 0, H, W);
    const size_t budget = (size_t)8 << 30;               // default window budget: 8 GiB
    long long win = (long long)(budget / per);
    if (win < 1) win = 1;
    if (win > B) win = B;
    // the lanes of the residual step lay out one fixed part each, and uneven splits give some lanes one angle more
    return per * (size_t)(win + 3) + 4 * project_fixed_bytes(F) + 1024;      // up to four lanes
}

int ctdr_project_forward(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                         const float* verts, int V, const int32_t* faces, int F,
                         const int32_t* labels2, const float* mus, int n_mus, int H, int W,
                         double spacing_x, double spacing_y, double max_sino_val,
                         float* im, int32_t* cnt, void* workspace, size_t workspace_bytes,
                         int64_t* stats, unsigned flags, ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !labels2 || !mus || !im || !workspace) return fail(CTDR_E_NULL, "ctdr_project_forward: null pointer");
    ProjectWs ws;
    const bool gather = tuning().gather != 0 && !(flags & CTDR_FLAG_NO_PREFILTER);   // the debug kernels read the face arrays
    if (int rc = project_ws_layout(B, V, F, H, W, false, workspace, workspace_bytes, &ws, !gather)) return rc;
    const size_t hw = (size_t)H * W;
    for (int b0 = 0; b0 < B; b0 += ws.window) {
        const int nb = (B - b0 < ws.window) ? (B - b0) : ws.window;
        VertexParams VP = vertex_params(geom_kind, geom, nangles, idx_angles + b0, nb, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
        if (int rc = launch_angle_table(VP, ws.atab, stream)) return rc;
        if (int rc = launch_vertex_forward_staged(VP, ws.vt, labels2, ws.p6, ws.len3, ws.ff, ws.rws.face_boxes, ws.rws.boxes, stream, !gather, ws.fnorm)) return rc;
        RasterParams P = base_params(ws.p6, ws.ff, ws.len3, labels2, mus, n_mus, nb, F, H, W, (float)max_sino_val, stats);
        if (gather) use_vertex_records(P, ws.vt, faces, V);
        P.im = im + (size_t)b0 * hw;
        P.cnt = cnt ? cnt + (size_t)b0 * hw : nullptr;
        fill_decomposition(P, ws.rws.boxes, ws.rws.face_boxes, ws.rws.slist, ws.rws.scount);
        if (int rc = launch_chunk_bbox(P, ws.rws.boxes, ws.rws.face_boxes, stream, true)) return rc;
        if (int rc = launch_raster<MODE_FWD>(P, flags, stream)) return rc;
    }
    return 0;
}

int ctdr_project_forward_multi(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                               const float* verts, int V, const int32_t* faces, int F,
                               const int32_t* labels2, const float* mus_kxn, int n_out, int n_mus, int H, int W,
                               double spacing_x, double spacing_y, double max_sino_val,
                               float* im_kxbxhxw, const float* target, double* gram,
                               void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                               ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (int rc = check_multi(n_out, n_mus, target, gram)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !labels2 || !mus_kxn || !im_kxbxhxw || !workspace)
        return fail(CTDR_E_NULL, "ctdr_project_forward_multi: null pointer");
    ProjectWs ws;
    if (int rc = project_ws_layout(B, V, F, H, W, false, workspace, workspace_bytes, &ws)) return rc;
    if (gram) CTDR_CUDA(cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)(n_out * n_out + n_out), stream), "gram memset");
    const size_t hw = (size_t)H * W;
    for (int b0 = 0; b0 < B; b0 += ws.window) {
        const int nb = (B - b0 < ws.window) ? (B - b0) : ws.window;
        VertexParams VP = vertex_params(geom_kind, geom, nangles, idx_angles + b0, nb, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
        if (int rc = launch_angle_table(VP, ws.atab, stream)) return rc;
        if (int rc = launch_vertex_forward_staged(VP, ws.vt, labels2, ws.p6, ws.len3, ws.ff, ws.rws.face_boxes, ws.rws.boxes, stream)) return rc;
        RasterParams P = base_params(ws.p6, ws.ff, ws.len3, labels2, mus_kxn, n_mus, nb, F, H, W, (float)max_sino_val, stats);
        P.im = im_kxbxhxw + (size_t)b0 * hw; P.n_out = n_out; P.out_stride = (long long)B * (long long)hw;
        P.target = target ? target + (size_t)b0 * hw : nullptr; P.gram = gram;
        fill_decomposition(P, ws.rws.boxes, ws.rws.face_boxes, ws.rws.slist, ws.rws.scount);
        if (int rc = launch_chunk_bbox(P, ws.rws.boxes, ws.rws.face_boxes, stream, true)) return rc;
        if (int rc = launch_raster<MODE_MULTI>(P, flags, stream)) return rc;
    }
    return 0;
}

static int project_backward_impl(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                                 const float* verts, int V, const int32_t* faces, int F,
                                 const int32_t* labels2, const float* mus, int n_mus, int H, int W,
                                 double spacing_x, double spacing_y, double max_sino_val,
                                 const float* grad_im, const float* im, const float* target, double* out2,
                                 float* grad_verts, float* grad_mus, const int32_t* vadj_start, const int32_t* vadj_corner,
                                 void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                                 cudaStream_t stream) {
    const bool det = (flags & CTDR_FLAG_DETERMINISTIC) != 0;
    if (det && (!vadj_start || !vadj_corner))
        return fail(CTDR_E_NULL, "CTDR_FLAG_DETERMINISTIC needs the vertex->corner adjacency (vadj_start, vadj_corner)");
    ProjectWs ws;
    // gather mode: corners from the vertex records, cotangents onto the vertices; the single-writer (deterministic) and
    // debug kernels, and the per-corner gradient fallback, keep the [B,F] face arrays
    const bool gather = !det && tuning().gather != 0 && tuning().grad_per_vertex != 0 && !(flags & CTDR_FLAG_NO_PREFILTER);
    if (int rc = project_ws_layout(B, V, F, H, W, true, workspace, workspace_bytes, &ws, !gather)) return rc;
    const size_t hw = (size_t)H * W;
    for (int b0 = 0; b0 < B; b0 += ws.window) {
        const int nb = (B - b0 < ws.window) ? (B - b0) : ws.window;
        VertexParams VP = vertex_params(geom_kind, geom, nangles, idx_angles + b0, nb, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
        if (int rc = launch_angle_table(VP, ws.atab, stream)) return rc;
        if (int rc = launch_vertex_forward_staged(VP, ws.vt, labels2, ws.p6, ws.len3, ws.ff, ws.rws.face_boxes, ws.rws.boxes, stream, !gather, ws.fnorm)) return rc;
        RasterParams P = base_params(ws.p6, ws.ff, ws.len3, labels2, mus, n_mus, nb, F, H, W, (float)max_sino_val, stats);
        if (gather) use_vertex_records(P, ws.vt, faces, V);
        P.im_in = im + (size_t)b0 * hw;
        P.grad_p6 = ws.gp6; P.grad_len3 = ws.glen3; P.grad_mus = grad_mus;
        const bool pervertex = !det && tuning().grad_per_vertex != 0;
        if (!pervertex) {   // single-writer kernels (and the tuning fallback) keep the per-corner layout
            CTDR_CUDA(cudaMemsetAsync(ws.gp6, 0, (size_t)nb * F * 24, stream), "memset grad_p6");
            CTDR_CUDA(cudaMemsetAsync(ws.glen3, 0, (size_t)nb * F * 12, stream), "memset grad_len3");
        } else {            // default: cotangents summed per (angle, vertex) by the raster backward itself
            CTDR_CUDA(cudaMemsetAsync(ws.gvt, 0, (size_t)nb * V * sizeof(float4), stream), "memset grad_vt");
            P.grad_vt = ws.gvt; P.faces = faces; P.V = V;
        }
        if (det) {
            if (target) { P.target = target + (size_t)b0 * hw; P.out2 = out2; P.mask_hi = (float)(max_sino_val - 1e-6); }
            else P.grad_im = grad_im + (size_t)b0 * hw;
            if (int rc = deterministic_window(P, VP, ws, vadj_start, vadj_corner, grad_verts, target != nullptr, flags, stream)) return rc;
            continue;
        }
        fill_decomposition(P, ws.rws.boxes, ws.rws.face_boxes, ws.rws.slist, ws.rws.scount);
        if (int rc = launch_chunk_bbox(P, ws.rws.boxes, ws.rws.face_boxes, stream, true)) return rc;
        if (target) {
            P.target = target + (size_t)b0 * hw;
            P.out2 = out2;
            P.mask_hi = (float)(max_sino_val - 1e-6);
            if (int rc = launch_raster<MODE_STEP>(P, flags, stream)) return rc;
        } else {
            P.grad_im = grad_im + (size_t)b0 * hw;
            if (int rc = launch_raster<MODE_BWD>(P, flags, stream)) return rc;
        }
        if (pervertex) { if (int rc = launch_vertex_backward_pervertex(VP, ws.gvt, grad_verts, stream)) return rc; }
        else if (int rc = launch_vertex_backward(VP, ws.gp6, ws.glen3, grad_verts, stream)) return rc;
    }
    return 0;
}

int ctdr_project_backward(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                          const float* verts, int V, const int32_t* faces, int F,
                          const int32_t* labels2, const float* mus, int n_mus, int H, int W,
                          double spacing_x, double spacing_y, double max_sino_val,
                          const float* grad_im, const float* im, float* grad_verts, float* grad_mus,
                          const int32_t* vadj_start, const int32_t* vadj_corner,
                          void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                          ctdr_stream_t stream) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !labels2 || !mus || !grad_im || !im || !grad_verts || !grad_mus || !workspace)
        return fail(CTDR_E_NULL, "ctdr_project_backward: null pointer");
    return project_backward_impl(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, labels2, mus, n_mus, H, W,
                                 spacing_x, spacing_y, max_sino_val, grad_im, im, nullptr, nullptr,
                                 grad_verts, grad_mus, vadj_start, vadj_corner, workspace, workspace_bytes, stats, flags, stream);
}

// The residual step's arguments, as the angle-window loop needs them
struct ResidualArgs {
    int geom_kind; const float* geom; int nangles; const int64_t* idx_angles;
    const float* verts; int V; const int32_t* faces; int F;
    const int32_t* labels2; const float* mus; int n_mus; int H, W;
    double spacing_x, spacing_y, max_sino_val;
    const float* target; float* phat; float* grad_verts; float* grad_mus; double* out2;
    const int32_t* vadj_start; const int32_t* vadj_corner;
    int64_t* stats; unsigned flags;
    bool det, gather;
    int concurrent_B;       // angles in flight over both lanes (fill_decomposition)
};

// angles [b_begin, b_end) of the step through the windows of workspace layout `ws`, enqueued on `stream`
static int residual_windows(const ResidualArgs& A, ProjectWs& ws, int b_begin, int b_end, cudaStream_t stream, StageTimer* tm) {
    const int geom_kind = A.geom_kind, nangles = A.nangles, V = A.V, F = A.F, n_mus = A.n_mus, H = A.H, W = A.W;
    const float* geom = A.geom; const int64_t* idx_angles = A.idx_angles; const float* verts = A.verts; const int32_t* faces = A.faces;
    const int32_t* labels2 = A.labels2; const float* mus = A.mus;
    const double spacing_x = A.spacing_x, spacing_y = A.spacing_y, max_sino_val = A.max_sino_val;
    const float* target = A.target; float* phat = A.phat; float* grad_verts = A.grad_verts; float* grad_mus = A.grad_mus; double* out2 = A.out2;
    const int32_t* vadj_start = A.vadj_start; const int32_t* vadj_corner = A.vadj_corner;
    int64_t* stats = A.stats; const unsigned flags = A.flags;
    const bool det = A.det, gather = A.gather;
    const size_t hw = (size_t)H * W;
    for (int b0 = b_begin; b0 < b_end; b0 += ws.window) {
        const int nb = (b_end - b0 < ws.window) ? (b_end - b0) : ws.window;
        VertexParams VP = vertex_params(geom_kind, geom, nangles, idx_angles + b0, nb, verts, V, faces, F, H, W, spacing_x, spacing_y, max_sino_val);
        if (int rc = launch_angle_table(VP, ws.atab, stream)) return rc;
        if (int rc = launch_vertex_forward_staged(VP, ws.vt, labels2, ws.p6, ws.len3, ws.ff, ws.rws.face_boxes, ws.rws.boxes, stream, !gather, ws.fnorm)) return rc;
        RasterParams P = base_params(ws.p6, ws.ff, ws.len3, labels2, mus, n_mus, nb, F, H, W, (float)max_sino_val, stats);
        if (gather) use_vertex_records(P, ws.vt, faces, V);
        fill_decomposition(P, ws.rws.boxes, ws.rws.face_boxes, ws.rws.slist, ws.rws.scount, A.concurrent_B);
        if (int rc = launch_chunk_bbox(P, ws.rws.boxes, ws.rws.face_boxes, stream, true)) return rc;
        CTDR_MARK(tm, CTDR_STAGE_VERTEX_FORWARD, stream);
        P.im = phat + (size_t)b0 * hw;
        if (!det) {             // the forward sweep records its face batches for the residual/backward sweep
            P.bmode = 1; P.brec = ws.brec; P.bhead = ws.bhead; P.btail = ws.btail; P.bflag = ws.bflag; P.bcursor = ws.bcursor; P.bcap = ws.bcap;
            P.xlist = ws.xlist; P.xcount = ws.bcursor - 1;
            if (tuning().no_batch_records || (flags & CTDR_FLAG_NO_PREFILTER)) P.bmode = 0;   // tuning/debug: no linking
            if (P.bmode) CTDR_CUDA(cudaMemsetAsync(ws.bcursor - 1, 0, ((size_t)nb + 1) * sizeof(int), stream), "memset record cursors");
        }
        const bool linked = P.bmode == 1;
        CTDR_MARK(tm, CTDR_STAGE_MEMSET, stream);
        if (int rc = launch_raster<MODE_FWD>(P, flags, stream, linked)) return rc;
        CTDR_MARK(tm, CTDR_STAGE_RASTER_FORWARD, stream);
        if (linked) P.bmode = 2;
        const bool pervertex = !det && tuning().grad_per_vertex != 0;
        if (!pervertex) {   // single-writer kernels (and the tuning fallback) keep the per-corner layout
            CTDR_CUDA(cudaMemsetAsync(ws.gp6, 0, (size_t)nb * F * 24, stream), "memset grad_p6");
            CTDR_CUDA(cudaMemsetAsync(ws.glen3, 0, (size_t)nb * F * 12, stream), "memset grad_len3");
        } else {        // default: cotangents summed per (angle, vertex) by the raster backward itself
            CTDR_CUDA(cudaMemsetAsync(ws.gvt, 0, (size_t)nb * V * sizeof(float4), stream), "memset grad_vt");
        }
        CTDR_MARK(tm, CTDR_STAGE_MEMSET, stream);
        P.im = nullptr;
        P.im_in = phat + (size_t)b0 * hw;
        P.target = target + (size_t)b0 * hw;
        P.out2 = out2;
        P.mask_hi = (float)(max_sino_val - 1e-6);
        P.grad_p6 = ws.gp6; P.grad_len3 = ws.glen3; P.grad_mus = grad_mus;
        if (pervertex) { P.grad_vt = ws.gvt; P.faces = faces; P.V = V; }
        P.stats = nullptr;                                   // counters describe the forward sweep only
        if (det) {
            if (int rc = deterministic_window(P, VP, ws, vadj_start, vadj_corner, grad_verts, true, flags, stream)) return rc;
            CTDR_MARK(tm, CTDR_STAGE_RASTER_STEP, stream);
            continue;
        }
        // linked sweeps: the replay-only kernel takes the recorded regions, the ordinary one the rest
        if (linked) {
            if (int rc = launch_raster<MODE_STEP>(P, flags, stream, true)) return rc;
            // the few faces whose coverage could not be recorded as row intervals
            if (P.vt) k_step_facelist<true><<<XLIST_CAP / 128, 128, 0, stream>>>(P);
            else k_step_facelist<false><<<XLIST_CAP / 128, 128, 0, stream>>>(P);
            CTDR_LAUNCHED(1, "k_step_facelist launch");
        }
        if (int rc = launch_raster<MODE_STEP>(P, flags, stream)) return rc;
        CTDR_MARK(tm, CTDR_STAGE_RASTER_STEP, stream);
        if (pervertex) { if (int rc = launch_vertex_backward_pervertex(VP, ws.gvt, grad_verts, stream)) return rc; }
        else if (int rc = launch_vertex_backward(VP, ws.gp6, ws.glen3, grad_verts, stream)) return rc;
        CTDR_MARK(tm, CTDR_STAGE_VERTEX_BACKWARD, stream);
    }
    return 0;
}

// A second, lowest-priority stream per host thread and device for the step's second lane (below), with the two
// events of the fork/join.  Created on first use; event record / wait / kernel launches on it are capturable.
constexpr int MAX_LANES = 4;
struct SideLanes { cudaStream_t s[MAX_LANES - 1] = {}; cudaEvent_t fork = nullptr, join[MAX_LANES - 1] = {}; bool tried = false, ok = false; };
static SideLanes* side_lanes() {
    constexpr int MAX_DEV = 64;
    thread_local SideLanes lanes[MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    SideLanes& L = lanes[dev];
    if (!L.tried) {
        L.tried = true;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = numerically largest = lowest priority
        bool ok = cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; ok && k < MAX_LANES - 1; ++k)
            ok = cudaStreamCreateWithPriority(&L.s[k], cudaStreamNonBlocking, lo) == cudaSuccess &&
                 cudaEventCreateWithFlags(&L.join[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) cudaGetLastError();
        L.ok = ok;
    }
    return L.ok ? &L : nullptr;
}

static int residual_step_impl(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                              const float* verts, int V, const int32_t* faces, int F,
                              const int32_t* labels2, const float* mus, int n_mus, int H, int W,
                              double spacing_x, double spacing_y, double max_sino_val,
                              const float* target, float* phat, float* grad_verts, float* grad_mus, double* out2,
                              const int32_t* vadj_start, const int32_t* vadj_corner,
                              void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                              cudaStream_t stream, StageTimer* tm) {
    if (int rc = check_geom(geom_kind)) return rc;
    if (int rc = check_sizes(B, F, H, W, n_mus)) return rc;
    if (V <= 0) return fail(CTDR_E_SHAPE, "V=%d", V);
    if (!geom || !idx_angles || !verts || !faces || !labels2 || !mus || !target || !phat || !grad_verts || !grad_mus || !out2 || !workspace)
        return fail(CTDR_E_NULL, "ctdr_project_residual_step: null pointer");
    // per angle window: vertex stage once, forward raster (phat is an output the caller keeps), then the
    // residual/backward sweep over the same staged p6/len3/ff and face boxes
    const bool det = (flags & CTDR_FLAG_DETERMINISTIC) != 0;
    if (det && (!vadj_start || !vadj_corner))
        return fail(CTDR_E_NULL, "CTDR_FLAG_DETERMINISTIC needs the vertex->corner adjacency (vadj_start, vadj_corner)");
    // gather mode: corners from the vertex records, cotangents onto the vertices; the single-writer (deterministic) and
    // debug kernels, and the per-corner gradient fallback, keep the [B,F] face arrays
    const bool gather = !det && tuning().gather != 0 && tuning().grad_per_vertex != 0 && !(flags & CTDR_FLAG_NO_PREFILTER);
    ResidualArgs A = {geom_kind, geom, nangles, idx_angles, verts, V, faces, F, labels2, mus, n_mus, H, W,
                            spacing_x, spacing_y, max_sino_val, target, phat, grad_verts, grad_mus, out2,
                            vadj_start, vadj_corner, stats, flags, det, gather, 0};
    // Two lanes: the angles are cut in two halves that run the whole sequence (vertex stage, forward sweep, memsets,
    // residual/backward sweep, vertex backward) on two streams, each in its own half of the workspace.  Every kernel
    // of the sequence ends in a tail of few, long CTAs (and the vertex stage and memsets are bandwidth-bound while the
    // sweeps are issue-bound): the second lane — lowest stream priority — fills what the first leaves idle.  Both
    // lanes add into grad_verts / grad_mus / out2 with the atomics the kernels use anyway.  Not for the deterministic
    // mode (fixed summation order) nor the timed variant (its stage marks are on one stream).
    // small calls are launch-bound: the second lane's extra launches and the fork/join cost more than the overlap wins
    // (measured on config 3: +1.2 % at 20 angles = 21 Mpixel, -1.5 % at 45 angles = 47 Mpixel; +9 % on config 1)
    int nl = (tuning().lanes_explicit || (long long)B * H * W >= (1LL << 25)) ? tuning().lanes : 1;
    if (nl > B) nl = B;
    SideLanes* side = (!det && !tm && nl >= 2) ? side_lanes() : nullptr;
    if (side) {
        const size_t part = (workspace_bytes / nl) & ~(size_t)255;
        const size_t need = project_per_angle_bytes(V, F, true, H, W, !gather) + project_fixed_bytes(F);
        ProjectWs ws[MAX_LANES];
        int begin[MAX_LANES + 1];
        bool ok = part >= need;
        int inflight = 0;
        for (int k = 0; k <= nl; ++k) begin[k] = (int)((long long)B * k / nl);
        for (int k = 0; ok && k < nl; ++k) {
            const int nbk = begin[k + 1] - begin[k];
            ok = project_ws_layout(nbk, V, F, H, W, true, static_cast<char*>(workspace) + (size_t)k * part, part, &ws[k], !gather) == 0;
            if (ok) inflight += ws[k].window < nbk ? ws[k].window : nbk;
        }
        if (ok) {
            A.concurrent_B = inflight;
            CTDR_CUDA(cudaEventRecord(side->fork, stream), "lane fork");
            int rc = 0;
            for (int k = nl - 1; k >= 1; --k) {          // side lanes are enqueued first and scheduled last (priority)
                CTDR_CUDA(cudaStreamWaitEvent(side->s[k - 1], side->fork, 0), "lane fork wait");
                if (!rc) rc = residual_windows(A, ws[k], begin[k], begin[k + 1], side->s[k - 1], nullptr);
            }
            if (!rc) rc = residual_windows(A, ws[0], begin[0], begin[1], stream, nullptr);
            // join even after an error: the caller's stream must not run ahead of work already enqueued on a side lane
            for (int k = 1; k < nl; ++k) {
                cudaEventRecord(side->join[k - 1], side->s[k - 1]);
                cudaStreamWaitEvent(stream, side->join[k - 1], 0);
            }
            return rc;
        }
    }
    ProjectWs ws;
    if (int rc = project_ws_layout(B, V, F, H, W, true, workspace, workspace_bytes, &ws, !gather)) return rc;
    CTDR_MARK(tm, -1, stream);
    return residual_windows(A, ws, 0, B, stream, tm);
}

int ctdr_project_residual_step(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                               const float* verts, int V, const int32_t* faces, int F,
                               const int32_t* labels2, const float* mus, int n_mus, int H, int W,
                               double spacing_x, double spacing_y, double max_sino_val,
                               const float* target, float* phat, float* grad_verts, float* grad_mus, double* out2,
                               const int32_t* vadj_start, const int32_t* vadj_corner,
                               void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                               ctdr_stream_t stream) {
    return residual_step_impl(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, labels2, mus, n_mus, H, W,
                              spacing_x, spacing_y, max_sino_val, target, phat, grad_verts, grad_mus, out2,
                              vadj_start, vadj_corner, workspace, workspace_bytes, stats, flags, stream, nullptr);
}

int ctdr_project_residual_step_timed(int geom_kind, const float* geom, int nangles, const int64_t* idx_angles, int B,
                                     const float* verts, int V, const int32_t* faces, int F,
                                     const int32_t* labels2, const float* mus, int n_mus, int H, int W,
                                     double spacing_x, double spacing_y, double max_sino_val,
                                     const float* target, float* phat, float* grad_verts, float* grad_mus, double* out2,
                                     const int32_t* vadj_start, const int32_t* vadj_corner,
                                     void* workspace, size_t workspace_bytes, int64_t* stats, unsigned flags,
                                     ctdr_stream_t stream, float* stage_ms_host) {
    if (!stage_ms_host) return fail(CTDR_E_NULL, "ctdr_project_residual_step_timed: stage_ms_host is NULL");
    for (int k = 0; k < CTDR_STAGE_COUNT; ++k) stage_ms_host[k] = 0.0f;
    StageTimer tm;
    const int rc = residual_step_impl(geom_kind, geom, nangles, idx_angles, B, verts, V, faces, F, labels2, mus, n_mus, H, W,
                                      spacing_x, spacing_y, max_sino_val, target, phat, grad_verts, grad_mus, out2,
                                      vadj_start, vadj_corner, workspace, workspace_bytes, stats, flags, stream, &tm);
    const bool ok = tm.ok;
    tm.collect(stage_ms_host, CTDR_STAGE_COUNT);          // synchronises with the last recorded event
    if (rc) return rc;
    if (!ok) return fail(CTDR_E_SHAPE, "ctdr_project_residual_step_timed: more than %d stage marks (too many angle windows)", StageTimer::MAX_MARKS);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Peer blocks + the one-launch all-reduce over NVLink peer memory (ctdr_peer.cuh)
// ---------------------------------------------------------------------------------------------
static_assert(CTDR_PEER_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "IPC handle size");
static_assert(CTDR_PEER_MAX_RANKS == PEER_MAX_RANKS, "kernel and ABI disagree on the rank count");

static size_t peer_slot_bytes(size_t payload_bytes) { return align_up(payload_bytes, 256); }

size_t ctdr_peer_block_bytes(size_t payload_bytes) { return PEER_HEADER_BYTES + 2 * peer_slot_bytes(payload_bytes); }

int ctdr_peer_block_create(size_t payload_bytes, void** block, unsigned char* handle_host) {
    if (!block) return fail(CTDR_E_NULL, "ctdr_peer_block_create: block is NULL");
    if (payload_bytes == 0 || (payload_bytes & 15)) return fail(CTDR_E_SHAPE, "ctdr_peer_block_create: payload_bytes=%zu must be a positive multiple of 16", payload_bytes);
    void* p = nullptr;
    const size_t n = ctdr_peer_block_bytes(payload_bytes);
    CTDR_CUDA(cudaMalloc(&p, n), "cudaMalloc(peer block)");
    cudaError_t e = cudaMemset(p, 0, n);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && handle_host) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) memcpy(handle_host, &h, sizeof(h));
    }
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "ctdr_peer_block_create"); }
    *block = p;
    return 0;
}

int ctdr_peer_block_open(const unsigned char* handle_host, void** block) {
    if (!handle_host || !block) return fail(CTDR_E_NULL, "ctdr_peer_block_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    CTDR_CUDA(cudaIpcOpenMemHandle(block, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return 0;
}

int ctdr_peer_block_close(void* block) {
    if (!block) return 0;
    CTDR_CUDA(cudaIpcCloseMemHandle(block), "cudaIpcCloseMemHandle");
    return 0;
}

int ctdr_peer_block_destroy(void* block) {
    if (!block) return 0;
    CTDR_CUDA(cudaFree(block), "cudaFree(peer block)");
    return 0;
}

int ctdr_peer_block_status(const void* block, int* status_host) {
    if (!block || !status_host) return fail(CTDR_E_NULL, "ctdr_peer_block_status: NULL argument");
    unsigned v = 0;
    CTDR_CUDA(cudaMemcpy(&v, static_cast<const unsigned*>(block) + PEER_CTL_STATUS, sizeof(v), cudaMemcpyDeviceToHost), "ctdr_peer_block_status");
    *status_host = (int)v;
    return 0;
}

int ctdr_peer_allreduce(void* const* blocks_host, int world, int rank, size_t payload_bytes,
                        void* local, size_t f32_bytes, size_t f64_bytes, ctdr_stream_t stream) {
    if (!blocks_host || !local) return fail(CTDR_E_NULL, "ctdr_peer_allreduce: NULL argument");
    if (world < 1 || world > PEER_MAX_RANKS || rank < 0 || rank >= world)
        return fail(CTDR_E_SHAPE, "ctdr_peer_allreduce: world=%d rank=%d (at most %d ranks)", world, rank, PEER_MAX_RANKS);
    if ((f32_bytes & 15) || (f64_bytes & 15) || f32_bytes + f64_bytes == 0 || f32_bytes + f64_bytes > payload_bytes || (reinterpret_cast<uintptr_t>(local) & 15))
        return fail(CTDR_E_SHAPE, "ctdr_peer_allreduce: f32_bytes=%zu f64_bytes=%zu must be multiples of 16 within payload_bytes=%zu, local 16-byte aligned", f32_bytes, f64_bytes, payload_bytes);
    PeerTable T = {};
    for (int r = 0; r < world; ++r) {
        if (!blocks_host[r]) return fail(CTDR_E_NULL, "ctdr_peer_allreduce: block of rank %d is NULL", r);
        T.block[r] = static_cast<unsigned char*>(blocks_host[r]);
    }
    const long long nvec = (long long)((f32_bytes + f64_bytes) / 16);
    int slices = (int)((nvec + 2 * PEER_THREADS - 1) / (2 * PEER_THREADS));     // about two vectors per thread
    // at most 48 CTAs (<= PEER_MAX_SLICES): eight emulated ranks of the single-GPU test are then still all resident on one B200
    slices = slices < 1 ? 1 : (slices > 48 ? 48 : slices);
    k_peer_allreduce<<<slices, PEER_THREADS, 0, (cudaStream_t)stream>>>(T, world, rank, static_cast<float4*>(local),
                                                                        (long long)(f32_bytes / 16), (int)(f64_bytes / 16),
                                                                        peer_slot_bytes(payload_bytes));
    CTDR_LAUNCHED(1, "k_peer_allreduce");
    return 0;
}

unsigned long long ctdr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void ctdr_tuning_reload(void) { tuning().load(); }

}  // extern "C"
