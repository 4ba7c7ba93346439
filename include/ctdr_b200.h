/* ctdr_b200.h — C ABI of the B200-native differentiable mesh forward projector.
 *
 * Drop-in boundary for ONE path of jakeoung/ShapeFromProjections (package `ctdr`): the native
 * module `cuda_diff_fp` (ctdr/cuda/diff_fp.cpp:131-134) behind `ctdr.function.rasterizer.Rasterizer`
 * and the per-angle vertex stage of `ctdr.nn.fp_vecs.FP` / `ctdr.nn.fp_opengl.FP`.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer on the current
 *     CUDA device unless its name ends in `_host`;
 *   - the caller owns and allocates every buffer (as in the reference, diff_fp.cpp:21-26 /
 *     rasterizer.py:36-46); the library allocates nothing (one exception: ctdr_peer_block_create, whose
 *     memory must be exportable through CUDA IPC). Scratch memory is a caller-provided
 *     `workspace` whose size comes from the matching `*_workspace_bytes` query;
 *   - work is enqueued on `stream` (a cudaStream_t) and returns without synchronising
 *     (the reference launches on the legacy default stream: diff_fp_for.cu:253);
 *   - return value: 0 ok, < 0 argument error (CTDR_E_*), > 0 a cudaError_t; a thread-local
 *     message is available from ctdr_last_error();
 *   - re-entrant; process state is limited to that thread-local string, the launch counter, the tuning overrides
 *     read once from the environment, and — per host thread and device — one internal low-priority stream with two
 *     events on which ctdr_project_residual_step runs the second half of its angles (joined back into `stream`
 *     before the call returns control of it; capturable in a CUDA graph);
 *   - float32 entry points are the tuned path; the reference also dispatches double (diff_fp_for.cu:251):
 *     the *_f64 entry points below are that instantiation on the device;
 *   - sinogram / counts are [B,H,W] row-major (the reference's [B,H,W,1]); pixel offsets are 64-bit.
 *
 * Deviations from the reference kernels, all documented in DESIGN.md:
 *   - the live device assert `bary > 0` (diff_fp_for.cu:183) is replaced by a counter
 *     (stats[CTDR_STAT_NONPOSITIVE_LEN]);
 *   - faces with label_in < label_out are skipped in every pass (the reference records them in
 *     pass 2 only and overruns the pixel's CSR slot: diff_fp_for.cu:161-163 vs :208-218);
 *   - `im` / `cnt` need not arrive zeroed for ctdr_raster_forward (every pixel is stored once).
 */
#ifndef CTDR_B200_H
#define CTDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTDR_ABI_VERSION 5   /* 2: + ctdr_raster_forward_multi, ctdr_project_forward_multi; 3: + ctdr_launch_count,
                                ctdr_project_residual_step_timed; 4: + the float64 instantiation (*_f64);
                                5: + peer blocks and the one-launch NVLink all-reduce (ctdr_peer_*) */

/* argument errors */
#define CTDR_E_NULL      (-1)   /* a required pointer is NULL */
#define CTDR_E_SHAPE     (-2)   /* a size is out of range (B,F,H,W,n_mus <= 0, H or W > 32767 ...) */
#define CTDR_E_WORKSPACE (-3)   /* workspace too small */
#define CTDR_E_GEOMETRY  (-4)   /* unknown geometry kind */

/* flags (bit field) */
#define CTDR_FLAG_NO_PREFILTER 1u   /* evaluate the fp64 barycentric divide for every bbox candidate
                                       (debug; results are identical with and without) */
#define CTDR_FLAG_DETERMINISTIC 2u  /* backward: face-owner kernel, single writer per gradient
                                       element, fixed summation order (slower, bit-reproducible) */

/* indices into the optional `stats` array (device, int64[CTDR_STAT_COUNT], caller zeroes it) */
#define CTDR_STAT_NONPOSITIVE_LEN 0 /* fragments whose interpolated length was <= 0 (reference: trap) */
#define CTDR_STAT_FRAGMENTS       1 /* covered (pixel, face) pairs == the reference's `vf` (rasterizer.py:64) */
#define CTDR_STAT_CANDIDATES      2 /* bbox (pixel, face) pairs evaluated */
#define CTDR_STAT_LIST_SEGMENTS   3 /* extra chunk-list segments taken (0 unless > capacity chunks hit a region) */
#define CTDR_STAT_FACE_BATCHES    4 /* batches of <= 32 queued faces staged in shared memory */
#define CTDR_STAT_CHUNK_HITS      5 /* (tile, chunk) pairs in which at least one face box overlapped the tile */
#define CTDR_STAT_COUNT           8

/* geometry kinds for the vertex stage */
#define CTDR_GEOM_PARALLEL3D     0  /* ctdr/nn/fp_opengl.py (type "parallel3d"); geom = angles[nangles] */
#define CTDR_GEOM_PARALLEL3D_VEC 1  /* ctdr/nn/fp_vecs.py  (type "parallel3d_vec"); geom = Vectors[nangles,12] */
#define CTDR_GEOM_CONE_VEC       2  /* ctdr/nn/fp_vecs.py  (type "cone_vec");       geom = Vectors[nangles,12] */

typedef struct CUstream_st* ctdr_stream_t;   /* == cudaStream_t */

int         ctdr_abi_version(void);
const char* ctdr_last_error(void);

/* Number of kernels this library has launched so far in the process (all threads, all streams); measurement
 * aid: the difference around a timed region is the region's launch count. */
unsigned long long ctdr_launch_count(void);

/* Tuning / debug overrides come from the environment (CTDR_REGION_TILES, CTDR_REGION_MAJOR, CTDR_DIRECT_MAX_BOX,
 * CTDR_NO_BATCH_RECORDS, CTDR_LANES, CTDR_REPLAY_TILES, CTDR_FACE_NORMALS, CTDR_GATHER, CTDR_GRAD_PER_VERTEX), read once
 * at first use; this re-reads them (tests and tuning scripts). */
void ctdr_tuning_reload(void);

/* Self-test of the one arithmetic shortcut the kernels take: fl32(k1/(k3+1e-15)) through the
 * per-face reciprocal + one fused correction step vs the IEEE fp64 division of the reference
 * (diff_fp_for.cu:144-145), over `samples` pseudo-random float pairs.  *mismatches (device uint64,
 * caller zeroes it) receives the number of bitwise differences; it must stay 0. */
int ctdr_selftest_quotient(unsigned long long samples, unsigned seed, unsigned long long* mismatches,
                           ctdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Rasterizer boundary — replaces cuda_diff_fp.forward / backward
 * (diff_fp.cpp:21-67,77-129 -> diff_fp_for.cu:226-274, diff_fp_back.cu:297-350) together with the
 * host logic of Rasterizer.forward/backward around them (rasterizer.py:23-32 bbox, :63-73 CSR).
 * Inputs are the tensors the reference passes to Rasterizer.apply (rasterizer.py:13):
 *   p6      [B,F,6]  projected vertices in detector pixels (ax,ay,bx,by,cx,cy)
 *   ff      [B,F]    front-facing value; only its sign is used (diff_fp_for.cu:94-96)
 *   len3    [B,F,3]  per-vertex source-to-surface lengths
 *   labels2 [F,2]    int32 (label_in, label_out)
 *   mus     [n_mus]  attenuation per label
 * ------------------------------------------------------------------------------------------- */

/* bytes of scratch needed by ctdr_raster_forward / ctdr_raster_backward for these sizes
 * (chunk and face pixel boxes; plus B*F floats used only by the deterministic backward) */
size_t ctdr_raster_workspace_bytes(int B, int F, int H, int W);
/* bytes of scratch needed by ctdr_raster_fragments (adds B*H*W cursors only for meshes whose
 * region chunk lists may need several segments, F > 10240) */
size_t ctdr_raster_fragments_workspace_bytes(int B, int F, int H, int W);

/* Pass 1 (is_the_first_pass=1): im[B,H,W] = sum over covering faces of
 * sign*(mu_in - mu_out')*interpolated length, accumulated per pixel in ascending face order with
 * the reference's exact arithmetic (bit-identical `im`); cnt[B,H,W] (may be NULL) = number of
 * covering faces == cntvisible_bxhxwx1.  stats may be NULL. */
int ctdr_raster_forward(const float* p6, const float* ff, const float* len3,
                        const int32_t* labels2, const float* mus, int n_mus,
                        int B, int F, int H, int W, float max_sino_val,
                        float* im, int32_t* cnt,
                        void* workspace, size_t workspace_bytes,
                        int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* n_out (1..CTDR_MAX_OUTPUTS) sinograms in ONE raster pass: the projector is linear in mus
 * (diff_fp_for.cu:174-190), so renders that differ only in mus share all coverage work — the
 * 2-3 one-hot renders of Model.estimate_mu (ctdr/model/vanilla.py:186-265) become one pass.
 *   mus_kxn     [n_out, n_mus]      one attenuation vector per output
 *   im_kxbxhxw  [n_out, B, H, W]    output k is bit-identical to ctdr_raster_forward with mus_kxn[k]
 *   target      [B,H,W] or NULL, gram float64 [n_out*n_out + n_out] or NULL (zeroed by the callee):
 *               gram[i*n_out + j] = sum_pixels im_i*im_j for i <= j (upper triangle),
 *               gram[n_out*n_out + i] = sum_pixels im_i*target — the normal equations of
 *               vanilla.py:208-216,245-256, accumulated in float64 while the tiles are on chip. */
#define CTDR_MAX_OUTPUTS 4
int ctdr_raster_forward_multi(const float* p6, const float* ff, const float* len3,
                              const int32_t* labels2, const float* mus_kxn, int n_out, int n_mus,
                              int B, int F, int H, int W, float max_sino_val,
                              float* im_kxbxhxw, const float* target, double* gram,
                              void* workspace, size_t workspace_bytes,
                              int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* Pass 2 (is_the_first_pass=0): fragment list in the reference's CSR layout, for pixels whose
 * `im` passes the reference's validity test (diff_fp_for.cu:69-73):
 *   imidx_vf[firstidx[pix]+k] = face id + 1, imwei_3vf[3*(firstidx[pix]+k)+{0,1,2}] = (w0,w1,w2),
 * k-th covering face in ascending face order.  Parity/debug output: the product backward
 * recomputes coverage instead of reading this list. */
int ctdr_raster_fragments(const float* p6, const float* ff, const float* len3,
                          const int32_t* labels2, const float* mus, int n_mus,
                          int B, int F, int H, int W, float max_sino_val,
                          const float* im, const int32_t* firstidx,
                          int32_t* imidx_vf, float* imwei_3vf,
                          void* workspace, size_t workspace_bytes,
                          unsigned flags, ctdr_stream_t stream);

/* Backward (diff_fp_back.cu:47-295): grad_im = dL/d im [B,H,W], im = forward output (decides
 * which pixels carry gradient, diff_fp_for.cu:69-73).  ACCUMULATES into grad_p6 [B,F,6],
 * grad_len3 [B,F,3], grad_mus [n_mus], which must arrive zeroed (rasterizer.py:97-99).
 * Default: warp-aggregated reductions + one atomic per (tile, face, component);
 * CTDR_FLAG_DETERMINISTIC: single-writer kernel, no atomics on grad_p6/grad_len3. */
int ctdr_raster_backward(const float* grad_im, const float* im,
                         const float* p6, const float* ff, const float* len3,
                         const int32_t* labels2, const float* mus, int n_mus,
                         int B, int F, int H, int W, float max_sino_val,
                         float* grad_p6, float* grad_len3, float* grad_mus,
                         void* workspace, size_t workspace_bytes,
                         int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Vertex stage — replaces the ~40 torch ops of FP.forward (fp_opengl.py:87-158, fp_vecs.py:85-196)
 * and their autograd.
 *   geom       angles[nangles] (PARALLEL3D) or Vectors[nangles,12] (the *_VEC kinds), float32,
 *              exactly the array the reference uploads (fp_opengl.py:63, fp_vecs.py:64)
 *   idx_angles [B] int64 rows of `geom` to project; a row outside [0, nangles) turns that angle's outputs into NaN
 *              (the reference raises IndexError, fp_vecs.py:101; Python callers validate host indices beforehand)
 *   verts [V,3] float32, faces [F,3] int32
 *   spacing_x/y: DetectorSpacingX/Y as the Python doubles of proj_geom (PARALLEL3D only: they
 *              define the orthographic matrix, projection.py:65-80, evaluated in double and
 *              stored as float32 like the reference); max_sino_val: the double computed by the FP
 *              constructor (fp_opengl.py:42, fp_vecs.py:59-62), narrowed to float32 on use.
 * Outputs p6 [B,F,6], len3 [B,F,3], ff [B,F].
 * ------------------------------------------------------------------------------------------- */
int ctdr_vertex_forward(int geom_kind, const float* geom, int nangles,
                        const int64_t* idx_angles, int B,
                        const float* verts, int V, const int32_t* faces, int F,
                        int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                        float* p6, float* len3, float* ff, ctdr_stream_t stream);

/* Transposed Jacobian of ctdr_vertex_forward: ACCUMULATES into grad_verts [V,3] (arrives zeroed
 * or holding a running sum).  front_facing carries no gradient (rasterizer.py:107). */
int ctdr_vertex_backward(int geom_kind, const float* geom, int nangles,
                         const int64_t* idx_angles, int B,
                         const float* verts, int V, const int32_t* faces, int F,
                         int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                         const float* grad_p6, const float* grad_len3,
                         float* grad_verts, ctdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * float64 instantiation — the reference dispatches its kernels over float AND double
 * (AT_DISPATCH_FLOATING_TYPES, diff_fp_for.cu:251, diff_fp_back.cu:326) and FP / Model take a dtype
 * (fp_vecs.py:53-56, vanilla.py:27,31).  Same arguments as the float32 entry points with double
 * arrays (geometry tables as float64: the float32 ProjectionAngles widened, Vectors as they are);
 * same workspace (ctdr_raster_workspace_bytes).  Tile-binned like the float32 kernels, every pixel
 * accumulated in ascending face order with IEEE divisions; the backward reduces three moments per
 * (tile, face) and adds them with double atomics.  n_mus <= 2048.  This is the high-precision
 * referee / drop-in for dtype=torch.float64, not the tuned path.
 * ------------------------------------------------------------------------------------------- */
int ctdr_raster_forward_f64(const double* p6, const double* ff, const double* len3,
                            const int32_t* labels2, const double* mus, int n_mus,
                            int B, int F, int H, int W, double max_sino_val,
                            double* im, int32_t* cnt,
                            void* workspace, size_t workspace_bytes,
                            int64_t* stats, unsigned flags, ctdr_stream_t stream);
int ctdr_raster_backward_f64(const double* grad_im, const double* im,
                             const double* p6, const double* ff, const double* len3,
                             const int32_t* labels2, const double* mus, int n_mus,
                             int B, int F, int H, int W, double max_sino_val,
                             double* grad_p6, double* grad_len3, double* grad_mus,
                             void* workspace, size_t workspace_bytes,
                             int64_t* stats, unsigned flags, ctdr_stream_t stream);
int ctdr_vertex_forward_f64(int geom_kind, const double* geom, int nangles,
                            const int64_t* idx_angles, int B,
                            const double* verts, int V, const int32_t* faces, int F,
                            int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                            double* p6, double* len3, double* ff, ctdr_stream_t stream);
int ctdr_vertex_backward_f64(int geom_kind, const double* geom, int nangles,
                             const int64_t* idx_angles, int B,
                             const double* verts, int V, const int32_t* faces, int F,
                             int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                             const double* grad_p6, const double* grad_len3,
                             double* grad_verts, ctdr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused projector — mesh + geometry in, sinogram out; replaces FP.forward end to end
 * (vertex stage -> Rasterizer -> get_valid_mask, fp_vecs.py:80-206) without materialising any
 * [B,F,*] tensor in the caller.  Angles are processed in windows sized to `workspace_bytes`.
 * ------------------------------------------------------------------------------------------- */
size_t ctdr_project_workspace_bytes(int B, int V, int F, int H, int W, int need_backward);

int ctdr_project_forward(int geom_kind, const float* geom, int nangles,
                         const int64_t* idx_angles, int B,
                         const float* verts, int V, const int32_t* faces, int F,
                         const int32_t* labels2, const float* mus, int n_mus,
                         int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                         float* im, int32_t* cnt,
                         void* workspace, size_t workspace_bytes,
                         int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* Fused projector, n_out sinograms in one pass (see ctdr_raster_forward_multi): what
 * Model.estimate_mu needs (vanilla.py:186-265) from a single sweep over the mesh. */
int ctdr_project_forward_multi(int geom_kind, const float* geom, int nangles,
                               const int64_t* idx_angles, int B,
                               const float* verts, int V, const int32_t* faces, int F,
                               const int32_t* labels2, const float* mus_kxn, int n_out, int n_mus,
                               int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                               float* im_kxbxhxw, const float* target, double* gram,
                               void* workspace, size_t workspace_bytes,
                               int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* grad_im [B,H,W], im [B,H,W] (forward output) -> ACCUMULATES grad_verts [V,3], grad_mus [n_mus].
 * vadj_start [V+1] / vadj_corner [3F] (int32, may be NULL unless CTDR_FLAG_DETERMINISTIC): for every
 * vertex the face corners (3*face + corner) that reference it, grouped by vertex in a fixed order
 * (CSR).  With CTDR_FLAG_DETERMINISTIC the sweep uses single-writer kernels and fixed summation
 * orders only (face-owner raster backward, adjacency gather for the vertices, tree reductions for
 * grad_mus and the loss): bit-reproducible results, no atomics. */
int ctdr_project_backward(int geom_kind, const float* geom, int nangles,
                          const int64_t* idx_angles, int B,
                          const float* verts, int V, const int32_t* faces, int F,
                          const int32_t* labels2, const float* mus, int n_mus,
                          int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                          const float* grad_im, const float* im,
                          float* grad_verts, float* grad_mus,
                          const int32_t* vadj_start, const int32_t* vadj_corner,
                          void* workspace, size_t workspace_bytes,
                          int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* Projection + masked-L2 residual step of the optimiser loop (ctdr/optimize.py:162-168,190 with
 * get_valid_mask, fp_vecs.py:10-13): phat = projector(mesh); valid = phat > -1e-6 && phat <=
 * max_sino_val - 1e-6; loss_sum = sum_valid (target - phat)^2, n_valid = #valid; the UNNORMALISED
 * gradient of loss_sum is accumulated into grad_verts / grad_mus (divide by n_valid — after the
 * cross-GPU all-reduce — to get d(mean)/d·).  out2 = device float64[2] {loss_sum, n_valid}
 * (accumulated; arrives zeroed).  phat [B,H,W] is written (needed by the reference's logging).
 * The forward sweep records the face batches of every tile in the workspace and the residual/backward
 * sweep replays them (no second binning); the workspace must come from
 * ctdr_project_workspace_bytes(..., need_backward = 1). */
int ctdr_project_residual_step(int geom_kind, const float* geom, int nangles,
                               const int64_t* idx_angles, int B,
                               const float* verts, int V, const int32_t* faces, int F,
                               const int32_t* labels2, const float* mus, int n_mus,
                               int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                               const float* target, float* phat,
                               float* grad_verts, float* grad_mus, double* out2,
                               const int32_t* vadj_start, const int32_t* vadj_corner,
                               void* workspace, size_t workspace_bytes,
                               int64_t* stats, unsigned flags, ctdr_stream_t stream);

/* Profiling variant of ctdr_project_residual_step: same work, same results, plus the device time of every
 * stage, measured with CUDA events recorded on `stream` between the stages (summed over the angle windows).
 * It SYNCHRONISES with the stream before returning (the one entry point that does) and writes
 * stage_ms_host[CTDR_STAGE_COUNT] (HOST memory, milliseconds). */
#define CTDR_STAGE_VERTEX_FORWARD  0  /* vertex transform + face assembly (p6/len3/ff, pixel boxes) + super-region lists */
#define CTDR_STAGE_RASTER_FORWARD  1  /* forward sweep (records its face batches) */
#define CTDR_STAGE_MEMSET          2  /* gradient-window and record-cursor memsets */
#define CTDR_STAGE_RASTER_STEP     3  /* residual + backward sweep: replay kernel + the walk kernel for unrecorded regions */
#define CTDR_STAGE_VERTEX_BACKWARD 4  /* transposed vertex stage into grad_verts */
#define CTDR_STAGE_COUNT           5
int ctdr_project_residual_step_timed(int geom_kind, const float* geom, int nangles,
                                     const int64_t* idx_angles, int B,
                                     const float* verts, int V, const int32_t* faces, int F,
                                     const int32_t* labels2, const float* mus, int n_mus,
                                     int H, int W, double spacing_x, double spacing_y, double max_sino_val,
                                     const float* target, float* phat,
                                     float* grad_verts, float* grad_mus, double* out2,
                                     const int32_t* vadj_start, const int32_t* vadj_corner,
                                     void* workspace, size_t workspace_bytes,
                                     int64_t* stats, unsigned flags, ctdr_stream_t stream,
                                     float* stage_ms_host);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU exchange of the angle-sharded residual step (the reference has no distributed code; this is the
 * one collective of the path: SUM over ranks of [grad_verts | grad_mus] and {loss_sum, n_valid}).
 * One process per GPU.  Every rank creates a peer block, hands its 64-byte IPC handle to the other ranks (any
 * transport: torch.distributed.all_gather_object, MPI, a file), opens theirs, and then sums a local payload
 * over all ranks with ONE kernel launch that reads the peers' blocks over NVLink (one-shot all-reduce; sized
 * for payloads up to ~1 MB, i.e. meshes up to ~80k vertices — a library all-reduce wins beyond that).
 * Sums are formed in rank order on every rank: all ranks end with bit-identical results.
 * ------------------------------------------------------------------------------------------- */
#define CTDR_PEER_HANDLE_BYTES 64
#define CTDR_PEER_MAX_RANKS    8
/* device bytes behind a peer block for payloads of up to payload_bytes (header + two payload slots) */
size_t ctdr_peer_block_bytes(size_t payload_bytes);
/* cudaMalloc + zero a peer block on the current device; handle_host (64 bytes, HOST memory, may be NULL when all
 * ranks live in one process) receives its CUDA IPC handle.  Synchronises the device once. */
int ctdr_peer_block_create(size_t payload_bytes, void** block, unsigned char* handle_host);
/* map another process's peer block into this one (peer access is enabled on demand) / unmap it */
int ctdr_peer_block_open(const unsigned char* handle_host, void** block);
int ctdr_peer_block_close(void* block);
int ctdr_peer_block_destroy(void* block);
/* 0 = fine, 1 = some launch gave up waiting for a peer (> ~2 s): its sums are invalid.  Sticky; synchronous copy. */
int ctdr_peer_block_status(const void* block, int* status_host);
/* local (device, 16-byte aligned) = f32_bytes of float32 followed by f64_bytes of float64 (both multiples of 16,
 * together <= the payload_bytes the blocks were created with); overwritten by the sums over all ranks.
 * blocks_host[r] = rank r's peer block as mapped in this process (blocks_host[rank] = the one this rank created).
 * Every rank must make the same sequence of calls with the same sizes.  Enqueued on `stream`; capturable. */
int ctdr_peer_allreduce(void* const* blocks_host, int world, int rank, size_t payload_bytes,
                        void* local, size_t f32_bytes, size_t f64_bytes, ctdr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CTDR_B200_H */
