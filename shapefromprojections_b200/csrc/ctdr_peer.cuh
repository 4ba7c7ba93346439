// ctdr_peer.cuh — the one exchange of the angle-sharded residual step: SUM over ranks of
// [grad_verts | grad_mus] (fp32) and {loss_sum, n_valid} (fp64), done by ONE kernel over NVLink peer memory
// instead of a library all-reduce.  The reference has no multi-GPU code; this replaces what a port would write
// as dist.all_reduce (two NCCL launches, ~0.1 ms for 0.6 MB on 8 GPUs — 6 % of a 1.7 ms step).
//
// Every rank owns a "peer block" (cudaMalloc'd by ctdr_peer_block_create, mapped into the other ranks' address
// spaces through CUDA IPC): a header of arrival flags and two payload slots used alternately.
//   one-shot all-reduce:  copy my payload into my block  ->  tell every peer (release store of the epoch into ITS
//   flag array, straight over NVLink)  ->  wait until every peer has told me  ->  read all blocks and add them
//   in rank order 0..N-1 (every rank forms the same sums in the same order: results are bit-identical across
//   ranks)  ->  write the sums over my local payload.
// The payload is cut into slices, one CTA each, with one flag per (slice, rank): a slice starts adding as soon as
// ITS pieces have arrived everywhere — no grid-wide barrier, no cooperative launch.  Every CTA signals before it
// waits, so CTAs never wait on each other in a cycle, and all CTAs of the grid are resident (grid <= 64).
// Two payload slots (epoch parity): a rank can only reach epoch e+1's copy after every peer has signalled epoch
// e, i.e. after every peer finished READING epoch e-1 (stream order on the peer) — the slot being overwritten.
// The epoch lives in the block (advanced by the last CTA to finish), not in a kernel argument: the launch is
// identical every step and can be captured in a CUDA graph.
// A wait that lasts longer than ~2 s (a peer died, kernels serialised by a debugger) gives up, sets the sticky
// status word of the block and lets the kernel finish with garbage instead of hanging the GPU.
#pragma once
#include "ctdr_common.cuh"

namespace ctdr {

constexpr int    PEER_MAX_RANKS  = 8;      // one NVSwitch domain of B200s
constexpr int    PEER_MAX_SLICES = 64;     // CTAs per launch (all resident)
constexpr int    PEER_THREADS    = 512;
constexpr size_t PEER_HEADER_BYTES = 4096; // flags [PEER_MAX_SLICES][PEER_MAX_RANKS] u32 (2 KB) | control words
constexpr int    PEER_CTL_EPOCH = 512, PEER_CTL_DONE = 513, PEER_CTL_STATUS = 514;   // u32 indices into the header

struct PeerTable {
    unsigned char* block[PEER_MAX_RANKS];  // block[r] = rank r's peer block as mapped in THIS process
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// local: this rank's payload, n_f32_vec float4 followed by n_f64_vec double2 (16-byte aligned); overwritten by the sums.
// slot_bytes: size of one payload slot of a peer block (>= 16 * (n_f32_vec + n_f64_vec)).
__global__ void __launch_bounds__(PEER_THREADS)
k_peer_allreduce(const PeerTable T, int world, int rank, float4* __restrict__ local,
                 long long n_f32_vec, int n_f64_vec, size_t slot_bytes) {
    __shared__ int s_timeout;
    const int tid = threadIdx.x, s = blockIdx.x, nsl = gridDim.x;
    unsigned char* mine = T.block[rank];
    unsigned* hdr = reinterpret_cast<unsigned*>(mine);
    const unsigned epoch = ld_relaxed_gpu(hdr + PEER_CTL_EPOCH) + 1u;      // advanced by the last CTA of the previous launch
    const size_t slot = PEER_HEADER_BYTES + (size_t)(epoch & 1u) * slot_bytes;
    const long long nvec = n_f32_vec + n_f64_vec;
    const long long per = (nvec + nsl - 1) / nsl;
    const long long lo = min((long long)s * per, nvec), hi = min(lo + per, nvec);
    if (tid == 0) s_timeout = 0;

    // 1. my slice into my block
    float4* out = reinterpret_cast<float4*>(mine + slot);
    for (long long i = lo + tid; i < hi; i += PEER_THREADS) out[i] = local[i];
    __threadfence_system();
    __syncthreads();
    // 2. tell every rank (myself included) that slice s of rank `rank` is in place
    if (tid < world)
        st_release_sys(reinterpret_cast<unsigned*>(T.block[tid]) + s * PEER_MAX_RANKS + rank, epoch);
    // 3. wait for slice s of every rank
    if (tid < world) {
        const unsigned* f = hdr + s * PEER_MAX_RANKS + tid;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000LL) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (s_timeout && tid == 0) atomicExch(hdr + PEER_CTL_STATUS, 1u);
    // 4. sums in rank order, over my local payload (peer memory is read past L1: .cg)
    for (long long i = lo + tid; i < hi; i += PEER_THREADS) {
        if (i < n_f32_vec) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < PEER_MAX_RANKS; ++r) {
                if (r < world) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(T.block[r] + slot) + i);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
            local[i] = acc;
        } else {
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < PEER_MAX_RANKS; ++r) {
                if (r < world) {
                    const double2 v = __ldcg(reinterpret_cast<const double2*>(T.block[r] + slot) + i);
                    acc.x += v.x; acc.y += v.y;
                }
            }
            reinterpret_cast<double2*>(local)[i] = acc;
        }
    }
    // 5. the last CTA to finish advances the block's epoch for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(hdr + PEER_CTL_DONE, 1u) == (unsigned)(nsl - 1)) {
            hdr[PEER_CTL_DONE] = 0u;
            __threadfence();
            atomicExch(hdr + PEER_CTL_EPOCH, epoch);
        }
    }
}

}  // namespace ctdr
