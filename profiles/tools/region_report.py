"""Share of executed warp-instructions and stall samples per part of k_raster, from an ncu report that was
captured with --import-source on (joins the SASS page with `nvdisasm -g` line info, see sass_by_line.py).

    python profiles/tools/region_report.py <report.ncu-rep> <libctdr_b200.so> <ctdr_raster.cuh>
"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_by_line as S  # noqa: E402

# (first source line containing the marker) -> region name; regions run until the next marker
MARKERS = [
    ("__device__ __forceinline__ bool box_overlap(", "box tests (chunk / face boxes vs tile)"),
    ("__device__ __forceinline__ void seg_reduce3(", "backward: moments -> per-face gradients"),
    ("struct NoCount {", "statistics / grad_mus cache"),
    ("__device__ __forceinline__ void process_batch(const RasterParams& P", "ordered walk (fragment-list mode)"),
    ("__device__ __forceinline__ bool cover_slot(", "exact coverage test (fp64 quotient)"),
    ("__device__ __noinline__ void redo_conflicted(", "forward fix-up of multiply-hit pixels"),
    ("__device__ __forceinline__ float* multi_xacc(", "multi-output pass"),
    ("struct SpanSetup {", "row spans: set-up"),
    ("__device__ __forceinline__ bool row_span(", "row spans: per row"),
    ("__device__ __forceinline__ void process_batch_blocked(", "batch: face loads + edge set-up"),
    ("    // ---- direct walk:", "batch: direct walk (fine-mesh variant)"),
    ("    // ---- general path:", "row spans: row loop"),
    ("    const unsigned nz = __ballot_sync(FULL, ncand > 0);", "batch: prefix sums, slot staging, range search"),
    ("    float M0 = 0.0f, A1 = 0.0f, A2 = 0.0f;", "candidate loop (walk + accumulate)"),
    ("        // Faces split over several lanes:", "backward: merge of partial moments"),
    ("    if (RECSPAN) {\n        // trim the conservative spans", "recording: trim spans to the exact coverage"),
    ("    if (RECSPAN) {\n        // the batch's exact coverage goes into its record", "recording: coverage -> batch record"),
    ("    if constexpr (IS_MULTI) {", "hit-mask check / fix-up / per-face RED"),
    ("constexpr int PREFIX_STRIDE = TILE + 1;", "backward rows: prefix sums of g"),
    ("struct SpanSetupB {", "backward rows: span set-up"),
    ("__device__ __noinline__ float4 bwd_row_exact(", "backward rows: exact rows (rare)"),
    ("__device__ __forceinline__ void process_batch_rows(", "backward rows: face loads + edge set-up"),
    ("            for (int r = 0; r < bhmax; ++r, rf += 1.0f, rowbase += PREFIX_STRIDE) {", "backward rows: row loop"),
    ("        if (__any_sync(FULL, exact_rows != 0u)) {", "backward rows: exact-row dispatch"),
    ("    st.cand += nfrag; st.frag += nfrag;", "backward rows: per-face gradients + RED"),
    ("// Backward batch of the replay kernel: lane per face, coverage READ from the batch record.", "replay batch: rows from recorded coverage"),
    ("    st.cand += nfrag; st.frag += nfrag;\n    if (queued && (M0 != 0.0", "replay batch: per-face gradients + RED"),
    ("// tile load / store helpers", "tile loads / stores"),
    ("struct ReplayCtaSmem {", "replay kernel: prologue, empty regions"),
    ("        const int tiles_x = (P.W + TILE - 1) / TILE, tiles_y = (P.H + TILE - 1) / TILE;\n        const int ntiles", "replay kernel: tile dispatch + residual prologue"),
    ("            // Pipeline: while batch k is processed out of one landing zone", "replay kernel: g tile, prefix sums, record pipeline"),
    ("    st.mus_flush(gmus);\n    __syncthreads();\n    for (int k = tid; k < P.n_mus; k += THREADS) {\n        const float v = gmus[k];\n        if (v != 0.0f) atomicAdd(P.grad_mus + k, v);\n    }\n    // masked L2", "replay kernel: tail"),
    ("k_raster(const RasterParams P) {", "kernel prologue + shared-address rebuilds"),
    ("        // ---- phase A:", "phase A (region list) + empty regions"),
    ("        // ---- phase B:", "tile dispatch + prologue"),
    ("            // walk the region's chunk list", "list walk + face queue"),
    ("            // tile epilogue", "tile epilogue + kernel tail"),
]


def main(rep, lib, src_path):
    text = open(src_path).read()
    src = text.split("\n")
    bounds = []
    for marker, name in MARKERS:
        key = marker.split("\n")[0]
        line = next((i + 1 for i, l in enumerate(src) if key in l), None)
        if line is None:
            raise SystemExit(f"marker not found: {marker!r}")
        bounds.append((line, name))
    bounds.sort()

    def region(key):
        if key is None:
            return "unattributed"
        f, l = key
        if f != os.path.basename(src_path):
            return {"ctdr_common.cuh": "edge set-up / helpers (ctdr_common.cuh)"}.get(f, "intrinsics headers (shuffles, ballots, loads, atomics)")
        name = "header"
        for ln, n in bounds:
            if l >= ln:
                name = n
        return name

    # the two kernels of the residual step: forward (records batches) and the replay-only residual/backward
    for ksub_ncu, ksub in (("k_raster<(int)0, (bool)0, (bool)0, (bool)0, (bool)1>", "k_rasterILi0ELb0ELb0ELb0ELb1EE"),
                           ("k_raster<(int)3, (bool)0, (bool)0, (bool)0, (bool)1", "k_rasterILi3ELb0ELb0ELb0ELb1"),
                           ("k_step_replay<(bool)0, (bool)1", "k_step_replayILb0ELb1"),
                           ("k_step_replay<(bool)0>", "k_step_replayILb0EE"),
                           ("k_raster<(int)0, (bool)0, (bool)0, (bool)0, (bool)1, (bool)1, (bool)1", "k_rasterILi0ELb0ELb0ELb0ELb1ELb1ELb1")):
        try:
            blk = S.ncu_rows(rep, ksub_ncu)
        except SystemExit:
            continue
        sass = S.sass_lines(lib, ksub)
        h = blk["hdr"]
        ie, sm = h.index("Instructions Executed"), h.index("# Samples")
        te, pe = h.index("Thread Instructions Executed"), h.index("Predicated-On Thread Instructions Executed")
        regions = defaultdict(lambda: [0, 0, 0, 0])
        tot = [0, 0, 0, 0]
        for r, s in zip(blk["rows"], sass):
            k = region(s[2])
            for j, col in enumerate((ie, sm, te, pe)):
                regions[k][j] += int(r[col]); tot[j] += int(r[col])
        print(f"{blk['name']}: {tot[0]:.4g} warp-instructions, {tot[1]} stall samples ({len(blk['rows'])} SASS instructions); "
              f"threads per instruction {tot[2] / tot[0]:.1f} active, {tot[3] / tot[0]:.1f} predicated-on "
              f"(smsp__thread_inst_executed_per_inst_executed)")
        for k, (i, s, t, p) in sorted(regions.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:58s} inst {100 * i / tot[0]:6.2f} %   samples {100 * s / tot[1]:6.2f} %   "
                  f"threads/inst {t / max(i, 1):5.1f} active {p / max(i, 1):5.1f} pred-on")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
