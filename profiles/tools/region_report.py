"""Per-region instruction / stall-sample shares of k_raster from an ncu report (uses sass_by_line)."""
import csv, io, subprocess, sys, os
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_by_line as S

def main(rep, lib, src_path):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt))); hdr = rows[0]
    keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
            'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
    ni = hdr.index('Kernel Name')
    print('kernels:', [r[ni][:40] for r in rows[2:]])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f'{k:90s}', [r[i] for r in rows[2:]])
    src = open(src_path).read().split('\n')
    def find(tag): return next(i + 1 for i, l in enumerate(src) if tag in l)
    marks = [('1 batch setup', find('__device__ __forceinline__ void process_batch(')),
             ('2 cand decode + slot loads', find('for (int base = 0; base < T; base += 32)')),
             ('3 cand coverage', find('// coverage: diff_fp_for.cu:132-151')),
             ('4 cand accumulate', find('if (act) st.cand++;')),
             ('5 bwd per-face epilogue', find('if (IS_BWD && mine) {')),
             ('6 tile load/store', find('// tile load / store helpers')),
             ('7 phase A + empty fill', find('k_raster(const RasterParams P)')),
             ('8 tile dispatch + prologue', find('// ---- phase B')),
             ('9 list walk + enqueue', find("// walk the region's chunk list")),
             ('a tile epilogue / sync / flush', find('// tile epilogue'))]
    first = marks[0][1]
    def region(key):
        if key is None: return 'unknown'
        f, l = key
        if f == 'ctdr_common.cuh': return 'common (edge_setup etc.)'
        if f != 'ctdr_raster.cuh': return f
        if l < first: return '0 helpers (seg reduce, grads, bbox kernel)'
        name = marks[0][0]
        for n, ln in marks:
            if l >= ln: name = n
        return name
    for ksub_ncu, ksub in (('k_raster<(int)0', 'k_rasterILi0ELb0'), ('k_raster<(int)3', 'k_rasterILi3ELb0')):
        try:
            blk = S.ncu_rows(rep, ksub_ncu)
        except SystemExit:
            continue
        sass = S.sass_lines(lib, ksub)
        h = blk['hdr']; ie = h.index('Instructions Executed'); sm = h.index('# Samples')
        regions = defaultdict(lambda: [0, 0]); tot = [0, 0]
        for r, s in zip(blk['rows'], sass):
            k = region(s[2]); regions[k][0] += int(r[ie]); regions[k][1] += int(r[sm]); tot[0] += int(r[ie]); tot[1] += int(r[sm])
        print(blk['name'], 'warp-inst', tot[0], 'samples', tot[1], '(sass rows', len(blk['rows']), len(sass), ')')
        for k, (i, s) in sorted(regions.items()):
            print(f'  {k:45s} inst {100*i/tot[0]:6.2f}%  samples {100*s/tot[1]:6.2f}%')

if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3])
