"""Markdown summary of an `ncu --set full` report: one row per captured kernel with the metrics the
north star asks for (HBM GB/s, L2 atomic/reduction traffic, FP32 pipe utilisation) next to the B200 peaks.

    python profiles/tools/summarize.py <report.ncu-rep> [hbm_peak_gbs]
"""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    mp = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
    peak = json.load(open(mp))["hbm_gbs"] if os.path.exists(mp) else 6650.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); h, units = rows[0], rows[1]

def col(r, name, default=float("nan")):
    if name not in h:
        return default
    v = r[h.index(name)].replace(",", "")
    try:
        x = float(v)
    except ValueError:
        return default
    u = units[h.index(name)]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
             "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return x * scale

print(f"| kernel | time ms | DRAM read MB | DRAM write MB | HBM GB/s | of {peak:.0f} GB/s | L2 RED/ATOM requests | L2 atomic unit % of peak | "
      "FP32 (fma) pipe % | FP64 pipe % | LSU pipe % | issue active % | warps active % | regs | warp-inst |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("ctdr::", "")
    t = col(r, "gpu__time_duration.sum")
    rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
    red = col(r, "lts__t_requests_srcunit_tex_op_red.sum", 0.0) + col(r, "lts__t_requests_srcunit_tex_op_atom.sum", 0.0)
    atom_pct = col(r, "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    print(f"| `{name}` | {t*1e3:.3f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/t/1e9:.0f} | {(rd+wr)/t/1e9/peak*100:.1f} % | "
          f"{red:.3g} | {atom_pct:.1f} | {col(r,'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{col(r,'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{col(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{col(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{col(r,'launch__registers_per_thread'):.0f} | {col(r,'smsp__inst_executed.sum'):.3g} |")
