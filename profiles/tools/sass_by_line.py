"""Aggregate an ncu SASS source page by CUDA source line.

    python profiles/tools/sass_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top]

ncu's CSV export of the source page is SASS-only; line numbers come from `nvdisasm -g` of the same
cubin (joined by instruction order).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
    out, cur, line, inside = [], None, None, False
    for ln in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            inside = kernel_sub in m.group(1)
            continue
        if ln.startswith("//--------------------- .text") or ln.startswith("//--------------------- ."):
            if inside and out:
                break
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def ncu_rows(rep, kernel_sub):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    for b in blocks:
        if kernel_sub in b["name"]:
            return b
    raise SystemExit(f"kernel {kernel_sub} not in report: {[b['name'] for b in blocks]}")


def main():
    rep, lib, ksub_ncu, ksub_sass = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    sass = sass_lines(lib, ksub_sass)
    blk = ncu_rows(rep, ksub_ncu)
    h = blk["hdr"]
    ie, sm = h.index("Instructions Executed"), h.index("# Samples")
    rows = blk["rows"]
    if len(rows) != len(sass):
        print(f"warning: {len(rows)} ncu instructions vs {len(sass)} disassembled", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r, s in zip(rows, sass):
        key = s[2]
        agg[key][0] += int(r[ie]); agg[key][1] += int(r[sm]); agg[key][2] += 1
        tot_i += int(r[ie]); tot_s += int(r[sm])
    print(f"kernel {blk['name']}: {tot_i} warp-instructions, {tot_s} stall samples")
    print(f"{'file:line':32s} {'inst%':>7s} {'samples%':>9s} {'#sass':>6s}")
    for key, (i, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        name = f"{key[0]}:{key[1]}" if key else "?"
        print(f"{name:32s} {100*i/max(tot_i,1):7.2f} {100*s/max(tot_s,1):9.2f} {n:6d}")


if __name__ == "__main__":
    main()
