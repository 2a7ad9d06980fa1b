"""Aggregate an ncu SASS source page by CUDA source line.

ncu's `--page source --csv` gives per-SASS-instruction counters but (in CSV) no source correlation; nvdisasm
`--print-line-info` gives the line of every instruction of the same cubin.  Both list the kernel's instructions in
address order, so they are joined by index.

    cuobjdump -xelf all HNSWIndex.Native.so          # -> <unit>.sm_100a.cubin
    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_by_line.py src.csv build_kernels.sm_100a.cubin build_select_kernelILi0 [top]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    src_csv, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, inside = [], ("?", 0), False
    for ln in dis:
        if ln.startswith("//--------------------- .text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((cur, m.group(2)))
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = rows[hdr_i + 1:]
    col = {h: i for i, h in enumerate(hdr)}
    if len(body) != len(lines):
        print(f"warning: {len(body)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])  # inst executed, samples, not-issued samples
    tot = [0, 0, 0]
    for (loc, _), r in zip(lines, body):
        ie = int(r[col["Instructions Executed"]] or 0)
        sm = int(r[col["# Samples"]] or 0)
        ni = int(r[col["Warp Stall Sampling (Not-issued Samples)"]] or 0)
        for k, v in enumerate((ie, sm, ni)):
            agg[loc][k] += v
            tot[k] += v
    print(f"total: inst {tot[0]}  samples {tot[1]}  not-issued {tot[2]}")
    print(f"{'file:line':32s} {'inst%':>7s} {'samp%':>7s} {'stall%':>7s}")
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{loc[0] + ':' + str(loc[1]):32s} {100 * v[0] / max(1, tot[0]):7.2f} {100 * v[1] / max(1, tot[1]):7.2f} {100 * v[2] / max(1, tot[2]):7.2f}")
    # per-file totals
    byfile = defaultdict(lambda: [0, 0])
    for loc, v in agg.items():
        byfile[loc[0]][0] += v[0]
        byfile[loc[0]][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
        print(f"{f:32s} inst {100 * v[0] / max(1, tot[0]):6.2f}%  samples {100 * v[1] / max(1, tot[1]):6.2f}%")


if __name__ == "__main__":
    main()
