#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples from an ncu report captured with --import-source on.

  python scripts/ncu_lines.py gpurun_out/tc_full.ncu-rep [top_n]
Aggregates the `--page source --print-source cuda,sass` view: for every kernel, warp instructions executed and
stall samples per CUDA source line (top N), plus per-file totals."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
kern = fpath = None
hdr = None
data = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))   # kernel -> (file,line) -> [inst, samples, text]
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
    elif row[0] == "Function Name":
        kern = row[1].split("(")[0].split("::")[-1] + ("<1>" if "(bool)1" in row[1] else "<0>" if "(bool)0" in row[1] else "")
    elif row[0] == "Line No":
        hdr = row
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and row[0].isdigit():
        try:
            d = data[kern][(fpath, int(row[0]))]
            d[0] += int(row[ii]); d[1] += int(row[si]); d[2] = row[1].strip()
        except ValueError:
            pass
for k, lines in data.items():
    tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
    print(f"== {k}: {tot/1e6:.1f} M warp instructions, {ts} samples")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {v[0]/tot*100:5.1f}% inst {v[1]/max(ts,1)*100:5.1f}% smp  {f}:{ln:<4d} {v[2][:110]}")
