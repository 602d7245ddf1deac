#!/usr/bin/env python
"""Shared-memory wavefront budget per CUDA source line from an ncu report captured with --set full --import-source on.

  python scripts/ncu_smem_lines.py gpurun_out/tc_full.ncu-rep [chunks_per_launch] [top_n]

For every kernel in the report: total L1 shared-memory wavefronts (actual / ideal / excess = bank conflicts) and L1 tag
requests of global accesses, then the source lines that own them, normalised per chunk when `chunks_per_launch` is
given (B*H*T/16 = 32768 at config c2).  The chunked WKV kernels are bound by this pipe, so this
table is their budget."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
chunks = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
kern = fpath = hdr = None
data = defaultdict(lambda: defaultdict(lambda: [0, 0, 0, 0, ""]))   # kernel -> (file, line) -> [wave, ideal, excess, gtag, text]
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
    elif row[0] == "Function Name":
        kern = row[1].split("(")[0].split("::")[-1] + ("<1>" if "(bool)1" in row[1] else "<0>" if "(bool)0" in row[1] else "")
    elif row[0] == "Line No":
        hdr = row
        cols = [hdr.index(n) for n in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive",
                                       "L1 Tag Requests Global")]
    elif hdr and row[0].isdigit():
        d = data[kern][(fpath, int(row[0]))]
        for j, c in enumerate(cols):
            try:
                d[j] += int(row[c])
            except ValueError:
                pass
        d[4] = row[1].strip()
norm = chunks if chunks > 0 else 1.0
unit = "per chunk" if chunks > 0 else "total"
for k, lines in data.items():
    w, i, e, g = (sum(v[j] for v in lines.values()) for j in range(4))
    print(f"== {k}: shared wavefronts {w / norm:.1f} ({unit}), ideal {i / norm:.1f}, excess {e / norm:.1f}; "
          f"global L1 tag requests {g / norm:.1f}")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -(kv[1][0] + kv[1][3]))[:top]:
        if v[0] + v[3] == 0:
            break
        print(f"  {v[0] / norm:8.1f} wave {v[1] / norm:8.1f} ideal {v[3] / norm:7.1f} gtag  {f}:{ln:<4d} {v[4][:100]}")
