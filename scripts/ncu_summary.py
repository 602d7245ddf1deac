#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python scripts/ncu_summary.py full     gpurun_out/prof.ncu-rep            > profiles/rNN_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    for r in rows[1:]:
        k = r[ki].split("(")[0][-70:]
        tot[k] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        cnt[k] += 1
    T = sum(tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source: {path}")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k:70s} {cnt[k]:5d} {tot[k]:12.1f} {tot[k] / T:7.3f} {tot[k] / cnt[k]:10.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on ; source: {path}")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")].split("(")[0])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:85s} {r[i]:>18s} {units[i]}")
        print()


def traffic(path):
    """profiles/r02_wkv_traffic.json: dram bytes per launch of the two training kernels (what bench.py's roofline.traffic
    reads): python scripts/ncu_summary.py traffic gpurun_out/tc_full.ncu-rep > profiles/r02_wkv_traffic.json"""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {"source": f"ncu --set full --clock-control none capture of scripts/run_pair.py at config c2 ({path})"}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        if "tc_bwd" in name:
            res["bwd_bytes"] = tot
        elif "tc_fwd" in name and ("<1" in name or "true" in name or "Lb1" in name):
            res.setdefault("fwd_bytes", tot)
        res.setdefault("kernels", []).append({"name": name.split("(")[0], "dram_bytes": tot})
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
