#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/.
  python tools/ncu_summary.py launches <launches.csv>          # per-kernel time share
  python tools/ncu_summary.py full <prof.ncu-rep>              # key metrics per captured kernel
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ni = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        if r[ni] != "gpu__time_duration.sum":
            continue
        agg.setdefault(r[ki][:100], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# per-kernel device time (ncu gpu__time_duration.sum, serialised, cold cache): total {tot/1e3:.1f} us")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v)/1e3:10.1f} us {100*sum(v)/tot:5.1f}%  n={len(v):3d} avg={sum(v)/len(v)/1e3:8.1f} us  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("----", r[idx["Kernel Name"]], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
        for k in KEYS:
            if k in idx:
                print(f"  {k:88s} {r[idx[k]]} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
