"""Summarise ncu outputs into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv      # per-kernel share of a launch list
  python tools/ncu_summary.py report gpurun_out/shade.ncu-rep       # key metrics of one --set full capture
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        name = r[ki].split("(")[0]
        tot[name][0] += 1
        tot[name][1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total/1e3:.3f} ms summed (ncu: serialised, cold cache -- compare SHARES)")
    print(f"{'kernel':70s} {'launches':>8s} {'us total':>12s} {'us avg':>10s} {'share':>7s}")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {t:12.1f} {t/n:10.1f} {100*t/total:6.1f}%")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"# {path}: {name}")
        for k in KEYS:
            if k in hdr:
                print(f"{k:90s} {vals[hdr.index(k)]} {rows[1][hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
