#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) or a launch-list CSV into markdown.

    python tools/ncu_summary.py report.ncu-rep  > profiles/xyz.md
    python tools/ncu_summary.py --launches launches.csv > profiles/xyz_launches.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA (FP32) pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst (of 32)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("sass__inst_executed_local_loads", "local loads (inst)"),
    ("sass__inst_executed_local_stores", "local stores (inst)"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("b2pt::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print("| kernel | launches | total us | share % | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"| `{k[:80]}` | {cnt[k]} | {v:.1f} | {100 * v / T:.1f} | {v / cnt[k]:.1f} |")


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "") for r in data]
    print("| metric | unit | " + " | ".join(f"#{i} `{n[:28]}`" for i, n in enumerate(names)) + " |")
    print("|---|---|" + "---:|" * len(names))
    for key, label in KEYS:
        if key in col:
            i = col[key]
            print(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1])
