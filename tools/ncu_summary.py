#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` into one row per kernel (mean over the captured launches) with the
metrics DESIGN.md quotes.  usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py > profiles/xxx.csv"""
import csv
import sys
from collections import OrderedDict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx["Kernel Name"]
    cols = [w for w in WANT if w in idx]
    agg = OrderedDict()
    for r in data:
        key = (r[name_i].split("(")[0].replace("void ", "").strip(), r[idx["launch__grid_size"]])
        a = agg.setdefault(key, [0, [0.0] * len(cols)])
        a[0] += 1
        for k, c in enumerate(cols):
            try:
                a[1][k] += float(r[idx[c]].replace(",", ""))
            except ValueError:
                pass
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "grid", "launches"] + ["%s [%s]" % (c, units[idx[c]]) for c in cols])
    for (name, grid), (n, sums) in agg.items():
        w.writerow([name, grid, n] + ["%.4g" % (s / n) for s in sums])


if __name__ == "__main__":
    main()
