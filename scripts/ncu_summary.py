#!/usr/bin/env python
"""Summarise ncu outputs into the CSVs kept under profiles/.

  ncu_summary.py launches <launch list csv from `ncu --metrics gpu__time_duration.sum --csv`> <out.csv>
      -> per kernel: launches, summed / mean duration, share of the sum
  ncu_summary.py full <file.ncu-rep> <out.csv>
      -> one row per captured launch with the metrics the roofline discussion uses
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

FULL = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "nvlink__bytes_received.sum", "nvltx__bytes.sum", "nvlrx__bytes.sum"]


def short(name: str) -> str:
    m = re.search(r"(k_[a-z0-9_]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    i_unit = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        a = agg.setdefault(short(r[i_name]), [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "sum_ms", "mean_ms", "share_of_kernel_time"])
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{ms:.3f}", f"{ms / n:.4f}", f"{ms / total:.4f}"])
        w.writerow(["TOTAL", sum(a[0] for a in agg.values()), f"{total:.3f}", "", "1.0"])


def full(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in FULL if c in idx]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{c} [{units[idx[c]]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[idx["Kernel Name"]])] + [r[idx[c]] for c in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
