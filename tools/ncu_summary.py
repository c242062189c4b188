#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report (read here, on the CPU box): tools/ncu_summary.py <report.ncu-rep> [title]
columns: us | DRAM read MB | DRAM write MB | L2 hit % | L1 hit % | regs | warps active % | eligible warps/cycle |
top stall (warps per issue) | issue slots busy % | warp instructions (M) | kernel"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]


def f(d, k, scale=1.0):
  try:
    return float(d[k]) * scale
  except Exception:
    return float("nan")


print(f"# {title}")
print("# ncu --set full --clock-control none, B200.  Cold-cache, serialised launches: read the times as SHARES; bench.py times the")
print("# same chain with CUDA events.")
print("# columns: us | DRAM read MB | DRAM write MB | L2 hit % | L1 hit % | regs | warps active % | eligible warps/cycle | top stall "
      "(warps per issue) | issue busy % | warp inst (M) | kernel")
tot = 0.0
n = 0
units = dict(zip(hdr, rows[1]))
for r in rows[2:]:
  d = dict(zip(hdr, r))
  us = f(d, "gpu__time_duration.sum") / (1000.0 if units.get("gpu__time_duration.sum", "ns") in ("ns", "nsecond") else 1.0)
  rd, wr = f(d, "dram__bytes_read.sum"), f(d, "dram__bytes_write.sum")
  ur, uw = units.get("dram__bytes_read.sum", "byte"), units.get("dram__bytes_write.sum", "byte")
  sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
  stalls = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: f(d, k) for k in hdr
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "selected" not in k}
  top = max(stalls.items(), key=lambda kv: kv[1] if kv[1] == kv[1] else -1)
  name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void mono::", "").replace("mono::", "").replace("(anonymous namespace)::", "")
  print(f"{us:8.1f} | {rd * sc.get(ur, 1e-6):8.1f} | {wr * sc.get(uw, 1e-6):8.1f} | {f(d, 'lts__t_sector_hit_rate.pct'):5.1f} | "
        f"{f(d, 'l1tex__t_sector_hit_rate.pct'):5.1f} | {int(f(d, 'launch__registers_per_thread')):3d} | "
        f"{f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} | {f(d, 'smsp__warps_eligible.avg.per_cycle_active'):5.2f} | "
        f"{top[0]}={top[1]:.1f} | {f(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} | "
        f"{f(d, 'smsp__inst_executed.sum') / 1e6:6.1f} | {name}")
  tot += us
  n += 1
print(f"{tot:8.1f} us total of {n} launches")
