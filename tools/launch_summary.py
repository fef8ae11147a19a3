"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, total, mean, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launch_list_summary.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[ki]).replace("smle::", "")
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += us
tot = sum(c[1] for c in agg.values())
print(f"# ncu launch list: `{sys.argv[2] if len(sys.argv) > 2 else ''}`")
print("# (cold-cache, serialised: compare SHARES with bench.py's live CUDA-event kernel_ms, not absolutes)")
print("kernel,launches,total_us,avg_us,share_pct")
for k, (n, t) in agg.items():
    print(f"{k},{n},{t:.1f},{t / n:.2f},{100 * t / tot:.1f}")
