#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total/avg us, share.
    python tools/ncu_summary.py launches.csv [--one-step]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v))
if "--one-step" in sys.argv:      # the launches between the last two dino_loss_kernel launches = one training step (phase-shifted)
    marks = [i for i, (n, _) in enumerate(rows) if "dino_loss_kernel" in n]
    if len(marks) >= 2:
        rows = rows[marks[-2]:marks[-1]]
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in rows)
print(f"# {len(rows)} launches, {tot/1e3:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:8d} {t:12.1f} {t/c:10.2f} {100*t/tot:6.2f}%")
