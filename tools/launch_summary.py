#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and total time per step.
usage: launch_summary.py launches.csv [n_steps]   (n_steps: how many bench steps the capture covered; default = number of
grind_finalize launches / 3, i.e. the three PoW grinds of one prove_shard)"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "")
    rows.append((name, float(r["Metric Value"]) / 1e6, r["Grid Size"], r["Block Size"]))
tot = defaultdict(float); cnt = defaultdict(int)
for n, ms, g, b in rows:
    tot[n] += ms; cnt[n] += 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, cnt.get("grind_finalize_kernel", 3) // 3)
allms = sum(tot.values())
print(f"# {path}: {len(rows)} launches, {allms:.1f} ms of kernel time, {steps} steps -> {allms / steps:.2f} ms / step (serialised, cold cache)")
print(f"{'kernel':44s} {'launches/step':>13s} {'ms/step':>9s} {'share':>6s} {'avg us':>9s}")
for n in sorted(tot, key=lambda k: -tot[k]):
    print(f"{n:44s} {cnt[n] / steps:13.1f} {tot[n] / steps:9.3f} {100 * tot[n] / allms:5.1f}% {1e3 * tot[n] / cnt[n]:9.1f}")
