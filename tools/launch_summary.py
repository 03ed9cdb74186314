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
rd = defaultdict(float); wr = defaultdict(float)   # optional: dram__bytes_read.sum / dram__bytes_write.sum captured in the same pass
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
def _bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "")
    if r.get("Metric Name") == "dram__bytes_read.sum": rd[name] += _bytes(r["Metric Value"], r.get("Metric Unit", "byte"))
    if r.get("Metric Name") == "dram__bytes_write.sum": wr[name] += _bytes(r["Metric Value"], r.get("Metric Unit", "byte"))
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    unit = r.get("Metric Unit", "ns")
    rows.append((name, float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6), r["Grid Size"], r["Block Size"]))
tot = defaultdict(float); cnt = defaultdict(int)
for n, ms, g, b in rows:
    tot[n] += ms; cnt[n] += 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, cnt.get("grind_finalize_kernel", 3) // 3)
allms = sum(tot.values())
print(f"# {path}: {len(rows)} launches, {allms:.1f} ms of kernel time, {steps} steps -> {allms / steps:.2f} ms / step (serialised, cold cache)")
print(f"{'kernel':44s} {'launches/step':>13s} {'ms/step':>9s} {'share':>6s} {'avg us':>9s}" + (f" {'dram rd GB':>10s} {'dram wr GB':>10s}" if rd else ""))
for n in sorted(tot, key=lambda k: -tot[k]):
    print(f"{n:44s} {cnt[n] / steps:13.1f} {tot[n] / steps:9.3f} {100 * tot[n] / allms:5.1f}% {1e3 * tot[n] / cnt[n]:9.1f}" +
          (f" {rd[n] / steps / 1e9:10.3f} {wr[n] / steps / 1e9:10.3f}" if rd else ""))
