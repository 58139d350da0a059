#!/usr/bin/env python
"""per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold caches, serialised launches:
compare SHARES, not absolutes).  usage: scripts/launch_summary.py raw.csv [header comment]"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    agg.setdefault(row["Kernel Name"], []).append(v)
ours = {k: v for k, v in agg.items() if "csdr::" in k or " k_" in (" " + k) or k.startswith("k_")}
tot = sum(sum(v) for v in ours.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("kernel,launches,avg_us,min_us,max_us,share_of_csdr_time")
for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
    print(f"\"{k}\",{len(v)},{sum(v)/len(v):.1f},{min(v):.1f},{max(v):.1f},{sum(v)/tot:.4f}")
