"""Summarise an ncu --csv launch list (gpu__time_duration.sum) into per-kernel count / total / share."""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
agg = OrderedDict()
for k, v in rows:
    c = agg.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += v
tot = sum(v[1] for v in agg.values())
print("kernel launches under ncu (gpu__time_duration.sum, ns; cold-cache serialised: compare SHARES) — " + " ".join(sys.argv[2:]))
print("%-72s %6s %14s %7s" % ("kernel", "count", "total_ns", "share"))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %14d %6.1f%%" % (k[:70], c, v, 100.0 * v / tot))
