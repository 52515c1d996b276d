"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and
share of the step.  Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/...txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = OrderedDict()
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
    total += ns
print("# %d launches, %.3f ms summed device time (serialised, cold-cache: compare SHARES)" %
      (sum(a[0] for a in agg.values()), total / 1e6))
print("%-90s %8s %12s %8s" % ("kernel", "launches", "total_us", "share"))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-90s %8d %12.1f %7.2f%%" % (name[:90], n, ns / 1e3, 100.0 * ns / total))
