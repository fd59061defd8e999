"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    rows.append((name, us))
agg = OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("# %s: %d launches, %.1f us total (ncu serialised, cold-cache: compare SHARES)" % (path, len(rows), tot))
print("%-60s %7s %12s %10s %7s" % ("kernel", "count", "total us", "mean us", "share"))
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %7d %12.1f %10.2f %6.1f%%" % (n[:60], c, us, us / c, 100 * us / tot))
