#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches, total and mean us."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += us
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
total = sum(t[1] for t in tot.values())
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%4d total=%10.1f us mean=%9.1f us  %5.1f%%" % (k[:60], n, us, us / n, 100 * us / total))
print("TOTAL %.1f us" % total)
