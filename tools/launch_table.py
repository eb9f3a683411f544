"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table by kernel.
usage: launch_table.py launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("unnamed>::", "").strip()
    us = float(r[-1]) / (1000.0 if r[-2] in ("nsecond", "ns") else 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | avg us | total us | share |")
print("|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k[:90], n, t / n, t, 100.0 * t / tot))
print("\n%d launches, %.1f us in total" % (len(rows), tot))
