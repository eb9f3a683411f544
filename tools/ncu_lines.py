"""Per-source-line instruction counts / stall samples of one kernel of an .ncu-rep (needs
-lineinfo and --import-source on).  usage: ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if r and ("Source" in r) and ("Instructions Executed" in r):
        hdr = r
        start = i + 1
        break
if hdr is None:
    print(raw[:2000])
    sys.exit(1)
# rows with a line number are CUDA source lines carrying the totals of their SASS
iline, isrc = 0, 1
ie = hdr.index("Instructions Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
items = []
tot = tots = 0
for r in rows[start:]:
    if len(r) <= ie or not r[iline].strip().isdigit():
        continue
    try:
        e, st = int(r[ie] or 0), int(r[ist] or 0)
    except ValueError:
        continue
    items.append((e, st, int(r[iline]), r[isrc].strip()))
    tot += e
    tots += st
print("executed %d, stall samples %d" % (tot, tots))
for e, st, ln, src in sorted(items, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  L%-4d %s" % (100.0 * e / max(tot, 1), 100.0 * st / max(tots, 1), ln, src[:110]))
