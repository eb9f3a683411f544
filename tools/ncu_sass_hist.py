"""Opcode histogram (instructions executed, stall samples) of one kernel of an .ncu-rep.
usage: ncu_sass_hist.py report.ncu-rep kernel-regex"""
import csv
import subprocess
import sys
from collections import Counter

rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
c, cs = Counter(), Counter()
tot = tots = n = 0
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break          # first kernel instance only
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    src = r[ix["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    e, s = int(r[ix["Instructions Executed"]]), int(r[ix["Warp Stall Sampling (All Samples)"]])
    c[op] += e
    cs[op] += s
    tot += e
    tots += s
    n += 1
print("SASS instructions %d, executed %d, stall samples %d" % (n, tot, tots))
for op, k in c.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 28):
    print("%-10s %12d %5.1f%%   stall %5.1f%%" % (op, k, 100.0 * k / tot, 100.0 * cs[op] / max(tots, 1)))
