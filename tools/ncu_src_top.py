"""Top stall sites of one kernel from an ncu report: `python tools/ncu_src_top.py rep.ncu-rep kernel_regex [N]`.
Prints the N SASS lines with the most warp-stall samples (in program order) with their main stall reasons."""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows if len(r) == len(hdr) and r[isamp].isdigit()]
    # the report may list several launches of the kernel back to back: keep the first listing
    seen, first = set(), []
    for r in data:
        if r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    data = first
    tot = sum(int(r[isamp]) for r in data)
    print("samples %d, warp instructions %d, SASS lines %d" % (tot, sum(int(r[iex]) for r in data), len(data)))
    agg = {h: sum(int(r[i]) for r in data) for h, i in stalls}
    print("stalls:", {h: v for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 50 > tot})
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
    for i in sorted(top):
        r = data[i]
        why = sorted(((int(r[j]), h) for h, j in stalls), reverse=True)[:2]
        print("%5d %-64s %7s %9s  %s" % (i, r[isrc].strip()[:64], r[isamp], r[iex],
                                       " ".join("%s=%d" % (h[6:], v) for v, h in why if v)))


if __name__ == "__main__":
    main()
