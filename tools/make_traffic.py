"""profiles/r02_traffic_C2.json from an ncu --set full capture of the C2 step:
`python tools/make_traffic.py gpurun_out/prof_r02c.ncu-rep > profiles/r02_traffic_C2.json`.
DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) keyed by the bench's stage
names, plus the hashes of the kernel source files the capture was taken from: bench.py reports
`roofline.traffic` of a kernel only while its file's hash matches the tree (stale: null)."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = {"reads_to_global_kernel": "index_map", "sp_split_kernel": "sp_split", "sp_group_kernel": "sp_sort",
            "sp_wtile_kernel": "sp_tile", "bin_mean_kernel": "prof_bin"}


FILE_OF = {"index_map": "index.cu", "sp_split": "coverage_split.cu", "sp_sort": "coverage_split.cu",
           "sp_tile": "coverage_split.cu", "prof_bin": "profile.cu"}


def source_hash(name):
    """sha256 of one kernel source file (bench.py recomputes it)"""
    return hashlib.sha256(open(os.path.join(ROOT, "recoup_b200", "csrc", name), "rb").read()).hexdigest()


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        v = float(r[ix[key]].replace(",", ""))
        u = units[ix[key]]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

    out = {"source_hash": {f: source_hash(f) for f in sorted(set(FILE_OF.values()))}, "file_of": FILE_OF,
           "capture": os.path.basename(rep),
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, C2 step, one B200"}
    for r in data:
        name = r[ix["Kernel Name"]]
        for k, stage in STAGE_OF.items():
            if k in name and stage not in out:
                out[stage] = int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
