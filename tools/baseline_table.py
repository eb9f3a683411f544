"""BASELINE.md section 5 from bench lines: `python tools/baseline_table.py bench_1gpu.json [bench_2gpu.json ...]`
prints the markdown tables (per-config results on one GPU, scaling over N)."""
import json
import sys


def last_line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def frac(st, k):
    return "%.2f" % st[k]["frac"] if st and k in st else "—"


def main():
    one = last_line(sys.argv[1])
    rows = []
    st = one["roofline"]["stages"]
    rows.append(("C2 (headline)", one["config"]["reads_per_gpu"], one["ms_per_step"], one["value"],
                 one["region_bins_per_s"], frac(st, "reads_map"), frac(st, "coverage"), frac(st, "profile"),
                 one["config"]["coverage_path_used"]))
    cfg = one.get("configs") or {}
    if "C1" in cfg:
        c = cfg["C1"]
        rows.insert(0, ("C1 (host API, copies in)", 200000, c["ms_per_pass"], c["reads_per_s"], None, "—", "—", "—", "split"))
    for k in ("C3", "C4"):
        if k in cfg:
            c = cfg[k]
            rows.append((k + (" (per sample, 8 samples)" if k == "C3" else ""), c["reads_per_sample"], c["ms_per_sample"],
                         c["reads_per_s"], c["region_bins_per_s"], frac(c.get("stages"), "reads_map"),
                         frac(c.get("stages"), "coverage"), frac(c.get("stages"), "profile"), c["coverage_path_used"]))
    if "C5" in cfg:
        c = cfg["C5"]
        rows.append(("C5 (device-resident part)", 200000000, c.get("ms_device_resident", 0.0),
                     c.get("reads_per_s_device_resident", 0.0), None, "—", "—", "—", c["coverage_path_used"]))
        rows.append(("C5 (with the 8 GB matrix download)", 200000000, c["ms_per_step"], c["reads_per_s"], None, "—", "—",
                     "—", c["coverage_path_used"]))
    print("| Config (1 × B200) | reads / step | ms / step | reads/s | region-bins/s | map | coverage | profile | path |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---|")
    for r in rows:
        print("| %s | %.3g | %.3f | %.3g | %s | %s | %s | %s | %s |" % (
            r[0], r[1], r[2], r[3], "—" if r[4] is None else "%.3g" % r[4], r[5], r[6], r[7], r[8]))
    fu = one.get("fused")
    if fu:
        print("\nFused C2 (`rcp_coverage_profile`): %.3f ms/step = %.3g reads/s, matrix equals the two-stage one: %s."
              % (fu["ms_per_step"], fu["reads_per_s"], fu["matrix_equals_two_stage"]))
    for k in ("e2e", "e2e_bam_order", "e2e_dense_arrays"):
        if k in one:
            e = one[k]
            print("`%s`: %.2f ms/step = %.3g reads/s, %d MB H2D + %d MB D2H per step (%s)." % (
                k, e["ms_per_step"], e["value"], e["h2d_bytes_per_step"] // 10**6, e["d2h_bytes_per_step"] // 10**6,
                e.get("inputs", "")))
    im = cfg.get("import")
    if im:
        print("Read import (decode only, %d records / lines in HBM): BAM records %.3g records/s (%.0f GB/s), "
              "spliceAction split %.3g records/s, BED text %.3g lines/s (%.0f GB/s); from host memory %.1f ms / %.1f ms "
              "per file; host walk of the BAM record chain %.1f ms." % (
                  im["records"], im["bam_keep"]["records_per_s"], im["bam_keep"]["GB_per_s"],
                  im["bam_split"]["records_per_s"], im["bed"]["lines_per_s"], im["bed"]["GB_per_s"],
                  im["bam_keep"]["ms_from_host"], im["bed"]["ms_from_host"], im["bam_index_host_ms"]))
        if "bgzf_inflate_host" in im:
            b = im["bgzf_inflate_host"]
            print("BGZF inflate on the host (zlib, %d threads): %.2f GB/s of inflated bytes." % (b["threads"], b["GB_per_s_inflated"]))
    cb = one.get("cpu_baseline")
    if cb:
        print("CPU arm (%s, %d cores): %.3g reads/s on %s." % (cb["kind"], cb["cores"], cb["value"], cb["sample"][:80]))
    lines = [one] + [last_line(p) for p in sys.argv[2:]]
    print("\n| GPUs | C2 replicas: ms/step | reads/s (all GPUs) | efficiency | e2e ms/step | C5 one problem: device-resident ms | with download ms | C3 one sample, one problem: device-resident ms | with download ms |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    base = lines[0]["value"]
    for d in lines:
        s = d.get("strong") or (d.get("configs") or {}).get("C5") or {}
        s3 = d.get("strong_C3") or {}
        print("| %d | %.3f | %.3g | %.2f | %.2f | %s | %s | %s | %s |" % (
            d["n_gpus"], d["ms_per_step"], d["value"], d["value"] / (base * d["n_gpus"]), d["e2e"]["ms_per_step"],
            "%.2f" % s["ms_device_resident"] if "ms_device_resident" in s else "—",
            "%.1f" % s["ms_per_step"] if "ms_per_step" in s else "—",
            "%.2f" % s3["ms_device_resident"] if "ms_device_resident" in s3 else "—",
            "%.2f" % s3["ms_per_step"] if "ms_per_step" in s3 else "—"))


if __name__ == "__main__":
    main()
