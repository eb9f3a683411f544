"""Minimal device-resident step loop for ncu: rcp_reads_load -> rcp_coverage -> rcp_profile_matrix
on one workload, inputs already in HBM.  usage: prof_step.py [workload] [scale] [steps] [path]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import recoup_b200 as rb  # noqa: E402
import workloads as W  # noqa: E402
from recoup_b200 import _lib  # noqa: E402
from recoup_b200.ranges import getRegionalRanges  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
path = sys.argv[4] if len(sys.argv) > 4 else "auto"
rb.init(0)
rb.set_coverage_path(path)
L = _lib.lib
dev = torch.device("cuda", 0)
w = W.CONFIGS[wl](scale=scale, seed=1001)
assert w["region"] != "rna"
genes = rb.GRanges(w["region_chrom"], w["region_start"], w["region_end"], strand=w["region_strand"],
                   seqlevels=w["chrom_names"])
win = getRegionalRanges(genes, w["region"], w["flank"])
R, N = len(win), len(w["read_start"])
clen = np.ascontiguousarray(w["chrom_len"], dtype=np.int64)
d = [torch.from_numpy(np.ascontiguousarray(w[k])).to(dev)
     for k in ("read_chrom", "read_start", "read_end", "read_strand")]
dw = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (win.seqnames, win.start, win.end, win.strand)]
vp = lambda t: C.c_void_p(t.data_ptr())
f1, f2 = w["flank"]
bp = w["bin_params"]
eq = 1 if w["region"] in ("tss", "tes", "custom") else 0
out = None
for it in range(steps):
    h, cov = C.c_int(0), C.c_int(0)
    _lib.check(L.rcp_reads_load(N, vp(d[0]), vp(d[1]), vp(d[2]), vp(d[3]), clen.shape[0],
                                clen.ctypes.data_as(C.POINTER(C.c_int64)), int(w["frag_len"]),
                                _lib.MEM_DEVICE, C.byref(h)))
    _lib.check(L.rcp_coverage(h.value, R, vp(dw[0]), vp(dw[1]), vp(dw[2]), vp(dw[3]), 1,
                              _lib.STRAND_ANY, _lib.MEM_DEVICE, C.byref(cov)))
    if out is None:
        nc = C.c_int64(0)
        _lib.check(L.rcp_profile_ncols(cov.value, eq, f1, f2, bp["flankBinSize"], bp["regionBinSize"],
                                       C.byref(nc)))
        out = torch.empty((nc.value, R), dtype=torch.float64, device=dev)
    _lib.check(L.rcp_profile_matrix(cov.value, eq, f1, f2, bp["flankBinSize"], bp["regionBinSize"],
                                    _lib.STAT[bp["sumStat"]], _lib.INTERP[bp["interpolation"]], 42, 0,
                                    vp(out), R, _lib.MEM_DEVICE))
    L.rcp_coverage_free(cov.value)
    L.rcp_reads_free(h.value)
L.rcp_sync()
print("ok", wl, N, R, int(L.rcp_launch_count(0)))
