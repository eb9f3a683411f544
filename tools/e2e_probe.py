"""Host-side timing of the public API calls on the C2 workload (where does e2e time go?)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (pinned host arrays)

import recoup_b200 as rb  # noqa: E402
import workloads as W  # noqa: E402
from recoup_b200 import _lib, profile as P  # noqa: E402
from recoup_b200.ranges import getRegionalRanges  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rb.init(0)
L = _lib.lib
w = W.CONFIGS["C2"](scale=scale, seed=1001)
genes = rb.GRanges(w["region_chrom"], w["region_start"], w["region_end"], strand=w["region_strand"],
                   seqlevels=w["chrom_names"])
win = getRegionalRanges(genes, w["region"], w["flank"])
clen = np.ascontiguousarray(w["chrom_len"], dtype=np.int64)


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


pins = [pinned(w[k]) for k in ("read_chrom", "read_start", "read_end", "read_strand")]
hv = [p.numpy() for p in pins]
bp = w["bin_params"]
for it in range(4):
    T = [time.perf_counter()]
    lap = lambda: T.append(time.perf_counter())
    reads = rb.GRanges(hv[0], hv[1], hv[2], strand=hv[3], seqlevels=w["chrom_names"], seqlengths=clen)
    lap()
    rb.device_reads(reads, w["frag_len"]); L.rcp_sync(); lap()
    cov = rb.calcCoverage(reads, win, frag_len=w["frag_len"]); L.rcp_sync(); lap()
    eq = P.haveEqualLengths(cov); lap()
    nc = C.c_int64(0)
    L.rcp_profile_ncols(cov.handle, int(eq), 5000, 5000, 0, 200, C.byref(nc)); lap()
    out = P._out(len(cov), nc.value); lap()
    _lib.check(L.rcp_profile_matrix(cov.handle, int(eq), 5000, 5000, 0, 200, 0, 0, 42, 0,
                                    out.ctypes.data_as(C.c_void_p), out.shape[0], _lib.MEM_HOST)); lap()
    m = P.ProfileMatrix(out, rownames=cov.names); lap()
    cov.free()
    for dr in reads._device.values():
        dr.free()
    reads._device.clear(); lap()
    names = ["GRanges()", "device_reads", "calcCoverage", "haveEqualLengths", "profile_ncols", "_out(pinned)",
             "rcp_profile_matrix", "ProfileMatrix()", "free"]
    print("iter %d: " % it + "  ".join("%s=%.3f" % (n, 1e3 * (T[i + 1] - T[i])) for i, n in enumerate(names)))
    del m, out
