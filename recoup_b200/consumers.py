"""Host mirror of the reductions /root/reference/R/plot.R runs over finished `$profile` matrices,
over librecoup_b200.so:

    calcPlotProfiles(input, opts, ...)   average curve + band per sample        plot.R:949-990
    orderProfiles(input, opts)           heat-map row order (sum / max / avg)   plot.R:1035-1150
    heatmapScale(input, ...)             upper colour limit from quantiles      plot.R:513-545

`opts` is the nested dict of the reference (`opts["plotParams"]["sumStat"]`, ...).  Matrices are
the `ProfileMatrix` objects profileMatrix() returns (host, column-major); they are staged to the
GPU by the library, so a device pointer can be passed through the C ABI directly instead
(`mem = RCP_MEM_DEVICE`) when the matrix never left the device.
"""
import ctypes as C
import re

import numpy as np

from . import _lib

_ROW = {"sum": 0, "max": 1, "avg": 2}


def _f(mat):
    m = np.asfortranarray(mat, dtype=np.float64)
    if m.ndim != 2:
        raise ValueError("a profile must be a matrix")
    return m


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def colProfile(mat, avgfun="mean", scale="natural"):
    """apply(x, 2, avgfun) and apply(x, 2, sd | mad), after log2(x + 1) for scale "log2"."""
    if avgfun not in ("mean", "median"):
        raise ValueError("sumStat must be mean or median")
    m = _f(mat)
    _lib.ensure_init()
    center = np.empty(m.shape[1], dtype=np.float64)
    spread = np.empty(m.shape[1], dtype=np.float64)
    _lib.check(_lib.lib.rcp_matrix_col_profile(_vp(m), m.shape[0], m.shape[1], max(m.shape[0], 1),
                                               _lib.STAT[avgfun], 1 if scale == "log2" else 0,
                                               _lib.MEM_HOST, _vp(center), _vp(spread)))
    return center, spread


def calcPlotProfiles(input, opts, sdim=2, rc=None):
    """plot.R:949-990 without the smoothing branch: one dict(profile, upper, lower) per sample.
    smooth = TRUE (smooth.spline + its confidence band) is plotting-side statistics and not part
    of this library."""
    pp = opts["plotParams"]
    if pp.get("smooth"):
        raise NotImplementedError("plotParams$smooth: smooth.spline is outside the accelerated path")
    if sdim != 2:
        raise NotImplementedError("only column profiles (sdim = 2) are used by recoup")
    out = []
    for x in input:
        center, spread = colProfile(x["profile"], pp.get("sumStat", "mean"), pp.get("signalScale", "natural"))
        out.append({"profile": center, "upper": center + spread, "lower": center - spread})
    return out


def rowStat(mat, what):
    """apply(x, 1, sum | max | mean)"""
    m = _f(mat)
    _lib.ensure_init()
    out = np.empty(m.shape[0], dtype=np.float64)
    _lib.check(_lib.lib.rcp_matrix_row_stat(_vp(m), m.shape[0], m.shape[1], max(m.shape[0], 1),
                                            _ROW[what], _lib.MEM_HOST, _vp(out)))
    return out


def sortIndex(v, decreasing=False):
    """sort(v, decreasing, index.return=TRUE): dict(x = sorted values, ix = 1-based positions)"""
    v = np.ascontiguousarray(v, dtype=np.float64)
    _lib.ensure_init()
    ix = np.empty(v.shape[0], dtype=np.int32)
    n_out = C.c_int64(0)
    _lib.check(_lib.lib.rcp_order(_vp(v), v.shape[0], 1 if decreasing else 0, _lib.MEM_HOST, _vp(ix),
                                  C.byref(n_out)))
    ix = ix[:n_out.value]
    return {"x": v[ix - 1], "ix": ix}


def orderProfiles(input, opts, rc=None):
    """plot.R:1035-1150.  orderBy$what = "sum" | "max" | "avg" followed by the 1-based number of
    the reference sample, or by "a" for all samples together; anything else keeps the input
    order; orderBy$custom overrides everything."""
    ob = opts["orderBy"]
    dec = ob.get("order") == "descending"
    if ob.get("custom") is not None:
        return sortIndex(ob["custom"], dec)
    what = ob.get("what", "none")
    m = re.match(r"^(sum|max|avg)", what)
    if not m:
        n = np.asarray(input[0]["profile"]).shape[0]
        return {"ix": np.arange(1, n + 1, dtype=np.int32)}
    kind = m.group(1)
    ref = 1
    last = what[-1]
    if last == "a":
        ref = 0
    elif last.isdigit():
        ref = int(last)
    if ref == 0:
        per = np.stack([rowStat(x["profile"], kind) for x in input], axis=1)
        val = rowStat(per, kind)
    else:
        val = rowStat(input[ref - 1]["profile"], kind)
    return sortIndex(val, dec)


def matrixQuantile(mat, probs):
    """quantile(x, probs), type 7, over every cell"""
    m = _f(mat)
    _lib.ensure_init()
    p = np.ascontiguousarray(np.atleast_1d(probs), dtype=np.float64)
    out = np.empty(p.shape[0], dtype=np.float64)
    _lib.check(_lib.lib.rcp_matrix_quantile(_vp(m), m.shape[0], m.shape[1], max(m.shape[0], 1), _vp(p),
                                            p.shape[0], _lib.MEM_HOST, _vp(out)))
    return out


_QS = (0.95, 0.96, 0.97, 0.98, 0.99, 0.995, 0.999)


def heatmapScale(input, heatmapScale="common", heatmapFactor=1.0):
    """Upper colour limit of every sample's heat map (plot.R:513-545): "each" walks up the
    quantile ladder until it is non-zero; "common" takes the largest 95 % quantile."""
    if heatmapScale == "each":
        out = []
        for x in input:
            qs = matrixQuantile(x["profile"], _QS)
            nz = [q for q in qs if q != 0]
            out.append(heatmapFactor * (nz[0] if nz else 0.0))
        return out
    sup = max(float(matrixQuantile(x["profile"], [0.95])[0]) for x in input)
    return [heatmapFactor * sup] * len(input)
