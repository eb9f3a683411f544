"""Host mirror of /root/reference/R/profile.R over librecoup_b200.so.

    profileMatrix(input, flank, binParams, rc=NULL)                          profile.R:1-98
    binCoverageMatrix(cvrg, binSize, stat, interpolation, flank, where, rc)  profile.R:153-212
    baseCoverageMatrix(cvrg, flank, where, rc)                               profile.R:100-151

The `$profile` of a sample is a double matrix [regions x bins] with rownames = region names
(contract T2); here a Fortran-ordered numpy array (`ProfileMatrix`) carrying
`rownames` / `colnames`.  `rc` is accepted and ignored.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from .coverage import CoverageList, _message

R_SEED = 42   # splitVector(..., seed=42), util.R:15


class ProfileMatrix(np.ndarray):
    def __new__(cls, array, rownames=None, colnames=None):
        obj = np.asfortranarray(array).view(cls)
        obj.rownames = rownames
        obj.colnames = colnames
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.rownames = getattr(obj, "rownames", None)
        self.colnames = getattr(obj, "colnames", None)


def _release_pinned(ptr):
    try:
        _lib.lib.rcp_host_free(C.c_void_p(ptr))
    except Exception:
        pass


def _out(n_rows, n_cols):
    """Column-major fp64 result matrix in page-locked host memory (rcp_host_alloc): the
    device->host copy runs at PCIe rate and the buffer returns to the library's pool when the
    last numpy view of it dies."""
    n = int(n_rows) * int(n_cols)
    if n == 0:
        return np.zeros((n_rows, n_cols), dtype=np.float64, order="F")
    _lib.ensure_init()
    ptr = C.c_void_p(0)
    _lib.check(_lib.lib.rcp_host_alloc(n * 8, C.byref(ptr)))
    raw = (C.c_double * n).from_address(ptr.value)
    weakref.finalize(raw, _release_pinned, ptr.value)
    return np.frombuffer(raw, dtype=np.float64).reshape((n_rows, n_cols), order="F")


def _ld(mat):
    return max(mat.shape[0], 1)


def _sample_kind(kind):
    return _lib.SAMPLE_KIND[kind]


def binCoverageMatrix(cvrg, binSize=1000, stat="mean", interpolation="auto", flank=None,
                      where="center", rc=None, seed=R_SEED, sample_kind="Rejection"):
    """profile.R:153-212 (+ splitVector, util.R:15-85)."""
    if isinstance(stat, (list, tuple)):
        stat = stat[0]
    if isinstance(interpolation, (list, tuple)):
        interpolation = interpolation[0]
    if not isinstance(cvrg, CoverageList):
        raise TypeError("cvrg must be the CoverageList returned by calcCoverage")
    f1, f2 = (0, 0) if flank is None else (int(flank[0]), int(flank[1]))
    w = _lib.WHERE["whole"] if flank is None else _lib.WHERE[where]
    out = _out(len(cvrg), int(binSize))
    _lib.check(_lib.lib.rcp_bin_matrix(cvrg.handle, w, f1, f2, int(binSize), _lib.STAT[stat],
                                       _lib.INTERP[interpolation], int(seed),
                                       _sample_kind(sample_kind),
                                       out.ctypes.data_as(C.c_void_p), _ld(out), _lib.MEM_HOST))
    return ProfileMatrix(out, rownames=cvrg.names,
                         colnames=[str(i + 1) for i in range(int(binSize))])


def baseCoverageMatrix(cvrg, flank=None, where="upstream", rc=None):
    """profile.R:100-151."""
    if not isinstance(cvrg, CoverageList):
        raise TypeError("cvrg must be the CoverageList returned by calcCoverage")
    if flank is None:
        lens = cvrg.lengths()
        nz = lens[lens > 0]
        size = int(nz[0]) if nz.size else 0                              # profile.R:103-111
        f1 = f2 = 0
        w = _lib.WHERE["whole"]
    else:
        f1, f2 = int(flank[0]), int(flank[1])
        size = f1 if where == "upstream" else f2                          # profile.R:128,135
        w = _lib.WHERE[where]
    out = _out(len(cvrg), size)
    if size > 0 and len(cvrg) > 0:
        _lib.check(_lib.lib.rcp_base_matrix(cvrg.handle, w, f1, f2, size,
                                            out.ctypes.data_as(C.c_void_p), _ld(out),
                                            _lib.MEM_HOST))
    return ProfileMatrix(out, rownames=cvrg.names)


def haveEqualLengths(cvrg):
    """profile.R:6-10: lengths of the first sample's coverages, zero-length (NULL) dropped."""
    lens = cvrg.lengths()
    lens = lens[lens != 0]
    return bool(np.all(lens == lens[0])) if lens.size else True


def _profile_one(cvrg, equal, flank, binParams, seed, sample_kind):
    f1, f2 = int(flank[0]), int(flank[1])
    fbs = int(binParams.get("flankBinSize", 0))
    rbs = int(binParams.get("regionBinSize", 0))
    stat = binParams.get("sumStat", "mean")
    interp = binParams.get("interpolation", "auto")
    if isinstance(stat, (list, tuple)):
        stat = stat[0]
    if isinstance(interp, (list, tuple)):
        interp = interp[0]
    ncols = C.c_int64(0)
    _lib.check(_lib.lib.rcp_profile_ncols(cvrg.handle, int(equal), f1, f2, fbs, rbs, C.byref(ncols)))
    out = _out(len(cvrg), ncols.value)
    if out.size:
        _lib.check(_lib.lib.rcp_profile_matrix(cvrg.handle, int(equal), f1, f2, fbs, rbs,
                                               _lib.STAT[stat], _lib.INTERP[interp], int(seed),
                                               _sample_kind(sample_kind),
                                               out.ctypes.data_as(C.c_void_p), _ld(out),
                                               _lib.MEM_HOST))
    return ProfileMatrix(out, rownames=cvrg.names)


def profileMatrix(input, flank, binParams, rc=None, seed=R_SEED, sample_kind="Rejection"):
    """profile.R:1-98.  Fills `profile` of every sample dict; returns `input` unchanged when all
    samples already have one (profile.R:2-4).  The equal-length decision is taken on the FIRST
    sample only, as in the reference (profile.R:6)."""
    if all(x.get("profile") is not None for x in input):
        return input
    equal = haveEqualLengths(input[0]["coverage"])
    for x in input:
        _message("Calculating profile for ", x.get("name"))
        x["profile"] = _profile_one(x["coverage"], equal, flank, binParams, seed, sample_kind)
    return input


def coverageProfile(input, mask, binSize, strand=None, ignore_strand=True, scale=1.0, frag_len=0,
                    seed=R_SEED, sample_kind="Rejection"):
    """coverageRef + profileMatrix of ONE sample over equal-length windows in one call
    (coverage.R:1-42 followed by profile.R:83-96, sumStat = "mean"), through rcp_coverage_profile:
    with binSize >= 1 the per-base coverage never reaches HBM (the tile kernel hands its bin sums
    on); binSize = 0 gives the per-base matrix.  Returns (ProfileMatrix, is_null) -- the rows of
    NULL coverages are zero, as in profile.R:191-197.  Windows of different lengths are an error
    (use calcCoverage + profileMatrix, whose unequal-length branch this call does not have)."""
    from .coverage import _as_reads, _map_chrom, _ptr, device_reads
    from .ranges import GRanges, strand_to_code
    reads = _as_reads(input)
    if reads is None or not isinstance(mask, GRanges):
        raise ValueError("coverageProfile takes decoded reads (GRanges) and a GRanges mask")
    strand_filter = int(strand_to_code(strand)[0]) if strand is not None else _lib.STRAND_ANY
    dr = device_reads(reads, frag_len)
    chrom = np.ascontiguousarray(_map_chrom(mask, reads), dtype=np.int32)
    start = np.ascontiguousarray(mask.start, dtype=np.int32)
    end = np.ascontiguousarray(mask.end, dtype=np.int32)
    bad = chrom < 0
    if bad.any():
        chrom = np.where(bad, 0, chrom).astype(np.int32)
        start = np.where(bad, -1, start).astype(np.int32)
        end = np.where(bad, -1, end).astype(np.int32)
    R = len(mask)
    if int(binSize) > 0:
        ncols = int(binSize)
    else:
        w = (end.astype(np.int64) - start + 1)[~bad]
        ncols = int(w[0]) if w.size else 0
    out = _out(R, ncols)
    is_null = np.zeros(R, dtype=np.uint8)
    if R > 0 and ncols > 0:
        _lib.check(_lib.lib.rcp_coverage_profile(
            dr.handle, R, _ptr(chrom), _ptr(start), _ptr(end), _ptr(mask.strand), int(bool(ignore_strand)),
            strand_filter, int(binSize), int(seed), _sample_kind(sample_kind), float(scale),
            out.ctypes.data_as(C.c_void_p), _ld(out), is_null.ctypes.data_as(C.c_void_p), _lib.MEM_HOST))
    cols = [str(i + 1) for i in range(ncols)] if int(binSize) > 0 else None
    return ProfileMatrix(out, rownames=mask.names, colnames=cols), is_null.astype(bool)
