"""ORACLE: ctypes wrapper around oracle/liboracle_recoup.so (oracle/recoup_oracle.c).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl
reference).  Returns the same python structures as oracle/recoup_oracle.py so the two
restatements can be compared with each other and with the CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import recoup_oracle as O
from .r_rng import rank_table

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_recoup.so")


def build(force=False):
    src = os.path.join(_HERE, "recoup_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_threads.restype = C.c_int
        L.orc_index_build.restype = C.c_void_p
        L.orc_index_build.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_int]
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_coverage.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]
        L.orc_coverage_list.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bin_matrix.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_double, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_base_matrix.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_void_p,
                                      C.c_int64]
        _lib = L
    return _lib


def threads():
    return lib().orc_threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Index:
    def __init__(self, chrom, start, end, strand, chrom_len, frag_len=0):
        self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.end = np.ascontiguousarray(end, dtype=np.int32)
        self.strand = None if strand is None else np.ascontiguousarray(strand, dtype=np.int8)
        self.chrom_len = np.ascontiguousarray(chrom_len, dtype=np.int64)
        self.ptr = lib().orc_index_build(self.start.shape[0], _p(self.chrom), _p(self.start),
                                         _p(self.end), _p(self.strand), self.chrom_len.shape[0],
                                         _p(self.chrom_len), int(frag_len))

    def __del__(self):
        try:
            if self.ptr:
                lib().orc_index_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class DenseCoverage:
    """Dense int32 buffer + offsets/len/is_null, as the C functions fill it."""

    def __init__(self, cov, off, length, is_null):
        self.cov, self.off, self.len, self.is_null = cov, off, length, is_null

    def to_list(self):
        out = []
        for r in range(self.len.shape[0]):
            if self.is_null[r]:
                out.append(None)
            else:
                o = int(self.off[r])
                out.append(self.cov[o:o + int(self.len[r])].astype(np.int64))
        return out


def _strand_filter(strand):
    return 2 if strand is None else int(strand)


def coverage(index, chrom, start, end, strand, ignore_strand=True, strand_filter=None):
    chrom = np.ascontiguousarray(chrom, dtype=np.int32)
    start = np.ascontiguousarray(start, dtype=np.int32)
    end = np.ascontiguousarray(end, dtype=np.int32)
    strand = np.ascontiguousarray(strand, dtype=np.int8)
    R = start.shape[0]
    L = np.maximum(end.astype(np.int64) - start.astype(np.int64) + 1, 0)
    off = np.concatenate([[0], np.cumsum(L)]).astype(np.int64)
    cov = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    length = np.zeros(R, dtype=np.int32)
    is_null = np.ones(R, dtype=np.uint8)
    lib().orc_coverage(index.ptr, R, _p(chrom), _p(start), _p(end), _p(strand), int(ignore_strand),
                       _strand_filter(strand_filter), _p(off), _p(cov), _p(length), _p(is_null))
    return DenseCoverage(cov, off, length, is_null)


def coverage_list(index, ptr, chrom, start, end, strand, ignore_strand=True, strand_filter=None):
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    chrom = np.ascontiguousarray(chrom, dtype=np.int32)
    start = np.ascontiguousarray(start, dtype=np.int32)
    end = np.ascontiguousarray(end, dtype=np.int32)
    strand = np.ascontiguousarray(strand, dtype=np.int8)
    G = ptr.shape[0] - 1
    w = np.maximum(end.astype(np.int64) - start.astype(np.int64) + 1, 0)
    csum = np.concatenate([[0], np.cumsum(w)])
    L = csum[ptr[1:]] - csum[ptr[:-1]]
    off = np.concatenate([[0], np.cumsum(L)]).astype(np.int64)
    cov = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    length = np.zeros(G, dtype=np.int32)
    is_null = np.ones(G, dtype=np.uint8)
    lib().orc_coverage_list(index.ptr, G, _p(ptr), _p(chrom), _p(start), _p(end), _p(strand),
                            int(ignore_strand), _strand_filter(strand_filter), _p(off), _p(cov),
                            _p(length), _p(is_null))
    return DenseCoverage(cov, off, length, is_null)


def concat3(left, center, right):
    """coverage.R:115-120 on the list form."""
    out = []
    for l, c, r in zip(left, center, right):
        out.append(None if (l is None or c is None or r is None) else np.concatenate([l, c, r]))
    return out


_WHERE = {None: 0, "whole": 0, "center": 1, "upstream": 2, "downstream": 3}


def bin_matrix(dense, n, stat="mean", interp="auto", flank=None, where="center", scale=1.0,
               seed=42, sample_kind="Rejection"):
    R = dense.len.shape[0]
    f1, f2 = (0, 0) if flank is None else (int(flank[0]), int(flank[1]))
    w = 0 if flank is None else _WHERE[where]
    rank = np.asarray(rank_table(n, seed, sample_kind), dtype=np.int32)
    out = np.zeros((R, n), dtype=np.float64, order="F")
    short = np.zeros(R, dtype=np.uint8)
    lib().orc_bin_matrix(R, _p(dense.cov), _p(dense.off), _p(dense.len), _p(dense.is_null), w, f1, f2,
                         n, _p(rank), 0 if stat == "mean" else 1, float(scale), _p(out), max(R, 1),
                         _p(short))
    for r in np.flatnonzero(short):          # interpolation rows: numpy restatement (util.R:17-73)
        o = int(dense.off[r])
        x = dense.cov[o:o + int(dense.len[r])].astype(np.float64)
        seg = x if flank is None else O._slice_where(x, (f1, f2), where)
        out[r] = scale * O.split_vector(seg, n, interp, stat, seed, sample_kind)
    return out


def base_matrix(dense, n_cols, flank=None, where="upstream", scale=1.0):
    R = dense.len.shape[0]
    f1, f2 = (0, 0) if flank is None else (int(flank[0]), int(flank[1]))
    w = 0 if flank is None else _WHERE[where]
    out = np.zeros((R, n_cols), dtype=np.float64, order="F")
    lib().orc_base_matrix(R, _p(dense.cov), _p(dense.off), _p(dense.len), _p(dense.is_null), w, f1, f2,
                          n_cols, float(scale), _p(out), max(R, 1))
    return out


def profile_matrix(dense, flank, bin_params, equal_lengths, scale=1.0, seed=42,
                   sample_kind="Rejection"):
    """profileMatrix (profile.R:1-98) over the dense C representation."""
    fbs = int(bin_params.get("flankBinSize", 0))
    rbs = int(bin_params.get("regionBinSize", 0))
    stat = bin_params.get("sumStat", "mean")
    interp = bin_params.get("interpolation", "auto")
    f1, f2 = int(flank[0]), int(flank[1])
    if equal_lengths:
        if rbs != 0:
            return bin_matrix(dense, rbs, stat, "auto", None, "center", scale, seed, sample_kind)
        nz = dense.len[dense.len > 0]
        return base_matrix(dense, int(nz[0]) if nz.size else 0, None, "upstream", scale)
    parts = []
    tot = float(f1 + f2)
    if fbs != 0:
        if f1:
            parts.append(bin_matrix(dense, O.r_round(2 * fbs * (f1 / tot)), stat, interp, (f1, f2),
                                    "upstream", scale, seed, sample_kind))
    elif f1:
        parts.append(base_matrix(dense, f1, (f1, f2), "upstream", scale))
    parts.append(bin_matrix(dense, rbs, stat, interp, (f1, f2), "center", scale, seed, sample_kind))
    if fbs != 0:
        if f2:
            parts.append(bin_matrix(dense, O.r_round(2 * fbs * (f2 / tot)), stat, interp, (f1, f2),
                                    "downstream", scale, seed, sample_kind))
    elif f2:
        parts.append(base_matrix(dense, f2, (f1, f2), "downstream", scale))
    return np.hstack(parts)
