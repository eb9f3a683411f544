"""Host mirror of the reference's file readers (ranges.R:102-146) over the device decoders
(SURVEY 8f N3):

    readBam(bam, sa = c("keep", "remove", "split"), sq = 0.75)        ranges.R:111-134
    readBed(bed, bg)                                                  ranges.R:136-146

What stays on the host: opening the file, the BGZF inflate (rcp_bgzf_inflate: zlib, one thread per
core, C++ inside the library), the BAM header (text + reference list) and the walk of the
length-prefixed record chain (rcp_bam_index, also inside the library).  Everything per read -- flag
filter, CIGAR walk, N-split, trim, text parsing, compaction in file order -- runs on the device (rcp_bam_decode / rcp_bed_decode); the decoded reads stay in
HBM and go to the hot path without a host round trip (rcp_reads_load_decoded).  Their coordinates
are copied to the host only when somebody reads them.
"""
import ctypes as C
import gzip
import struct
import weakref

import numpy as np

from . import _lib
from .coverage import _message
from .preprocess import SelectedGRanges, widthQuantile
from .ranges import GRanges


def _free_decoded(h):
    try:
        _lib.lib.rcp_decoded_free(h)
    except Exception:
        pass


class DecodedGRanges(GRanges):
    """The reads of one file, decoded on the device (a `decoded` handle)."""

    def __init__(self, handle, n, seqlevels, seqlengths):
        self.decoded_handle = int(handle)
        self._n = int(n)
        self.seqlevels = list(seqlevels)
        self.seqlengths = None if seqlengths is None else np.ascontiguousarray(seqlengths, dtype=np.int64)
        self.names = None
        self.seqnames_rle = None
        self.fixed_width = None
        self._device = {}
        self._host = None
        self._fin = weakref.finalize(self, _free_decoded, self.decoded_handle)

    def __len__(self):
        return self._n

    def _materialise(self):
        if self._host is None:
            n = self._n
            chrom, start, end = (np.zeros(n, dtype=np.int32) for _ in range(3))
            strand = np.zeros(n, dtype=np.int8)
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            _lib.check(_lib.lib.rcp_decoded_fetch(self.decoded_handle, p(chrom), p(start), p(end), p(strand), n))
            self._host = (chrom, start, end, strand)
        return self._host

    seqnames = property(lambda self: self._materialise()[0])
    start = property(lambda self: self._materialise()[1])
    end = property(lambda self: self._materialise()[2])
    strand = property(lambda self: self._materialise()[3])

    @property
    def width(self):
        return self.end.astype(np.int64) - self.start.astype(np.int64) + 1


def bgzfInflate(data, threads=0):
    """The inflated bytes of a BGZF file as a uint8 array: the library's multi-threaded inflate
    (rcp_bgzf_inflate), or python's gzip for data that is gzip but not BGZF."""
    buf = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    total, nblk = C.c_int64(0), C.c_int64(0)
    if _lib.lib.rcp_bgzf_size(vp(buf), buf.shape[0], C.byref(total), C.byref(nblk)) != _lib.RCP_OK:
        return np.frombuffer(gzip.decompress(bytes(data)), dtype=np.uint8)
    out = np.empty(total.value, dtype=np.uint8)
    _lib.check(_lib.lib.rcp_bgzf_inflate(vp(buf), buf.shape[0], vp(out), out.shape[0], int(threads)))
    return out


def bam_header(raw):
    """(seqlevels, seqlengths, offset of the first alignment record) of an inflated BAM."""
    raw = memoryview(raw)
    if bytes(raw[:4]) != b"BAM\x01":
        raise ValueError("not a BAM file (magic %r)" % bytes(raw[:4]))
    l_text, = struct.unpack_from("<i", raw, 4)
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, p)
    p += 4
    names, lens = [], []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, p)
        names.append(bytes(raw[p + 4:p + 4 + l_name - 1]).decode("ascii"))
        ln, = struct.unpack_from("<i", raw, p + 4 + l_name)
        lens.append(ln)
        p += 8 + l_name
    return names, np.asarray(lens, dtype=np.int64), p


def decodeBam(raw, split=False):
    """Inflated BAM bytes -> DecodedGRanges: readGAlignments + as(., "GRanges") (or
    unlist(grglist(.)) when `split`) + trim."""
    _lib.ensure_init()
    names, lens, first = bam_header(raw)
    rec = np.frombuffer(raw, dtype=np.uint8, offset=first)
    n_rec = C.c_int64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib.rcp_bam_index(vp(rec), rec.shape[0], C.byref(n_rec), None, 0))
    off = np.zeros(n_rec.value + 1, dtype=np.int64)
    _lib.check(_lib.lib.rcp_bam_index(vp(rec), rec.shape[0], C.byref(n_rec), vp(off), off.shape[0]))
    h, n_out = C.c_int(0), C.c_int64(0)
    _lib.check(_lib.lib.rcp_bam_decode(vp(rec), rec.shape[0], vp(off), n_rec.value, len(names),
                                       lens.ctypes.data_as(C.POINTER(C.c_int64)), 1 if split else 0,
                                       _lib.MEM_HOST, C.byref(h), C.byref(n_out)))
    return DecodedGRanges(h.value, n_out.value, names, lens)


def readBam(bam, sa="keep", sq=0.75, params=None):
    """ranges.R:111-134.  `bam`: a path, or the bytes of the file."""
    if sa not in ("keep", "remove", "split"):
        raise ValueError("sa must be one of keep, remove, split")
    if not 0 <= sq <= 1:
        raise ValueError("sq must be in [0, 1]")
    data = bam if isinstance(bam, (bytes, bytearray, memoryview)) else open(bam, "rb").read()
    raw = bgzfInflate(data)                     # BGZF blocks inflated in parallel (host zlib)
    reads = decodeBam(raw, split=(sa == "split"))
    if sa != "remove" or len(reads) == 0:
        return reads
    qu, n_kept = widthQuantile(reads, sq)
    _message("  Excluded ", len(reads) - n_kept, " reads")
    return SelectedGRanges(reads, n_kept, max_width=qu)


def readBed(bed, seqlevels, seqlengths=None):
    """ranges.R:136-146: import.bed(bed, trackLine = FALSE); the reference then looks the
    chromosome lengths of the genome up (UCSC): here the caller passes them.  `bed`: a path or the
    bytes of the file."""
    _lib.ensure_init()
    data = bed if isinstance(bed, (bytes, bytearray, memoryview)) else open(bed, "rb").read()
    if bytes(data[:2]) == b"\x1f\x8b":
        data = gzip.decompress(bytes(data))
    text = np.frombuffer(bytes(data), dtype=np.uint8)
    arr = (C.c_char_p * len(seqlevels))(*[s.encode("ascii") for s in seqlevels])
    h, n_out = C.c_int(0), C.c_int64(0)
    _lib.check(_lib.lib.rcp_bed_decode(text.ctypes.data_as(C.c_void_p), text.shape[0], len(seqlevels), arr,
                                       _lib.MEM_HOST, C.byref(h), C.byref(n_out)))
    return DecodedGRanges(h.value, n_out.value, seqlevels, seqlengths)


def readRangesFile(input, format, sa="keep", sq=0.75, seqlevels=None, seqlengths=None):
    """readRanges (ranges.R:102-109) for files; "bigwig" has no reads (NULL)."""
    if format == "bam":
        return readBam(input, sa, sq)
    if format == "bed":
        return readBed(input, seqlevels, seqlengths)
    if format == "bigwig":
        return None
    raise ValueError("format must be one of bam, bed, bigwig")
