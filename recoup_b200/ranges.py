"""Host-side range containers and region geometry.

Mirrors the part of /root/reference/R/ranges.R that is on the coverage path:
`getRegionalRanges` (ranges.R:67-91) and `getFlankingRanges` (ranges.R:93-100), expressed over
plain integer arrays.  The GenomicRanges verbs they call (`promoters`, `resize`, `flank`) are
third-party code that is not vendored with the reference; their documented semantics are
restated in `_promoters` / `_resize` / `_flank`.

`GRanges` / `GRangesList` are the minimum stand-ins for the Bioconductor classes the reference
API takes: integer arrays, no metadata columns.
"""
import numpy as np

_STRAND_CODE = {"+": 1, "-": -1, "*": 0}
_STRAND_CHAR = {1: "+", -1: "-", 0: "*"}


def strand_to_code(strand, n=None):
    """'+','-','*' (scalar or sequence) or +1/-1/0 integers -> int8 array."""
    if strand is None:
        return np.zeros(0 if n is None else n, dtype=np.int8)
    if isinstance(strand, str):
        return np.full(1 if n is None else n, _STRAND_CODE[strand], dtype=np.int8)
    arr = np.asarray(strand)
    if arr.dtype == np.int8:
        return arr          # already coded; the library only looks at the sign
    if arr.dtype.kind in "US":
        out = np.zeros(arr.shape, dtype=np.int8)
        out[arr == "+"] = 1
        out[arr == "-"] = -1
        bad = ~np.isin(arr, ["+", "-", "*"])
        if bad.any():
            raise ValueError("strand values must be '+', '-' or '*'")
        return out
    return np.sign(arr).astype(np.int8)


class Rle:
    """Run-length encoded vector (S4Vectors::Rle) -- how a GRanges holds its seqnames.  `values[r]`
    repeated `lengths[r]` times."""

    def __init__(self, values, lengths):
        self.values = np.asarray(values)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        if self.values.shape[0] != self.lengths.shape[0]:
            raise ValueError("values and lengths must have the same length")
        if self.lengths.shape[0] and int(self.lengths.min()) < 0:
            raise ValueError("negative run length")
        self._n = int(self.lengths.astype(np.int64).sum())

    def __len__(self):
        return self._n

    @property
    def nrun(self):
        return self.lengths.shape[0]

    def decode(self):
        return np.repeat(self.values, self.lengths)

    @staticmethod
    def encode(x):
        x = np.asarray(x)
        if x.shape[0] == 0:
            return Rle(x, np.zeros(0, dtype=np.int32))
        cut = np.flatnonzero(x[1:] != x[:-1]) + 1
        first = np.concatenate(([0], cut))
        return Rle(x[first], np.diff(np.concatenate((first, [x.shape[0]]))))


class GRanges:
    """seqnames are stored as int32 ids into `seqlevels` (dense, or as an `Rle` when given as
    one: `seqnames_rle`; the dense `seqnames` is then decoded on first use); `seqlengths[i]` is
    the length of seqlevels[i] (required for reads: the coverage vector of a chromosome has that
    length, coverage.R:201)."""

    def __init__(self, seqnames, start, end=None, width=None, strand=None, seqlevels=None,
                 seqlengths=None, names=None):
        start = np.ascontiguousarray(start, dtype=np.int32)
        n = start.shape[0]
        # ONE width for every range (fixed-length reads, or a GRanges after resize()): kept as a
        # number, as a constant-width IRanges is known to be; `end` is materialised only on demand
        # and the reads upload sends start + width (rcp_reads_load_width)
        self.fixed_width = None
        if end is None:
            if width is None:
                raise ValueError("give end or width")
            if np.ndim(width) == 0:
                if int(width) < 1:
                    raise ValueError("width must be >= 1")
                self.fixed_width = int(width)
            else:
                end = start.astype(np.int64) + np.asarray(width, dtype=np.int64) - 1
        if end is not None:
            end = np.ascontiguousarray(end, dtype=np.int32)
        run_len = None
        if isinstance(seqnames, Rle):
            if len(seqnames) != n:
                raise ValueError("seqnames, start, end and strand must have the same length")
            run_len = seqnames.lengths
            seqnames = seqnames.values
        sn = np.asarray(seqnames)
        if sn.dtype.kind in "US":
            if seqlevels is None:
                seqlevels = list(dict.fromkeys(sn.tolist()))
            lut = {s: i for i, s in enumerate(seqlevels)}
            try:
                ids = np.fromiter((lut[s] for s in sn.tolist()), dtype=np.int32, count=sn.shape[0])
            except KeyError as e:
                raise ValueError("seqname %s is not in seqlevels" % e)
        else:
            ids = np.ascontiguousarray(sn, dtype=np.int32)
            if seqlevels is None:
                seqlevels = ["chr%d" % (i + 1) for i in range(int(ids.max()) + 1 if n else 0)]
        if run_len is None and ids.shape[0] == 1 and n != 1:
            ids = np.full(n, ids[0], dtype=np.int32)
        st = strand_to_code(strand, n) if strand is not None else np.zeros(n, dtype=np.int8)
        if st.shape[0] == 1 and n != 1:
            st = np.full(n, st[0], dtype=np.int8)
        if not ((run_len is not None or ids.shape[0] == n) and st.shape[0] == n
                and (end is None or end.shape[0] == n)):
            raise ValueError("seqnames, start, end and strand must have the same length")
        if run_len is None:
            self.seqnames_rle = None
            self._seqnames = ids
        else:
            self.seqnames_rle = Rle(ids, run_len)
            self._seqnames = None
        self.start = start
        self._end = end
        self.strand = np.ascontiguousarray(st, dtype=np.int8)
        self.seqlevels = list(seqlevels)
        self.seqlengths = (None if seqlengths is None
                           else np.ascontiguousarray(seqlengths, dtype=np.int64))
        self.names = None if names is None else [str(x) for x in names]
        self._device = {}    # frag_len -> device reads handle (recoup_b200.coverage)

    def __len__(self):
        return self.start.shape[0]

    @property
    def seqnames(self):
        if self._seqnames is None:
            self._seqnames = np.ascontiguousarray(self.seqnames_rle.decode(), dtype=np.int32)
        return self._seqnames

    @property
    def end(self):
        if self._end is None:
            e = self.start.astype(np.int64) + (self.fixed_width - 1)
            if e.size and int(e.max()) > np.iinfo(np.int32).max:
                raise OverflowError("start + width - 1 leaves the int32 range")
            self._end = np.ascontiguousarray(e, dtype=np.int32)
        return self._end

    @property
    def width(self):
        if self.fixed_width is not None:
            return np.full(len(self), self.fixed_width, dtype=np.int64)
        return self.end.astype(np.int64) - self.start.astype(np.int64) + 1

    def subset(self, idx):
        idx = np.asarray(idx)
        names = None
        if self.names is not None:
            sel = np.flatnonzero(idx) if idx.dtype == bool else idx
            names = [self.names[int(i)] for i in sel]
        return GRanges(self.seqnames[idx], self.start[idx], self.end[idx], strand=self.strand[idx],
                       seqlevels=self.seqlevels, seqlengths=self.seqlengths, names=names)

    def with_ranges(self, start, end):
        return GRanges(self.seqnames, start, end, strand=self.strand, seqlevels=self.seqlevels,
                       seqlengths=self.seqlengths, names=self.names)

    def strand_chars(self):
        return [_STRAND_CHAR[int(s)] for s in self.strand]


class GRangesList:
    """`ptr[g]:ptr[g+1]` are the ranges of element g inside `unlisted`."""

    def __init__(self, unlisted, ptr, names=None):
        self.unlisted = unlisted
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        if self.ptr.shape[0] < 1 or self.ptr[0] != 0 or self.ptr[-1] != len(unlisted):
            raise ValueError("ptr must run from 0 to length(unlisted)")
        self.names = None if names is None else [str(x) for x in names]

    def __len__(self):
        return self.ptr.shape[0] - 1


# ------------------------------------------------------------------------------------------------
# GenomicRanges verbs (documented semantics)
# ------------------------------------------------------------------------------------------------
def _promoters(start, end, strand, upstream, downstream):
    minus = strand < 0
    s = np.where(minus, end - downstream + 1, start - upstream)
    e = np.where(minus, end + upstream, start + downstream - 1)
    return s, e


def _resize(start, end, strand, width, fix="start"):
    minus = strand < 0
    keep_start = ~minus if fix == "start" else minus
    s = np.where(keep_start, start, end - width + 1)
    e = np.where(keep_start, start + width - 1, end)
    return s, e


def _flank_downstream(start, end, strand, width):
    minus = strand < 0
    s = np.where(minus, start - width, end + 1)
    e = np.where(minus, start - 1, end + width)
    return s, e


def _as64(gr):
    return gr.start.astype(np.int64), gr.end.astype(np.int64), gr.strand.astype(np.int64)


def _checked(gr, s, e):
    lim = np.iinfo(np.int32)
    if s.size and (s.min() < lim.min or e.max() > lim.max):
        raise OverflowError("flanked coordinates leave the int32 range")
    return gr.with_ranges(s.astype(np.int32), e.astype(np.int32))


def getRegionalRanges(ranges, region, flank):
    """ranges.R:67-91."""
    start, end, strand = _as64(ranges)
    f1, f2 = int(flank[0]), int(flank[1])
    w = end - start + 1
    if region == "custom":
        region = "tss" if bool(np.all(w == 1)) else "genebody"          # ranges.R:81-89
    if region == "genebody":                                            # ranges.R:69-73
        s, e = _promoters(start, end, strand, f1, 0)
        s, e = _resize(s, e, strand, w + f1 + f2, "start")
    elif region == "tss":                                               # ranges.R:74-76
        s, e = _promoters(start, end, strand, f1, f2)
    elif region == "tes":                                               # ranges.R:77-80
        s, e = _resize(start, end, strand, 1, "end")
        s, e = _promoters(s, e, strand, f1, f2)
    else:
        raise ValueError("region must be one of tss, tes, genebody, custom")
    return _checked(ranges, s, e)


def getFlankingRanges(ranges, flank, dir="upstream"):
    """ranges.R:93-100."""
    start, end, strand = _as64(ranges)
    if dir == "upstream":
        s, e = _promoters(start, end, strand, int(flank), 0)
    elif dir == "downstream":
        s, e = _flank_downstream(start, end, strand, int(flank))
    else:
        return None                                                     # R falls through: NULL
    return _checked(ranges, s, e)
