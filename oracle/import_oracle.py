"""TEST INFRASTRUCTURE -- CPU restatement of the reference's file readers (readBam / readBed,
/root/reference/R/ranges.R:111-146).  Only tests/ may import this; the product never does.

**Parity unpinned.**  The arithmetic lives in third-party packages that are not vendored with the
reference and cannot run here (no R): GenomicAlignments (`readGAlignments`, `granges`, `grglist`),
Rsamtools (`scanBam`), rtracklayer (`import.bed`), GenomicRanges (`trim`).  What is restated is
their documented behaviour:

* `readGAlignments(file)` with no `param`: `ScanBamParam(flag = scanBamFlag(isUnmappedQuery =
  FALSE))` -- unmapped records are dropped, everything else (secondary, duplicate, QC-fail) is
  kept (GenomicAlignments man page `readGAlignments`, section "Arguments: param").
* `as(galn, "GRanges")` = `granges(galn)`: one range per alignment, start = POS, width =
  `cigarWidthAlongReferenceSpace(cigar)` = the summed lengths of M, D, N, =, X (man page
  `cigar-utils`); strand '-' when flag 0x10 is set, else '+'.
* `grglist(galn)` (defaults `drop.D.ranges = FALSE`): `cigarRangesAlongReferenceSpace(cigar, pos =
  start, ops = c("M", "=", "X", "D"), drop.empty.ranges = TRUE, reduce.ranges = TRUE)` -- the
  alignment cut at its N operations; I, S, H, P take no reference space, so the M / = / X / D runs
  either side of them are adjacent and `reduce.ranges` merges them; empty ranges are dropped.
  `unlist()` concatenates the per-alignment ranges in alignment order.
* `trim(gr)`: out-of-bound ranges of a sequence of known, non-circular length L are cut to
  [max(start, 1), min(end, L)].
* `import.bed(file, trackLine = FALSE)`: tab- (or blank-) separated chrom, chromStart, chromEnd
  [, name, score, strand]; start = chromStart + 1, end = chromEnd; strand '.' or missing -> '*';
  comment ('#'), `track` and `browser` lines are not data (UCSC BED specification; rtracklayer man
  page `BEDFile-class`).

The BAM record layout follows the SAM/BAM specification (SAMv1 section 4.2): little-endian
block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID, next_pos, tlen,
read_name, cigar (op_len << 4 | op, ops "MIDNSHP=X"), seq, qual, tags.
"""
import gzip
import struct

import numpy as np

REF_OPS = (0, 2, 3, 7, 8)      # M D N = X take reference space
BLOCK_OPS = (0, 2, 7, 8)       # M D = X make up the ranges of grglist()


def bgzf_inflate(data):
    return gzip.decompress(bytes(data))


def bam_header(raw):
    assert raw[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", raw, 4)
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, p)
    p += 4
    names, lens = [], []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, p)
        names.append(raw[p + 4:p + 4 + l_name - 1].decode("ascii"))
        lens.append(struct.unpack_from("<i", raw, p + 4 + l_name)[0])
        p += 8 + l_name
    return names, np.asarray(lens, dtype=np.int64), p


def bam_record_offsets(rec):
    """Walk of the record chain: offsets of the records and of the end."""
    off, p = [], 0
    while p < len(rec):
        bs, = struct.unpack_from("<i", rec, p)
        off.append(p)
        p += 4 + bs
    off.append(p)
    return np.asarray(off, dtype=np.int64)


def _trim(s, e, L):
    s2, e2 = max(s, 1), min(e, L)
    if e2 < s2 - 1:
        if s > L:
            s2, e2 = L + 1, L
        else:
            e2 = s2 - 1
    return s2, e2


def bam_decode(rec, ref_len, split=False):
    """Alignment records (inflated, header stripped) -> chrom, start, end, strand in file order."""
    chrom, start, end, strand = [], [], [], []
    p = 0
    while p < len(rec):
        bs, ref, pos, l_name, _mapq, _bin, n_cig, flag, _l_seq = struct.unpack_from("<iiiBBHHHi", rec, p)
        q = p + 36 + l_name
        ops = struct.unpack_from("<%dI" % n_cig, rec, q)
        p += 4 + bs
        if (flag & 4) or ref < 0 or pos < 0:
            continue
        if n_cig == 0:
            raise ValueError("a mapped record has no CIGAR")
        if any((c & 15) > 8 for c in ops):
            raise ValueError("CIGAR operation code above 8")
        st = -1 if flag & 16 else 1
        L = int(ref_len[ref])
        if not split:
            w = sum(c >> 4 for c in ops if (c & 15) in REF_OPS)
            s, e = _trim(pos + 1, pos + w, L)
            chrom.append(ref); start.append(s); end.append(e); strand.append(st)
            continue
        cur, blk = pos + 1, 0
        for c in ops:
            op, ln = c & 15, c >> 4
            if op in BLOCK_OPS:
                blk += ln
            elif op == 3:
                if blk > 0:
                    s, e = _trim(cur, cur + blk - 1, L)
                    chrom.append(ref); start.append(s); end.append(e); strand.append(st)
                cur += blk + ln
                blk = 0
        if blk > 0:
            s, e = _trim(cur, cur + blk - 1, L)
            chrom.append(ref); start.append(s); end.append(e); strand.append(st)
    return (np.asarray(chrom, dtype=np.int32), np.asarray(start, dtype=np.int32),
            np.asarray(end, dtype=np.int32), np.asarray(strand, dtype=np.int8))


def bed_decode(text, seqlevels):
    """BED text (bytes) -> chrom id, start, end, strand in file order."""
    lut = {s: i for i, s in enumerate(seqlevels)}
    chrom, start, end, strand = [], [], [], []
    for line in bytes(text).split(b"\n"):
        if line.endswith(b"\r"):
            line = line[:-1]
        f = line.replace(b"\t", b" ").split()
        if not f or f[0].startswith(b"#") or f[0] in (b"track", b"browser"):
            continue
        if len(f) < 3:
            raise ValueError("a data line has fewer than three fields")
        name = f[0].decode("ascii")
        if name not in lut:
            raise ValueError("unknown chromosome %s" % name)
        if not (f[1].isdigit() and f[2].isdigit()):
            raise ValueError("bad coordinate")
        st = 0
        if len(f) >= 6:
            if f[5] not in (b"+", b"-", b".", b"*"):
                raise ValueError("bad strand")
            st = 1 if f[5] == b"+" else (-1 if f[5] == b"-" else 0)
        chrom.append(lut[name]); start.append(int(f[1]) + 1); end.append(int(f[2])); strand.append(st)
    return (np.asarray(chrom, dtype=np.int32), np.asarray(start, dtype=np.int32),
            np.asarray(end, dtype=np.int32), np.asarray(strand, dtype=np.int8))
