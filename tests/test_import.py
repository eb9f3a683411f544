"""Read import: BAM records and BED text decoded on the device against the oracle's
restatement of readBam / readBed (oracle/import_oracle.py), on synthetic files the tests write
themselves (every CIGAR operation, unmapped / reverse / empty alignments, spliced reads) and on a
real file cut out of the reference's own sample (tests/golden/WT_H4K20me1_5kr.bam).  Integer
results: BIT-EXACT."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import import_oracle as IO
from oracle import recoup_oracle as O
from tests.bam_writer import bam_file, bam_record

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_BAM = os.path.join(HERE, "golden", "WT_H4K20me1_5kr.bam")
REFS = [("chr1", 50000), ("chr2", 30000), ("chrM", 900)]
CIGARS = ["36M", "10M500N26M", "5S20M3I11M", "12M2D8M1000N4M4D10M", "3H10=1X25=", "20M1P16M",
          "8M30N8M40N8M50N12M", "36S", "6I", "10M5N2I5N10M", "4D", "15M2000N", "7N20M"]


def _synthetic_bam(rng, n, clean=False):
    """clean: no alignment of reference width 0 and none that trim() cuts to width 0 (the reads
    loader refuses zero-width reads, like a GRanges holding them passed as arrays)"""
    recs, truth = [], []
    for i in range(n):
        ref = int(rng.integers(0, 2 if clean else len(REFS)))
        cigar = CIGARS[int(rng.integers(0, len(CIGARS)))]
        while clean and cigar in ("36S", "6I"):
            cigar = CIGARS[int(rng.integers(0, len(CIGARS)))]
        pos0 = int(rng.integers(0, REFS[ref][1] - (2200 if clean else 10)))
        flag = int(rng.choice([0, 16, 4, 20, 256, 1024 + 16, 99, 147]))
        if i % 97 == 0:
            ref, pos0, flag = -1, -1, 4                 # unplaced
        recs.append(bam_record(ref, pos0, flag, cigar, name=b"read%d" % i, l_seq=int(rng.integers(0, 40)),
                               tags=b"NMC\x01" if i % 3 == 0 else b""))
        truth.append((ref, pos0, flag, cigar))
    return bam_file(REFS, recs), truth


def _assert_same(got, want):
    for g, w, what in zip(got, want, ("chrom", "start", "end", "strand")):
        assert g.shape == w.shape, "%s: %s != %s" % (what, g.shape, w.shape)
        assert np.array_equal(g, w), what


# ------------------------------------------------------------------------------- CPU ------------
def test_oracle_decodes_known_alignments():
    """hand-computed answers (GenomicAlignments man page `cigar-utils` semantics)"""
    recs = [bam_record(0, 99, 0, "10M500N26M"), bam_record(1, 0, 16, "5S20M3I11M"),
            bam_record(0, 9, 4, "36M"), bam_record(2, 880, 0, "36M"), bam_record(0, 49, 0, "12M2D8M1000N4M4D10M"),
            bam_record(0, 10, 16, "6I"), bam_record(0, 10, 0, "7N20M")]
    raw, bgzf = bam_file(REFS, recs)
    assert IO.bgzf_inflate(bgzf) == raw
    names, lens, first = IO.bam_header(raw)
    assert names == ["chr1", "chr2", "chrM"] and lens.tolist() == [50000, 30000, 900]
    rec = raw[first:]
    assert IO.bam_record_offsets(rec).shape[0] == len(recs) + 1
    c, s, e, st = IO.bam_decode(rec, lens)                     # keep: one range per mapped alignment
    assert c.tolist() == [0, 1, 2, 0, 0, 0]
    assert s.tolist() == [100, 1, 881, 50, 11, 11]
    assert e.tolist() == [635, 31, 900, 1089, 10, 37]          # chrM read trimmed at 900; "6I" is empty
    assert st.tolist() == [1, -1, 1, 1, -1, 1]
    c, s, e, st = IO.bam_decode(rec, lens, split=True)         # split: cut at N, empty ranges dropped
    assert list(zip(c.tolist(), s.tolist(), e.tolist())) == [
        (0, 100, 109), (0, 610, 635), (1, 1, 31), (2, 881, 900), (0, 50, 71), (0, 1072, 1089), (0, 18, 37)]


# The example alignments of the SAM specification (SAMv1.pdf, section 1.1 "An example"): reference
# `ref` of 45 bases; POS, FLAG and CIGAR as printed there, ends and blocks from its picture.
SAM_SPEC_EXAMPLE = [          # (name, flag, pos, cigar, keep range, split ranges)
    (b"r001", 99, 7, "8M2I4M1D3M", (7, 22), [(7, 22)]),
    (b"r002", 0, 9, "3S6M1P1I4M", (9, 18), [(9, 18)]),
    (b"r003", 0, 9, "5S6M", (9, 14), [(9, 14)]),
    (b"r004", 0, 16, "6M14N5M", (16, 40), [(16, 21), (36, 40)]),
    (b"r003", 2064, 29, "6H5M", (29, 33), [(29, 33)]),
    (b"r001", 147, 37, "9M", (37, 45), [(37, 45)]),
]


def _sam_spec_bam():
    return bam_file([("ref", 45)], [bam_record(0, pos - 1, flag, cig, name=nm) for nm, flag, pos, cig, _, _ in SAM_SPEC_EXAMPLE])


def test_oracle_on_the_sam_specification_example():
    raw, _ = _sam_spec_bam()
    names, lens, first = IO.bam_header(raw)
    c, s, e, st = IO.bam_decode(raw[first:], lens)
    assert list(zip(s.tolist(), e.tolist())) == [k for _, _, _, _, k, _ in SAM_SPEC_EXAMPLE]
    assert st.tolist() == [1, 1, 1, 1, -1, -1]                  # FLAG 0x10: r003 (supplementary) and r001/2
    c, s, e, st = IO.bam_decode(raw[first:], lens, split=True)
    assert list(zip(s.tolist(), e.tolist())) == [r for *_, sp in SAM_SPEC_EXAMPLE for r in sp]
    assert st.tolist() == [1, 1, 1, 1, 1, -1, -1]


def test_oracle_bed_known_lines():
    text = (b"track name=x\n# c\nchr1\t0\t10\tn\t0\t+\nchr2 5 9\r\n\nbrowser position\n"
            b"chrM\t3\t4\t.\t1\t.\nchr1\t7\t20\tq\t9\t-")
    c, s, e, st = IO.bed_decode(text, ["chr1", "chr2", "chrM"])
    assert c.tolist() == [0, 1, 2, 0] and s.tolist() == [1, 6, 4, 8] and e.tolist() == [10, 9, 4, 20]
    assert st.tolist() == [1, 0, 0, -1]


def test_golden_bam_through_the_oracle():
    raw = IO.bgzf_inflate(open(GOLDEN_BAM, "rb").read())
    names, lens, first = IO.bam_header(raw)
    assert names == ["chr12"] and lens.tolist() == [121257530]
    c, s, e, st = IO.bam_decode(raw[first:], lens)
    assert c.shape[0] == 5000 and (e - s + 1 == 180).all() and set(st.tolist()) == {1, -1}
    assert int(s[0]) == 3000989          # first record of the reference's file, POS + 1


def test_bam_index_is_host_code_and_matches_the_walk():
    from recoup_b200 import _lib
    rng = np.random.default_rng(5)
    (raw, _), _ = _synthetic_bam(rng, 500)
    _, _, first = IO.bam_header(raw)
    rec = np.frombuffer(raw, dtype=np.uint8, offset=first)
    n = C.c_int64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib.rcp_bam_index(vp(rec), rec.shape[0], C.byref(n), None, 0))
    assert n.value == 500
    off = np.zeros(501, dtype=np.int64)
    _lib.check(_lib.lib.rcp_bam_index(vp(rec), rec.shape[0], C.byref(n), vp(off), 501))
    assert np.array_equal(off, IO.bam_record_offsets(raw[first:]))
    with pytest.raises(_lib.RecoupError):
        _lib.check(_lib.lib.rcp_bam_index(vp(rec), rec.shape[0], C.byref(n), vp(off), 500))
    with pytest.raises(_lib.RecoupError):        # a chain that runs past the end
        _lib.check(_lib.lib.rcp_bam_index(vp(rec), rec.shape[0] - 3, C.byref(n), None, 0))


def test_bgzf_inflate_is_host_code_and_matches_gzip():
    """rcp_bgzf_inflate (zlib, block-parallel) against python's gzip on the reference's BAM, on a
    multi-block file of the test writer and on damaged / foreign data."""
    import gzip
    import zlib
    from recoup_b200 import _lib
    from recoup_b200.readers import bgzfInflate
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(8)
    (raw, bgzf), _ = _synthetic_bam(rng, 6000)                  # ~ 600 KB: ten blocks and the EOF marker
    for data, want in ((open(GOLDEN_BAM, "rb").read(), None), (bgzf, raw)):
        want = gzip.decompress(data) if want is None else want
        buf = np.frombuffer(data, dtype=np.uint8)
        total, nblk = C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.lib.rcp_bgzf_size(vp(buf), buf.shape[0], C.byref(total), C.byref(nblk)))
        assert total.value == len(want) and nblk.value >= 2
        for threads in (1, 3, 0):
            assert bgzfInflate(data, threads).tobytes() == want
    buf = np.frombuffer(bgzf, dtype=np.uint8)
    out = np.empty(len(raw), dtype=np.uint8)
    assert _lib.lib.rcp_bgzf_inflate(vp(buf), buf.shape[0], vp(out), len(raw) - 1, 1) == _lib.RCP_ERR_ARG
    bad = buf.copy()
    bad[200] ^= 0x55                                            # inside the first block's deflate stream
    assert _lib.lib.rcp_bgzf_inflate(vp(bad), bad.shape[0], vp(out), out.shape[0], 2) == _lib.RCP_ERR_DATA
    assert _lib.lib.rcp_bgzf_inflate(vp(buf), buf.shape[0] - 5, vp(out), out.shape[0], 1) == _lib.RCP_ERR_DATA
    plain = gzip.compress(raw)                                  # gzip, but not BGZF: python's gzip takes over
    pbuf = np.frombuffer(plain, dtype=np.uint8)
    total = C.c_int64(0)
    assert _lib.lib.rcp_bgzf_size(vp(pbuf), pbuf.shape[0], C.byref(total), None) == _lib.RCP_ERR_DATA
    assert bgzfInflate(plain).tobytes() == raw
    assert bgzfInflate(b"").shape[0] == 0
    assert zlib.crc32(bgzfInflate(bgzf).tobytes()) == zlib.crc32(raw)


# ------------------------------------------------------------------------------- GPU ------------
@pytest.fixture(scope="module")
def rb():
    import recoup_b200 as rb
    rb.init(0)
    rb.set_coverage_path("auto")
    return rb


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 255, 4000])
@pytest.mark.parametrize("sa", ["keep", "split"])
def test_bam_decode_matches_the_oracle(rb, n, sa):
    rng = np.random.default_rng(100 + n)
    (raw, bgzf), _ = _synthetic_bam(rng, n)
    names, lens, first = IO.bam_header(raw)
    want = IO.bam_decode(raw[first:], lens, split=(sa == "split"))
    got = rb.readBam(bgzf, sa=sa)
    assert got.seqlevels == names and len(got) == want[0].shape[0]
    _assert_same((got.seqnames, got.start, got.end, got.strand), want)


@pytest.mark.gpu
def test_bam_decode_on_the_sam_specification_example(rb):
    _, bgzf = _sam_spec_bam()
    got = rb.readBam(bgzf)
    assert got.seqlevels == ["ref"] and got.seqlengths.tolist() == [45]
    assert list(zip(got.start.tolist(), got.end.tolist())) == [k for _, _, _, _, k, _ in SAM_SPEC_EXAMPLE]
    assert got.strand.tolist() == [1, 1, 1, 1, -1, -1]
    got = rb.readBam(bgzf, sa="split")
    assert list(zip(got.start.tolist(), got.end.tolist())) == [r for *_, sp in SAM_SPEC_EXAMPLE for r in sp]
    # coverage of the example straight from the decoded reads: the spliced r004 covers its gap under
    # "keep" (as(galn, "GRanges") spans it) and not under "split"
    from tests.helpers import assert_coverage_equal
    od = dict(chrom=np.zeros(1, dtype=np.int64), start=np.asarray([1]), end=np.asarray([45]), strand=np.zeros(1, dtype=np.int64))
    mask = rb.GRanges(np.zeros(1, dtype=np.int32), [1], [45], strand=np.zeros(1, dtype=np.int8), seqlevels=["ref"])
    for sa in ("keep", "split"):
        reads = rb.readBam(bgzf, sa=sa)
        cov = rb.calcCoverage(reads, mask).to_list()[0]
        want = np.zeros(45, dtype=np.int64)
        for *_, keep, split in SAM_SPEC_EXAMPLE:
            for a, b in ([keep] if sa == "keep" else split):
                want[a - 1:b] += 1
        assert np.array_equal(cov.astype(np.int64), want)
        assert (cov[25] == 1) == (sa == "keep")                 # position 26: inside r004's N gap only


@pytest.mark.gpu
def test_golden_bam_decode_and_coverage_without_a_host_round_trip(rb, fixture_data):
    from tests.helpers import assert_coverage_equal, fixture_genes
    data = open(GOLDEN_BAM, "rb").read()
    raw = IO.bgzf_inflate(data)
    names, lens, first = IO.bam_header(raw)
    want = IO.bam_decode(raw[first:], lens)
    reads = rb.readBam(data)
    assert reads._host is None                       # nothing fetched yet
    o_genes, g_genes = fixture_genes(fixture_data)
    inp = [dict(id="s", name="s", ranges=reads)]
    rb.coverageRef(inp, g_genes, "tss", (2000, 2000))
    assert reads._host is None                       # rcp_reads_load_decoded: the reads never left HBM
    o_reads = O.Reads(want[0], want[1].astype(np.int64), want[2].astype(np.int64), want[3], lens)
    assert_coverage_equal(inp[0]["coverage"].to_list(), O.coverage_ref(o_reads, o_genes, "tss", (2000, 2000)))
    _assert_same((reads.seqnames, reads.start, reads.end, reads.strand), want)


@pytest.mark.gpu
def test_read_bam_remove_and_preprocess_ranges_from_files(rb, tmp_path):
    rng = np.random.default_rng(77)
    files = []
    for k in range(2):
        (raw, bgzf), _ = _synthetic_bam(rng, 3000 + 500 * k, clean=True)
        p = tmp_path / ("s%d.bam" % k)
        p.write_bytes(bgzf)
        files.append((str(p), raw))
    # spliceAction "remove": reads wider than the 0.75 quantile of the widths go (ranges.R:125-133)
    path, raw = files[0]
    names, lens, first = IO.bam_header(raw)
    c, s, e, st = IO.bam_decode(raw[first:], lens)
    keep = O.splice_remove(s, e, 0.75)[0]
    got = rb.readBam(path, sa="remove", sq=0.75)
    assert len(got) == int(keep.sum())
    _assert_same((got.seqnames, got.start, got.end, got.strand), (c[keep], s[keep], e[keep], st[keep]))
    # preprocessRanges straight from the files, spliceAction "split" + down-sampling
    inp = [dict(id="a", name="a", file=files[0][0], format="bam"), dict(id="b", name="b", file=files[1][0], format="bam")]
    rb.preprocessRanges(inp, dict(normalize="downsample", spliceAction="split", seed=42))
    sizes = []
    for x, (_, raw) in zip(inp, files):
        names, lens, first = IO.bam_header(raw)
        sizes.append(IO.bam_decode(raw[first:], lens, split=True)[0].shape[0])
    assert len(inp[0]["ranges"]) == len(inp[1]["ranges"]) == min(sizes)
    idx = O.downsample_indices(sizes, "downsample", seed=42)
    for x, (_, raw), ix in zip(inp, files, idx):
        names, lens, first = IO.bam_header(raw)
        c, s, e, st = IO.bam_decode(raw[first:], lens, split=True)
        g = x["ranges"]
        # the selection is applied on the device to the decoded reads (rcp_reads_load_decoded_select)
        from tests.helpers import assert_coverage_equal
        o_mask, g_mask = both_regions_([0, 1, 2, 0], [1, 100, 1, 20000], [50000, 25000, 900, 29999], [1, -1, 0, 1])
        o_reads = O.Reads(c[ix - 1], s[ix - 1].astype(np.int64), e[ix - 1].astype(np.int64), st[ix - 1], lens)
        assert g._host is None and g.parent._host is None
        assert_coverage_equal(rb.calcCoverage(g, g_mask, ignore_strand=False).to_list(),
                              O.calc_coverage(o_reads, o_mask, None, False))
        assert g._host is None and g.parent._host is None          # still nothing fetched
        _assert_same((g.seqnames, g.start, g.end, g.strand), (c[ix - 1], s[ix - 1], e[ix - 1], st[ix - 1]))


@pytest.mark.gpu
def test_zero_width_alignments_are_decoded_but_refused_by_the_reads_loader(rb):
    from recoup_b200 import _lib
    raw, bgzf = bam_file(REFS, [bam_record(0, 5, 0, "10M"), bam_record(0, 50, 16, "6I")])
    reads = rb.readBam(bgzf)                           # as(galn, "GRanges") keeps the empty alignment
    assert reads.start.tolist() == [6, 51] and reads.end.tolist() == [15, 50]
    o_mask, g_mask = both_regions_([0], [1], [100], [1])
    with pytest.raises(_lib.RecoupError) as ei:        # the same answer a GRanges with such a read gets
        rb.calcCoverage(reads, g_mask)
    assert ei.value.code == _lib.RCP_ERR_DATA
    assert len(rb.readBam(bgzf, sa="split")) == 1      # grglist drops empty ranges


def both_regions_(chrom, start, end, strand):
    """mask on the BAM's own seqlevels (calcCoverage matches chromosomes by NAME, coverage.R:182,189)"""
    import recoup_b200 as rb
    od = dict(chrom=np.asarray(chrom, dtype=np.int64), start=np.asarray(start, dtype=np.int64),
              end=np.asarray(end, dtype=np.int64), strand=np.asarray(strand, dtype=np.int64))
    return od, rb.GRanges(np.asarray(chrom, dtype=np.int32), start, end, strand=np.asarray(strand, dtype=np.int8),
                          seqlevels=[r[0] for r in REFS])


@pytest.mark.gpu
def test_bam_decode_errors(rb):
    from recoup_b200 import _lib
    raw, bgzf = bam_file(REFS, [bam_record(0, 5, 0, "10M"), bam_record(3, 5, 0, "10M")])
    with pytest.raises(_lib.RecoupError) as ei:
        rb.readBam(bgzf)
    assert ei.value.code == _lib.RCP_ERR_DATA                      # refID outside the header
    raw, bgzf = bam_file(REFS, [bam_record(0, 5, 0, "*")])
    with pytest.raises(_lib.RecoupError):
        rb.readBam(bgzf)                                           # mapped, no CIGAR
    # offsets that do not follow the chain
    raw, _ = bam_file(REFS, [bam_record(0, 5, 0, "10M"), bam_record(1, 7, 16, "4M")])
    _, lens, first = IO.bam_header(raw)
    rec = np.frombuffer(raw, dtype=np.uint8, offset=first)
    off = IO.bam_record_offsets(raw[first:]).copy()
    off[1] += 1
    h, n = C.c_int(0), C.c_int64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib.lib.rcp_bam_decode(vp(rec), rec.shape[0], vp(off), 2, 3, lens.ctypes.data_as(C.POINTER(C.c_int64)), 0,
                                 _lib.MEM_HOST, C.byref(h), C.byref(n))
    assert rc == _lib.RCP_ERR_DATA
    assert _lib.lib.rcp_decoded_free(12345) == _lib.RCP_ERR_HANDLE


def _bed_text(rng, n, trailing_newline=True):
    names = ["chr1", "chr10", "chr2", "chrX", "chr1_gl000191_random"]
    lines = [b"track name=\"reads\" description=\"x\"", b"# comment", b""]
    for i in range(n):
        nm = names[int(rng.integers(0, len(names)))].encode()
        a = int(rng.integers(0, 2_000_000_000 if i % 50 == 0 else 100000))
        b = a + int(rng.integers(0, 300))
        kind = i % 7
        if kind == 0:
            ln = b"%s\t%d\t%d" % (nm, a, b)
        elif kind == 1:
            ln = b"%s %d  %d name%d 0 -" % (nm, a, b, i)
        elif kind == 2:
            ln = b"%s\t%d\t%d\tname\t960\t.\r" % (nm, a, b)
        elif kind == 3:
            ln = b"%s\t%d\t%d\tn\t1" % (nm, a, b)
        elif kind == 4:
            ln = b"%s\t%d\t%d\tn\t1\t*\t%d\t%d\t255,0,0" % (nm, a, b, a, b)
        else:
            ln = b"%s\t%d\t%d\tn\t1\t%s" % (nm, a, b, b"+" if i % 2 else b"-")
        lines.append(ln)
        if i % 401 == 0:
            lines.append(b"browser position chr1:1-100")
    text = b"\n".join(lines)
    return (text + b"\n" if trailing_newline else text), names


@pytest.mark.gpu
@pytest.mark.parametrize("n,nl", [(0, True), (1, False), (1, True), (4097, True), (30000, False)])
def test_bed_decode_matches_the_oracle(rb, n, nl):
    rng = np.random.default_rng(300 + n)
    text, names = _bed_text(rng, n, nl)
    levels = list(reversed(names)) + ["chrUn"]              # ids follow the caller's seqlevels
    want = IO.bed_decode(text, levels)
    got = rb.readBed(text, levels, seqlengths=np.full(len(levels), 2_100_000_000, dtype=np.int64))
    assert len(got) == n
    _assert_same((got.seqnames, got.start, got.end, got.strand), want)


@pytest.mark.gpu
def test_bed_decode_empty_text_and_errors(rb):
    from recoup_b200 import _lib
    assert len(rb.readBed(b"", ["chr1"])) == 0
    assert len(rb.readBed(b"\n\n# only comments\n", ["chr1"])) == 0
    for bad in (b"chr9\t1\t2\n", b"chr1\t1\n", b"chr1\tx\t5\n", b"chr1\t1\t5\tn\t0\t?\n", b"chr1\t1\t99999999999\n"):
        with pytest.raises(_lib.RecoupError) as ei:
            rb.readBed(b"chr1\t0\t5\n" + bad, ["chr1", "chr2"])
        assert ei.value.code == _lib.RCP_ERR_DATA
