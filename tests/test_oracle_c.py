"""The C twin of the oracle (oracle/recoup_oracle.c) against the numpy restatement: two
independently written restatements of the same R sources must agree bit for bit."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import recoup_oracle as O
from tests.helpers import (assert_coverage_equal, assert_matrix_close, fixture_exons,
                           fixture_genes, fixture_reads, synth_reads)


def _random_regions(rng, n, chrom_len, max_len=3000):
    chrom_len = np.asarray(chrom_len)
    chrom = rng.integers(0, chrom_len.shape[0], size=n)
    L = rng.integers(1, max_len, size=n)
    start = (rng.random(n) * (chrom_len[chrom] + 200)).astype(np.int64) - 100   # some leave the chromosome
    strand = rng.choice(np.array([1, -1, 0]), size=n)
    return chrom, start, start + L - 1, strand


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("ignore,filt", [(True, None), (False, None), (True, 1), (False, -1), (True, 0)])
def test_c_coverage_matches_numpy(seed, ignore, filt):
    rng = np.random.default_rng(seed)
    clen = [5000, 12000, 800]
    chrom, s, e, st = synth_reads(rng, 3000, clen, width=(1, 300))
    reads = O.Reads(chrom, s, e, st, clen)
    rc, rs, re_, rst = _random_regions(rng, 120, clen)
    rs[:3] = [0, -3, 1]
    want = O.calc_coverage(reads, dict(chrom=rc, start=rs, end=re_, strand=rst), filt, ignore)
    ix = CO.Index(chrom, s, e, st, clen)
    got = CO.coverage(ix, rc, rs, re_, rst, ignore, filt).to_list()
    assert_coverage_equal(got, want)
    assert any(w is None for w in want) and any(w is not None for w in want)


@pytest.mark.parametrize("seed", [4, 5])
@pytest.mark.parametrize("ignore", [True, False])
def test_c_list_coverage_matches_numpy(seed, ignore):
    rng = np.random.default_rng(seed)
    clen = [20000, 9000]
    chrom, s, e, st = synth_reads(rng, 4000, clen, width=(20, 900))     # long reads span exons
    reads = O.Reads(chrom, s, e, st, clen)
    ptr, xc, xs, xe, xst = [0], [], [], [], []
    for g in range(40):
        c = int(rng.integers(0, 2))
        ne = int(rng.integers(1, 8))
        pos = int(rng.integers(1, clen[c] - 5000))
        gst = int(rng.choice([1, -1, 0]))
        for _ in range(ne):
            w = int(rng.integers(1, 400))
            xc.append(c); xs.append(pos); xe.append(pos + w - 1); xst.append(gst)
            pos += w + int(rng.integers(1, 500))
        ptr.append(len(xs))
    ptr.append(len(xs))                                                    # an empty element
    mask = dict(ptr=ptr, chrom=xc, start=xs, end=xe, strand=xst)
    want = O.calc_coverage(reads, mask, None, ignore)
    ix = CO.Index(chrom, s, e, st, clen)
    got = CO.coverage_list(ix, ptr, xc, xs, xe, xst, ignore, None).to_list()
    assert_coverage_equal(got, want)
    assert want[-1] is None


def test_c_fixture_and_profiles(fixture_data):
    z = fixture_data
    reads, _ = fixture_reads(z, 0)
    genes, _ = fixture_genes(z)
    ix = CO.Index(reads.chrom, reads.start, reads.end, reads.strand, reads.chrom_len)
    for region, flank, bp in [("tss", (2000, 2000), dict(flankBinSize=0, regionBinSize=100)),
                              ("tes", (1000, 3000), dict(flankBinSize=0, regionBinSize=0)),
                              ("genebody", (2000, 2000), dict(flankBinSize=50, regionBinSize=150,
                                                              sumStat="mean", interpolation="auto")),
                              ("genebody", (2000, 1000), dict(flankBinSize=0, regionBinSize=60,
                                                              sumStat="median", interpolation="spline"))]:
        s, e = O.get_regional_ranges(genes["start"], genes["end"], genes["strand"], region, flank)
        want = O.calc_coverage(reads, dict(chrom=genes["chrom"], start=s, end=e, strand=genes["strand"]))
        dense = CO.coverage(ix, genes["chrom"], s, e, genes["strand"])
        assert_coverage_equal(dense.to_list(), want)
        eq = O.have_equal_lengths(want)
        wm = O.profile_matrix(want, flank, bp)
        gm = CO.profile_matrix(dense, flank, bp, eq)
        assert_matrix_close(gm, wm)
    ex, _ = fixture_exons(z)
    want = O.calc_coverage(reads, ex)
    got = CO.coverage_list(ix, ex["ptr"], ex["chrom"], ex["start"], ex["end"], ex["strand"]).to_list()
    assert_coverage_equal(got, want)


def test_c_fragment_extension():
    rng = np.random.default_rng(9)
    clen = [3000]
    chrom, s, e, st = synth_reads(rng, 500, clen, width=(36, 36))
    es, ee = O.extend_fragments(s, e, st, 200, chrom, clen)
    assert ((ee - es + 1) <= 200).all() and (es >= 1).all() and (ee <= 3000).all()
    assert (es[st >= 0] == s[st >= 0]).all() and (ee[st < 0] == e[st < 0]).all()
    want = O.calc_coverage(O.Reads(chrom, es, ee, st, clen),
                           dict(chrom=[0, 0], start=[1, 2500], end=[600, 3000], strand=[1, -1]))
    ix = CO.Index(chrom, s, e, st, clen, frag_len=200)
    got = CO.coverage(ix, [0, 0], [1, 2500], [600, 3000], [1, -1]).to_list()
    assert_coverage_equal(got, want)
