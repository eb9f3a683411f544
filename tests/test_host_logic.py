"""Host-side mirror logic (no GPU): geometry against the oracle, argument errors of the
reference's closures."""
import numpy as np
import pytest

import recoup_b200 as rb
from oracle import recoup_oracle as O


def _random_ranges(rng, n):
    start = rng.integers(50_000, 1_000_000, size=n)
    width = rng.integers(1, 5000, size=n)
    strand = rng.choice(np.array([1, -1, 0], dtype=np.int8), size=n)
    return start, start + width - 1, strand


@pytest.mark.parametrize("region", ["tss", "tes", "genebody", "custom"])
@pytest.mark.parametrize("flank", [(2000, 2000), (0, 500), (1000, 0), (0, 0), (5000, 300)])
def test_regional_ranges_match_oracle(region, flank):
    rng = np.random.default_rng(5)
    s, e, st = _random_ranges(rng, 200)
    gr = rb.GRanges(np.zeros(200, np.int32), s, e, strand=st, seqlevels=["c0"])
    got = rb.getRegionalRanges(gr, region, flank)
    ws, we = O.get_regional_ranges(s, e, st, region, flank)
    assert np.array_equal(got.start, ws) and np.array_equal(got.end, we)
    if region in ("tss", "tes"):
        assert np.all(got.width == flank[0] + flank[1])      # promoters(): upstream+downstream
    if region == "genebody":
        assert np.all(got.width == (e - s + 1) + flank[0] + flank[1])


def test_custom_one_bp_ranges_behave_like_tss():
    s = np.array([1000, 2000, 3000])
    st = np.array([1, -1, 0], dtype=np.int8)
    gr = rb.GRanges(np.zeros(3, np.int32), s, s, strand=st, seqlevels=["c0"])
    a = rb.getRegionalRanges(gr, "custom", (500, 500))
    b = rb.getRegionalRanges(gr, "tss", (500, 500))
    assert np.array_equal(a.start, b.start) and np.array_equal(a.end, b.end)
    # documented promoters() formulas: '+' [s-f1, s+f2-1], '-' [e-f2+1, e+f1]
    assert (a.start[0], a.end[0]) == (500, 1499)
    assert (a.start[1], a.end[1]) == (1501, 2500)


@pytest.mark.parametrize("direction", ["upstream", "downstream"])
def test_flanking_ranges_match_oracle(direction):
    rng = np.random.default_rng(6)
    s, e, st = _random_ranges(rng, 100)
    gr = rb.GRanges(np.zeros(100, np.int32), s, e, strand=st, seqlevels=["c0"])
    got = rb.getFlankingRanges(gr, 1000, direction)
    ws, we = O.get_flanking_ranges(s, e, st, 1000, direction)
    assert np.array_equal(got.start, ws) and np.array_equal(got.end, we)
    assert np.all(got.width == 1000)


def test_calccoverage_argument_errors_follow_the_reference():
    gr = rb.GRanges(np.zeros(1, np.int32), [1], [10], seqlevels=["c0"], seqlengths=[100])
    with pytest.raises(ValueError, match="mask argument must be a GRanges or GRangesList"):
        rb.calcCoverage(gr, mask=[1, 2, 3])                                  # coverage.R:131-132
    with pytest.raises(ValueError, match="input argument must be a GenomicRanges"):
        rb.calcCoverage(12345, mask=gr)                                      # coverage.R:127-130
    with pytest.raises(NotImplementedError):
        rb.calcCoverage("sample.bam", mask=gr)


def test_coverage_ref_early_return_when_all_samples_have_coverage():
    sentinel = object()
    inp = [dict(id="a", coverage=sentinel), dict(id="b", coverage=sentinel)]
    assert rb.coverageRef(inp, None, "tss") is inp                           # coverage.R:4-6
    assert rb.coverageRnaRef(inp, None, None, (10, 10)) is inp               # coverage.R:81-83
    inp2 = [dict(id="a", profile=sentinel)]
    assert rb.profileMatrix(inp2, (1, 1), {}) is inp2                        # profile.R:2-4


def test_granges_container():
    gr = rb.GRanges(["chrB", "chrA", "chrB"], [5, 1, 9], width=[3, 3, 3], strand=["+", "-", "*"])
    assert gr.seqlevels == ["chrB", "chrA"]
    assert gr.seqnames.tolist() == [0, 1, 0]
    assert gr.end.tolist() == [7, 3, 11]
    assert gr.strand.tolist() == [1, -1, 0]
    sub = gr.subset(np.array([True, False, True]))
    assert len(sub) == 2 and sub.start.tolist() == [5, 9]
    with pytest.raises(ValueError):
        rb.GRangesList(gr, [0, 2])


def test_granges_with_one_width_keeps_it_as_a_number():
    gr = rb.GRanges([0, 0, 1], [5, 1, 9], width=36, strand=["+", "-", "*"], seqlevels=["a", "b"])
    assert gr.fixed_width == 36 and gr._end is None          # nothing materialised
    assert gr.width.tolist() == [36, 36, 36]
    assert gr.end.tolist() == [40, 36, 44]
    sub = gr.subset(np.array([0, 2]))
    assert sub.end.tolist() == [40, 44] and sub.fixed_width is None
    assert rb.GRanges([0], [5], width=[7]).fixed_width is None
    with pytest.raises(ValueError):
        rb.GRanges([0], [5], width=0)
    with pytest.raises(OverflowError):
        rb.GRanges([0], [2**31 - 5], width=36).end


def test_rle_seqnames_round_trip():
    from recoup_b200 import GRanges, Rle
    x = np.array([0, 0, 0, 2, 2, 1, 1, 1, 1], dtype=np.int32)
    r = Rle.encode(x)
    assert r.values.tolist() == [0, 2, 1] and r.lengths.tolist() == [3, 2, 4] and len(r) == 9
    assert np.array_equal(r.decode(), x)
    assert len(Rle.encode(np.zeros(0, np.int32))) == 0
    g = GRanges(r, np.arange(1, 10), np.arange(1, 10) + 5, seqlevels=["a", "b", "c"])
    assert g.seqnames_rle.nrun == 3 and np.array_equal(g.seqnames, x)
    g2 = GRanges(Rle(["a", "c", "b"], [3, 2, 4]), np.arange(1, 10), width=np.full(9, 6),
                 seqlevels=["a", "b", "c"])
    assert np.array_equal(g2.seqnames, x) and np.array_equal(g2.end, g.end)
    sub = g2.subset(np.array([0, 4, 8]))
    assert sub.seqnames.tolist() == [0, 2, 1] and sub.seqnames_rle is None
    with pytest.raises(ValueError):
        GRanges(Rle([0], [8]), np.arange(1, 10), np.arange(1, 10))
    with pytest.raises(ValueError):
        Rle([0, 1], [3, -1])


def test_downsampling_indices_follow_one_r_stream():
    """sort(sample(libSize, s)) per sample after ONE set.seed (ranges.R:38-41): the library's
    host RNG against the oracle's, including sample.int's hash variant for n > 1e7"""
    from oracle.r_rng import RRandom
    for kind in ("Rejection", "Rounding"):
        r = RRandom(7, kind)
        want = [np.sort(r.sample_int(n, 400)) for n in (1000, 20000001, 400)]
        got = rb.sampleSorted([1000, 20000001, 400], 400, 7, kind)
        for w_, g_ in zip(want, got):
            assert np.array_equal(w_, g_)
    assert [a.tolist() for a in rb.sampleSorted([10, 10], 3, 42)] == [[1, 5, 10], [2, 4, 9]]
    idx = O.downsample_indices([1000, 600], "downsample", seed=7)
    assert np.array_equal(idx[0], rb.sampleSorted([1000, 600], 600, 7)[0]) and len(idx[1]) == 600
    assert O.downsample_indices([5, 6], "linear") == [None, None]
    with pytest.raises(rb.RecoupError):
        rb.sampleSorted([10], 11, 42)


def test_selected_granges_and_preprocess_host_logic():
    gr = rb.GRanges(np.zeros(6, np.int32), [1, 10, 20, 30, 40, 50], [5, 29, 24, 34, 90, 54],
                    strand=[1, -1, 1, 0, 1, -1], seqlevels=["c"], seqlengths=[100])
    sel = rb.SelectedGRanges(gr, 4, max_width=5.0, idx=[1, 3, 4])     # widths 5 20 5 5 51 5
    assert len(sel) == 3 and sel.start.tolist() == [1, 30, 50] and sel.strand.tolist() == [1, 0, -1]
    assert sel.seqlengths.tolist() == [100] and sel.width.tolist() == [5, 5, 5]
    keep, qu = O.splice_remove(gr.start, gr.end, 0.75)
    assert qu == 16.25 and keep.tolist() == [True, False, True, True, False, True]
    done = [{"ranges": gr}]
    assert rb.preprocessRanges(done, {"normalize": "downsample"}) is done     # ranges.R:2-4
    with pytest.raises(ValueError):
        rb.preprocessRanges([{"name": "a"}], {"normalize": "none"})
    with pytest.raises(ValueError):
        rb.preprocessRanges([{"name": "a"}], {"normalize": "bogus"}, reader=lambda x: gr)
    out = rb.preprocessRanges([{"name": "a"}], {"normalize": "none"}, reader=lambda x: gr)
    assert out[0]["ranges"] is gr
