"""Known-answer tests that pin the oracle to PUBLISHED base-R / Bioconductor behaviour.

R is not installed here, so none of these values was produced by running R in this image.  Each
case takes its INPUT from a published man page / vignette and its EXPECTED value either from output
that the same document (or R's well-known seeded streams) prints, or -- where the document prints
no output -- from the definition the man page states, worked by hand in the comment.  The source is
named next to every case; `tools/make_r_golden.R` regenerates all of them (and the C1 coverage /
matrices of the real package) with a real R, and `tests/test_r_golden.py` picks that file up.
"""
import numpy as np
import pytest

from oracle import recoup_oracle as O
from oracle.r_rng import RRandom, r_sample


# ---- base R: set.seed / runif / sample (Mersenne-Twister, Inversion, Rejection | Rounding) -----
def test_runif_streams_of_well_known_seeds():
    # `set.seed(1); runif(3)` and `set.seed(123); runif(3)`: the first lines of countless R
    # tutorials (e.g. ?set.seed examples worked in "R for Data Science"); R >= 1.7 default RNG.
    r = RRandom(1)
    got = [r.unif_rand() for _ in range(3)]
    np.testing.assert_allclose(got, [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    r = RRandom(123)
    got = [r.unif_rand() for _ in range(3)]
    np.testing.assert_allclose(got, [0.2875775, 0.7883051, 0.4089769], atol=5e-8)


def test_sample_permutations_both_sample_kinds():
    # R >= 3.6.0 NEWS: "sample.kind = 'Rejection'" became the default; the permutations before
    # and after the change are the ones every migration note quotes.
    assert r_sample(10, 10, seed=1, sample_kind="Rounding") == [3, 4, 5, 7, 2, 8, 9, 6, 10, 1]
    assert r_sample(10, 10, seed=1, sample_kind="Rejection") == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    assert r_sample(10, 10, seed=123, sample_kind="Rounding") == [3, 8, 4, 7, 6, 1, 10, 9, 2, 5]
    assert r_sample(10, 10, seed=123, sample_kind="Rejection") == [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]


# ---- base R: round (IEC 60559 half-even), ?Round ------------------------------------------------
def test_round_half_even_as_documented():
    # ?Round: "round(0.5) is 0 and round(-1.5) is -2" (go to the even digit); profile.R:36,51
    assert O.r_round(0.5) == 0
    assert O.r_round(-1.5) == -2
    assert O.r_round(2.5) == 2
    assert O.r_round(3.5) == 4


# ---- stats::quantile(type = 7), ?quantile -------------------------------------------------------
def test_quantile_type7_definition():
    # ?quantile, "Type 7: m = 1 - p, p[k] = (k - 1) / (n - 1)": Q(p) = x[j] + g (x[j+1] - x[j]),
    # j = floor((n-1)p) + 1, g = (n-1)p - floor((n-1)p).
    # x = 1:10, p = 0.95: (n-1)p = 8.55 -> x[9] + 0.55 = 9.55;  p = 0.25: 2.25 -> 3.25
    np.testing.assert_allclose(O.r_quantile7(np.arange(1, 11), [0.95, 0.25, 0.5, 0, 1]),
                               [9.55, 3.25, 5.5, 1.0, 10.0], rtol=1e-15)
    # x = c(10, 20, 30, 40), p = 0.95: 3 * 0.95 = 2.85 -> 30 + 0.85 * 10 = 38.5
    np.testing.assert_allclose(O.r_quantile7([40, 10, 30, 20], [0.95]), [38.5], rtol=1e-15)


# ---- stats::spline(method = "fmm"), ?splinefun --------------------------------------------------
def test_fmm_spline_reproduces_cubics_exactly():
    # ?splinefun: method "fmm" "(Forsythe, Malcolm and Moler) ... an exact cubic is fitted through
    # the four points at each end of the data, and this is used to determine the end conditions".
    # Hence data sampled from ONE cubic are reproduced by that cubic everywhere.
    L = 9
    x = np.arange(1, L + 1, dtype=np.float64)
    f = lambda t: 0.5 * t ** 3 - 4.0 * t ** 2 + 3.0 * t + 7.0      # noqa: E731
    for n in (9, 17, 40):
        xout = np.linspace(1.0, float(L), n)
        np.testing.assert_allclose(O.r_spline(f(x), n), f(xout), rtol=1e-12, atol=1e-10)
    # and through the knots for arbitrary data (interpolating spline)
    y = np.array([3, 0, 7, 7, 2, 9, 1, 4, 4], dtype=np.float64)
    np.testing.assert_allclose(O.r_spline(y, 2 * L - 1)[::2], y, rtol=0, atol=1e-12)


# ---- IRanges::coverage, vignette "An Overview of the IRanges package" ---------------------------
IR_START = np.array([1, 8, 14, 15, 19, 34, 40])
IR_WIDTH = np.array([12, 6, 6, 15, 6, 2, 7])


def test_iranges_vignette_coverage_example():
    # the vignette's running example `ir <- IRanges(c(1, 8, 14, 15, 19, 34, 40),
    # width = c(12, 6, 6, 15, 6, 2, 7))`; `coverage(ir)` prints an integer-Rle of length 46 with
    # 11 runs, Lengths 7 5 2 4 1 5 5 4 2 4 7, Values 1 2 1 2 3 2 1 0 1 0 1.
    end = IR_START + IR_WIDTH - 1
    cov = O.brute_coverage(IR_START, end, 1, 46)
    lengths = [7, 5, 2, 4, 1, 5, 5, 4, 2, 4, 7]
    values = [1, 2, 1, 2, 3, 2, 1, 0, 1, 0, 1]
    assert np.array_equal(cov, np.repeat(values, lengths))
    # the same through the oracle's coverageFromRanges restatement: one region = the whole vector
    n = IR_START.shape[0]
    reads = O.Reads(np.zeros(n, dtype=np.int32), IR_START, end, np.zeros(n, dtype=np.int8), np.array([46]))
    got = O.coverage_from_ranges(reads, 0, [1], [46], [0])
    assert np.array_equal(got, np.repeat(values, lengths))
    # `[start:end]` then `rev` for a '-' region (coverage.R:209-213)
    got = O.coverage_from_ranges(reads, 0, [10], [20], [-1])
    assert np.array_equal(got, np.repeat(values, lengths)[9:20][::-1])


# ---- IRanges / GenomicRanges intra-range methods: ?flank, ?promoters, ?resize -------------------
def test_flank_promoters_resize_as_documented():
    # ?flank (IRanges intra-range-methods) example input: ir3 <- IRanges(c(2,5,1), c(3,7,3)).
    # Definition (same page): flank(x, width, start=TRUE) = [start - width, start - 1];
    # start=FALSE: [end + 1, end + width].  GRanges: "start" means the 5' end, so on the '-'
    # strand the two swap.  ranges.R:93-100 uses promoters(x, f, 0) upstream and
    # flank(x, f, start=FALSE) downstream.
    s = np.array([2, 5, 1])
    e = np.array([3, 7, 3])
    plus = np.array([1, 1, 1])
    minus = -plus
    us, ue = O.get_flanking_ranges(s, e, plus, 2, "upstream")      # == flank(ir3, 2)
    assert us.tolist() == [0, 3, -1] and ue.tolist() == [1, 4, 0]
    ds, de = O.get_flanking_ranges(s, e, plus, 2, "downstream")    # == flank(ir3, 2, start=FALSE)
    assert ds.tolist() == [4, 8, 4] and de.tolist() == [5, 9, 5]
    us, ue = O.get_flanking_ranges(s, e, minus, 2, "upstream")     # 5' end of '-' is `end`
    assert us.tolist() == [4, 8, 4] and ue.tolist() == [5, 9, 5]
    ds, de = O.get_flanking_ranges(s, e, minus, 2, "downstream")
    assert ds.tolist() == [0, 3, -1] and de.tolist() == [1, 4, 0]
    # ?promoters: "promoters(x, upstream=2000, downstream=200)": the result spans
    # [start - upstream, start + downstream - 1] on '+' / '*' and [end - downstream + 1,
    # end + upstream] on '-', width upstream + downstream.  Example input of the same page:
    # ir4 <- IRanges(20:23, width=3); promoters(ir4, upstream=0, downstream=1) is "start value
    # only", promoters(ir4, upstream=1, downstream=0) "single upstream nucleotide".
    s4 = np.arange(20, 24)
    e4 = s4 + 2
    p = np.ones(4, dtype=np.int64)
    ps, pe = O._promoters(s4, e4, p, 0, 1)
    assert ps.tolist() == [20, 21, 22, 23] and pe.tolist() == [20, 21, 22, 23]
    ps, pe = O._promoters(s4, e4, p, 1, 0)
    assert ps.tolist() == [19, 20, 21, 22] and pe.tolist() == [19, 20, 21, 22]
    ps, pe = O._promoters(s4, e4, -p, 1, 0)                        # '-': upstream of `end`
    assert ps.tolist() == [23, 24, 25, 26] and pe.tolist() == [23, 24, 25, 26]
    # getRegionalRanges(region="tss") = promoters(x, f1, f2) (ranges.R:70-73): width f1 + f2
    ts, te = O.get_regional_ranges(np.array([1000, 5000]), np.array([2000, 9000]), np.array([1, -1]),
                                   "tss", (300, 200))
    assert ts.tolist() == [700, 8801] and te.tolist() == [1199, 9300]
    # ?resize: fix="end" keeps the 3' end: '+' -> [end - width + 1, end], '-' -> [start, start + width - 1]
    rs, re_ = O._resize(np.array([10, 10]), np.array([19, 19]), np.array([1, -1]), 4, fix="end")
    assert rs.tolist() == [16, 10] and re_.tolist() == [19, 13]
    rs, re_ = O._resize(np.array([10, 10]), np.array([19, 19]), np.array([1, -1]), 4, fix="start")
    assert rs.tolist() == [10, 16] and re_.tolist() == [13, 19]


# ---- splitVector's bin layout on a worked case (util.R:74-84) -----------------------------------
def test_split_vector_layout_from_the_seed42_permutation():
    # set.seed(42); sample(1:10, 3) is the head of sample(1:10) = 1 5 10 8 2 4 6 9 7 3 (R >= 3.6,
    # SURVEY 8c): L = 23, n = 10 -> bin.size 2, dif 3, bins 1, 5 and 10 get the extra base.
    sizes = O.bin_layout(23, 10)
    assert list(sizes) == [3, 2, 2, 2, 3, 2, 2, 2, 2, 3]
    x = np.arange(1, 24, dtype=np.float64)
    got = O.split_vector(x, 10)
    want = [np.mean(c) for c in np.split(x, np.cumsum(sizes)[:-1])]
    np.testing.assert_allclose(got, want, rtol=1e-15)
