"""The oracle against (1) the reference's decoded fixture and the invariants SURVEY.md section 4
derives from it, (2) a brute-force definition of coverage (property tests), (3) hand-worked
cases of every NULL / strand / multiplicity rule."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import recoup_oracle as O
from tests.helpers import fixture_exons, fixture_genes, fixture_reads


def test_fixture_shapes(fixture_data):
    z = fixture_data
    for k in range(2):
        assert z["reads_%d_start" % k].shape == (100000,)
        assert set(np.unique(z["reads_%d_width" % k])) == {180}
    assert int(z["chrom_len"][0]) == 121257530
    assert z["gene_start"].shape == (100,)
    assert z["exon_ptr"][-1] == 1125 and z["exon_ptr"].shape == (101,)
    w = z["gene_end"] - z["gene_start"] + 1
    assert w.min() == 54 and w.max() == 823499


@pytest.mark.parametrize("k,n_null", [(0, 6), (1, 4)])
def test_fixture_tss_invariants(fixture_data, k, n_null):
    reads, _ = fixture_reads(fixture_data, k)
    genes, _ = fixture_genes(fixture_data)
    cov = O.coverage_ref(reads, genes, "tss", (2000, 2000))
    assert sum(c is None for c in cov) == n_null          # zero-read windows -> NULL rows
    assert {c.shape[0] for c in cov if c is not None} == {4000}
    m = O.profile_matrix(cov, (2000, 2000), dict(flankBinSize=0, regionBinSize=100))
    assert m.shape == (100, 100)
    null_rows = [i for i, c in enumerate(cov) if c is None]
    assert np.all(m[null_rows] == 0)
    # 40 bp / bin, d = 0: RNG-independent -> plain reshape mean
    for i, c in enumerate(cov):
        if c is not None:
            np.testing.assert_allclose(m[i], c.reshape(100, 40).mean(axis=1))


def test_fixture_genebody_invariants(fixture_data):
    reads, _ = fixture_reads(fixture_data, 0)
    genes, _ = fixture_genes(fixture_data)
    cov = O.coverage_ref(reads, genes, "genebody", (2000, 2000))
    assert all(c is not None for c in cov)
    assert sum(c.shape[0] for c in cov) == 7779665
    m = O.profile_matrix(cov, (2000, 2000), dict(flankBinSize=50, regionBinSize=150,
                                                 sumStat="mean", interpolation="auto"))
    assert m.shape == (100, 250) and not np.isnan(m).any() and (m >= 0).all()


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_coverage_matches_brute_force(data):
    clen = data.draw(st.integers(50, 400))
    n = data.draw(st.integers(0, 40))
    starts = [data.draw(st.integers(1, clen)) for _ in range(n)]
    ends = [data.draw(st.integers(s, min(clen, s + 60))) for s in starts]
    strands = [data.draw(st.sampled_from([1, -1, 0])) for _ in range(n)]
    rs = data.draw(st.integers(-5, clen + 5))
    re_ = data.draw(st.integers(rs, rs + 80))
    rst = data.draw(st.sampled_from([1, -1, 0]))
    ignore = data.draw(st.booleans())
    reads = O.Reads(np.zeros(n, int), starts, ends, strands, [clen])
    got = O.coverage_from_ranges(reads, 0, [rs], [re_], [rst], ignore)
    sa, ea, sta = np.array(starts, int), np.array(ends, int), np.array(strands, int)
    ok = np.ones(n, bool) if ignore or rst == 0 else ((sta == rst) | (sta == 0))
    hit = ok & (sa <= re_) & (ea >= rs) if n else np.zeros(0, bool)
    if not hit.any() or rs < 0 or re_ > clen:
        assert got is None
        return
    lo = max(rs, 1)
    want = O.brute_coverage(sa[hit], ea[hit], lo, re_)
    if rst < 0:
        want = want[::-1]
    assert np.array_equal(got, want)


def test_null_rules():
    reads = O.Reads([0, 0], [10, 30], [19, 39], [1, -1], [100])
    f = lambda s, e, st=1, ig=True: O.coverage_from_ranges(reads, 0, [s], [e], [st], ig)
    assert f(50, 60) is None                       # no overlapping read (coverage.R:198,224)
    assert f(95, 101) is None                      # end past the chromosome, even w/o reads
    reads2 = O.Reads([0], [95], [100], [1], [100])
    assert O.coverage_from_ranges(reads2, 0, [95], [101], [1]) is None   # out of bounds -> tryCatch
    assert O.coverage_from_ranges(reads2, 0, [-2], [96], [1]) is None    # mixed-sign subscript
    got = O.coverage_from_ranges(O.Reads([0], [1], [5], [1], [100]), 0, [0], [6], [1])
    assert got.tolist() == [1, 1, 1, 1, 1, 0]      # zero index dropped: length L-1
    assert O.coverage_from_ranges(reads, 1, [10], [20], [1]) is None     # chromosome absent
    assert f(15, 34).tolist() == [1] * 5 + [0] * 10 + [1] * 5
    assert f(15, 34, st=-1).tolist() == ([1] * 5 + [0] * 10 + [1] * 5)[::-1]
    assert f(15, 34, st=1, ig=False).tolist() == [1] * 5 + [0] * 15     # '-' read not counted
    assert f(25, 34, st=1, ig=False) is None


def test_rna_multiplicity_and_merge():
    # one read spanning both exons of a gene is counted twice at every base it covers
    reads = O.Reads([0, 0], [12, 40], [33, 44], [0, 0], [200])
    cov = O.coverage_from_ranges(reads, 0, [10, 30], [14, 34], [1, 1])
    assert cov.tolist() == [0, 0, 2, 2, 2] + [2, 2, 2, 2, 0]
    exons = dict(ptr=[0, 2], chrom=[0, 0], start=[10, 30], end=[14, 34], strand=[1, 1])
    genes = dict(chrom=[0], start=[10], end=[34], strand=[1])
    merged = O.coverage_rna_ref(reads, exons, genes, (3, 5))
    assert merged[0] is None                       # upstream flank [7,9] has no read -> NULL
    reads3 = O.Reads([0, 0, 0], [8, 12, 36], [9, 33, 44], [0, 0, 0], [200])
    merged = O.coverage_rna_ref(reads3, exons, genes, (3, 5))
    assert merged[0].tolist() == [0, 1, 1] + [0, 0, 2, 2, 2, 2, 2, 2, 2, 0] + [0, 1, 1, 1, 1]
    genes_m = dict(chrom=[0], start=[10], end=[34], strand=[-1])
    exons_m = dict(exons, strand=[-1, -1])
    mm = O.coverage_rna_ref(reads3, exons_m, genes_m, (3, 5))
    # '-' gene: upstream flank is [35,37] (width 3), downstream [5,9]
    assert mm[0].tolist() == [1, 1, 0] + [0, 2, 2, 2, 2, 2, 2, 2, 0, 0] + [1, 1, 0, 0, 0]


def test_bin_layout_and_split_vector():
    fac = O.bin_layout(10, 4)                      # 10 = 4*2 + 2 -> two bins of 3
    assert fac.sum() == 10 and sorted(fac.tolist()) == [2, 2, 3, 3]
    from oracle.r_rng import r_sample
    add = r_sample(4, 2)
    assert all(fac[a - 1] == 3 for a in add)
    x = np.arange(10, dtype=float)
    m = O.split_vector(x, 4)
    edges = np.concatenate([[0], np.cumsum(fac)])
    np.testing.assert_allclose(m, [x[edges[i]:edges[i + 1]].mean() for i in range(4)])
    med = O.split_vector(np.array([5, 1, 9, 3, 7, 7, 2, 8.0]), 2, stat="median")
    np.testing.assert_allclose(med, [4.0, 7.0])


def test_spline_reproduces_cubics_and_endpoints():
    # the fmm end conditions make the spline exact for cubic data
    xs = np.arange(1, 13, dtype=float)
    y = 0.5 * xs ** 3 - 2 * xs ** 2 + xs + 3
    out = O.r_spline(y, 40)
    xo = np.linspace(1, 12, 40)
    np.testing.assert_allclose(out, 0.5 * xo ** 3 - 2 * xo ** 2 + xo + 3, rtol=1e-10)
    assert O.r_spline([4.0], 5).tolist() == [4.0] * 5
    np.testing.assert_allclose(O.r_spline([1.0, 3.0], 5), [1, 1.5, 2, 2.5, 3])
    scipy = pytest.importorskip("scipy.interpolate")
    yy = np.array([3, 0, 7, 2, 2, 9, 4, 1.0])
    cs = scipy.CubicSpline(np.arange(1, 9), yy, bc_type="not-a-knot")
    # fmm != not-a-knot in general, but both interpolate the knots
    np.testing.assert_allclose(O.r_spline(yy, 8), yy, atol=1e-12)
    assert np.allclose(cs(np.arange(1, 9)), yy)


def test_interpolation_paths():
    x = np.array([0, 2, 4, 4, 1.0])
    sp = O.split_vector(x, 20, "auto")             # (20-5)/20 >= 0.2 -> spline, clamped at 0
    assert sp.shape == (20,) and (sp >= 0).all() and sp[0] == 0 and sp[-1] == 1
    x2 = np.arange(1, 18, dtype=float)             # (20-17)/20 < 0.2 -> neighbourhood
    nb = O.split_vector(x2, 20, "auto")
    assert nb.shape == (20,) and nb[0] == 1 and nb[1] == 2 and nb[-1] == 17 and nb[-2] == 16
    assert np.all(np.diff(nb[~np.isnan(nb)]) >= 0)
    with pytest.raises(ValueError, match="dead code"):
        O.split_vector(x, 20, "linear")            # util.R:49 'inear'


def test_profile_matrix_paths():
    cov = [np.arange(20), None, np.arange(20)[::-1].copy()]
    m = O.profile_matrix(cov, (5, 5), dict(flankBinSize=0, regionBinSize=4))
    assert m.shape == (3, 4) and (m[1] == 0).all()
    pb = O.profile_matrix(cov, (5, 5), dict(flankBinSize=0, regionBinSize=0))
    assert pb.shape == (3, 20) and (pb[0] == np.arange(20)).all() and (pb[1] == 0).all()
    cov2 = [np.arange(30), np.arange(26), None]
    u = O.profile_matrix(cov2, (5, 5), dict(flankBinSize=2, regionBinSize=4, sumStat="mean",
                                            interpolation="auto"))
    assert u.shape == (3, 2 + 4 + 2)               # round(2*2*0.5) = 2 bins per flank
    k = int(O.bin_layout(5, 2)[0])                 # first upstream bin has k of the 5 flank bases
    np.testing.assert_allclose(u[0, :2], [np.arange(5)[:k].mean(), np.arange(5)[k:].mean()])
    u2 = O.profile_matrix(cov2, (5, 5), dict(flankBinSize=0, regionBinSize=4))
    assert u2.shape == (3, 5 + 4 + 5) and (u2[0, :5] == np.arange(5)).all()
    assert (u2[1, -5:] == np.arange(21, 26)).all() and (u2[2] == 0).all()
    assert O.r_round(2.5) == 2 and O.r_round(3.5) == 4 and O.r_round(33.33) == 33


def test_consumer_restatements_known_answers():
    """R known answers (base R semantics): sort(index.return) is stable and drops NA; quantile is
    type 7; mad uses the constant 1.4826; mean refines its first pass."""
    v = np.array([3, 1, 2, 1, np.nan, 3.0])
    assert O.r_sort_index(v).tolist() == [2, 4, 3, 1, 6]
    assert O.r_sort_index(v, decreasing=True).tolist() == [1, 6, 3, 2, 4]
    # quantile(1:10, c(.1,.5,.95)) = 1.90 5.50 9.55;  quantile(c(0,0,0,4), .95) = 3.4
    assert np.allclose(O.r_quantile7(np.arange(1, 11), [0.1, 0.5, 0.95]), [1.9, 5.5, 9.55])
    assert np.allclose(O.r_quantile7([0, 0, 0, 4], [0.95]), [3.4])
    with pytest.raises(ValueError):
        O.r_quantile7([1.0, np.nan], [0.5])
    # x = matrix(c(1,2,3,4, 10,20,30,50), 4): colMeans 2.5 27.5; sd 1.290994 17.07825;
    # median 2.5 25; mad 1.4826 14.826
    x = np.array([[1, 10], [2, 20], [3, 30], [4, 50.0]])
    p = O.plot_profile(x, "mean")
    assert np.allclose(p["profile"], [2.5, 27.5])
    assert np.allclose(p["upper"] - p["profile"], [1.2909944487, 17.0782512766])
    p = O.plot_profile(x, "median")
    assert np.allclose(p["profile"], [2.5, 25.0]) and np.allclose(p["profile"] - p["lower"], [1.4826, 14.826])
    p = O.plot_profile(x, "mean", "log2")
    assert np.allclose(p["profile"], np.log2(x + 1).mean(0))
    assert O.row_order_values(x, "sum").tolist() == [11, 22, 33, 54]
    assert O.row_order_values(x, "max").tolist() == [10, 20, 30, 50]
    assert np.allclose(O.row_order_values(x, "avg"), [5.5, 11, 16.5, 27])
    assert O.r_median([5, 1, 3]) == 3 and O.r_median([4, 1, 3, 2]) == 2.5
