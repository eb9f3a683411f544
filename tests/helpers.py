"""Shared builders for the parity tests: the same seeded inputs go to the oracle and to the
CUDA library."""
import numpy as np

from oracle import recoup_oracle as O
import recoup_b200 as rb


def fixture_reads(z, k):
    s = z["reads_%d_start" % k].astype(np.int64)
    w = z["reads_%d_width" % k].astype(np.int64)
    st = z["reads_%d_strand" % k]
    chrom = np.zeros(s.shape[0], dtype=np.int32)
    oracle = O.Reads(chrom, s, s + w - 1, st, z["chrom_len"])
    gr = rb.GRanges(chrom, s, s + w - 1, strand=st, seqlevels=["chr12"], seqlengths=z["chrom_len"])
    return oracle, gr


def fixture_genes(z):
    n = z["gene_start"].shape[0]
    od = dict(chrom=np.zeros(n, dtype=np.int64), start=z["gene_start"], end=z["gene_end"],
              strand=z["gene_strand"])
    gr = rb.GRanges(np.zeros(n, dtype=np.int32), z["gene_start"], z["gene_end"],
                    strand=z["gene_strand"], seqlevels=["chr12"], names=list(z["gene_names"]))
    return od, gr


def fixture_exons(z):
    ne = z["exon_start"].shape[0]
    od = dict(ptr=z["exon_ptr"], chrom=np.zeros(ne, dtype=np.int64), start=z["exon_start"],
              end=z["exon_end"], strand=z["exon_strand"])
    u = rb.GRanges(np.zeros(ne, dtype=np.int32), z["exon_start"], z["exon_end"],
                   strand=z["exon_strand"], seqlevels=["chr12"])
    grl = rb.GRangesList(u, z["exon_ptr"], names=list(z["exon_gene_names"]))
    return od, grl


def synth_reads(rng, n, chrom_len, width=(20, 120), strands=(1, -1, 0), sort=False):
    chrom_len = np.asarray(chrom_len, dtype=np.int64)
    chrom = rng.integers(0, chrom_len.shape[0], size=n).astype(np.int32)
    w = rng.integers(width[0], width[1] + 1, size=n).astype(np.int64)
    w = np.minimum(w, chrom_len[chrom])
    start = (rng.random(n) * (chrom_len[chrom] - w + 1)).astype(np.int64) + 1
    end = start + w - 1
    strand = rng.choice(np.asarray(strands, dtype=np.int8), size=n)
    if sort:
        o = np.lexsort((start, chrom))
        chrom, start, end, strand = chrom[o], start[o], end[o], strand[o]
    return chrom, start, end, strand


def both_reads(chrom, start, end, strand, chrom_len):
    lv = ["c%d" % i for i in range(len(chrom_len))]
    return (O.Reads(chrom, start, end, strand, chrom_len),
            rb.GRanges(chrom, start, end, strand=strand, seqlevels=lv, seqlengths=chrom_len))


def both_regions(chrom, start, end, strand, n_chrom):
    lv = ["c%d" % i for i in range(n_chrom)]
    od = dict(chrom=np.asarray(chrom, dtype=np.int64), start=np.asarray(start, dtype=np.int64),
              end=np.asarray(end, dtype=np.int64), strand=np.asarray(strand, dtype=np.int64))
    gr = rb.GRanges(np.asarray(chrom, dtype=np.int32), start, end, strand=strand, seqlevels=lv)
    return od, gr


def assert_coverage_equal(got_list, want_list):
    """Coverage parity is BIT-EXACT (integer counts)."""
    assert len(got_list) == len(want_list)
    for i, (g, w) in enumerate(zip(got_list, want_list)):
        if w is None:
            assert g is None, "region %d: oracle NULL, CUDA length %d" % (i, len(g))
        else:
            assert g is not None, "region %d: CUDA NULL, oracle length %d" % (i, len(w))
            assert g.shape[0] == w.shape[0], "region %d: length %d != %d" % (i, g.shape[0], w.shape[0])
            assert np.array_equal(g.astype(np.int64), w), "region %d: coverage differs" % i


MATRIX_RTOL = 1e-6   # north_star: profile-matrix bin means within 1e-6 relative


def assert_matrix_close(got, want):
    got = np.asarray(got)
    assert got.shape == want.shape, "%s != %s" % (got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(want), rtol=MATRIX_RTOL, atol=1e-12)
