"""Parity against outputs of the REAL reference, when somebody has produced them.

`Rscript tools/make_r_golden.R` (on a machine with R + Bioconductor + recoup) writes
tests/golden/r_outputs/; this module then compares the oracle -- and, under `-m gpu`, the CUDA
path -- with it: coverage bit-exact, matrices within 1e-6 relative.  Without that directory the
tests SKIP and the parity of the repository stays "unpinned" (DESIGN.md section 2): nothing here
fabricates R output.
"""
import os

import numpy as np
import pytest

from oracle import recoup_oracle as O
from tests.helpers import (assert_coverage_equal, assert_matrix_close, fixture_exons, fixture_genes,
                           fixture_reads)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r_outputs")
needs_r = pytest.mark.skipif(not os.path.isdir(GOLD),
                             reason="no tests/golden/r_outputs (run tools/make_r_golden.R with a real R): "
                                    "parity unpinned")
SAMPLES = ["WT_H4K20me1", "Set8KO_H4K20me1"]
BP = {"tss": dict(flankBinSize=0, regionBinSize=100, sumStat="mean", interpolation="auto"),
      "genebody": dict(flankBinSize=50, regionBinSize=150, sumStat="mean", interpolation="auto"),
      "rna": dict(flankBinSize=50, regionBinSize=150, sumStat="mean", interpolation="auto")}


def read_cov(path):
    out = []
    for line in open(path):
        name, rest = line.rstrip("\n").split(",", 1)
        if rest == "NULL":
            out.append(None)
            continue
        n, runs = rest.split(",", 1)
        v = np.concatenate([np.full(int(r.split(":")[1]), int(r.split(":")[0]), dtype=np.int64)
                            for r in runs.split()]) if runs else np.zeros(0, dtype=np.int64)
        assert v.shape[0] == int(n)
        out.append(v)
    return out


def read_mat(path):
    return np.array([[float(x) for x in line.rstrip("\n").split(",")[1:]] for line in open(path)])


def oracle_case(z, case, k):
    reads, _ = fixture_reads(z, k)
    genes, _ = fixture_genes(z)
    if case == "rna":
        exons, _ = fixture_exons(z)
        cov = O.coverage_rna_ref(reads, exons, genes, (2000, 2000))
    else:
        cov = O.coverage_ref(reads, genes, case, (2000, 2000))
    return cov, O.profile_matrix(cov, (2000, 2000), BP[case])


@needs_r
@pytest.mark.parametrize("case", ["tss", "genebody", "rna"])
@pytest.mark.parametrize("k", [0, 1])
def test_oracle_matches_r(fixture_data, case, k):
    cov, mat = oracle_case(fixture_data, case, k)
    assert_coverage_equal(cov, read_cov(os.path.join(GOLD, "%s_%s_coverage.csv" % (case, SAMPLES[k]))))
    assert_matrix_close(mat, read_mat(os.path.join(GOLD, "%s_%s_profile.csv" % (case, SAMPLES[k]))))


@needs_r
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tss", "genebody", "rna"])
def test_cuda_matches_r(gpu, fixture_data, case):
    rb = gpu
    _, g_reads = fixture_reads(fixture_data, 0)
    _, g_genes = fixture_genes(fixture_data)
    inp = [dict(id="WT", name="WT", ranges=g_reads)]
    if case == "rna":
        _, grl = fixture_exons(fixture_data)
        rb.coverageRnaRef(inp, grl, g_genes, (2000, 2000))
    else:
        rb.coverageRef(inp, g_genes, case, (2000, 2000))
    rb.profileMatrix(inp, (2000, 2000), BP[case])
    assert_coverage_equal(inp[0]["coverage"].to_list(),
                          read_cov(os.path.join(GOLD, "%s_%s_coverage.csv" % (case, SAMPLES[0]))))
    assert_matrix_close(inp[0]["profile"], read_mat(os.path.join(GOLD, "%s_%s_profile.csv" % (case, SAMPLES[0]))))


def test_the_loader_round_trips(tmp_path):
    """The CSV reader understands exactly what make_r_golden.R's write_cov / write_mat emit."""
    p = tmp_path / "c.csv"
    p.write_text("g1,5,0:2 3:1 1:2\ng2,NULL\n")
    cov = read_cov(str(p))
    assert cov[1] is None and cov[0].tolist() == [0, 0, 3, 1, 1]
    m = tmp_path / "m.csv"
    m.write_text("g1,0.5,1.25\ng2,0,0\n")
    assert read_mat(str(m)).tolist() == [[0.5, 1.25], [0.0, 0.0]]
