#!/usr/bin/env python
"""Decode the reference's bundled data set into a small integer fixture.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_fixtures.py

Reads  /root/reference/data/recoup_test_data.rda  (man/recoup_test_data.Rd:1-39) and writes
`tests/golden/recoup_test_data.npz` holding, for config C1 and the reference's own smoke test
(inst/unitTests/test_recoup.R:1-32):

  reads_<k>_start / _width / _strand   k = 0 (WT_H4K20me1), 1 (Set8KO_H4K20me1); strand +1/-1/0
  sample_ids, chrom_names, chrom_len   single chromosome chr12
  gene_start / gene_end / gene_strand / gene_names          test.genome (100 genes)
  exon_ptr / exon_start / exon_end / exon_strand / exon_gene_names   test.exons (GRangesList)

The reference's tests assert no numbers (test_recoup.R:28-31), so this fixture pins INPUTS only;
outputs are pinned by the oracle (see oracle/README.md, "parity unpinned").
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import rdx2  # noqa: E402

SRC = "/root/reference/data/recoup_test_data.rda"


def rle_expand(rle):
    values = rdx2.val(rle["values"])
    lengths = rdx2.val(rle["lengths"])
    levels = rdx2.attr(rle["values"], "levels")
    out = np.repeat(np.asarray(values), np.asarray(lengths))
    return out, levels


def strand_codes(rle):
    codes, levels = rle_expand(rle)  # factor codes, 1-based into levels ("+","-","*")
    lut = {"+": 1, "-": -1, "*": 0}
    table = np.array([lut[l] for l in levels], dtype=np.int8)
    return table[codes - 1]


def main():
    d = rdx2.load_rda(SRC)
    out = {}

    samples = rdx2.val(d["test.input"])
    ids = []
    chrom_names = None
    for k, s in enumerate(samples):
        fields = dict(zip(rdx2.attr(s, "names"), rdx2.val(s)))
        ids.append(fields["id"][0])
        gr = fields["ranges"]
        start = np.asarray(gr["ranges"]["start"], dtype=np.int32)
        width = np.asarray(gr["ranges"]["width"], dtype=np.int32)
        strand = strand_codes(gr["strand"])
        seq, seq_levels = rle_expand(gr["seqnames"])
        assert len(set(seq.tolist())) == 1, "fixture expected on one chromosome"
        si = gr["seqinfo"]
        names = list(si["seqnames"])
        lens = np.asarray(si["seqlengths"], dtype=np.int64)
        chrom = seq_levels[int(seq[0]) - 1]
        clen = int(lens[names.index(chrom)])
        if chrom_names is None:
            chrom_names, chrom_len = [chrom], [clen]
        assert chrom_names == [chrom] and chrom_len == [clen]
        assert start.shape == width.shape == strand.shape
        out["reads_%d_start" % k] = start
        out["reads_%d_width" % k] = width
        out["reads_%d_strand" % k] = strand
    out["sample_ids"] = np.array(ids)
    out["chrom_names"] = np.array(chrom_names)
    out["chrom_len"] = np.array(chrom_len, dtype=np.int64)

    g = d["test.genome"]
    cols = dict(zip(rdx2.attr(g, "names"), rdx2.val(g)))
    chrom_f = cols["chromosome"]
    lv = rdx2.attr(chrom_f, "levels")
    assert all(lv[c - 1] == chrom_names[0] for c in rdx2.val(chrom_f))
    out["gene_start"] = np.asarray(rdx2.val(cols["start"]), dtype=np.int32)
    out["gene_end"] = np.asarray(rdx2.val(cols["end"]), dtype=np.int32)
    st = cols["strand"]
    st_lv = rdx2.attr(st, "levels")
    lut = {"+": 1, "-": -1, "*": 0}
    if st_lv is not None:
        out["gene_strand"] = np.array([lut[st_lv[c - 1]] for c in rdx2.val(st)], dtype=np.int8)
    else:
        out["gene_strand"] = np.array([lut[c] for c in rdx2.val(st)], dtype=np.int8)
    gn = cols["gene_name"]
    gn_lv = rdx2.attr(gn, "levels")
    gene_names = [gn_lv[c - 1] for c in rdx2.val(gn)] if gn_lv is not None else list(rdx2.val(gn))
    out["gene_names"] = np.array(gene_names)
    out["gene_row_names"] = np.array([str(x) for x in rdx2.attr(g, "row.names")])

    ex = d["test.exons"]
    ends = np.asarray(ex["partitioning"]["end"], dtype=np.int64)
    ptr = np.concatenate([[0], ends]).astype(np.int32)
    u = ex["unlistData"]
    es = np.asarray(u["ranges"]["start"], dtype=np.int32)
    ew = np.asarray(u["ranges"]["width"], dtype=np.int32)
    out["exon_ptr"] = ptr
    out["exon_start"] = es
    out["exon_end"] = (es + ew - 1).astype(np.int32)
    out["exon_strand"] = strand_codes(u["strand"])
    out["exon_gene_names"] = np.array(list(ex["partitioning"]["NAMES"]))
    seq, seq_levels = rle_expand(u["seqnames"])
    assert all(seq_levels[c - 1] == chrom_names[0] for c in seq)

    dst = os.path.join(HERE, "recoup_test_data.npz")
    np.savez_compressed(dst, **out)
    for k, v in out.items():
        print("%-22s %-8s %s" % (k, v.dtype, v.shape))
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
