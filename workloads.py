"""Synthetic workloads of BASELINE.json's configs: seeded generators that
bench.py and the parity tests share.  Pure numpy; nothing here touches the product or the oracle.

Every generator takes `scale` (0 < scale <= 1) shrinking reads AND regions by the same factor, so
tests can run the named shapes at sizes the oracle finishes in seconds.
"""
import numpy as np

HG19 = [("chr1", 249250621), ("chr2", 243199373), ("chr3", 198022430), ("chr4", 191154276),
        ("chr5", 180915260), ("chr6", 171115067), ("chr7", 159138663), ("chr8", 146364022),
        ("chr9", 141213431), ("chr10", 135534747), ("chr11", 135006516), ("chr12", 133851895),
        ("chr13", 115169878), ("chr14", 107349540), ("chr15", 102531392), ("chr16", 90354753),
        ("chr17", 81195210), ("chr18", 78077248), ("chr19", 59128983), ("chr20", 63025520),
        ("chr21", 48129895), ("chr22", 51304566), ("chrX", 155270560), ("chrY", 59373566)]
HG19_NAMES = [n for n, _ in HG19]
HG19_LEN = np.array([l for _, l in HG19], dtype=np.int64)


def _uniform_positions(rng, n, chrom_len):
    """n positions uniform over the genome: (chrom id, 1-based position)."""
    cum = np.concatenate([[0], np.cumsum(chrom_len)])
    g = (rng.random(n) * cum[-1]).astype(np.int64)
    chrom = np.searchsorted(cum, g, side="right") - 1
    pos = g - cum[chrom] + 1
    return chrom.astype(np.int32), pos


def chipseq_tss(scale=1.0, seed=1001, n_reads=50_000_000, n_regions=20_000, flank=5000,
                read_len=36, frag_len=200, chrom_len=HG19_LEN):
    """C2: single-end ChIP-seq reads (36 bp, extended to fragLen 200 on load) over TSS +-flank.
    70 % uniform background, 30 % Gaussian (sigma 300) around TSSs; 0.1 % of the TSSs sit at a
    chromosome edge so their window leaves the chromosome (NULL rule); reads in random order."""
    rng = np.random.default_rng(seed)
    N = max(int(n_reads * scale), 1000)
    R = max(int(n_regions * scale), 50)
    rchrom, tss = _uniform_positions(rng, R, chrom_len)
    tss = np.clip(tss, flank + 1, chrom_len[rchrom] - flank)
    n_edge = max(R // 1000, 1)
    edge = rng.choice(R, size=n_edge, replace=False)
    tss[edge[::2]] = rng.integers(1, flank, size=edge[::2].shape[0])
    tss[edge[1::2]] = chrom_len[rchrom[edge[1::2]]] - rng.integers(0, flank - 1, size=edge[1::2].shape[0])
    rstrand = rng.choice(np.array([1, -1], dtype=np.int8), size=R)
    n_peak = int(0.3 * N)
    chrom_bg, pos_bg = _uniform_positions(rng, N - n_peak, chrom_len)
    which = rng.integers(0, R, size=n_peak)
    pos_pk = tss[which] + np.rint(rng.normal(0.0, 300.0, size=n_peak)).astype(np.int64)
    chrom = np.concatenate([chrom_bg, rchrom[which]])
    pos = np.concatenate([pos_bg, pos_pk])
    clen = chrom_len[chrom]
    start = np.clip(pos, 1, clen - read_len + 1)
    perm = rng.permutation(N)
    chrom, start = chrom[perm], start[perm]
    strand = rng.choice(np.array([1, -1], dtype=np.int8), size=N)
    return dict(
        name="C2 synthetic ChIP-seq: %d single-end reads (fragLen %d) over %d hg19 TSS +-%d bp"
             % (N, frag_len, R, flank),
        chrom_names=HG19_NAMES, chrom_len=np.asarray(chrom_len, dtype=np.int64),
        read_chrom=chrom.astype(np.int32), read_start=start.astype(np.int32),
        read_end=(start + read_len - 1).astype(np.int32), read_strand=strand, frag_len=frag_len,
        region_chrom=rchrom, region_start=tss.astype(np.int32), region_end=tss.astype(np.int32),
        region_strand=rstrand, region="tss", flank=(flank, flank),
        bin_params=dict(flankBinSize=0, regionBinSize=200, sumStat="mean", interpolation="auto"))


def gene_bodies(scale=1.0, seed=1003, n_reads=30_000_000, n_regions=60_000, flank=2000,
                read_len=200, chrom_len=HG19_LEN):
    """C3 (one sample): reads over gene bodies with heavy-tailed lengths + 2 kb flanks,
    50 + 150 + 50 bins (lengths almost never divide the bin count: seeded bin layout)."""
    rng = np.random.default_rng(seed)
    N = max(int(n_reads * scale), 1000)
    R = max(int(n_regions * scale), 50)
    glen = np.clip(np.exp(rng.normal(np.log(2e4), 1.3, size=R)), 60, 2e6).astype(np.int64)
    rchrom, gs = _uniform_positions(rng, R, chrom_len)
    gs = np.clip(gs, flank + 1, np.maximum(chrom_len[rchrom] - glen - flank, flank + 1))
    ge = np.minimum(gs + glen - 1, chrom_len[rchrom] - flank)
    rstrand = rng.choice(np.array([1, -1], dtype=np.int8), size=R)
    chrom, pos = _uniform_positions(rng, N, chrom_len)
    start = np.clip(pos, 1, chrom_len[chrom] - read_len + 1)
    strand = rng.choice(np.array([1, -1], dtype=np.int8), size=N)
    return dict(
        name="C3 synthetic gene bodies: %d reads over %d genes +-%d bp, 50+150+50 bins" % (N, R, flank),
        chrom_names=HG19_NAMES, chrom_len=np.asarray(chrom_len, dtype=np.int64),
        read_chrom=chrom, read_start=start.astype(np.int32),
        read_end=(start + read_len - 1).astype(np.int32), read_strand=strand, frag_len=0,
        region_chrom=rchrom, region_start=gs.astype(np.int32), region_end=ge.astype(np.int32),
        region_strand=rstrand, region="genebody", flank=(flank, flank),
        bin_params=dict(flankBinSize=50, regionBinSize=150, sumStat="mean", interpolation="auto"))


def rnaseq(scale=1.0, seed=1004, n_reads=100_000_000, n_genes=20_000, exons_per_gene=10,
           flank=1000, read_len=100, chrom_len=HG19_LEN):
    """C4: spliced reads over merged exons (10 per gene, log-normal exon/intron lengths).
    30 % of reads span a junction and arrive split into two blocks; 2 % are unspliced ranges
    covering >= 2 exons (multiplicity quirk, coverage.R:190-192); 5 % fall outside exons."""
    rng = np.random.default_rng(seed)
    N = max(int(n_reads * scale), 1000)
    G = max(int(n_genes * scale), 20)
    E = exons_per_gene
    exon_len = np.clip(np.exp(rng.normal(np.log(150), 0.6, size=(G, E))), 20, 5000).astype(np.int64)
    intron_len = np.clip(np.exp(rng.normal(np.log(1500), 1.0, size=(G, E - 1))), 50, 100000).astype(np.int64)
    span = exon_len.sum(1) + intron_len.sum(1)
    gchrom, gs = _uniform_positions(rng, G, chrom_len)
    gs = np.clip(gs, flank + 2, np.maximum(chrom_len[gchrom] - span - flank - 2, flank + 2))
    rel_start = np.zeros((G, E), dtype=np.int64)
    rel_start[:, 1:] = np.cumsum(exon_len[:, :-1] + intron_len, axis=1)
    ex_start = gs[:, None] + rel_start
    ex_end = ex_start + exon_len - 1
    gstrand = rng.choice(np.array([1, -1], dtype=np.int8), size=G)
    tx_off = np.concatenate([np.zeros((G, 1), dtype=np.int64), np.cumsum(exon_len, axis=1)], axis=1)
    tx_len = tx_off[:, -1]

    def tx_to_genome(g, t):
        """transcript coordinate t (0-based) of gene g -> genomic position"""
        j = (tx_off[g, 1:] <= t[:, None]).sum(1)
        j = np.minimum(j, E - 1)
        return ex_start[g, j] + (t - tx_off[g, j]), j

    g_of = rng.integers(0, G, size=N)
    t0 = (rng.random(N) * np.maximum(tx_len[g_of] - read_len, 1)).astype(np.int64)
    t1 = np.minimum(t0 + read_len - 1, tx_len[g_of] - 1)
    p0, j0 = tx_to_genome(g_of, t0)
    p1, j1 = tx_to_genome(g_of, t1)
    kind = rng.random(N)
    unspliced = (kind < 0.02)
    outside = (kind >= 0.02) & (kind < 0.07)
    spliced = (j1 != j0) & ~unspliced & ~outside
    chrom_r = gchrom[g_of]
    # block 1 / block 2 of spliced reads (split at the first junction; deeper junctions keep the
    # rest of the read in block 2's exon span as one range, like a 2-block GAlignments)
    b1s, b1e = p0.copy(), np.where(spliced, ex_end[g_of, j0], p1)
    b2s, b2e = ex_start[g_of, np.minimum(j0 + 1, E - 1)], p1
    keep2 = spliced & (b2e >= b2s)
    # unspliced: one range from p0 to the matching position one exon further (covers the intron)
    nxt = np.minimum(j0 + 1, E - 1)
    b1e = np.where(unspliced, np.maximum(ex_start[g_of, nxt] + 10, b1s + read_len - 1), b1e)
    # outside: the 5 % land in the two flanks (alternating), so the flank coverages are not NULL
    up = (t0 & 1) == 0
    ge_gene = ex_end[g_of, E - 1]
    pos_out = np.where(up, gs[g_of] - 1 - (t0 % flank), ge_gene + 1 + (t0 % flank))
    b1s = np.where(outside, np.maximum(pos_out - read_len // 2, 1), b1s)
    b1e = np.where(outside, b1s + read_len - 1, b1e)
    b1e = np.minimum(b1e, chrom_len[chrom_r])
    chrom = np.concatenate([chrom_r, chrom_r[keep2]])
    start = np.concatenate([b1s, b2s[keep2]])
    end = np.concatenate([b1e, b2e[keep2]])
    end = np.maximum(end, start)
    st = gstrand[g_of]
    strand = np.concatenate([st, st[keep2]])
    perm = rng.permutation(chrom.shape[0])
    ptr = np.arange(0, G * E + 1, E, dtype=np.int64)
    return dict(
        name="C4 synthetic RNA-seq: %d spliced reads (%d ranges) over %d exons stitched to %d genes"
             % (N, chrom.shape[0], G * E, G),
        chrom_names=HG19_NAMES, chrom_len=np.asarray(chrom_len, dtype=np.int64),
        read_chrom=chrom[perm].astype(np.int32), read_start=start[perm].astype(np.int32),
        read_end=end[perm].astype(np.int32), read_strand=strand[perm].astype(np.int8), frag_len=0,
        exon_ptr=ptr, exon_chrom=np.repeat(gchrom, E).astype(np.int32),
        exon_start=ex_start.reshape(-1).astype(np.int32), exon_end=ex_end.reshape(-1).astype(np.int32),
        exon_strand=np.repeat(gstrand, E),
        region_chrom=gchrom, region_start=gs.astype(np.int32),
        region_end=ex_end[:, -1].astype(np.int32), region_strand=gstrand, region="rna",
        flank=(flank, flank),
        bin_params=dict(flankBinSize=25, regionBinSize=50, sumStat="mean", interpolation="auto"))


def dnase_sites(scale=1.0, seed=1005, n_reads=200_000_000, n_regions=1_000_000, flank=500,
                read_len=50, chrom_len=HG19_LEN):
    """C5: width-1 motif sites +-500 bp at per-base resolution; half the reads cluster within
    +-100 bp of a site."""
    rng = np.random.default_rng(seed)
    N = max(int(n_reads * scale), 1000)
    R = max(int(n_regions * scale), 50)
    rchrom, site = _uniform_positions(rng, R, chrom_len)
    site = np.clip(site, flank + 1, chrom_len[rchrom] - flank)
    rstrand = rng.choice(np.array([1, -1, 0], dtype=np.int8), size=R)
    n_cl = N // 2
    chrom_bg, pos_bg = _uniform_positions(rng, N - n_cl, chrom_len)
    which = rng.integers(0, R, size=n_cl)
    pos_cl = site[which] + rng.integers(-100, 101, size=n_cl)
    chrom = np.concatenate([chrom_bg, rchrom[which]])
    pos = np.concatenate([pos_bg, pos_cl])
    start = np.clip(pos, 1, chrom_len[chrom] - read_len + 1)
    perm = rng.permutation(N)
    strand = rng.choice(np.array([1, -1], dtype=np.int8), size=N)
    return dict(
        name="C5 synthetic DNase-seq: %d reads over %d sites +-%d bp, per-base" % (N, R, flank),
        chrom_names=HG19_NAMES, chrom_len=np.asarray(chrom_len, dtype=np.int64),
        read_chrom=chrom[perm].astype(np.int32), read_start=start[perm].astype(np.int32),
        read_end=(start[perm] + read_len - 1).astype(np.int32), read_strand=strand, frag_len=0,
        region_chrom=rchrom, region_start=site.astype(np.int32), region_end=site.astype(np.int32),
        region_strand=rstrand, region="custom", flank=(flank, flank),
        bin_params=dict(flankBinSize=0, regionBinSize=0, sumStat="mean", interpolation="auto"))


def dnase_sites_part(part, n_parts=8, scale=1.0, seed=1005, n_reads=200_000_000, n_regions=1_000_000,
                     flank=500, read_len=50, chrom_len=HG19_LEN):
    """C5 for the strong-scaling run: the SAME sites as dnase_sites (same seed), the reads drawn in
    `n_parts` independent parts so that W ranks can each generate the parts {p : p % W == rank}
    and the union is the same read set for every W (checksums comparable across W)."""
    rng = np.random.default_rng(seed)
    N = max(int(n_reads * scale), 1000) // n_parts
    R = max(int(n_regions * scale), 50)
    rchrom, site = _uniform_positions(rng, R, chrom_len)
    site = np.clip(site, flank + 1, chrom_len[rchrom] - flank)
    rstrand = rng.choice(np.array([1, -1, 0], dtype=np.int8), size=R)
    prng = np.random.default_rng([seed, part + 1])
    n_cl = N // 2
    chrom_bg, pos_bg = _uniform_positions(prng, N - n_cl, chrom_len)
    which = prng.integers(0, R, size=n_cl)
    pos_cl = site[which] + prng.integers(-100, 101, size=n_cl)
    chrom = np.concatenate([chrom_bg, rchrom[which]])
    pos = np.concatenate([pos_bg, pos_cl])
    start = np.clip(pos, 1, chrom_len[chrom] - read_len + 1)
    perm = prng.permutation(N)
    strand = prng.choice(np.array([1, -1], dtype=np.int8), size=N)
    return dict(
        name="C5 synthetic DNase-seq: %d reads (part %d of %d) over %d sites +-%d bp, per-base"
             % (N, part, n_parts, R, flank),
        chrom_names=HG19_NAMES, chrom_len=np.asarray(chrom_len, dtype=np.int64),
        read_chrom=chrom[perm].astype(np.int32), read_start=start[perm].astype(np.int32),
        read_end=(start[perm] + read_len - 1).astype(np.int32), read_strand=strand, frag_len=0,
        region_chrom=rchrom, region_start=site.astype(np.int32), region_end=site.astype(np.int32),
        region_strand=rstrand, region="custom", flank=(flank, flank),
        bin_params=dict(flankBinSize=0, regionBinSize=0, sumStat="mean", interpolation="auto"))


CONFIGS = {"C2": chipseq_tss, "C3": gene_bodies, "C4": rnaseq, "C5": dnase_sites}
