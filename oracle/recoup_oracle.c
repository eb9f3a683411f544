/*
 * ORACLE (C twin) -- CPU restatement of recoup's coverage -> profile-matrix path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (medium-size parity), __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product never links or loads it.
 * PARITY UNPINNED by the reference (see oracle/recoup_oracle.py header): this file follows the
 * same R sources and is itself checked against the numpy restatement in tests/test_oracle_c.py.
 *
 * It deliberately uses a DIFFERENT algorithm from the CUDA library (start-sorted (start,end)
 * pairs + maximum read width, per-region difference array + cumsum -- the literal shape of
 * coverageFromRanges, /root/reference/R/coverage.R:176-226), so that agreement between the two
 * is evidence and not an echo.  Parallelism mirrors the reference's: an embarrassingly parallel
 * map over regions (cmclapply -> mclapply, R/util.R:364-382), here an OpenMP loop.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define STRAND_ANY 2

typedef struct {
    int64_t n;
    int n_chrom;
    int64_t* chrom_ptr;  /* n_chrom + 1: reads of chromosome c are [chrom_ptr[c], chrom_ptr[c+1]) */
    int32_t* start;      /* sorted by (chrom, start) */
    int32_t* end;
    int8_t* strand;
    int32_t* maxw;       /* per chromosome: maximum read width */
    int64_t* chrom_len;
} orc_index;

int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* frag_len > 0: trim(resize(reads, frag_len, fix="start")) -- absent from the reference
 * snapshot, named by the north star. */
orc_index* orc_index_build(int64_t n, const int32_t* chrom, const int32_t* start,
                           const int32_t* end, const int8_t* strand, int n_chrom,
                           const int64_t* chrom_len, int frag_len) {
    orc_index* ix = (orc_index*)calloc(1, sizeof(orc_index));
    ix->n = n;
    ix->n_chrom = n_chrom;
    ix->chrom_ptr = (int64_t*)calloc((size_t)n_chrom + 1, 8);
    ix->start = (int32_t*)malloc((size_t)(n ? n : 1) * 4);
    ix->end = (int32_t*)malloc((size_t)(n ? n : 1) * 4);
    ix->strand = (int8_t*)malloc((size_t)(n ? n : 1));
    ix->maxw = (int32_t*)calloc((size_t)n_chrom, 4);
    ix->chrom_len = (int64_t*)malloc((size_t)n_chrom * 8);
    memcpy(ix->chrom_len, chrom_len, (size_t)n_chrom * 8);
    /* key = chrom (16 bits) << 32 | start : 48 significant bits -> 3 passes of 16 */
    uint64_t* key = (uint64_t*)malloc((size_t)(n ? n : 1) * 8);
    uint32_t* idx = (uint32_t*)malloc((size_t)(n ? n : 1) * 4);
    int32_t* es = (int32_t*)malloc((size_t)(n ? n : 1) * 4);
    int32_t* ee = (int32_t*)malloc((size_t)(n ? n : 1) * 4);
    for (int64_t i = 0; i < n; i++) {
        int64_t s = start[i], e = end[i];
        const int st = strand ? strand[i] : 0;
        if (frag_len > 0) {
            if (st < 0) s = e - frag_len + 1; else e = s + frag_len - 1;
        }
        if (s < 1) s = 1;
        if (e > chrom_len[chrom[i]]) e = chrom_len[chrom[i]];
        es[i] = (int32_t)s;
        ee[i] = (int32_t)e;
        key[i] = ((uint64_t)(uint32_t)chrom[i] << 32) | (uint32_t)s;
        idx[i] = (uint32_t)i;
        ix->chrom_ptr[chrom[i] + 1]++;
    }
    for (int c = 0; c < n_chrom; c++) ix->chrom_ptr[c + 1] += ix->chrom_ptr[c];
    /* LSD radix sort of (key, index): 3 passes of 16 bits over the 48 significant bits */
    {
        uint64_t* k = key;
        uint32_t* id = idx;
        uint64_t* k2 = (uint64_t*)malloc((size_t)(n ? n : 1) * 8);
        uint32_t* i2 = (uint32_t*)malloc((size_t)(n ? n : 1) * 4);
        int64_t* cnt = (int64_t*)malloc(65537 * sizeof(int64_t));
        for (int pass = 0; pass < 3; pass++) {
            const int sh = pass * 16;
            memset(cnt, 0, 65537 * sizeof(int64_t));
            for (int64_t i = 0; i < n; i++) cnt[((k[i] >> sh) & 0xffff) + 1]++;
            for (int b = 0; b < 65536; b++) cnt[b + 1] += cnt[b];
            for (int64_t i = 0; i < n; i++) {
                const int64_t o = cnt[(k[i] >> sh) & 0xffff]++;
                k2[o] = k[i];
                i2[o] = id[i];
            }
            uint64_t* tk = k; k = k2; k2 = tk;
            uint32_t* ti = id; id = i2; i2 = ti;
        }
        for (int64_t i = 0; i < n; i++) {
            const uint32_t j = id[i];
            ix->start[i] = es[j];
            ix->end[i] = ee[j];
            ix->strand[i] = strand ? (int8_t)(strand[j] > 0 ? 1 : (strand[j] < 0 ? -1 : 0)) : 0;
            const int c = chrom[j];
            const int32_t w = ee[j] - es[j] + 1;
            if (w > ix->maxw[c]) ix->maxw[c] = w;
        }
        free(k);
        free(k2);
        free(id);
        free(i2);
        free(cnt);
    }
    free(es);
    free(ee);
    return ix;
}

void orc_index_free(orc_index* ix) {
    if (!ix) return;
    free(ix->chrom_ptr);
    free(ix->start);
    free(ix->end);
    free(ix->strand);
    free(ix->maxw);
    free(ix->chrom_len);
    free(ix);
}

/* first i in [lo, hi) with a[i] >= key */
static int64_t lower_bound(const int32_t* a, int64_t lo, int64_t hi, int64_t key) {
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if ((int64_t)a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static int strand_ok(int read_strand, int range_strand, int ignore_strand, int strand_filter) {
    if (strand_filter != STRAND_ANY && read_strand != strand_filter) return 0;   /* coverage.R:141-144 */
    if (ignore_strand || range_strand == 0 || read_strand == 0) return 1;        /* coverage.R:191 */
    return read_strand == range_strand;
}

/*
 * calcCoverage over a GRanges mask (coverage.R:126-226).  Region r writes its vector at
 * cov + off[r]; off[] is given by the caller (any layout with room for end-start+1 ints per
 * region).  len_out[r] = length written (0 for NULL), is_null[r] = 1 for NULL.
 */
int orc_coverage(const orc_index* ix, int64_t R, const int32_t* rchrom, const int32_t* rstart,
                 const int32_t* rend, const int8_t* rstrand, int ignore_strand, int strand_filter,
                 const int64_t* off, int32_t* cov, int32_t* len_out, uint8_t* is_null) {
#pragma omp parallel
    {
        int32_t* diff = NULL;
        int64_t diff_cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < R; r++) {
            const int c = rchrom[r];
            int64_t s = rstart[r], e = rend[r];
            const int st = rstrand ? rstrand[r] : 0;
            len_out[r] = 0;
            is_null[r] = 1;
            if (c < 0 || c >= ix->n_chrom) continue;              /* coverage.R:194-197 */
            if (e < s) continue;
            const int64_t lo = ix->chrom_ptr[c], hi = ix->chrom_ptr[c + 1];
            const int64_t a = lower_bound(ix->start, lo, hi, s - ix->maxw[c] + 1);
            const int64_t b = lower_bound(ix->start, lo, hi, e + 1);
            int64_t nov = 0;
            for (int64_t i = a; i < b; i++)
                if (ix->end[i] >= s && strand_ok(ix->strand[i], st, ignore_strand, strand_filter)) nov++;
            if (nov == 0) continue;                               /* coverage.R:198,224-225 */
            if (s < 0 || e > ix->chrom_len[c]) continue;          /* tryCatch -> NULL, coverage.R:217-222 */
            if (s == 0) s = 1;                                    /* zero subscript dropped */
            const int64_t L = e - s + 1;
            if (L <= 0) continue;
            if (L + 1 > diff_cap) {
                diff_cap = L + 1;
                diff = (int32_t*)realloc(diff, (size_t)diff_cap * 4);
            }
            memset(diff, 0, (size_t)(L + 1) * 4);
            for (int64_t i = a; i < b; i++) {
                if (ix->end[i] < s || !strand_ok(ix->strand[i], st, ignore_strand, strand_filter)) continue;
                const int64_t p0 = ix->start[i] > s ? ix->start[i] - s : 0;
                const int64_t p1 = ix->end[i] < e ? ix->end[i] - s + 1 : L;
                diff[p0] += 1;
                diff[p1] -= 1;
            }
            int32_t* dst = cov + off[r];
            int32_t run = 0;
            if (st < 0) {                                         /* coverage.R:212-213 */
                for (int64_t k = 0; k < L; k++) { run += diff[k]; dst[L - 1 - k] = run; }
            } else {
                for (int64_t k = 0; k < L; k++) { run += diff[k]; dst[k] = run; }
            }
            len_out[r] = (int32_t)L;
            is_null[r] = 0;
        }
        free(diff);
    }
    return 0;
}

/*
 * calcCoverage over a GRangesList mask (coverage.R:177-178,202-207): element g owns ranges
 * ptr[g]..ptr[g+1]-1; chromosome/strand of the FIRST range; a read overlapping k ranges counts
 * k times; the vector is stitched in list order.
 */
int orc_coverage_list(const orc_index* ix, int64_t G, const int64_t* ptr, const int32_t* xchrom,
                      const int32_t* xstart, const int32_t* xend, const int8_t* xstrand,
                      int ignore_strand, int strand_filter, const int64_t* off, int32_t* cov,
                      int32_t* len_out, uint8_t* is_null) {
#pragma omp parallel
    {
        int32_t* diff = NULL;
        int64_t diff_cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int64_t g = 0; g < G; g++) {
            const int64_t qa = ptr[g], qb = ptr[g + 1];
            len_out[g] = 0;
            is_null[g] = 1;
            if (qb <= qa) continue;
            const int c = xchrom[qa];
            const int st0 = xstrand ? xstrand[qa] : 0;
            if (c < 0 || c >= ix->n_chrom) continue;
            int64_t span_lo = INT64_MAX, span_hi = INT64_MIN, L = 0;
            int oob = 0;
            for (int64_t q = qa; q < qb; q++) {
                int64_t s = xstart[q], e = xend[q];
                if (s < 0 || e > ix->chrom_len[c]) oob = 1;
                if (s == 0) s = 1;
                if (e >= s) {
                    L += e - s + 1;
                    if (s < span_lo) span_lo = s;
                    if (e > span_hi) span_hi = e;
                }
            }
            if (L == 0) continue;
            const int64_t lo = ix->chrom_ptr[c], hi = ix->chrom_ptr[c + 1];
            const int64_t a = lower_bound(ix->start, lo, hi, span_lo - ix->maxw[c] + 1);
            const int64_t b = lower_bound(ix->start, lo, hi, span_hi + 1);
            if (L + 1 > diff_cap) {
                diff_cap = L + 1;
                diff = (int32_t*)realloc(diff, (size_t)diff_cap * 4);
            }
            memset(diff, 0, (size_t)(L + 1) * 4);
            int64_t hits = 0;
            for (int64_t i = a; i < b; i++) {
                const int64_t rs = ix->start[i], re = ix->end[i];
                if (re < span_lo) continue;
                int mult = 0;
                for (int64_t q = qa; q < qb; q++) {
                    int64_t s = xstart[q], e = xend[q];
                    if (s == 0) s = 1;
                    if (e < s) continue;
                    if (rs <= e && re >= s &&
                        strand_ok(ix->strand[i], xstrand ? xstrand[q] : 0, ignore_strand, strand_filter))
                        mult++;
                }
                if (!mult) continue;
                hits += mult;
                /* the selected reads (with multiplicity) cover the whole chromosome vector; the
                 * stitched index picks every range's bases, whether or not that range was hit */
                int64_t xo = 0;
                for (int64_t q = qa; q < qb; q++) {
                    int64_t s = xstart[q], e = xend[q];
                    if (s == 0) s = 1;
                    if (e < s) continue;
                    if (rs <= e && re >= s) {
                        const int64_t p0 = xo + (rs > s ? rs - s : 0);
                        const int64_t p1 = xo + (re < e ? re - s + 1 : e - s + 1);
                        diff[p0] += mult;
                        diff[p1] -= mult;
                    }
                    xo += e - s + 1;
                }
            }
            if (hits == 0 || oob) continue;
            /* every (read, range) interval is clipped to its range's segment, so one cumulative
             * sum over the stitched vector is exact */
            int32_t* dst = cov + off[g];
            int32_t run = 0;
            for (int64_t k = 0; k < L; k++) {
                run += diff[k];
                if (st0 < 0) dst[L - 1 - k] = run; else dst[k] = run;
            }
            len_out[g] = (int32_t)L;
            is_null[g] = 0;
        }
        free(diff);
    }
    return 0;
}

/*
 * binCoverageMatrix for segments with length >= n (profile.R:153-212, util.R:74-85); rows whose
 * segment is shorter than n (interpolation, util.R:17-73) are flagged in short_out and left to
 * the numpy restatement.  where: 0 whole, 1 center, 2 upstream, 3 downstream.  stat: 0 mean,
 * 1 median.  rank = seed-42 rank table of n.  out is column-major R x n with leading dim ld.
 */
static int cmp_i32(const void* a, const void* b) {
    const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

int orc_bin_matrix(int64_t R, const int32_t* cov, const int64_t* off, const int32_t* len,
                   const uint8_t* is_null, int where, int f1, int f2, int n, const int32_t* rank,
                   int stat, double scale, double* out, int64_t ld, uint8_t* short_out) {
#pragma omp parallel
    {
        int32_t* tmp = NULL;
        int64_t tmp_cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < R; r++) {
            short_out[r] = 0;
            if (is_null[r]) {                                     /* profile.R:191-197 */
                for (int i = 0; i < n; i++) out[(int64_t)i * ld + r] = 0.0;
                continue;
            }
            const int64_t L = len[r];
            int64_t a = 0, b = L;
            if (where == 1) { a = f1; b = L - f2; }
            else if (where == 2) { a = 0; b = f1; }
            else if (where == 3) { a = L - f2; b = L; }
            if (a < 0) a = 0;
            if (b > L) b = L;
            const int64_t Ls = b - a;
            if (Ls < n) { short_out[r] = 1; continue; }
            const int64_t bsz = Ls / n, dif = Ls - bsz * n;
            const int32_t* x = cov + off[r] + a;
            int64_t pos = 0;
            for (int i = 0; i < n; i++) {
                const int64_t cnt = bsz + (rank[i] <= dif ? 1 : 0);
                double v;
                if (stat == 0) {
                    int64_t s = 0;
                    for (int64_t k = 0; k < cnt; k++) s += x[pos + k];
                    v = (double)s / (double)cnt;
                } else {
                    if (cnt > tmp_cap) { tmp_cap = cnt; tmp = (int32_t*)realloc(tmp, (size_t)cnt * 4); }
                    memcpy(tmp, x + pos, (size_t)cnt * 4);
                    qsort(tmp, (size_t)cnt, 4, cmp_i32);
                    v = 0.5 * ((double)tmp[(cnt - 1) / 2] + (double)tmp[cnt / 2]);
                }
                out[(int64_t)i * ld + r] = scale * v;
                pos += cnt;
            }
        }
        free(tmp);
    }
    return 0;
}

/* baseCoverageMatrix (profile.R:100-151) */
int orc_base_matrix(int64_t R, const int32_t* cov, const int64_t* off, const int32_t* len,
                    const uint8_t* is_null, int where, int f1, int f2, int64_t n_cols, double scale,
                    double* out, int64_t ld) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < R; r++) {
        const int64_t L = len[r];
        int64_t a = 0, b = L;
        if (where == 2) { a = 0; b = f1; }
        else if (where == 3) { a = L - f2; b = L; }
        if (a < 0) a = 0;
        if (b > L) b = L;
        for (int64_t k = 0; k < n_cols; k++) {
            double v = 0.0;
            if (!is_null[r] && a + k < b) v = scale * (double)cov[off[r] + a + k];
            out[k * ld + r] = v;
        }
    }
    return 0;
}
