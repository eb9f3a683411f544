/*
 * recoup_b200.h -- plain C ABI of librecoup_b200.so
 *
 * B200-native (sm_100a) implementation of recoup's coverage -> profile-matrix hot path.
 * The reference (hjanime/recoup, /root/reference) is interpreted R with NO native interface
 * (NAMESPACE:1-33 has no useDynLib); the path sits behind exported R closures.  Each entry point
 * below names the reference closure (file:line under /root/reference) whose body it replaces.
 * The R-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - coordinates 1-based closed [start,end]; strand +1 '+', -1 '-', 0 '*'.
 *   - chromosomes are dense ids 0..n_chrom-1 (as.integer(seqnames)-1).
 *   - every function returns RCP_OK (0) or an RCP_ERR_* code; rcp_last_error() gives the text.
 *     Nothing longjmps / calls back into the host language.  Per-region failures of the
 *     reference ("Caught invalid genomic area!", coverage.R:217-222) are DATA (a NULL coverage),
 *     never errors.
 *   - the caller owns every pointer it passes; outputs are caller-allocated.  The library owns
 *     device memory behind integer handles (reads / coverage), released by the *_free calls or
 *     rcp_shutdown().
 *   - `mem` says where the caller's arrays live: RCP_MEM_HOST (pageable or pinned host memory)
 *     or RCP_MEM_DEVICE (pointers valid on the bound GPU; no copies are made).
 *   - one process drives one GPU (rcp_init(device)); calls are serialised by the caller.
 *   - there is NO CPU fallback: without a usable sm_100 GPU every compute call fails with
 *     RCP_ERR_NOGPU.
 */
#ifndef RECOUP_B200_H
#define RECOUP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCP_ABI_VERSION 1

enum {
    RCP_OK = 0,
    RCP_ERR_CUDA = 1,        /* a CUDA runtime call failed                      */
    RCP_ERR_ARG = 2,         /* bad argument (R's stop() in the reference)      */
    RCP_ERR_HANDLE = 3,      /* unknown / freed handle                          */
    RCP_ERR_NOGPU = 4,       /* no CUDA device, or rcp_init() not called        */
    RCP_ERR_UNSUPPORTED = 5, /* documented unsupported corner   */
    RCP_ERR_DATA = 6         /* input arrays violate the stated contract        */
};

#define RCP_MEM_HOST 0
#define RCP_MEM_DEVICE 1

/* `strand` argument of calcCoverage (coverage.R:126,141-144): NULL or one of "+","-","*" */
#define RCP_STRAND_ANY 2

/* `where` of binCoverageMatrix/baseCoverageMatrix (profile.R:100-101,153-155) */
#define RCP_WHERE_WHOLE 0      /* flank = NULL */
#define RCP_WHERE_CENTER 1
#define RCP_WHERE_UPSTREAM 2
#define RCP_WHERE_DOWNSTREAM 3

#define RCP_STAT_MEAN 0
#define RCP_STAT_MEDIAN 1

#define RCP_INTERP_AUTO 0
#define RCP_INTERP_SPLINE 1
#define RCP_INTERP_LINEAR 2          /* dead code in the reference (util.R:49) -> UNSUPPORTED */
#define RCP_INTERP_NEIGHBORHOOD 3

#define RCP_SAMPLE_REJECTION 0       /* R >= 3.6.0 sample.kind (default) */
#define RCP_SAMPLE_ROUNDING 1        /* R <  3.6.0                       */

/* ---------------------------------------------------------------- library / device -------- */
const char* rcp_last_error(void);
int rcp_abi_version(void);
/* Bind this process to CUDA device `device` (one process per GPU). Replaces the reference's
 * "grid launch" cmclapply (util.R:364-382): `rc` is accepted and ignored by the R shim. */
int rcp_init(int device);
int rcp_shutdown(void);
int rcp_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, int64_t* mem_bytes);
/* cudaStream_t on which every kernel of the library is launched (for external event timing). */
void* rcp_stream(void);
int rcp_sync(void);
/* Kernels launched by this library since rcp_init / the last reset (bench.py "gpu_launches"). */
int64_t rcp_launch_count(int reset);

/* Page-locked host buffers for outputs (profile matrices): device->host copies into them run at
 * full PCIe rate.  Released buffers are kept by the library and handed out again on an
 * exact-size match, so a loop over samples pays the page-locking cost once. */
int rcp_host_alloc(int64_t bytes, void** ptr_out);
int rcp_host_free(void* ptr);

/* Optional CUDA-event timing of the library's stages on its own stream (bench.py roofline).
 * rcp_timing_read synchronises, then reports accumulated milliseconds / launches per stage
 * (stage i is named rcp_timing_stage_name(i); NULL past the last stage). */
int rcp_timing_enable(int on);
int rcp_timing_read(int reset, int capacity, double* ms_out, int64_t* count_out);
const char* rcp_timing_stage_name(int stage);

/* How rcp_coverage (GRanges masks) finds the reads of each region.  All give identical results.
 *   RCP_PATH_SPLIT    ONE streaming pass over the unsorted reads: a block bitmap of the mask in
 *                     shared memory filters them, the survivors are packed into 32-bit words and
 *                     split into <= 1024 groups of the mask's blocks through shared-memory rings
 *                     (64-byte chunks, no histogram pass, no prefix sum over the reads); each
 *                     group is then sorted by 2-kb sub-bin and every output tile reads the
 *                     sub-bins under it.  The host synchronises once, right after the plan.
 *                     Needs reads narrower than the packed word allows (<= 8191 bp, less for
 *                     dense masks / stranded calls); otherwise the call uses the paths below;
 *   RCP_PATH_BUCKETS  two passes over the unsorted reads drop each read into the buckets of the
 *                     output tiles it overlaps (cell table + hit list, no sort);
 *   RCP_PATH_BLOCKS   the reads that pass a block bitmap are partitioned by 16-kb genome block
 *                     (two multisplit passes, no per-read random access) and every output tile
 *                     scans the candidates of the blocks under it;
 *   RCP_PATH_INDEX    the reads are radix-sorted once per handle and every region is served by
 *                     rank searches (cheaper when one handle serves many masks);
 *   RCP_PATH_AUTO     (default) the index when the handle already has one; else split when the
 *                     reads fit its packed word; else the index when the mask is dense and the
 *                     reads many, blocks for masks made of tiled regions (> 1024 bp: TSS windows,
 *                     gene bodies), buckets for masks dominated by short regions.
 * The reference has one path (findOverlaps per region, coverage.R:189-193); this is a tuning
 * knob with no counterpart there. */
#define RCP_PATH_AUTO 0
#define RCP_PATH_INDEX 1
#define RCP_PATH_BUCKETS 2
#define RCP_PATH_BLOCKS 3
#define RCP_PATH_SPLIT 4
int rcp_set_coverage_path(int path);
/* Which path produced a coverage (RCP_PATH_INDEX / BUCKETS / BLOCKS; 0 for GRangesList and
 * concatenated coverages) and, for the block path, how many reads passed its bitmap filter
 * (bench.py derives that path's algorithmic bytes from it). */
int rcp_coverage_path_info(int cov, int* path, int64_t* candidates);

/* Deferred validation of the reads (default off).  When on, rcp_reads_load* return as soon as
 * their work is enqueued -- no host synchronisation -- and a data error of the reads (bad
 * chromosome id, start < 1, ...) is reported by the FIRST call that uses the reads handle
 * (rcp_coverage, rcp_coverage_list) instead of by the load itself.  Host arrays passed to a
 * deferred load must stay unchanged until that call returns.  The reference validates inside
 * GRanges construction; which call raises is the only difference. */
int rcp_set_deferred_validation(int on);

/* ---------------------------------------------------------------- base-R RNG -------------- */
/* `set.seed(seed); sample(1:n, k)` -- the bin layout of splitVector (util.R:78-79). */
int rcp_r_sample(int n, int k, int seed, int sample_kind, int* out /* k */);
/* rank[i] = position (1-based) of bin i+1 in `set.seed(seed); sample(1:n, n)`; bin i gets an
 * extra base iff rank[i] <= dif (util.R:74-80). */
int rcp_r_rank_table(int n, int seed, int sample_kind, int* rank_out /* n */);

/* ---------------------------------------------------------------- reads ------------------- */
/* Upload decoded reads (the `ranges` GRanges produced by preprocessRanges, ranges.R:1-65),
 * validate them and map them to global coordinates on the device (the sorted index is built
 * lazily by the first call that needs it).  `strand` may be NULL (all '*').
 * chrom_len[c] is the seqlength; <= 0 means unknown (R's NA): coverage(reads)[[chr]] then ends at
 * the largest end among the reads that overlap a region (coverage.R:201), i.e. a window on such a
 * chromosome is NULL unless a read reaches its last position (the load then takes one extra pass
 * over the reads and one host synchronisation).  Reads must satisfy 1 <= start <= end and, after
 * the optional extension, are trimmed into [1, chrom_len] like readBam's trim() (ranges.R:117;
 * nothing to trim at on a chromosome of unknown length).  frag_len > 0 applies the fragment extension named by the north star
 * (`trim(resize(reads, frag_len, fix="start"))`; absent from the reference snapshot), 0 = off. */
int rcp_reads_load(int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end,
                   const int8_t* strand, int n_chrom, const int64_t* chrom_len, int frag_len,
                   int mem, int* reads_out);
/* The same with the chromosome ids as runs -- the form a GRanges holds them in (`seqnames(x)` is
 * a factor-Rle: run_chrom = as.integer(runValue(seqnames(x))) - 1L, run_len = runLength(...));
 * coordinate-sorted reads have one run per chromosome, so 4 of the 13 bytes per read never cross
 * PCIe.  The run lengths must be >= 0 and sum to n (< 2^32).  Same result as rcp_reads_load on the
 * expanded ids. */
int rcp_reads_load_rle(int64_t n, int64_t n_runs, const int32_t* run_chrom, const int32_t* run_len,
                       const int32_t* start, const int32_t* end, const int8_t* strand, int n_chrom,
                       const int64_t* chrom_len, int frag_len, int mem, int* reads_out);
/* The same for a library whose reads all have ONE width -- fixed-length single-end reads, or a
 * GRanges after `resize(x, w)`: an IRanges holds (start, width), and a constant width is one number
 * (`w <- width(x)[1L]` when `S4Vectors::isConstant(width(x))`, checked once at import).  Every read
 * is [start, start + width - 1]; the `end` array never crosses PCIe (9 -> 5 bytes per read with the
 * seqnames as runs).  chrom == NULL: the seqnames come as runs (n_runs, run_chrom, run_len, as in
 * rcp_reads_load_rle); otherwise dense ids and the run arguments are ignored.  Same result as
 * rcp_reads_load with end = start + width - 1. */
int rcp_reads_load_width(int64_t n, const int32_t* chrom, int64_t n_runs, const int32_t* run_chrom,
                         const int32_t* run_len, const int32_t* start, int width, const int8_t* strand,
                         int n_chrom, const int64_t* chrom_len, int frag_len, int mem, int* reads_out);
/* The read-import step right before the path (preprocessRanges / readBam, ranges.R:1-65,
 * 111-134), for reads that are already decoded:
 *
 * rcp_reads_width_quantile  quantile(width(reads), prob), type 7 -- the cut of spliceAction =
 *   "remove" (ranges.R:126-127) -- and how many reads are not wider than it (the library size
 *   that "downsample" / "sampleto" then draw from).
 * rcp_r_sample_sorted  set.seed(seed) ONCE, then sort(sample(n[c], k[c])) for c = 0..n_calls-1
 *   from the one RNG stream, as the lapply over samples does (ranges.R:38-41, 55-58); sample()
 *   dispatches like base R's sample.int: the hash variant (redraw duplicates, at most 100
 *   tries) when n > 1e7 and k <= n/2, else the partial Fisher-Yates loop.  Serial by nature
 *   (one Mersenne-Twister stream): host code.  out = the blocks one after the other, 1-based.
 * rcp_reads_load_select  reads[-which(width > max_width)][idx] (ranges.R:128-130 then 43,62) and
 *   rcp_reads_load of the result, the selection done on the device.  max_width < 0: no width
 *   filter; idx NULL: keep every surviving read, else k 1-based positions among the SURVIVING
 *   reads in their original order.  *n_kept_out = reads that survive the width filter. */
int rcp_reads_width_quantile(int64_t n, const int32_t* start, const int32_t* end, double prob, int mem,
                             double* quantile_out /* host */, int64_t* n_le_out /* host */);
int rcp_r_sample_sorted(int seed, int sample_kind, int n_calls, const int64_t* n, const int64_t* k,
                        int32_t* out /* host, sum(k) */);
int rcp_reads_load_select(int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end,
                          const int8_t* strand, double max_width, int64_t k, const int32_t* idx,
                          int n_chrom, const int64_t* chrom_len, int frag_len, int mem,
                          int64_t* n_kept_out, int* reads_out);
/* The decode itself (readBam / readBed, ranges.R:111-146), on the device.  The
 * result is a `decoded` handle: chrom id, start, end (1-based, closed), strand (+1 / -1 / 0) in
 * FILE ORDER, resident in HBM.
 *
 * rcp_bam_index   walks the length-prefixed alignment records of an INFLATED BAM (the bytes after
 *   the header and the reference list, e.g. from rcp_bgzf_inflate): *n_records_out
 *   = records; offsets_out (may be NULL; capacity >= records + 1) = byte offset of every record
 *   and of the end.  Host code: the chain is serial.
 * rcp_bam_decode  readGAlignments(file) with its default flag filter (unmapped records dropped),
 *   then as(., "GRanges") (splice_split = 0: one range per alignment, start = pos + 1, width =
 *   the CIGAR's extent on the reference, M D N = X) or unlist(grglist(.)) (splice_split = 1,
 *   spliceAction "split": the runs of M D = X between N operations, empty ones dropped), strand
 *   from flag 0x10, then trim() against ref_len.  rec / offsets: host or device memory (`mem`);
 *   ref_len: host, the header's reference lengths.
 * rcp_bed_decode  import.bed(file, trackLine = FALSE): lines of chrom, chromStart, chromEnd
 *   [, name, score, strand] separated by tabs or blanks; start = chromStart + 1; strand '.' or
 *   absent = '*'; empty lines, '#' comments, "track" and "browser" lines skipped.  names: the
 *   seqlevels (host), chrom id = index into them; a name not among them is RCP_ERR_DATA.
 * rcp_decoded_fetch  the arrays -> host (any pointer may be NULL).
 * rcp_reads_load_decoded  rcp_reads_load of the decoded reads without a host round trip. */
/* BGZF, the container of a BAM file (SAMv1 4.1): gzip members of <= 64 KB that carry their own
 * compressed size, inflated independently by n_threads host threads (0: one per core; zlib).
 * rcp_bgzf_size: inflated bytes and blocks of the file; rcp_bgzf_inflate: the inflated bytes (the
 * BAM header, then the alignment records).  Host code, no GPU needed; RCP_ERR_DATA when the data
 * is not BGZF (e.g. plain gzip) or a block fails its CRC32 / ISIZE. */
int rcp_bgzf_size(const uint8_t* data /* host */, int64_t n_bytes, int64_t* inflated_bytes, int64_t* n_blocks);
int rcp_bgzf_inflate(const uint8_t* data /* host */, int64_t n_bytes, uint8_t* out /* host */, int64_t capacity,
                     int n_threads);
int rcp_bam_index(const uint8_t* rec /* host */, int64_t n_bytes, int64_t* n_records_out,
                  int64_t* offsets_out /* host */, int64_t capacity);
int rcp_bam_decode(const uint8_t* rec, int64_t n_bytes, const int64_t* offsets, int64_t n_records,
                   int n_ref, const int64_t* ref_len /* host */, int splice_split, int mem,
                   int* decoded_out, int64_t* n_out);
int rcp_bed_decode(const char* text, int64_t n_bytes, int n_names, const char* const* names /* host */,
                   int mem, int* decoded_out, int64_t* n_out);
int rcp_decoded_fetch(int decoded, int32_t* chrom, int32_t* start, int32_t* end, int8_t* strand,
                      int64_t capacity);
int rcp_decoded_free(int decoded);
int rcp_reads_load_decoded(int decoded, int n_chrom, const int64_t* chrom_len, int frag_len,
                           int* reads_out);
/* rcp_reads_width_quantile / rcp_reads_load_select on a decoded handle (idx: host memory). */
int rcp_decoded_width_quantile(int decoded, double prob, double* quantile_out, int64_t* n_le_out);
int rcp_reads_load_decoded_select(int decoded, double max_width, int64_t k, const int32_t* idx /* host */,
                                  int n_chrom, const int64_t* chrom_len, int frag_len,
                                  int64_t* n_kept_out, int* reads_out);
int rcp_reads_info(int reads, int64_t* n, int* n_chrom, int64_t* device_bytes);
int rcp_reads_free(int reads);

/* ---------------------------------------------------------------- coverage ---------------- */
/* calcCoverage(input, mask = GRanges, strand, ignore.strand) (coverage.R:126-174 with
 * coverageFromRanges, coverage.R:176-226): one window per region, exact integer per-base
 * coverage, reversed on '-' regions, NULL when no read overlaps / window leaves the chromosome.
 * strand_filter: RCP_STRAND_ANY or +1/-1/0.  The result stays on the device behind *cov_out. */
int rcp_coverage(int reads, int64_t n_regions, const int32_t* chrom, const int32_t* start,
                 const int32_t* end, const int8_t* strand, int ignore_strand, int strand_filter,
                 int mem, int* cov_out);
/* calcCoverage(input, mask = GRangesList, ...) (coverage.R:177-178,202-207): element g owns
 * ranges ptr[g]..ptr[g+1]-1 (the exons of one gene, list order); coverage is stitched in list
 * order; chromosome and strand are those of the element's FIRST range (coverage.R:182,185);
 * a read overlapping k ranges of the element counts k times (coverage.R:190-192). */
int rcp_coverage_list(int reads, int64_t n_elements, const int64_t* ptr, const int32_t* chrom,
                      const int32_t* start, const int32_t* end, const int8_t* strand,
                      int ignore_strand, int strand_filter, int mem, int* cov_out);
/* The merge step of coverageRnaRef (coverage.R:115-120): c(left, center, right) per element,
 * NULL if any part is NULL. */
int rcp_coverage_concat3(int left, int center, int right, int* cov_out);
/* Linear normalisation (recoup.R:559-577): every coverage value is multiplied by `factor`. */
int rcp_coverage_set_scale(int cov, double factor);
int rcp_coverage_info(int cov, int64_t* n_regions, int64_t* total_len, int64_t* n_null,
                      double* scale);
/* lengths(coverage): 0 for NULL entries (profile.R:6). */
int rcp_coverage_lengths(int cov, int32_t* len_out /* n_regions */);
/* Copy the unscaled integer coverage of regions [first, first+count) to the host, packed back to
 * back (region i occupies len[i] ints).  `capacity` = ints available in `out`. */
int rcp_coverage_fetch(int cov, int64_t first, int64_t count, int32_t* out, int64_t capacity);
/* The same regions as integer run-length encodings -- the Rle objects calcCoverage returns
 * (coverage.R:171-173, contract T1 of SURVEY 8a: `input[[s]]$coverage` is a list of integer Rle
 * or NULL).  run_ptr_out (count + 1) receives the offset of each region's runs; region i owns
 * values / lengths [run_ptr[i], run_ptr[i+1]) with sum(lengths) = len[i]; NULL regions own none.
 * Call with values_out = lengths_out = NULL to size the buffers (run_ptr_out[count] = total runs),
 * then again with `capacity` entries in each.  The encoding runs on the device (run heads by
 * neighbour comparison, compaction by prefix sums): only the runs cross PCIe. */
int rcp_coverage_rle(int cov, int64_t first, int64_t count, int64_t* run_ptr_out,
                     int32_t* values_out, int32_t* lengths_out, int64_t capacity);
int rcp_coverage_free(int cov);

/* ---------------------------------------------------------------- profile matrix ---------- */
/* binCoverageMatrix(cvrg, binSize, stat, interpolation, flank, where) (profile.R:153-212) with
 * splitVector (util.R:15-85): writes an [n_regions x n_bins] block, column-major with leading
 * dimension ld >= n_regions, at `out`.  NULL coverage -> zero row (profile.R:191-197). */
int rcp_bin_matrix(int cov, int where, int f1, int f2, int n_bins, int stat, int interp,
                   int seed, int sample_kind, double* out, int64_t ld, int mem);
/* baseCoverageMatrix(cvrg, flank, where) (profile.R:100-151): [n_regions x n_cols] per-base
 * block, same layout.  RCP_WHERE_WHOLE: n_cols must equal the common region length. */
int rcp_base_matrix(int cov, int where, int f1, int f2, int64_t n_cols, double* out, int64_t ld,
                    int mem);
/* Number of columns profileMatrix produces (profile.R:13-96) for these parameters. */
int rcp_profile_ncols(int cov, int equal_lengths, int f1, int f2, int flank_bin_size,
                      int region_bin_size, int64_t* ncols_out);
/* profileMatrix for ONE sample (profile.R:1-98): `equal_lengths` is the decision the reference
 * takes on the first sample (profile.R:6-10).  Writes cbind(left, center, right). */
int rcp_profile_matrix(int cov, int equal_lengths, int f1, int f2, int flank_bin_size,
                       int region_bin_size, int stat, int interp, int seed, int sample_kind,
                       double* out, int64_t ld, int mem);

/* ---------------------------------------------------------------- fused path -------------- */
/* coverageRef + profileMatrix (mean) for fixed-width windows (coverage.R:1-42 + profile.R:83-96):
 * windows of equal length (else RCP_ERR_ARG), n_bins bins (0 = per base), the matrix scaled by
 * `scale`.  With n_bins >= 1, windows at least n_bins long and reads the split path accepts, the
 * coverage is NEVER materialised: the tile kernel adds each tile's bin sums to 64-bit
 * accumulators and a last kernel divides -- same integer sums, same fp64 divide, so the matrix
 * equals rcp_coverage + rcp_profile_matrix bit for bit.  Otherwise the two stages are composed
 * and the coverage is released at once.  is_null_out (n_regions, may be NULL) receives the NULL
 * flags. */
int rcp_coverage_profile(int reads, int64_t n_regions, const int32_t* chrom,
                         const int32_t* start, const int32_t* end, const int8_t* strand,
                         int ignore_strand, int strand_filter, int n_bins, int seed,
                         int sample_kind, double scale, double* out, int64_t ld,
                         uint8_t* is_null_out, int mem);

/* ---------------------------------------------------------------- diagnostics ------------- */
/* Sort n 32-bit keys whose values are < 2^key_bits with the library's index sort (in place).
 * Exposed so the hand-written sort can be tested directly against a host sort. */
int rcp_sort_keys_u32(uint32_t* keys, int64_t n, int key_bits, int mem);

/* ---------------------------------------------------------------- matrix consumers -------- */
/* The reductions recoup's plotting side runs over a finished `$profile`, on the
 * matrix where rcp_profile_matrix left it: column-major n_rows x n_cols, leading dimension ld,
 * `mem` says where m AND the outputs live (rcp_matrix_quantile's probs/out are always host).
 *
 * rcp_matrix_col_profile: the average curve and its band of calcPlotProfiles (plot.R:949-990):
 *   center = apply(x, 2, mean)   spread = apply(x, 2, sd)     (stat = RCP_STAT_MEAN)
 *   center = apply(x, 2, median) spread = apply(x, 2, mad)    (stat = RCP_STAT_MEDIAN; mad's
 *   constant 1.4826); log2_scale != 0 first maps x <- log2(x + 1) (plot.R:953-956).  The caller
 *   forms upper/lower = center +- spread.  sd of a single row is NaN (R: NA).
 * rcp_matrix_row_stat: apply(x, 1, sum | max | mean), the ordering values of orderProfiles
 *   (plot.R:1076-1081, 1110-1131, 1135-1146; the max variant's sample() among tied maxima
 *   returns the same value whichever it picks).
 * rcp_order: sort(v, decreasing, index.return=TRUE)$ix (plot.R:1038-1041,1100-1103): 1-based,
 *   ties keep their input order in both directions (R's default radix method), NaN/NA dropped:
 *   *n_out entries of ix are valid (R drops them: na.last = NA).
 * rcp_matrix_quantile: quantile(x, probs), type 7, over every cell (plot.R:519,536); a NaN
 *   anywhere is RCP_ERR_DATA like R's error without na.rm. */
#define RCP_ROW_SUM 0
#define RCP_ROW_MAX 1
#define RCP_ROW_MEAN 2
int rcp_matrix_col_profile(const double* m, int64_t n_rows, int64_t n_cols, int64_t ld, int stat,
                           int log2_scale, int mem, double* center /* n_cols */,
                           double* spread /* n_cols */);
int rcp_matrix_row_stat(const double* m, int64_t n_rows, int64_t n_cols, int64_t ld, int what, int mem,
                        double* out /* n_rows */);
int rcp_order(const double* v, int64_t n, int decreasing, int mem, int32_t* ix /* n */, int64_t* n_out);
int rcp_matrix_quantile(const double* m, int64_t n_rows, int64_t n_cols, int64_t ld, const double* probs,
                        int k, int mem, double* out /* k, host */);

/* ---------------------------------------------------------------- multi-GPU helper -------- */
/* Scatter a row block into the gathered matrix: dst[row_index[i] + c*ld_dst] =
 * src[i + c*ld_src] for i < n_rows, c < n_cols (all pointers on the device).  Used after the
 * NCCL gather of per-rank row blocks (the reference's do.call(rbind, ...), profile.R:150,208). */
int rcp_rows_scatter(const double* src, int64_t ld_src, int64_t n_rows, int64_t n_cols,
                     const int64_t* row_index, double* dst, int64_t ld_dst);

/* Place a row block into a matrix with the COPY ENGINE (one strided 2-D copy, no SM is used):
 * dst[i + c*ld_dst] = src[i + c*ld_src] for i < n_rows, c < n_cols.  `dst` may lie in another
 * GPU's memory mapped with rcp_shared_open (the block then crosses NVLink straight into its place
 * in the gathering rank's matrix -- the reference's do.call(rbind, ...), profile.R:150,208, without a
 * gather buffer or a placement pass) or in page-locked host memory.  `stream`: a cudaStream_t, or
 * NULL for the library's stream. */
int rcp_rows_put(const double* src, int64_t ld_src, int64_t n_rows, int64_t n_cols, double* dst,
                 int64_t ld_dst, void* stream);

/* Region-sharded runs (every GPU owns a region slice plus its overlapping reads; the
 * reference's analogue is the split of the regions over mclapply workers, coverage.R:148-154,
 * each of which subsets the reads with findOverlaps).  A rank holds an arbitrary share of the
 * reads ON ITS DEVICE; spans[(r * n_chrom + c) * 2 + {0, 1}] (host) is the closed interval
 * [lo, hi] that rank r's slice covers on chromosome c (lo > hi: nothing).  A read goes to every
 * rank whose interval it overlaps; a read with a chromosome id outside [0, n_chrom) goes to rank
 * 0, whose rcp_reads_load reports it.  _count: counts_out[r] (host) = reads bound for rank r (one
 * host synchronisation).  _pack: writes them into four device arrays (chrom, start, end int32,
 * strand int8; strand / strand_out may be NULL), rank r's run starting at offsets[r] (host: the
 * exclusive prefix sum of the counts) in each -- the send buffers of the all-to-all.  The order
 * inside a run is unspecified (coverage does not depend on it). */
int rcp_reads_route_count(int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end, int world,
                          int n_chrom, const int32_t* spans, int64_t* counts_out /* world, host */);
int rcp_reads_route_pack(int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end,
                         const int8_t* strand, int world, int n_chrom, const int32_t* spans,
                         const int64_t* offsets /* world, host */, int32_t* chrom_out, int32_t* start_out,
                         int32_t* end_out, int8_t* strand_out);

/* Device memory that the other processes of this box (one per GPU) can map, so that every rank's
 * rcp_profile_matrix / rcp_bin_matrix / rcp_base_matrix writes its row block STRAIGHT into the
 * matrix of the gathering rank over NVLink (out = base + first_row, ld = total rows): the
 * exchange is fused into the kernel's own stores and no gather / scatter pass runs.  The owner
 * allocates and exports a 64-byte handle (CUDA IPC), peers open it; completion is ordered by any
 * collective or barrier after the call.  rcp_shared_close undoes rcp_shared_open,
 * rcp_shared_free undoes rcp_shared_alloc. */
#define RCP_IPC_HANDLE_BYTES 64
int rcp_shared_alloc(int64_t bytes, void** ptr_out, unsigned char* handle_out /* 64 */);
int rcp_shared_open(const unsigned char* handle /* 64 */, void** ptr_out);
int rcp_shared_close(void* ptr);
int rcp_shared_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* RECOUP_B200_H */
