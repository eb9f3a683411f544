// Coverage without a sort (sm_100a): calcCoverage / coverageFromRanges of the reference
// (/root/reference/R/coverage.R:126-226) straight from the UNSORTED reads.
//
// Only reads that overlap a region contribute to that region's coverage, and a region's coverage
// does not depend on the order of its reads.  So instead of sorting every read by coordinate
// (index.cu + sort.cu) this path drops each read into the buckets of the OUTPUT TILES it
// overlaps -- the region-sorted read lists the coverage kernels consume:
//
//   1. plan          regions -> windows (geometry / NULL rules) -> tiles (<= 7168 outputs, cut
//                    in output space; regions <= 1024 bp are one warp-sized tile)
//   2. cell lists    the genome is cut into 1024-bp cells; cell -> tiles overlapping it (CSR,
//                    L2-resident: 4 B per cell + 4 B per (tile, cell) pair)
//   3. count pass    every read looks up the cells it touches and bumps the counter of each
//                    tile it overlaps (a read that hits no cell list costs two L2 loads)
//   4. NULL rule     a region with no read in any tile is NULL (coverage.R:198,224-225);
//                    offsets of the dense coverage and of the buckets by prefix sums
//   5. scatter pass  the same walk; the read is clipped to the tile and stored as ONE packed
//                    32-bit event pair (first covered output | one past the last) -- '-' regions
//                    are mirrored here, so the tile kernels are strand-agnostic
//   6. tile kernels  bucket -> shared-memory difference array (atomics) -> block prefix scan ->
//                    aligned 16-byte stores of the int32 coverage
//
// HBM traffic per read: 2 x 8-9 B (the two passes) + 2 x 4 B per (read, tile) hit, against
// >= 4 passes x 8 B per read for a radix sort of the whole read set.
#include "cov_common.cuh"
#include "rcp_internal.cuh"

namespace rcp {

using namespace covk;

namespace {

constexpr int CELL_SHIFT = 10;                       // 1024-bp cells
constexpr int CELLS_PER_BIG = ((TILE - 1) >> CELL_SHIFT) + 2;
constexpr int CELLS_PER_SMALL = ((SMALL_MAX - 1) >> CELL_SHIFT) + 2;
constexpr int RTPB = 256;

// tile record, first half: what the read passes need
//   x = global coordinate of the first genomic base of the tile
//   y = tlen (bits 0..15) | reverse (bit 16) | class mask (bits 17..19)
// second half: x = region, y = offset of the tile inside the region's output
struct Tiles {
    uint2* a;
    uint2* b;
};

__global__ void __launch_bounds__(CTA)
bkt_plan_kernel(int64_t R, const int32_t* __restrict__ chrom, const int32_t* __restrict__ start,
                const int32_t* __restrict__ end, const int8_t* __restrict__ strand,
                const uint32_t* __restrict__ chrom_off, const int64_t* __restrict__ chrom_len,
                int n_chrom, int ignore_strand, int strand_filter, uint32_t* __restrict__ gs_out,
                int32_t* __restrict__ plen, uint8_t* __restrict__ flags,
                int64_t* __restrict__ nbig, int64_t* __restrict__ nsmall,
                unsigned int* __restrict__ err) {
    const int64_t r = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (r >= R) return;
    const int st = strand ? (int)strand[r] : 0;
    uint32_t gs;
    int64_t L;
    const bool null = window_geometry(chrom[r], start[r], end[r], n_chrom, chrom_off, chrom_len,
                                      err, &gs, &L);
    const int32_t len = null ? 0 : (int32_t)L;
    gs_out[r] = gs;
    plen[r] = len;
    flags[r] = (uint8_t)((st < 0 ? 1u : 0u) | (class_mask(st, ignore_strand, strand_filter) << 1));
    nbig[r] = len > SMALL_MAX ? ((int64_t)len + TILE - 1) / TILE : 0;
    nsmall[r] = (len > 0 && len <= SMALL_MAX) ? 1 : 0;
}

// last r in [0, R) with off[r] <= t  (off is an exclusive prefix sum with off[R] = total > t)
__device__ __forceinline__ int64_t owner_of(const int64_t* __restrict__ off, int64_t R, int64_t t) {
    int64_t lo = 0, hi = R;              // answer in [lo, hi)
    while (hi - lo > 1) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(off + mid) <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

// One thread per tile: tile record + the count of every cell it overlaps.
__global__ void __launch_bounds__(CTA)
bkt_tiles_kernel(int64_t R, int64_t Tb, int64_t Ts, const int64_t* __restrict__ off_big,
                 const int64_t* __restrict__ off_small, const uint32_t* __restrict__ gs,
                 const int32_t* __restrict__ plen, const uint8_t* __restrict__ flags, Tiles tiles,
                 uint32_t* __restrict__ cell_cnt) {
    const int64_t t = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (t >= Tb + Ts) return;
    int64_t r;
    uint32_t o_lo, tlen, tstart;
    if (t < Tb) {
        r = owner_of(off_big, R, t);
        const int j = (int)(t - off_big[r]);
        const int L = plen[r];
        const int m = (L + TILE - 1) / TILE;
        const int tile_len = (((L + m - 1) / m) + ROW - 1) / ROW * ROW;
        o_lo = (uint32_t)(j * tile_len);
        tlen = (uint32_t)min(tile_len, L - (int)o_lo);
        // tiles are cut in OUTPUT space: on a '-' region tile j covers the mirrored positions
        const uint32_t q0 = (flags[r] & 1u) ? (uint32_t)L - o_lo - tlen : o_lo;
        tstart = gs[r] + q0;
    } else {
        r = owner_of(off_small, R, t - Tb);
        o_lo = 0;
        tlen = (uint32_t)plen[r];
        tstart = gs[r];
    }
    tiles.a[t] = make_uint2(tstart, tlen | ((uint32_t)flags[r] << 16));
    tiles.b[t] = make_uint2((uint32_t)r, o_lo);
    const uint32_t c1 = (tstart + tlen - 1u) >> CELL_SHIFT;
    for (uint32_t c = tstart >> CELL_SHIFT; c <= c1; c++) atomicAdd(cell_cnt + c, 1u);
}

// cell -> tiles lists.  cell_cnt counts down to zero while the slots are handed out.
__global__ void __launch_bounds__(CTA)
bkt_cells_kernel(int64_t T, Tiles tiles, uint32_t* __restrict__ cell_cnt,
                 const uint32_t* __restrict__ cell_ptr, uint32_t* __restrict__ cell_list) {
    const int64_t t = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (t >= T) return;
    const uint2 a = tiles.a[t];
    const uint32_t tlen = a.y & 0xffffu;
    const uint32_t c1 = (a.x + tlen - 1u) >> CELL_SHIFT;
    for (uint32_t c = a.x >> CELL_SHIFT; c <= c1; c++) {
        const uint32_t slot = atomicSub(cell_cnt + c, 1u) - 1u;
        cell_list[cell_ptr[c] + slot] = (uint32_t)t;
    }
}

// The walk both read passes share.  A (read, tile) hit is credited in exactly one cell: the one
// holding the first base of their intersection.
template <bool SCATTER, bool STRANDED>
__device__ __forceinline__ void visit_read(uint32_t s, uint32_t e1, int st,
                                           const uint32_t* __restrict__ cell_ptr,
                                           const uint32_t* __restrict__ cell_list,
                                           const uint2* __restrict__ tile_a,
                                           unsigned long long* __restrict__ tile_cnt,
                                           const int64_t* __restrict__ boff,
                                           uint32_t* __restrict__ bucket) {
    if (e1 <= s) return;
    const unsigned bit = st > 0 ? 0u : (st < 0 ? 1u : 2u);
    const uint32_t c1 = (e1 - 1u) >> CELL_SHIFT;
    for (uint32_t c = s >> CELL_SHIFT; c <= c1; c++) {
        const uint32_t ja = __ldg(cell_ptr + c), jb = __ldg(cell_ptr + c + 1);
        for (uint32_t j = ja; j < jb; j++) {
            const uint32_t t = __ldg(cell_list + j);
            const uint2 a = __ldg(tile_a + t);
            const uint32_t ts = a.x, tl = a.y & 0xffffu;
            if (!(ts < e1 && ts + tl > s)) continue;
            const uint32_t is = max(ts, s);
            if ((is >> CELL_SHIFT) != c) continue;
            if (STRANDED && !((a.y >> (17 + bit)) & 1u)) continue;
            if (!SCATTER) {
                atomicAdd(tile_cnt + t, 1ull);
            } else {
                const unsigned long long slot = atomicAdd(tile_cnt + t, ~0ull) - 1ull;
                uint32_t lo = is - ts, hi = min(e1, ts + tl) - ts;      // covered [lo, hi)
                if (a.y & 0x10000u) {                                     // '-' region: mirror
                    const uint32_t l2 = tl - hi;
                    hi = tl - lo;
                    lo = l2;
                }
                bucket[boff[t] + (int64_t)slot] = lo | (hi << 16);
            }
        }
    }
}

template <bool SCATTER, bool STRANDED>
__global__ void __launch_bounds__(RTPB)
bkt_reads_kernel(int64_t n, const uint32_t* __restrict__ g_start,
                 const uint32_t* __restrict__ g_end1, const int8_t* __restrict__ strand,
                 const uint32_t* __restrict__ cell_ptr, const uint32_t* __restrict__ cell_list,
                 const uint2* __restrict__ tile_a, unsigned long long* __restrict__ tile_cnt,
                 const int64_t* __restrict__ boff, uint32_t* __restrict__ bucket) {
    const int64_t stride = (int64_t)gridDim.x * RTPB;
    const int64_t n_vec = n >> 2;
    for (int64_t v = (int64_t)blockIdx.x * RTPB + threadIdx.x; v < n_vec; v += stride) {
        const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(g_start) + v);
        const uint4 e4 = __ldg(reinterpret_cast<const uint4*>(g_end1) + v);
        char4 t4 = make_char4(0, 0, 0, 0);
        if (STRANDED && strand) t4 = __ldg(reinterpret_cast<const char4*>(strand) + v);
        visit_read<SCATTER, STRANDED>(s4.x, e4.x, t4.x, cell_ptr, cell_list, tile_a, tile_cnt, boff, bucket);
        visit_read<SCATTER, STRANDED>(s4.y, e4.y, t4.y, cell_ptr, cell_list, tile_a, tile_cnt, boff, bucket);
        visit_read<SCATTER, STRANDED>(s4.z, e4.z, t4.z, cell_ptr, cell_list, tile_a, tile_cnt, boff, bucket);
        visit_read<SCATTER, STRANDED>(s4.w, e4.w, t4.w, cell_ptr, cell_list, tile_a, tile_cnt, boff, bucket);
    }
    for (int64_t i = n_vec * 4 + (int64_t)blockIdx.x * RTPB + threadIdx.x; i < n; i += stride)
        visit_read<SCATTER, STRANDED>(g_start[i], g_end1[i], (STRANDED && strand) ? (int)strand[i] : 0,
                                      cell_ptr, cell_list, tile_a, tile_cnt, boff, bucket);
}

// NULL rule (coverage.R:198,224-225): no overlapping read in any tile of the region.
__global__ void __launch_bounds__(CTA)
bkt_null_kernel(int64_t R, int64_t Tb, const int64_t* __restrict__ off_big,
                const int64_t* __restrict__ off_small, const int32_t* __restrict__ plen,
                const unsigned long long* __restrict__ tile_cnt, int32_t* __restrict__ len,
                uint8_t* __restrict__ is_null, int64_t* __restrict__ padded,
                unsigned long long* __restrict__ stats /* [0] n_null, [1] total_len */) {
    const int64_t r = (int64_t)blockIdx.x * CTA + threadIdx.x;
    unsigned long long my_null = 0, my_len = 0;
    if (r < R) {
        const int32_t L = plen[r];
        unsigned long long hits = 0;
        if (L > SMALL_MAX) {
            for (int64_t t = off_big[r]; t < off_big[r + 1]; t++) hits += tile_cnt[t];
        } else if (L > 0) {
            hits = tile_cnt[Tb + off_small[r]];
        }
        const bool null = hits == 0;
        const int32_t out = null ? 0 : L;
        len[r] = out;
        is_null[r] = null ? 1 : 0;
        padded[r] = ((int64_t)out + PAD - 1) / PAD * PAD;
        my_null = null ? 1 : 0;
        my_len = (unsigned long long)out;
    }
    for (int d = 16; d > 0; d >>= 1) {
        my_null += __shfl_xor_sync(0xffffffffu, my_null, d);
        my_len += __shfl_xor_sync(0xffffffffu, my_len, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_null) atomicAdd(&stats[0], my_null);
        if (my_len) atomicAdd(&stats[1], my_len);
    }
}

// One CTA per tile of a long region.
__global__ void __launch_bounds__(CTA)
bkt_tile_kernel(Tiles tiles, const int64_t* __restrict__ boff, const uint32_t* __restrict__ bucket,
                const uint8_t* __restrict__ is_null, const int64_t* __restrict__ off,
                int32_t* __restrict__ cov) {
    __shared__ __align__(16) int diff[TILE];
    __shared__ int rowpre[MAX_ROWS + 1];
    const int tid = threadIdx.x;
    const int64_t t = blockIdx.x;
    const uint2 b = tiles.b[t];
    if (is_null[b.x]) return;
    const int tlen = (int)(tiles.a[t].y & 0xffffu);
    const int nrows = (tlen + ROW - 1) / ROW;
    for (int i = tid; i < nrows * (ROW / 4); i += CTA)
        reinterpret_cast<int4*>(diff)[i] = make_int4(0, 0, 0, 0);
    __syncthreads();
    const int64_t b1 = boff[t + 1];
    for (int64_t i = boff[t] + tid; i < b1; i += CTA) {
        const uint32_t e = __ldg(bucket + i);
        const int lo = (int)(e & 0xffffu), hi = (int)(e >> 16);
        atomicAdd(diff + lo, 1);
        if (hi < tlen) atomicSub(diff + hi, 1);
    }
    __syncthreads();
    block_scan_store(diff, tlen, 0, false, rowpre, cov + off[b.x] + b.y);
}

// One warp per short region (<= SMALL_MAX bases): warp-private tile, __syncwarp only.
__global__ void __launch_bounds__(CTA)
bkt_small_kernel(int64_t Tb, int64_t Ts, Tiles tiles, const int64_t* __restrict__ boff,
                 const uint32_t* __restrict__ bucket, const uint8_t* __restrict__ is_null,
                 const int64_t* __restrict__ off, int32_t* __restrict__ cov) {
    __shared__ __align__(16) int sm[WARPS][SMALL_MAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ts = (int64_t)blockIdx.x * WARPS + warp;
    if (ts >= Ts) return;
    const int64_t t = Tb + ts;
    const uint32_t r = tiles.b[t].x;
    if (is_null[r]) return;
    const int L = (int)(tiles.a[t].y & 0xffffu);
    int* diff = sm[warp];
    const int nrows = (L + ROW - 1) / ROW;
    for (int i = lane; i < nrows * (ROW / 4); i += 32)
        reinterpret_cast<int4*>(diff)[i] = make_int4(0, 0, 0, 0);
    __syncwarp();
    const int64_t b1 = boff[t + 1];
    for (int64_t i = boff[t] + lane; i < b1; i += 32) {
        const uint32_t e = __ldg(bucket + i);
        const int lo = (int)(e & 0xffffu), hi = (int)(e >> 16);
        atomicAdd(diff + lo, 1);
        if (hi < L) atomicSub(diff + hi, 1);
    }
    __syncwarp();
    int32_t* dst = cov + off[r];
    int pre = 0;
    for (int row = 0; row < nrows; row++)
        pre += warp_row_scan_store(diff + row * ROW, pre, 0, false, L - row * ROW, dst + row * ROW);
}

inline unsigned reads_grid(int64_t n) {
    int64_t b = ((n + 3) / 4 + RTPB - 1) / RTPB;
    const int64_t cap = (int64_t)g_ctx.sm_count * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// every device array of one call; freed on every exit path
struct Work {
    uint32_t* gs = nullptr;
    int32_t* plen = nullptr;
    uint8_t* flags = nullptr;
    int64_t* nbig = nullptr;
    int64_t* nsmall = nullptr;
    int64_t* off_big = nullptr;
    int64_t* off_small = nullptr;
    int64_t* padded = nullptr;
    unsigned int* err = nullptr;
    unsigned long long* stats = nullptr;
    Tiles tiles = {nullptr, nullptr};
    uint32_t* cell_cnt = nullptr;
    uint32_t* cell_ptr = nullptr;
    uint32_t* cell_list = nullptr;
    unsigned long long* tile_cnt = nullptr;
    int64_t* boff = nullptr;
    uint32_t* bucket = nullptr;
    ~Work() {
        dfree(gs);
        dfree(plen);
        dfree(flags);
        dfree(nbig);
        dfree(nsmall);
        dfree(off_big);
        dfree(off_small);
        dfree(padded);
        dfree(err);
        dfree(stats);
        dfree(tiles.a);
        dfree(tiles.b);
        dfree(cell_cnt);
        dfree(cell_ptr);
        dfree(cell_list);
        dfree(tile_cnt);
        dfree(boff);
        dfree(bucket);
    }
};

template <bool STRANDED>
int launch_reads_passes(bool scatter, const ReadsIdx& rd, const Work& w) {
    const unsigned grid = reads_grid(rd.n);
    if (!scatter)
        bkt_reads_kernel<false, STRANDED><<<grid, RTPB, 0, g_ctx.stream>>>(
            rd.n, rd.g_start, rd.g_end1, rd.d_strand, w.cell_ptr, w.cell_list, w.tiles.a,
            w.tile_cnt, w.boff, w.bucket);
    else
        bkt_reads_kernel<true, STRANDED><<<grid, RTPB, 0, g_ctx.stream>>>(
            rd.n, rd.g_start, rd.g_end1, rd.d_strand, w.cell_ptr, w.cell_list, w.tiles.a,
            w.tile_cnt, w.boff, w.bucket);
    RCP_LAUNCHED();
    return RCP_OK;
}

}  // namespace

int coverage_ranges_bucketed(ReadsIdx& rd, int64_t R, const int32_t* chrom, const int32_t* start,
                             const int32_t* end, const int8_t* strand, int ignore_strand,
                             int strand_filter, int mem, Coverage* cv) {
    DevIn<int32_t> d_chrom, d_start, d_end;
    DevIn<int8_t> d_strand;
    RCP_TRY(d_chrom.init(chrom, (size_t)R, mem));
    RCP_TRY(d_start.init(start, (size_t)R, mem));
    RCP_TRY(d_end.init(end, (size_t)R, mem));
    RCP_TRY(d_strand.init(strand, (size_t)R, mem));
    // the class test is needed only when some rule can exclude a read (reads without a strand
    // array are all '*': a '+' / '-' filter then keeps nothing)
    const bool stranded = !((strand_filter == RCP_STRAND_ANY) && (ignore_strand || strand == nullptr));

    cv->n_regions = R;
    RCP_TRY(dalloc(&cv->off, (size_t)R + 1));
    RCP_TRY(dalloc(&cv->len, (size_t)R));
    RCP_TRY(dalloc(&cv->is_null, (size_t)R));
    Work w;
    RCP_TRY(dalloc(&w.gs, (size_t)R));
    RCP_TRY(dalloc(&w.plen, (size_t)R));
    RCP_TRY(dalloc(&w.flags, (size_t)R));
    RCP_TRY(dalloc(&w.nbig, (size_t)R));
    RCP_TRY(dalloc(&w.nsmall, (size_t)R));
    RCP_TRY(dalloc(&w.off_big, (size_t)R + 1));
    RCP_TRY(dalloc(&w.off_small, (size_t)R + 1));
    RCP_TRY(dalloc(&w.padded, (size_t)R));
    RCP_TRY(dalloc(&w.err, 1));
    RCP_TRY(dalloc(&w.stats, 2));
    RCP_CUDA(cudaMemsetAsync(w.err, 0, sizeof(unsigned int), g_ctx.stream));
    RCP_CUDA(cudaMemsetAsync(w.stats, 0, 2 * sizeof(unsigned long long), g_ctx.stream));

    struct Host {
        int64_t Tb, Ts, total_padded, hits;
        unsigned long long stats[2];
        unsigned int err;
    } h = {0, 0, 0, 0, {0, 0}, 0};

    // ---- 1. plan: windows and tile counts ------------------------------------------------
    {
        StageTimer t(ST_BKT_PLAN);
        if (R > 0) {
            bkt_plan_kernel<<<blocks_for(R, CTA), CTA, 0, g_ctx.stream>>>(
                R, d_chrom.ptr, d_start.ptr, d_end.ptr, d_strand.ptr, rd.d_chrom_off,
                rd.d_chrom_len, rd.n_chrom, ignore_strand, strand_filter, w.gs, w.plen, w.flags,
                w.nbig, w.nsmall, w.err);
            RCP_LAUNCHED();
        }
        RCP_TRY(exclusive_scan2_i64(w.nbig, w.off_big, w.off_big + R, w.nsmall, w.off_small,
                                    w.off_small + R, R));
    }
    RCP_CUDA(cudaMemcpyAsync(&h.Tb, w.off_big + R, 8, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(&h.Ts, w.off_small + R, 8, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(&h.err, w.err, 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaStreamSynchronize(g_ctx.stream));
    if (h.err & 1u) return fail(RCP_ERR_DATA, "a region has a chromosome id outside [0, n_chrom)");
    if (h.err & 2u) return fail(RCP_ERR_DATA, "a region has end < start - 1");
    const int64_t Tb = h.Tb, Ts = h.Ts, T = Tb + Ts;
    if (T > 0x7fffffff) return fail(RCP_ERR_UNSUPPORTED, "more than 2^31-1 coverage tiles");

    // ---- 2. tiles and cell lists ----------------------------------------------------------
    const int64_t n_cell = ((int64_t)rd.chrom_off[(size_t)rd.n_chrom] >> CELL_SHIFT) + 2;
    RCP_TRY(dalloc(&w.tiles.a, (size_t)T));
    RCP_TRY(dalloc(&w.tiles.b, (size_t)T));
    RCP_TRY(dalloc(&w.tile_cnt, (size_t)T + 1));
    RCP_TRY(dalloc(&w.boff, (size_t)T + 1));
    RCP_TRY(dalloc(&w.cell_cnt, (size_t)n_cell + 1));
    RCP_TRY(dalloc(&w.cell_ptr, (size_t)n_cell + 1));
    RCP_TRY(dalloc(&w.cell_list, (size_t)(Tb * CELLS_PER_BIG + Ts * CELLS_PER_SMALL)));
    {
        StageTimer t(ST_BKT_PLAN);
        RCP_CUDA(cudaMemsetAsync(w.cell_cnt, 0, ((size_t)n_cell + 1) * 4, g_ctx.stream));
        RCP_CUDA(cudaMemsetAsync(w.tile_cnt, 0, ((size_t)T + 1) * 8, g_ctx.stream));
        if (T > 0) {
            bkt_tiles_kernel<<<blocks_for(T, CTA), CTA, 0, g_ctx.stream>>>(
                R, Tb, Ts, w.off_big, w.off_small, w.gs, w.plen, w.flags, w.tiles, w.cell_cnt);
            RCP_LAUNCHED();
        }
        RCP_TRY(exclusive_scan_u32(w.cell_cnt, w.cell_ptr, n_cell + 1, nullptr));
        if (T > 0) {
            bkt_cells_kernel<<<blocks_for(T, CTA), CTA, 0, g_ctx.stream>>>(T, w.tiles, w.cell_cnt,
                                                                          w.cell_ptr, w.cell_list);
            RCP_LAUNCHED();
        }
    }
    // ---- 3. count pass ----------------------------------------------------------------------
    if (T > 0 && rd.n > 0) {
        StageTimer t(ST_BKT_COUNT);
        RCP_TRY(stranded ? launch_reads_passes<true>(false, rd, w)
                         : launch_reads_passes<false>(false, rd, w));
    }
    // ---- 4. NULL rule, offsets ---------------------------------------------------------------
    {
        StageTimer t(ST_BKT_PLAN);
        if (R > 0) {
            bkt_null_kernel<<<blocks_for(R, CTA), CTA, 0, g_ctx.stream>>>(
                R, Tb, w.off_big, w.off_small, w.plen, w.tile_cnt, cv->len, cv->is_null, w.padded,
                w.stats);
            RCP_LAUNCHED();
        }
        RCP_TRY(exclusive_scan_i64(w.padded, cv->off, R, cv->off + R));
        RCP_TRY(exclusive_scan_i64(reinterpret_cast<const int64_t*>(w.tile_cnt), w.boff, T,
                                   w.boff + T));
    }
    RCP_CUDA(cudaMemcpyAsync(&h.total_padded, cv->off + R, 8, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(&h.hits, w.boff + T, 8, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(h.stats, w.stats, 16, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaStreamSynchronize(g_ctx.stream));
    cv->total_padded = h.total_padded;
    cv->n_null = (int64_t)h.stats[0];
    cv->total_len = (int64_t)h.stats[1];
    RCP_TRY(dalloc(&cv->cov, (size_t)h.total_padded));
    if (h.hits == 0) return RCP_OK;                 // every region is NULL
    RCP_TRY(dalloc(&w.bucket, (size_t)h.hits));
    // ---- 5. scatter pass --------------------------------------------------------------------
    {
        StageTimer t(ST_BKT_SCATTER);
        RCP_TRY(stranded ? launch_reads_passes<true>(true, rd, w)
                         : launch_reads_passes<false>(true, rd, w));
    }
    // ---- 6. tiles -> coverage ---------------------------------------------------------------
    if (Tb > 0) {
        StageTimer t(ST_BKT_TILE);
        bkt_tile_kernel<<<(unsigned)Tb, CTA, 0, g_ctx.stream>>>(w.tiles, w.boff, w.bucket,
                                                                cv->is_null, cv->off, cv->cov);
        RCP_LAUNCHED();
    }
    if (Ts > 0) {
        StageTimer t(ST_BKT_SMALL);
        bkt_small_kernel<<<blocks_for(Ts, WARPS), CTA, 0, g_ctx.stream>>>(
            Tb, Ts, w.tiles, w.boff, w.bucket, cv->is_null, cv->off, cv->cov);
        RCP_LAUNCHED();
    }
    return RCP_OK;
}

}  // namespace rcp
