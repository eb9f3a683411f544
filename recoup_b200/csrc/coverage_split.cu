// SPLIT path (sm_100a): calcCoverage / coverageFromRanges of the reference
// (/root/reference/R/coverage.R:126-226) straight from the UNSORTED reads, with ONE streaming pass
// over the reads and no per-read random access outside shared memory.
//
//   plan     regions -> windows (geometry / NULL rules of coverage.R:209,217-222), tile counts,
//            storage offsets, and a 1-bit-per-16-kb BLOCK BITMAP of the mask.  A rank table over
//            the bitmap numbers the set blocks 0 .. n_set-1 ("compacted genome"): consecutive
//            ranks are cut into <= 1024 GROUPS of 2^gshift blocks.
//   split    every read is tested against the bitmap in shared memory; a survivor is clipped to
//            the set blocks it touches, packed into ONE 32-bit word (position inside its group |
//            strand class | width) and appended to its group's 32-entry ring in shared memory
//            (one shared-memory atomic).  Full 16-entry chunks leave as 64-byte stores into a
//            chunk pool; a chunk carries its group in a 2-byte tag.  No histogram pass, no global
//            atomics per read, no prefix sum over the reads.
//   sort     the chunk tags are counting-sorted by group (three tiny kernels over ~N/16 tags);
//            then one CTA per group sorts the group's candidates by 2-kb sub-bin of the compacted
//            genome (shared-memory histogram + cursors) into one dense candidate array and
//            writes the sub-bin offsets.
//   tiles    a tile (<= 7168 outputs of one region) reads the candidates of the sub-bins under
//            it (reaching back by the widest read), clips them, builds the difference array in
//            shared memory (atomics), scans it and writes the int32 coverage with TMA bulk
//            stores; regions <= 1024 bp are handled by one warp each.
//   NULL     a region none of whose tiles saw a read is NULL (coverage.R:198,224-225).  Its
//            storage is allocated up front (offsets do not depend on the reads), so the host
//            synchronises ONCE per call, right after the plan, and never in the read passes.
//
// HBM traffic: 8-9 B per read once, 4 B per candidate written and read twice (second time from
// L2), 4 B per covered base written once.
#include <algorithm>
#include <cstdlib>

#include "cov_common.cuh"
#include "rcp_internal.cuh"

namespace rcp {

using namespace covk;

namespace {

constexpr int BLK_SHIFT = 14;                         // 16384-bp blocks of the bitmap
constexpr uint32_t BLK_MASK = (1u << BLK_SHIFT) - 1u;
constexpr int SUB_SHIFT = 11;                         // 2048-bp sub-bins of the candidate sort
constexpr int SUBS = 1 << (BLK_SHIFT - SUB_SHIFT);    // sub-bins per block
constexpr int NG = 1024;                              // groups (at most)
constexpr int RING = 32;                              // ring entries per group (power of two)
constexpr int CH = 16;                                // entries per chunk (64 bytes)
constexpr int ST = 1024;                              // threads of the split kernel (== NG)
constexpr int SLAB = 64;                              // chunks a warp takes from the pool at a time
constexpr int MAX_WORDS = 8192;                       // bitmap words (2^32 positions)
constexpr int MIN_GSHIFT = 2;                         // >= 4 blocks per group: position field >= 16 bits
constexpr int MAX_GSHIFT = 8;
constexpr int GT = 512;                               // threads of the group kernel
constexpr int CS = 1024;                              // threads of the chunk-sort kernels
static_assert(ST == NG, "the flush phase maps thread t to group t");
static_assert(RING == 2 * CH, "a ring holds two chunks");
static_assert(SLAB >= 64, "one flush phase of a warp needs up to 64 chunks");

// ---------------------------------------------------------------------------------- plan ------
__global__ void __launch_bounds__(CTA)
sp_plan_kernel(int64_t R, const int32_t* __restrict__ chrom, const int32_t* __restrict__ start,
               const int32_t* __restrict__ end, const int8_t* __restrict__ strand,
               const uint32_t* __restrict__ chrom_off, const int64_t* __restrict__ chrom_len,
               int n_chrom, int ignore_strand, int strand_filter, uint32_t* __restrict__ gs_out,
               int32_t* __restrict__ plen, uint8_t* __restrict__ flags, int64_t* __restrict__ nbig,
               int64_t* __restrict__ nsmall, int64_t* __restrict__ padded,
               uint32_t* __restrict__ bitmap, unsigned int* __restrict__ err,
               unsigned long long* __restrict__ pstats /* [0] total len [1] max len */) {
    const int64_t r = (int64_t)blockIdx.x * CTA + threadIdx.x;
    unsigned long long my_len = 0;
    if (r < R) {
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
        padded[r] = ((int64_t)len + PAD - 1) / PAD * PAD;
        my_len = (unsigned long long)len;
        if (len > 0) {
            const uint32_t b0 = gs >> BLK_SHIFT, b1 = (gs + (uint32_t)len - 1u) >> BLK_SHIFT;
            for (uint32_t b = b0; b <= b1;) {
                const uint32_t w = b >> 5, hi = min(b1, (w << 5) + 31u);
                const uint32_t nb = hi - b + 1u;
                const uint32_t m = (nb == 32u ? 0xffffffffu : ((1u << nb) - 1u)) << (b & 31u);
                atomicOr(bitmap + w, m);
                b = hi + 1u;
            }
        }
    }
    unsigned long long my_max = my_len;
    for (int d = 16; d > 0; d >>= 1) {
        my_len += __shfl_xor_sync(0xffffffffu, my_len, d);
        my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, d));
    }
    if ((threadIdx.x & 31) == 0 && my_len) {
        atomicAdd(&pstats[0], my_len);
        atomicMax(&pstats[1], my_max);
    }
}

// bitmap -> (bits, set blocks before this word) per 32-block word; one CTA of 1024 threads
__global__ void __launch_bounds__(1024)
sp_rank_kernel(const uint32_t* __restrict__ bitmap, int words, uint2* __restrict__ tab,
               uint32_t* __restrict__ n_set) {
    __shared__ uint32_t wsum[32];
    const int t = threadIdx.x;
    constexpr int PER = MAX_WORDS / 1024;
    uint32_t bits[PER], c = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int w = t * PER + k;
        bits[k] = w < words ? bitmap[w] : 0u;
        c += __popc(bits[k]);
    }
    uint32_t inc = c;
    const unsigned lane = t & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) wsum[t >> 5] = inc;
    __syncthreads();
    uint32_t run = inc - c;
    for (int k = 0; k < (t >> 5); k++) run += wsum[k];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int w = t * PER + k;
        if (w < words) tab[w] = make_uint2(bits[k], run);
        run += __popc(bits[k]);
    }
    if (t == 1023) *n_set = run;
}

__device__ __forceinline__ int64_t sp_owner_of(const int64_t* __restrict__ off, int64_t R, int64_t t) {
    int64_t lo = 0, hi = R;
    while (hi - lo > 1) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(off + mid) <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

// rank of block b among the set blocks (for an unset block: the rank of the next set one)
__device__ __forceinline__ uint32_t sp_rank(const uint2* __restrict__ tab, uint32_t b, bool* set) {
    const uint2 w = __ldg(tab + (b >> 5));
    *set = (w.x >> (b & 31u)) & 1u;
    return w.y + __popc(w.x & ((1u << (b & 31u)) - 1u));
}

// Everything a tile kernel needs in one 32-byte record.
struct __align__(16) SpDesc {
    int64_t out;         // offset of the tile's first output in the dense coverage
    uint32_t c0;         // first candidate
    uint32_t n;          // candidates to look at
    int32_t tlen;        // outputs
    uint32_t cts;        // tile start in compacted coordinates, low P bits
    uint32_t flags;      // bit0 reverse, bits1..3 strand classes the region counts
    uint32_t region;
};

// One thread per tile: where it lies (tiles are cut in OUTPUT space, as in the other paths) and
// which candidates can reach it.  boff is read later (sp_range_kernel): it does not exist yet
// when the tiles are laid out, so the record keeps the sub-bin indices in c0 / n meanwhile.
__global__ void __launch_bounds__(CTA)
sp_tiles_kernel(int64_t R, int64_t Tb, int64_t Ts, const int64_t* __restrict__ off_big,
                const int64_t* __restrict__ off_small, const uint32_t* __restrict__ gs,
                const int32_t* __restrict__ plen, const uint8_t* __restrict__ flags,
                const int64_t* __restrict__ cov_off, const uint2* __restrict__ tab, uint32_t max_w,
                uint32_t pmask, SpDesc* __restrict__ desc) {
    const int64_t t = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (t >= Tb + Ts) return;
    int64_t r;
    uint32_t o_lo, tlen, tstart;
    if (t < Tb) {
        r = sp_owner_of(off_big, R, t);
        const int j = (int)(t - off_big[r]);
        const int L = plen[r];
        const int m = (L + TILE - 1) / TILE;
        const int tile_len = (((L + m - 1) / m) + ROW - 1) / ROW * ROW;
        o_lo = (uint32_t)(j * tile_len);
        tlen = (uint32_t)min(tile_len, L - (int)o_lo);
        const uint32_t q0 = (flags[r] & 1u) ? (uint32_t)L - o_lo - tlen : o_lo;
        tstart = gs[r] + q0;
    } else {
        r = sp_owner_of(off_small, R, t - Tb);
        o_lo = 0;
        tlen = (uint32_t)plen[r];
        tstart = gs[r];
    }
    // candidates that can reach [tstart, tstart + tlen): those homed from (tstart - widest + 1) on
    const uint32_t first = tstart >= max_w ? tstart - max_w + 1u : 0u;
    const uint32_t last = tstart + tlen - 1u;
    bool set;
    const uint32_t rf = sp_rank(tab, first >> BLK_SHIFT, &set);
    const uint32_t bin_lo = rf * SUBS + (set ? ((first & BLK_MASK) >> SUB_SHIFT) : 0u);
    const uint32_t rl = sp_rank(tab, last >> BLK_SHIFT, &set);          // a tile's own blocks are set
    const uint32_t bin_hi = rl * SUBS + ((last & BLK_MASK) >> SUB_SHIFT);
    const uint32_t rs = sp_rank(tab, tstart >> BLK_SHIFT, &set);
    SpDesc d;
    d.out = cov_off[r] + o_lo;
    d.c0 = bin_lo;
    d.n = bin_hi + 1u;
    d.tlen = (int32_t)tlen;
    d.cts = ((rs << BLK_SHIFT) | (tstart & BLK_MASK)) & pmask;
    d.flags = flags[r];
    d.region = (uint32_t)r;
    desc[t] = d;
}

// sub-bin indices -> candidate range
__global__ void __launch_bounds__(CTA)
sp_range_kernel(int64_t T, const uint32_t* __restrict__ boff, SpDesc* __restrict__ desc) {
    const int64_t t = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (t >= T) return;
    const uint32_t lo = desc[t].c0, hi1 = desc[t].n;
    const uint32_t c0 = __ldg(boff + lo), c1 = __ldg(boff + hi1);
    desc[t].c0 = c0;
    desc[t].n = c1 - c0;
}

// --------------------------------------------------------------------------------- split ------
// One read -> (group, packed word), or false when no block it touches is in the mask.  The read
// is clipped to the set blocks it touches (the part inside an unset block cannot reach any
// region), so that every candidate lies inside consecutive ranks of the compacted genome.
template <bool STRANDED>
__device__ __forceinline__ bool sp_classify(uint32_t s, uint32_t e1, int st, const uint2* tab,
                                            int gshift, int P, uint32_t* g, uint32_t* packed) {
    if (e1 <= s) return false;
    const uint32_t b0 = s >> BLK_SHIFT, b1 = (e1 - 1u) >> BLK_SHIFT;
    uint2 w = tab[b0 >> 5];
    bool set0 = (w.x >> (b0 & 31u)) & 1u;
    uint32_t home = b0;
    if (b1 != b0) {                                   // at most two blocks: width < 16384
        const uint2 w1 = tab[b1 >> 5];
        const bool set1 = (w1.x >> (b1 & 31u)) & 1u;
        if (set0) {
            if (!set1) e1 = b1 << BLK_SHIFT;
        } else {
            if (!set1) return false;
            s = b1 << BLK_SHIFT;
            home = b1;
            w = w1;
            set0 = true;
        }
    }
    if (!set0) return false;
    const uint32_t rank = w.y + __popc(w.x & ((1u << (home & 31u)) - 1u));
    *g = rank >> gshift;
    uint32_t word = ((rank & ((1u << gshift) - 1u)) << BLK_SHIFT) | (s & BLK_MASK);
    int sh = P;
    if (STRANDED) {
        word |= (st > 0 ? 0u : (st < 0 ? 1u : 2u)) << P;
        sh += 2;
    }
    *packed = word | ((e1 - s) << sh);
    return true;
}

struct SplitOut {
    uint32_t* pool;          // chunks of CH entries
    uint16_t* meta;          // per chunk: group << 5 | entries (0: unused slot)
    uint32_t* pool_next;     // chunks handed out so far (in SLABs)
};

template <bool STRANDED>
__global__ void __launch_bounds__(ST, 1)
sp_split_kernel(int64_t n, const uint32_t* __restrict__ g_start, const uint32_t* __restrict__ g_end1,
                const int8_t* __restrict__ strand, const uint2* __restrict__ tab_g, int words,
                int gshift, int P, SplitOut out) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(sp_smem);            // NG * RING
    uint32_t* cnt = ring + NG * RING;                                 // NG: ring start << 16 | fill
    uint2* tab = reinterpret_cast<uint2*>(cnt + NG);                  // words
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, lt = (1u << lane) - 1u;
    for (int i = tid; i < words; i += ST) tab[i] = tab_g[i];
    cnt[tid] = 0;
    __syncthreads();
    uint32_t slab_next = 0, slab_end = 0;          // this warp's share of the chunk pool

    // thread t owns group t here: every complete chunk of its ring leaves as one 64-byte store
    auto flush_phase = [&](bool final_pass) {
        const uint32_t w = cnt[tid];
        const uint32_t raw = w & 0xffffu, start = w >> 16;
        const uint32_t nn = min(raw, (uint32_t)RING);
        const uint32_t k = final_pass ? (nn > 0u ? 1u : 0u) : nn / CH;
        const unsigned m1 = __ballot_sync(0xffffffffu, k >= 1u), m2 = __ballot_sync(0xffffffffu, k >= 2u);
        const uint32_t total = __popc(m1) + __popc(m2);
        if (total) {                                // warp-uniform
            const uint32_t idx = __popc(m1 & lt) + __popc(m2 & lt);
            const uint32_t rem = slab_end - slab_next;
            uint32_t nb = 0;
            if (total > rem) {
                if (lane == 0) nb = atomicAdd(out.pool_next, (uint32_t)SLAB);
                nb = __shfl_sync(0xffffffffu, nb, 0);
            }
            for (uint32_t j = 0; j < k; j++) {
                const uint32_t i = idx + j;
                const uint32_t c = i < rem ? slab_next + i : nb + (i - rem);
                const uint4* src = reinterpret_cast<const uint4*>(ring + tid * RING + ((start + CH * j) & (RING - 1)));
                uint4* dst = reinterpret_cast<uint4*>(out.pool + (size_t)c * CH);
                const uint4 a = src[0], b = src[1], cc = src[2], d = src[3];
                __stcs(dst, a);
                __stcs(dst + 1, b);
                __stcs(dst + 2, cc);
                __stcs(dst + 3, d);
                out.meta[c] = (uint16_t)(((uint32_t)tid << 5) | (final_pass ? nn : (uint32_t)CH));
            }
            if (total > rem) {
                slab_next = nb + (total - rem);
                slab_end = nb + SLAB;
            } else {
                slab_next += total;
            }
        }
        if (k || raw > (uint32_t)RING)
            cnt[tid] = final_pass ? 0u : ((((start + CH * k) & (RING - 1)) << 16) | (nn - CH * k));
    };

    // up to four reads per thread and round; a read that finds its ring full waits for the flush
    auto insert_round = [&](const uint32_t* gk, const uint32_t* pk, uint32_t pend) {
        for (;;) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if ((pend >> k) & 1u) {
                    const uint32_t old = atomicAdd(&cnt[gk[k]], 1u);
                    const uint32_t slot = old & 0xffffu;
                    if (slot < (uint32_t)RING) {
                        ring[gk[k] * RING + (((old >> 16) + slot) & (RING - 1))] = pk[k];
                        pend &= ~(1u << k);
                    }
                }
            }
            const int any = __syncthreads_or(pend != 0u);
            flush_phase(false);
            __syncthreads();
            if (!any) break;
        }
    };

    const int64_t n_vec = n >> 2;
    const int64_t rounds = (n_vec + ST - 1) / ST;
    int64_t round = blockIdx.x;
    uint4 s4 = make_uint4(0, 0, 0, 0), e4 = make_uint4(0, 0, 0, 0);
    char4 t4 = make_char4(0, 0, 0, 0);
    auto load = [&](int64_t rd, uint4* s, uint4* e, char4* t) {
        const int64_t v = rd * ST + tid;
        *s = make_uint4(0, 0, 0, 0);
        *e = make_uint4(0, 0, 0, 0);
        *t = make_char4(0, 0, 0, 0);
        if (v < n_vec) {
            *s = __ldcs(reinterpret_cast<const uint4*>(g_start) + v);
            *e = __ldcs(reinterpret_cast<const uint4*>(g_end1) + v);
            if (STRANDED && strand) *t = __ldcs(reinterpret_cast<const char4*>(strand) + v);
        }
    };
    if (round < rounds) load(round, &s4, &e4, &t4);
    while (round < rounds) {
        const int64_t nr = round + gridDim.x;
        uint4 s4n, e4n;
        char4 t4n;
        if (nr < rounds) load(nr, &s4n, &e4n, &t4n);      // in flight while this round is split
        uint32_t gk[4], pk[4], pend = 0;
        if (sp_classify<STRANDED>(s4.x, e4.x, t4.x, tab, gshift, P, &gk[0], &pk[0])) pend |= 1u;
        if (sp_classify<STRANDED>(s4.y, e4.y, t4.y, tab, gshift, P, &gk[1], &pk[1])) pend |= 2u;
        if (sp_classify<STRANDED>(s4.z, e4.z, t4.z, tab, gshift, P, &gk[2], &pk[2])) pend |= 4u;
        if (sp_classify<STRANDED>(s4.w, e4.w, t4.w, tab, gshift, P, &gk[3], &pk[3])) pend |= 8u;
        insert_round(gk, pk, pend);
        round = nr;
        if (round < rounds) {
            s4 = s4n;
            e4 = e4n;
            t4 = t4n;
        }
    }
    if (blockIdx.x == 0) {              // the n % 4 tail
        uint32_t gk[4], pk[4], pend = 0;
        const int64_t i = n_vec * 4 + tid;
        if (i < n && sp_classify<STRANDED>(g_start[i], g_end1[i], (STRANDED && strand) ? (int)strand[i] : 0,
                                           tab, gshift, P, &gk[0], &pk[0]))
            pend = 1u;
        insert_round(gk, pk, pend);
    }
    flush_phase(true);                   // what is left in the rings: one partial chunk per group
}

// ---------------------------------------------------------------------- chunks by group -------
// Counting sort of the chunk tags: per-CTA histograms over a contiguous range of pool slots, a
// column-wise prefix by one CTA, then the placement with shared-memory cursors.
__device__ __forceinline__ void sp_slot_range(uint32_t n_slots, uint32_t* lo, uint32_t* hi) {
    const uint32_t per = ((n_slots + gridDim.x - 1) / gridDim.x + CS - 1) / CS * CS;
    *lo = min(n_slots, blockIdx.x * per);
    *hi = min(n_slots, *lo + per);
}

__global__ void __launch_bounds__(CS)
sp_chunk_hist_kernel(const uint16_t* __restrict__ meta, const uint32_t* __restrict__ pool_next,
                     uint32_t* __restrict__ col /* [gridDim.x][NG] */) {
    __shared__ uint32_t h[NG];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t lo, hi;
    sp_slot_range(*pool_next, &lo, &hi);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += CS) {
        const uint32_t m = meta[i];
        if (m & 31u) atomicAdd(&h[m >> 5], 1u);
    }
    __syncthreads();
    col[(size_t)blockIdx.x * NG + threadIdx.x] = h[threadIdx.x];
}

// one CTA, thread g = group g: cb[g] = first chunk-list entry of group g (cb[NG] = chunks in all);
// col[c][g] becomes the first entry CTA c writes for group g
__global__ void __launch_bounds__(NG)
sp_chunk_scan_kernel(uint32_t* __restrict__ col, int n_cols, uint32_t* __restrict__ cb) {
    __shared__ uint32_t wsum[NG / 32];
    const int g = threadIdx.x;
    uint32_t tot = 0;
    for (int c = 0; c < n_cols; c++) tot += col[(size_t)c * NG + g];
    uint32_t inc = tot;
    const unsigned lane = g & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) wsum[g >> 5] = inc;
    __syncthreads();
    uint32_t run = inc - tot;
    for (int k = 0; k < (g >> 5); k++) run += wsum[k];
    cb[g] = run;
    if (g == NG - 1) cb[NG] = run + tot;
    for (int c = 0; c < n_cols; c++) {
        const uint32_t v = col[(size_t)c * NG + g];
        col[(size_t)c * NG + g] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(CS)
sp_chunk_place_kernel(const uint16_t* __restrict__ meta, const uint32_t* __restrict__ pool_next,
                      const uint32_t* __restrict__ col, uint32_t* __restrict__ list /* slot << 5 | entries */) {
    __shared__ uint32_t cur[NG];
    cur[threadIdx.x] = col[(size_t)blockIdx.x * NG + threadIdx.x];
    __syncthreads();
    uint32_t lo, hi;
    sp_slot_range(*pool_next, &lo, &hi);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += CS) {
        const uint32_t m = meta[i];
        if (m & 31u) list[atomicAdd(&cur[m >> 5], 1u)] = (i << 5) | (m & 31u);
    }
}

// ------------------------------------------------------------------- candidates by sub-bin -----
// One CTA per group: histogram of the group's candidates over its sub-bins, prefix, placement.
// The group's output starts at CH * cb[g] (its chunk capacity: the partial chunks leave a hole of
// zero words -- width 0, never a hit -- at the end of the group's range).
__global__ void __launch_bounds__(GT)
sp_group_kernel(const uint32_t* __restrict__ pool, const uint32_t* __restrict__ list,
                const uint32_t* __restrict__ cb, int n_groups, int gshift, uint32_t pmask,
                uint32_t* __restrict__ cand, uint32_t* __restrict__ boff) {
    extern __shared__ uint32_t gsm[];
    const int nb = SUBS << gshift;                 // sub-bins of a group
    uint32_t* h = gsm;                             // nb counts, then cursors
    __shared__ uint32_t wsum[GT / 32];
    const int tid = threadIdx.x, g = blockIdx.x;
    for (int i = tid; i < nb; i += GT) h[i] = 0;
    __syncthreads();
    const uint32_t c0 = cb[g], c1 = cb[g + 1];
    const uint32_t sub = tid & (CH - 1), hw = tid / CH;           // half-warp = one chunk
    constexpr int HW = GT / CH, U = 4;                            // chunks in flight per half-warp
    for (uint32_t i0 = c0 + hw; i0 < c1; i0 += HW * U) {
        uint32_t le[U], e[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = i0 + (uint32_t)u * HW;
            le[u] = i < c1 ? __ldg(list + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            e[u] = sub < (le[u] & 31u) ? __ldcs(pool + (size_t)(le[u] >> 5) * CH + sub) : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < U; u++)
            if (sub < (le[u] & 31u)) atomicAdd(&h[(e[u] & pmask) >> SUB_SHIFT], 1u);
    }
    __syncthreads();
    // exclusive prefix of the nb counts (nb / GT consecutive bins per thread, nb <= 2048)
    const int per = (nb + GT - 1) / GT;
    uint32_t c[4], mine = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b = tid * per + q;
        c[q] = (q < per && b < nb) ? h[b] : 0u;
        mine += c[q];
    }
    uint32_t inc = mine;
    const unsigned lane = tid & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    uint32_t run = inc - mine;
    for (int k = 0; k < (tid >> 5); k++) run += wsum[k];
    const uint32_t base = c0 * CH;
    uint32_t total = 0;
    for (int k = 0; k < GT / 32; k++) total += wsum[k];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b = tid * per + q;
        if (q < per && b < nb) {
            h[b] = run;                                         // cursor
            boff[(size_t)g * nb + b] = base + run;
            run += c[q];
        }
    }
    if (g == n_groups - 1 && tid == 0) boff[(size_t)n_groups * nb] = base + total;
    // the hole between this group's candidates and the next group's first
    for (uint32_t i = base + total + tid; i < c1 * CH; i += GT) cand[i] = 0u;
    __syncthreads();
    for (uint32_t i0 = c0 + hw; i0 < c1; i0 += HW * U) {
        uint32_t le[U], e[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = i0 + (uint32_t)u * HW;
            le[u] = i < c1 ? __ldg(list + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            e[u] = sub < (le[u] & 31u) ? __ldcs(pool + (size_t)(le[u] >> 5) * CH + sub) : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < U; u++)
            if (sub < (le[u] & 31u)) cand[base + atomicAdd(&h[(e[u] & pmask) >> SUB_SHIFT], 1u)] = e[u];
    }
}

// --------------------------------------------------------------------------------- tiles ------
__device__ __forceinline__ SpDesc sp_load_desc(const SpDesc* __restrict__ p) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(p));
    const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
    SpDesc d;
    d.out = (int64_t)(((uint64_t)(uint32_t)a.y << 32) | (uint32_t)a.x);
    d.c0 = (uint32_t)a.z;
    d.n = (uint32_t)a.w;
    d.tlen = b.x;
    d.cts = (uint32_t)b.y;
    d.flags = (uint32_t)b.z;
    d.region = (uint32_t)b.w;
    return d;
}

// One candidate against one tile: adds its two events to the difference array (mirrored on '-'
// regions, so the scan only ever runs forwards).  Returns true on a hit.
template <bool STRANDED>
__device__ __forceinline__ bool sp_apply(uint32_t e, const SpDesc& d, int P, int* diff) {
    const int sh = 32 - P;
    const int rel = ((int)((e - d.cts) << sh)) >> sh;              // candidate start - tile start
    uint32_t w = e >> (STRANDED ? P + 2 : P);
    if (STRANDED && !((d.flags >> (1u + ((e >> P) & 3u))) & 1u)) w = 0;
    int lo = max(rel, 0), hi = min(rel + (int)w, d.tlen);
    if (lo >= hi) return false;
    if (d.flags & 1u) {
        const int l2 = d.tlen - hi;
        hi = d.tlen - lo;
        lo = l2;
    }
    atomicAdd(diff + lo, 1);
    if (hi < d.tlen) atomicSub(diff + hi, 1);
    return true;
}

template <int RPW, bool STRANDED>
__device__ __forceinline__ bool sp_tile_body(int* diff, int* wtot, const SpDesc& d,
                                             const uint32_t* __restrict__ cand, int P,
                                             int32_t* __restrict__ dst) {
    constexpr int B = 4;
    const uint32_t tid = threadIdx.x;
    const uint32_t n = d.n;
    const uint32_t* c = cand + d.c0;
    uint32_t e[B];
    auto load = [&](uint32_t i0) {
#pragma unroll
        for (int k = 0; k < B; k++) {
            const uint32_t i = i0 + (uint32_t)k * CTA + tid;
            e[k] = i < n ? __ldg(c + i) : 0u;               // 0: width 0, never a hit
        }
    };
    load(0);
#pragma unroll
    for (int k = 0; k < RPW; k++)
        reinterpret_cast<int4*>(diff)[k * CTA + tid] = make_int4(0, 0, 0, 0);
    __syncthreads();
    bool hit = false;
    for (uint32_t i0 = 0; i0 < n; i0 += B * CTA) {
        if (i0) load(i0);
#pragma unroll
        for (int k = 0; k < B; k++) hit |= sp_apply<STRANDED>(e[k], d, P, diff);
    }
    const bool any = __syncthreads_or(hit) != 0;
    block_scan_store_fwd<RPW, true>(diff, d.tlen, wtot, dst);
    return any;
}

template <bool STRANDED>
__device__ __forceinline__ bool sp_tile_dispatch(int* diff, int* wtot, const SpDesc& d,
                                                 const uint32_t* __restrict__ cand, int P,
                                                 int32_t* __restrict__ dst) {
    const int rpw = ((d.tlen + ROW - 1) / ROW + WARPS - 1) / WARPS;
    switch (rpw) {
        case 1: return sp_tile_body<1, STRANDED>(diff, wtot, d, cand, P, dst);
        case 2: return sp_tile_body<2, STRANDED>(diff, wtot, d, cand, P, dst);
        case 3: return sp_tile_body<3, STRANDED>(diff, wtot, d, cand, P, dst);
        case 4: return sp_tile_body<4, STRANDED>(diff, wtot, d, cand, P, dst);
        case 5: return sp_tile_body<5, STRANDED>(diff, wtot, d, cand, P, dst);
        case 6: return sp_tile_body<6, STRANDED>(diff, wtot, d, cand, P, dst);
        default: return sp_tile_body<7, STRANDED>(diff, wtot, d, cand, P, dst);
    }
}

template <bool STRANDED>
__global__ void __launch_bounds__(CTA, 5)
sp_tile_kernel(int64_t Tb, const SpDesc* __restrict__ desc, const uint32_t* __restrict__ cand, int P,
               int32_t* __restrict__ cov, uint8_t* __restrict__ region_hit) {
    __shared__ __align__(16) int diff[TILE];
    __shared__ int wtot[WARPS];
    const uint32_t tid = threadIdx.x;
    const int64_t step = gridDim.x;
    int64_t t = blockIdx.x;
    if (t >= Tb) return;
    SpDesc d = sp_load_desc(desc + t);
    for (;;) {
        SpDesc dn;
        dn.tlen = 0;
        if (t + step < Tb) dn = sp_load_desc(desc + t + step);
        const bool any = sp_tile_dispatch<STRANDED>(diff, wtot, d, cand, P, cov + d.out);
        if (any && tid == 0) region_hit[d.region] = 1;
        t += step;
        if ((tid & 31u) == 0) tma_store_wait_read();
        if (t >= Tb) break;
        __syncthreads();
        d = dn;
    }
}

template <bool STRANDED>
__global__ void __launch_bounds__(CTA)
sp_small_kernel(int64_t Ts, const SpDesc* __restrict__ desc, const uint32_t* __restrict__ cand, int P,
                int32_t* __restrict__ cov, uint8_t* __restrict__ region_hit) {
    __shared__ __align__(16) int sm[WARPS][SMALL_MAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ts_i = (int64_t)blockIdx.x * WARPS + warp;
    if (ts_i >= Ts) return;
    const SpDesc d = sp_load_desc(desc + ts_i);
    const int L = d.tlen;
    int* diff = sm[warp];
    const int nrows = (L + ROW - 1) / ROW;
    for (int i = lane; i < nrows * (ROW / 4); i += 32)
        reinterpret_cast<int4*>(diff)[i] = make_int4(0, 0, 0, 0);
    __syncwarp();
    bool hit = false;
    for (uint32_t i = lane; i < d.n; i += 32) hit |= sp_apply<STRANDED>(__ldg(cand + d.c0 + i), d, P, diff);
    hit = __any_sync(0xffffffffu, hit);
    __syncwarp();
    if (hit && lane == 0) region_hit[d.region] = 1;
    int32_t* dst = cov + d.out;
    int pre = 0;
    for (int row = 0; row < nrows; row++)
        pre += warp_row_scan_store(diff + row * ROW, pre, 0, false, L - row * ROW, dst + row * ROW);
}

// NULL rule (coverage.R:198,224-225): no overlapping read in any tile of the region.
__global__ void __launch_bounds__(CTA)
sp_null_kernel(int64_t R, const int32_t* __restrict__ plen, const uint8_t* __restrict__ region_hit,
               const uint32_t* __restrict__ cb, int32_t* __restrict__ len, uint8_t* __restrict__ is_null,
               unsigned long long* __restrict__ stats /* n_null, total_len, max_len, candidates */) {
    const int64_t r = (int64_t)blockIdx.x * CTA + threadIdx.x;
    unsigned long long my_null = 0, my_len = 0;
    if (r < R) {
        const bool null = plen[r] == 0 || region_hit[r] == 0;
        const int32_t out = null ? 0 : plen[r];
        len[r] = out;
        is_null[r] = null ? 1 : 0;
        my_null = null ? 1 : 0;
        my_len = (unsigned long long)out;
    }
    unsigned long long my_max = my_len;
    for (int d = 16; d > 0; d >>= 1) {
        my_null += __shfl_xor_sync(0xffffffffu, my_null, d);
        my_len += __shfl_xor_sync(0xffffffffu, my_len, d);
        my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_null) atomicAdd(&stats[0], my_null);
        if (my_len) atomicAdd(&stats[1], my_len);
        if (my_max) atomicMax(&stats[2], my_max);
    }
    if (r == 0) stats[3] = (unsigned long long)cb[NG] * CH;     // chunk capacity: >= the candidates
}

}  // namespace

// Same contract as coverage_ranges_bucketed.  RCP_SPLIT_NOT_APPLICABLE: the reads are too wide for
// the packed candidate word of this mask (or the genome too long for the shared-memory table);
// nothing has been produced and the caller uses another path.
int coverage_ranges_split(ReadsIdx& rd, int64_t R, const int32_t* chrom, const int32_t* start,
                          const int32_t* end, const int8_t* strand, int ignore_strand,
                          int strand_filter, int mem, Coverage* cv) {
    const bool stranded = !((strand_filter == RCP_STRAND_ANY) && (ignore_strand || strand == nullptr));
    const bool st_arr = stranded && rd.d_strand != nullptr;      // strandless reads are all '*'
    const int64_t span = (int64_t)rd.chrom_off[(size_t)rd.n_chrom];
    const int words = (int)(((span >> BLK_SHIFT) + 32) / 32);
    if (words > MAX_WORDS || rd.n >= 0x7ffffff0ll) return RCP_SPLIT_NOT_APPLICABLE;

    DevIn<int32_t> d_chrom, d_start, d_end;
    DevIn<int8_t> d_strand;
    RCP_TRY(d_chrom.init(chrom, (size_t)R, mem));
    RCP_TRY(d_start.init(start, (size_t)R, mem));
    RCP_TRY(d_end.init(end, (size_t)R, mem));
    RCP_TRY(d_strand.init(strand, (size_t)R, mem));

    // ---- plan -------------------------------------------------------------------------------
    Arena A;
    const size_t r = (size_t)R;
    RCP_TRY(A.reserve(Arena::pad(64) + Arena::pad((size_t)words * 4) + Arena::pad((size_t)words * 8) +
                      Arena::pad(r * 4) * 2 + Arena::pad(r) * 2 + Arena::pad(r * 8) * 3 +
                      Arena::pad((r + 1) * 8) * 2));
    // zero-initialised block first (ONE memset): status words, bitmap, region hit flags
    unsigned int* err = A.take<unsigned int>(16);         // [0] err [1] n_set; pstats at +8 bytes x2
    unsigned long long* pstats = reinterpret_cast<unsigned long long*>(err + 4);   // [0] total len [1] max len
    uint32_t* n_set = err + 1;
    uint32_t* bitmap = A.take<uint32_t>((size_t)words);
    uint8_t* region_hit = A.take<uint8_t>(r);
    const size_t zero_bytes = A.used;
    uint2* tab = A.take<uint2>((size_t)words);
    uint32_t* gs = A.take<uint32_t>(r);
    int32_t* plen = A.take<int32_t>(r);
    uint8_t* flags = A.take<uint8_t>(r);
    int64_t* nbig = A.take<int64_t>(r);
    int64_t* nsmall = A.take<int64_t>(r);
    int64_t* padded = A.take<int64_t>(r);
    int64_t* off_big = A.take<int64_t>(r + 1);
    int64_t* off_small = A.take<int64_t>(r + 1);
    if (A.used > A.cap) return fail(RCP_ERR_CUDA, "internal: split plan arena overrun");

    cv->n_regions = R;
    RCP_TRY(dalloc(&cv->off, r + 1));
    RCP_TRY(dalloc(&cv->len, r));
    RCP_TRY(dalloc(&cv->is_null, r));
    RCP_TRY(dalloc(&cv->d_stats, 4));
    {
        StageTimer t(ST_SP_PLAN);
        RCP_CUDA(cudaMemsetAsync(A.base, 0, zero_bytes, g_ctx.stream));
        RCP_CUDA(cudaMemsetAsync(cv->d_stats, 0, 32, g_ctx.stream));
        if (R > 0) {
            sp_plan_kernel<<<blocks_for(R, CTA), CTA, 0, g_ctx.stream>>>(
                R, d_chrom.ptr, d_start.ptr, d_end.ptr, d_strand.ptr, rd.d_chrom_off, rd.d_chrom_len,
                rd.n_chrom, ignore_strand, strand_filter, gs, plen, flags, nbig, nsmall, padded, bitmap,
                err, pstats);
            RCP_LAUNCHED();
        }
        RCP_TRY(exclusive_scan2_i64(nbig, off_big, off_big + R, nsmall, off_small, off_small + R, R));
        RCP_TRY(exclusive_scan_i64(padded, cv->off, R, cv->off + R));
        sp_rank_kernel<<<1, 1024, 0, g_ctx.stream>>>(bitmap, words, tab, n_set);
        RCP_LAUNCHED();
    }
    struct Host {
        int64_t Tb, Ts, total_padded;
        unsigned long long pstats[2];
        unsigned int err[2];
    } h = {0, 0, 0, {0, 0}, {0, 0}};
    {
        FetchItem items[8] = {{off_big + R, &h.Tb, 8}, {off_small + R, &h.Ts, 8}, {cv->off + R, &h.total_padded, 8},
                              {pstats, h.pstats, 16}, {err, h.err, 8}};
        int n_items = 5;
        reads_pending_items(rd, items, &n_items);       // a deferred rcp_reads_load is validated here
        RCP_TRY(fetch_and_sync(items, n_items));
        RCP_TRY(reads_finish(rd));
    }
    if (h.err[0] & 1u) return fail(RCP_ERR_DATA, "a region has a chromosome id outside [0, n_chrom)");
    if (h.err[0] & 2u) return fail(RCP_ERR_DATA, "a region has end < start - 1");
    const int64_t Tb = h.Tb, Ts = h.Ts, T = Tb + Ts;
    if (T > 0x7fffffff) return fail(RCP_ERR_UNSUPPORTED, "more than 2^31-1 coverage tiles");
    const uint32_t set_blocks = h.err[1];
    int gshift = MIN_GSHIFT;
    while (gshift < MAX_GSHIFT && (((int64_t)set_blocks + (1ll << gshift) - 1) >> gshift) > NG) gshift++;
    const int n_groups = (int)std::max<int64_t>(1, ((int64_t)set_blocks + (1ll << gshift) - 1) >> gshift);
    const int P = gshift + BLK_SHIFT;
    const int wbits = 32 - P - (stranded ? 2 : 0);
    const uint32_t max_w = rd.max_width > 0 ? rd.max_width : 1u;
    if (n_groups > NG || max_w > 8191u || max_w >= (1u << wbits)) {
        dfree(cv->off);
        dfree(cv->len);
        dfree(cv->is_null);
        dfree(cv->d_stats);
        return RCP_SPLIT_NOT_APPLICABLE;
    }
    const uint32_t pmask = (1u << P) - 1u;
    const int nb = SUBS << gshift;

    cv->path = RCP_PATH_SPLIT;
    cv->total_padded = h.total_padded;
    cv->total_len = (int64_t)h.pstats[0];       // upper bounds until the NULL rule has run
    cv->max_len = (int32_t)h.pstats[1];
    cv->n_null = 0;
    cv->stats_pending = true;
    RCP_TRY(dalloc(&cv->cov, (size_t)h.total_padded));

    // ---- scratch of the read passes ------------------------------------------------------------
    const int split_grid = g_ctx.sm_count;
    const int sort_grid = g_ctx.sm_count;
    const size_t pool_cap = (size_t)(rd.n / CH + 1) + (size_t)split_grid * NG +
                            (size_t)split_grid * (ST / 32) * SLAB + SLAB;
    Arena B;
    RCP_TRY(B.reserve(Arena::pad(pool_cap * CH * 4) * 2 + Arena::pad(pool_cap * 2) + Arena::pad(pool_cap * 4) +
                      Arena::pad(16) + Arena::pad((size_t)sort_grid * NG * 4) + Arena::pad((NG + 1) * 4) +
                      Arena::pad(((size_t)n_groups * nb + 1) * 4) + Arena::pad((size_t)(T + 1) * sizeof(SpDesc))));
    uint32_t* pool_next = B.take<uint32_t>(4);
    uint16_t* meta = B.take<uint16_t>(pool_cap);
    const size_t zero_b = B.used;
    uint32_t* pool = B.take<uint32_t>(pool_cap * CH);
    uint32_t* cand = B.take<uint32_t>(pool_cap * CH);
    uint32_t* list = B.take<uint32_t>(pool_cap);
    uint32_t* col = B.take<uint32_t>((size_t)sort_grid * NG);
    uint32_t* cb = B.take<uint32_t>(NG + 1);
    uint32_t* boff = B.take<uint32_t>((size_t)n_groups * nb + 1);
    SpDesc* desc = B.take<SpDesc>((size_t)T + 1);
    if (B.used > B.cap) return fail(RCP_ERR_CUDA, "internal: split arena overrun");
    {
        StageTimer t(ST_SP_PLAN);
        RCP_CUDA(cudaMemsetAsync(B.base, 0, zero_b, g_ctx.stream));
        if (T > 0) {
            sp_tiles_kernel<<<blocks_for(T, CTA), CTA, 0, g_ctx.stream>>>(
                R, Tb, Ts, off_big, off_small, gs, plen, flags, cv->off, tab, max_w, pmask, desc);
            RCP_LAUNCHED();
        }
    }
    {
        StageTimer t(ST_SP_SPLIT);
        const size_t smem = (size_t)NG * RING * 4 + (size_t)NG * 4 + (size_t)words * 8;
        SplitOut out = {pool, meta, pool_next};
        if (st_arr) {
            RCP_CUDA(cudaFuncSetAttribute(sp_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sp_split_kernel<true><<<split_grid, ST, smem, g_ctx.stream>>>(rd.n, rd.g_start, rd.g_end1, rd.d_strand,
                                                                         tab, words, gshift, P, out);
        } else if (stranded) {     // no strand array: every read is '*'
            RCP_CUDA(cudaFuncSetAttribute(sp_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sp_split_kernel<true><<<split_grid, ST, smem, g_ctx.stream>>>(rd.n, rd.g_start, rd.g_end1, nullptr, tab,
                                                                         words, gshift, P, out);
        } else {
            RCP_CUDA(cudaFuncSetAttribute(sp_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sp_split_kernel<false><<<split_grid, ST, smem, g_ctx.stream>>>(rd.n, rd.g_start, rd.g_end1, nullptr, tab,
                                                                          words, gshift, P, out);
        }
        RCP_LAUNCHED();
    }
    {
        StageTimer t(ST_SP_SORT);
        sp_chunk_hist_kernel<<<sort_grid, CS, 0, g_ctx.stream>>>(meta, pool_next, col);
        RCP_LAUNCHED();
        sp_chunk_scan_kernel<<<1, NG, 0, g_ctx.stream>>>(col, sort_grid, cb);
        RCP_LAUNCHED();
        sp_chunk_place_kernel<<<sort_grid, CS, 0, g_ctx.stream>>>(meta, pool_next, col, list);
        RCP_LAUNCHED();
        sp_group_kernel<<<n_groups, GT, (size_t)nb * 4, g_ctx.stream>>>(pool, list, cb, n_groups, gshift, pmask,
                                                                      cand, boff);
        RCP_LAUNCHED();
    }
    if (T > 0) {
        {
            StageTimer t(ST_SP_PLAN);
            sp_range_kernel<<<blocks_for(T, CTA), CTA, 0, g_ctx.stream>>>(T, boff, desc);
            RCP_LAUNCHED();
        }
        if (Tb > 0) {
            StageTimer t(ST_SP_TILE);
            int per_sm = 0;
            if (stranded) {
                RCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sp_tile_kernel<true>, CTA, 0));
                per_sm = std::max(per_sm, 1);
                sp_tile_kernel<true><<<(unsigned)std::min<int64_t>(Tb, (int64_t)g_ctx.sm_count * per_sm), CTA, 0,
                                       g_ctx.stream>>>(Tb, desc, cand, P, cv->cov, region_hit);
            } else {
                RCP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sp_tile_kernel<false>, CTA, 0));
                per_sm = std::max(per_sm, 1);
                sp_tile_kernel<false><<<(unsigned)std::min<int64_t>(Tb, (int64_t)g_ctx.sm_count * per_sm), CTA, 0,
                                        g_ctx.stream>>>(Tb, desc, cand, P, cv->cov, region_hit);
            }
            RCP_LAUNCHED();
        }
        if (Ts > 0) {
            StageTimer t(ST_SP_SMALL);
            if (stranded)
                sp_small_kernel<true><<<blocks_for(Ts, WARPS), CTA, 0, g_ctx.stream>>>(Ts, desc + Tb, cand, P, cv->cov,
                                                                                      region_hit);
            else
                sp_small_kernel<false><<<blocks_for(Ts, WARPS), CTA, 0, g_ctx.stream>>>(Ts, desc + Tb, cand, P, cv->cov,
                                                                                       region_hit);
            RCP_LAUNCHED();
        }
    }
    if (R > 0) {
        StageTimer t(ST_SP_PLAN);
        sp_null_kernel<<<blocks_for(R, CTA), CTA, 0, g_ctx.stream>>>(R, plen, region_hit, cb, cv->len, cv->is_null,
                                                                    cv->d_stats);
        RCP_LAUNCHED();
    }
    return RCP_OK;
}

}  // namespace rcp
