// Read import, the decode half: what readBam / readBed of the reference
// (/root/reference/R/ranges.R:111-146) hand to the hot path, produced on the device.
//
//   BAM   readGAlignments(file) then as(., "GRanges") (spliceAction keep / remove) or
//         unlist(grglist(.)) (spliceAction split), then trim().  The caller inflates the BGZF
//         blocks (host zlib) and passes the alignment-record section; the record chain is walked
//         once on the host (rcp_bam_index: the records are length-prefixed), everything else -- flag
//         filter, CIGAR walk, N-split, trim, compaction -- is one thread per record on the device.
//   BED   trim(import.bed(file, trackLine = FALSE)): text lines -> chrom id, start + 1, end, strand.
//         Line starts by a counting pass + prefix sum over the bytes, then one thread per line.
//
// Both run count -> prefix sum -> write, so that the output keeps the file order (the seeded
// down-sampling of preprocessRanges indexes the reads by position).  The result stays on the
// device behind a `decoded` handle: rcp_reads_load_decoded feeds it to the hot path without a
// host round trip, rcp_decoded_fetch hands the arrays back.
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>

#include "cov_common.cuh"
#include "rcp_internal.cuh"

namespace rcp {

using namespace covk;

namespace {

struct Decoded {
    int64_t n = 0;
    int32_t* chrom = nullptr;
    int32_t* start = nullptr;
    int32_t* end = nullptr;
    int8_t* strand = nullptr;
};
std::map<int, std::unique_ptr<Decoded>> g_decoded;
int g_next_decoded = 1 << 20;

void decoded_release(Decoded& d) {
    dfree(d.chrom);
    dfree(d.start);
    dfree(d.end);
    dfree(d.strand);
    d.n = 0;
}

struct Out {
    int32_t* chrom;
    int32_t* start;
    int32_t* end;
    int8_t* strand;
};

__device__ __forceinline__ uint32_t ld_u32(const uint8_t* __restrict__ p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t ld_u16(const uint8_t* __restrict__ p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
}

// trim() of a GRanges with known seqlengths: [max(start, 1), min(end, L)]; a range that lies
// wholly outside becomes the zero-width range at the boundary it left by
__device__ __forceinline__ void trim_range(int64_t s, int64_t e, int64_t L, int32_t* so, int32_t* eo) {
    int64_t s2 = max(s, (int64_t)1), e2 = min(e, L);
    if (e2 < s2 - 1) {
        if (s > L) {
            s2 = L + 1;
            e2 = L;
        } else {
            e2 = s2 - 1;
        }
    }
    *so = (int32_t)s2;
    *eo = (int32_t)e2;
}

// ------------------------------------------------------------------------------- BAM ----------
// err bits: 1 record chain / truncated record, 2 refID outside the header's references,
//           4 CIGAR operation code > 8, 8 mapped record without a CIGAR
// One thread per alignment record.  WRITE = false: counts[r] = ranges the record yields.
template <bool WRITE>
__global__ void __launch_bounds__(CTA)
bam_records_kernel(const uint8_t* __restrict__ rec, int64_t n_bytes, const int64_t* __restrict__ off, int64_t n,
                   int n_ref, const int64_t* __restrict__ ref_len, int split, uint32_t* __restrict__ counts,
                   const uint32_t* __restrict__ where, Out out, unsigned int* __restrict__ err) {
    const int64_t r = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (r >= n) return;
    const int64_t o = off[r], o1 = off[r + 1];
    uint32_t cnt = 0;
    do {
        if (o < 0 || o1 > n_bytes || o1 - o < 36) {
            atomicOr(err, 1u);
            break;
        }
        const uint8_t* p = rec + o;
        if ((int64_t)ld_u32(p) + 4 != o1 - o) {
            atomicOr(err, 1u);
            break;
        }
        const int ref = (int)ld_u32(p + 4), pos = (int)ld_u32(p + 8);
        const uint32_t l_name = p[12], n_cig = ld_u16(p + 16), flag = ld_u16(p + 18);
        if ((flag & 4u) || ref < 0 || pos < 0) break;          // unmapped: readGAlignments drops it
        if (ref >= n_ref) {
            atomicOr(err, 2u);
            break;
        }
        if (n_cig == 0) {
            atomicOr(err, 8u);
            break;
        }
        const int64_t cig = 36 + (int64_t)l_name;
        if (cig + 4 * (int64_t)n_cig > o1 - o) {
            atomicOr(err, 1u);
            break;
        }
        const int64_t L = ref_len[ref];
        const int8_t st = (flag & 16u) ? -1 : 1;
        uint32_t w = WRITE ? where[r] : 0u;
        int64_t cur = (int64_t)pos + 1, blk = 0;          // open block [cur, cur + blk)
        bool bad = false;
        auto emit = [&]() {
            if (WRITE) {
                trim_range(cur, cur + blk - 1, L, out.start + w, out.end + w);
                out.chrom[w] = ref;
                out.strand[w] = st;
                w++;
            }
            cnt++;
        };
        for (uint32_t k = 0; k < n_cig; k++) {
            const uint32_t c = ld_u32(p + cig + 4 * k), op = c & 15u;
            const int64_t len = (int64_t)(c >> 4);
            if (op > 8u) bad = true;
            if (op == 0u || op == 2u || op == 7u || op == 8u) {      // M D = X
                blk += len;
            } else if (op == 3u) {                                      // N
                if (split) {
                    if (blk > 0) emit();
                    cur += blk + len;
                    blk = 0;
                } else {
                    blk += len;
                }
            }
        }
        if (bad) {
            atomicOr(err, 4u);
            cnt = 0;
            break;
        }
        // as(., "GRanges") keeps an alignment of reference width 0; grglist drops empty ranges
        if (blk > 0 || !split) emit();
    } while (false);
    if (!WRITE) counts[r] = cnt;
}

// ------------------------------------------------------------------------------- BED ----------
__device__ __forceinline__ int nl_count16(uint4 v, int valid) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int c = 0;
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (k < valid && ((w[k >> 2] >> ((k & 3) * 8)) & 0xffu) == (uint32_t)'\n') c++;
    return c;
}

// thread i looks at the bytes [16 i, 16 i + 16); block totals of the newline count
__global__ void __launch_bounds__(CTA)
bed_nl_count_kernel(const uint4* __restrict__ text, int64_t n_bytes, uint32_t* __restrict__ block_cnt) {
    __shared__ int wsum[WARPS];
    const int64_t i = (int64_t)blockIdx.x * CTA + threadIdx.x;
    const int64_t lo = i * 16;
    int c = 0;
    if (lo < n_bytes) c = nl_count16(__ldg(text + i), (int)min((int64_t)16, n_bytes - lo));
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < WARPS; k++) t += wsum[k];
        block_cnt[blockIdx.x] = (uint32_t)t;
    }
}

// starts[0] = 0, starts[k] = position after the k-th newline
__global__ void __launch_bounds__(CTA)
bed_nl_write_kernel(const uint4* __restrict__ text, int64_t n_bytes, const uint32_t* __restrict__ block_base,
                    int64_t* __restrict__ starts) {
    __shared__ int wsum[WARPS];
    const int64_t i = (int64_t)blockIdx.x * CTA + threadIdx.x;
    const int64_t lo = i * 16;
    const int valid = lo < n_bytes ? (int)min((int64_t)16, n_bytes - lo) : 0;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (valid) v = __ldg(text + i);
    const int c = nl_count16(v, valid);
    const int lane = threadIdx.x & 31;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    int base = inc - c;
    for (int k = 0; k < (int)(threadIdx.x >> 5); k++) base += wsum[k];
    int64_t at = (int64_t)block_base[blockIdx.x] + base + 1;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (k < valid && ((w[k >> 2] >> ((k & 3) * 8)) & 0xffu) == (uint32_t)'\n') starts[at++] = lo + k + 1;
    if (i == 0) starts[0] = 0;
}

struct NameTable {
    const char* blob;        // the names, sorted, one after the other
    const int32_t* name_off; // n + 1
    const int32_t* id;       // seqlevel index of the sorted name
    int n;
};

__device__ __forceinline__ bool is_sep(uint8_t c) { return c == '\t' || c == ' '; }

__device__ __forceinline__ int name_cmp(const uint8_t* a, int la, const char* b, int lb) {
    const int m = min(la, lb);
    for (int k = 0; k < m; k++) {
        const int d = (int)a[k] - (int)(uint8_t)b[k];
        if (d) return d;
    }
    return la - lb;
}

// err bits: 1 chromosome name not among the seqlevels, 2 a data line with fewer than three
//           fields or a bad number, 4 a strand field other than + - . *
// One thread per line.  Skipped: empty lines, comments (#), "track" and "browser" lines.
template <bool WRITE>
__global__ void __launch_bounds__(CTA)
bed_lines_kernel(const uint8_t* __restrict__ text, int64_t n_bytes, const int64_t* __restrict__ starts,
                 int64_t n_nl, int64_t n_lines, NameTable names, uint32_t* __restrict__ counts,
                 const uint32_t* __restrict__ where, Out out, unsigned int* __restrict__ err) {
    const int64_t j = (int64_t)blockIdx.x * CTA + threadIdx.x;
    if (j >= n_lines) return;
    int64_t a = starts[j], b = j < n_nl ? starts[j + 1] - 1 : n_bytes;
    if (b > a && text[b - 1] == '\r') b--;
    while (a < b && is_sep(text[a])) a++;
    uint32_t cnt = 0;
    do {
        if (a >= b || text[a] == '#') break;
        int64_t f0 = a, f1 = a;
        while (f1 < b && !is_sep(text[f1])) f1++;
        const int l0 = (int)(f1 - f0);
        if ((l0 == 5 && name_cmp(text + f0, 5, "track", 5) == 0) ||
            (l0 == 7 && name_cmp(text + f0, 7, "browser", 7) == 0))
            break;
        int lo = 0, hi = names.n - 1, found = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int c = name_cmp(text + f0, l0, names.blob + names.name_off[mid],
                                   names.name_off[mid + 1] - names.name_off[mid]);
            if (c == 0) {
                found = mid;
                break;
            }
            if (c < 0) hi = mid - 1;
            else lo = mid + 1;
        }
        if (found < 0) {
            atomicOr(err, 1u);
            break;
        }
        int64_t num[2] = {0, 0};
        bool bad = false;
        int64_t q = f1;
        for (int f = 0; f < 2; f++) {
            while (q < b && is_sep(text[q])) q++;
            const int64_t q0 = q;
            int64_t v = 0;
            while (q < b && !is_sep(text[q])) {
                const int dgt = (int)text[q] - '0';
                if (dgt < 0 || dgt > 9 || v > 0x7fffffff) bad = true;
                v = v * 10 + dgt;
                q++;
            }
            if (q == q0 || v > 0x7ffffffe) bad = true;
            num[f] = v;
        }
        if (bad) {
            atomicOr(err, 2u);
            break;
        }
        int8_t st = 0;
        for (int f = 3; f <= 5; f++) {              // name, score, strand
            while (q < b && is_sep(text[q])) q++;
            const int64_t q0 = q;
            while (q < b && !is_sep(text[q])) q++;
            if (f == 5 && q > q0) {
                const uint8_t c = text[q0];
                if (q - q0 != 1 || !(c == '+' || c == '-' || c == '.' || c == '*')) atomicOr(err, 4u);
                st = c == '+' ? 1 : (c == '-' ? -1 : 0);
            }
        }
        if (WRITE) {
            const uint32_t w = where[j];
            out.chrom[w] = names.id[found];
            out.start[w] = (int32_t)(num[0] + 1);   // BED starts are 0-based, ends exclusive
            out.end[w] = (int32_t)num[1];
            out.strand[w] = st;
        }
        cnt = 1;
    } while (false);
    if (!WRITE) counts[j] = cnt;
}

int new_decoded(Decoded** d, int* h) {
    *h = g_next_decoded++;
    g_decoded[*h] = std::unique_ptr<Decoded>(new Decoded());
    *d = g_decoded[*h].get();
    return RCP_OK;
}

int alloc_out(Decoded& d, int64_t n) {
    d.n = n;
    RCP_TRY(dalloc(&d.chrom, (size_t)n));
    RCP_TRY(dalloc(&d.start, (size_t)n));
    RCP_TRY(dalloc(&d.end, (size_t)n));
    RCP_TRY(dalloc(&d.strand, (size_t)n));
    return RCP_OK;
}

}  // namespace

void decoded_release_all() {
    for (auto& kv : g_decoded) decoded_release(*kv.second);
    g_decoded.clear();
}

}  // namespace rcp

using namespace rcp;

extern "C" {

int rcp_bam_index(const uint8_t* rec, int64_t n_bytes, int64_t* n_records_out, int64_t* offsets_out,
                  int64_t capacity) {
    if (n_bytes < 0 || (n_bytes > 0 && rec == nullptr) || n_records_out == nullptr)
        return fail(RCP_ERR_ARG, "rcp_bam_index: bad argument");
    int64_t p = 0, n = 0;
    while (p < n_bytes) {
        if (n_bytes - p < 4) return fail(RCP_ERR_DATA, "BAM records: truncated block_size at byte %lld", (long long)p);
        int32_t bs;
        memcpy(&bs, rec + p, 4);
        if (bs < 32 || (int64_t)bs + 4 > n_bytes - p)
            return fail(RCP_ERR_DATA, "BAM records: record %lld at byte %lld has block_size %d", (long long)n,
                        (long long)p, (int)bs);
        if (offsets_out) {
            if (n + 1 >= capacity)      // room for this record's offset and for the end
                return fail(RCP_ERR_ARG, "rcp_bam_index: offsets_out has %lld entries, the file more records", (long long)capacity);
            offsets_out[n] = p;
        }
        p += 4 + (int64_t)bs;
        n++;
    }
    if (offsets_out) {
        if (capacity < 1) return fail(RCP_ERR_ARG, "rcp_bam_index: offsets_out needs records + 1 entries");
        offsets_out[n] = p;
    }
    *n_records_out = n;
    return RCP_OK;
}

int rcp_bam_decode(const uint8_t* rec, int64_t n_bytes, const int64_t* offsets, int64_t n_records, int n_ref,
                   const int64_t* ref_len, int splice_split, int mem, int* decoded_out, int64_t* n_out) {
    RCP_TRY(require_ready());
    if (n_bytes < 0 || n_records < 0 || n_ref < 0 || decoded_out == nullptr || n_out == nullptr ||
        (n_records > 0 && (rec == nullptr || offsets == nullptr)) || (n_ref > 0 && ref_len == nullptr))
        return fail(RCP_ERR_ARG, "rcp_bam_decode: bad argument");
    if (mem != RCP_MEM_HOST && mem != RCP_MEM_DEVICE) return fail(RCP_ERR_ARG, "bad mem kind");
    if (n_records >= 0x7ffffff0ll) return fail(RCP_ERR_UNSUPPORTED, "more than 2^31 - 17 alignment records");
    DevIn<uint8_t> d_rec;
    DevIn<int64_t> d_off, d_len;
    RCP_TRY(d_rec.init(rec, (size_t)n_bytes, mem));
    RCP_TRY(d_off.init(offsets, (size_t)n_records + 1, mem));
    RCP_TRY(d_len.init(ref_len, (size_t)std::max(n_ref, 1), RCP_MEM_HOST));
    Arena A;
    const size_t n = (size_t)n_records;
    RCP_TRY(A.reserve(Arena::pad(16) + Arena::pad((n + 1) * 4) * 2));
    unsigned int* err = A.take<unsigned int>(4);        // [0] err [1] total
    uint32_t* counts = A.take<uint32_t>(n + 1);
    uint32_t* where = A.take<uint32_t>(n + 1);
    RCP_CUDA(cudaMemsetAsync(err, 0, 16, g_ctx.stream));
    const Out none = {nullptr, nullptr, nullptr, nullptr};
    if (n_records > 0) {
        bam_records_kernel<false><<<blocks_for(n_records, CTA), CTA, 0, g_ctx.stream>>>(
            d_rec.ptr, n_bytes, d_off.ptr, n_records, n_ref, d_len.ptr, splice_split ? 1 : 0, counts, nullptr, none,
            err);
        RCP_LAUNCHED();
        RCP_TRY(exclusive_scan_u32(counts, where, n_records, err + 1));
    }
    unsigned int h_err = 0, h_total = 0;
    FetchItem items[2] = {{err, &h_err, 4}, {err + 1, &h_total, 4}};
    RCP_TRY(fetch_and_sync(items, 2));
    if (h_err & 1u) return fail(RCP_ERR_DATA, "BAM records: the offsets do not follow the record chain, or a record is truncated");
    if (h_err & 2u) return fail(RCP_ERR_DATA, "BAM records: a refID is outside the header's %d references", n_ref);
    if (h_err & 4u) return fail(RCP_ERR_DATA, "BAM records: a CIGAR operation code is above 8");
    if (h_err & 8u) return fail(RCP_ERR_DATA, "BAM records: a mapped record has no CIGAR (GAlignments needs one)");
    Decoded* d;
    int h;
    RCP_TRY(new_decoded(&d, &h));
    int rc = alloc_out(*d, (int64_t)h_total);
    if (rc == RCP_OK && n_records > 0) {
        const Out out = {d->chrom, d->start, d->end, d->strand};
        bam_records_kernel<true><<<blocks_for(n_records, CTA), CTA, 0, g_ctx.stream>>>(
            d_rec.ptr, n_bytes, d_off.ptr, n_records, n_ref, d_len.ptr, splice_split ? 1 : 0, nullptr, where, out, err);
        g_ctx.launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(RCP_ERR_CUDA, "bam_records_kernel launch failed");
    }
    if (rc != RCP_OK) {
        decoded_release(*d);
        g_decoded.erase(h);
        return rc;
    }
    *decoded_out = h;
    *n_out = (int64_t)h_total;
    return RCP_OK;
}

int rcp_bed_decode(const char* text, int64_t n_bytes, int n_names, const char* const* names, int mem,
                   int* decoded_out, int64_t* n_out) {
    RCP_TRY(require_ready());
    if (n_bytes < 0 || (n_bytes > 0 && text == nullptr) || n_names < 1 || names == nullptr ||
        decoded_out == nullptr || n_out == nullptr)
        return fail(RCP_ERR_ARG, "rcp_bed_decode: bad argument");
    if (mem != RCP_MEM_HOST && mem != RCP_MEM_DEVICE) return fail(RCP_ERR_ARG, "bad mem kind");
    // the seqlevels, sorted by name for the device's binary search
    std::vector<int32_t> order((size_t)n_names);
    for (int i = 0; i < n_names; i++) {
        if (names[i] == nullptr) return fail(RCP_ERR_ARG, "rcp_bed_decode: names[%d] is NULL", i);
        order[(size_t)i] = i;
    }
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        const size_t lx = strlen(names[x]), ly = strlen(names[y]);
        const int c = memcmp(names[x], names[y], std::min(lx, ly));
        return c != 0 ? c < 0 : lx < ly;
    });
    std::string blob;
    std::vector<int32_t> noff((size_t)n_names + 1);
    for (int i = 0; i < n_names; i++) {
        noff[(size_t)i] = (int32_t)blob.size();
        blob += names[order[(size_t)i]];
        if (i > 0 && blob.compare((size_t)noff[(size_t)i - 1], (size_t)(noff[(size_t)i] - noff[(size_t)i - 1]),
                                  names[order[(size_t)i]]) == 0)
            return fail(RCP_ERR_ARG, "rcp_bed_decode: seqlevel '%s' appears twice", names[order[(size_t)i]]);
    }
    noff[(size_t)n_names] = (int32_t)blob.size();
    // the text, 16-byte aligned with 16 readable bytes past the end (our own copy unless the
    // caller's device pointer is aligned; the tail of the last vector is masked by n_bytes)
    const size_t padded = ((size_t)n_bytes + 15) / 16 * 16 + 16;
    uint8_t* d_text = nullptr;
    RCP_TRY(dalloc(&d_text, padded));
    struct Free {
        uint8_t*& p;
        ~Free() { dfree(p); }
    } free_text{d_text};
    if (n_bytes > 0)
        RCP_CUDA(cudaMemcpyAsync(d_text, text, (size_t)n_bytes,
                                 mem == RCP_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, g_ctx.stream));
    const int64_t n_vec = (n_bytes + 15) / 16;
    const unsigned blocks = std::max(1u, blocks_for(n_vec, CTA));
    Arena A;
    RCP_TRY(A.reserve(Arena::pad(16) + Arena::pad(((size_t)blocks + 1) * 4) * 2 + Arena::pad(blob.size() + 1) +
                      Arena::pad(((size_t)n_names + 1) * 4) * 2));
    unsigned int* err = A.take<unsigned int>(4);        // [0] err [1] newlines [2] lines kept
    uint32_t* bcnt = A.take<uint32_t>((size_t)blocks + 1);
    uint32_t* bbase = A.take<uint32_t>((size_t)blocks + 1);
    char* d_blob = A.take<char>(blob.size() + 1);
    int32_t* d_noff = A.take<int32_t>((size_t)n_names + 1);
    int32_t* d_id = A.take<int32_t>((size_t)n_names + 1);
    RCP_CUDA(cudaMemsetAsync(err, 0, 16, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(d_noff, noff.data(), noff.size() * 4, cudaMemcpyHostToDevice, g_ctx.stream));
    RCP_CUDA(cudaMemcpyAsync(d_id, order.data(), order.size() * 4, cudaMemcpyHostToDevice, g_ctx.stream));
    bed_nl_count_kernel<<<blocks, CTA, 0, g_ctx.stream>>>(reinterpret_cast<const uint4*>(d_text), n_bytes, bcnt);
    RCP_LAUNCHED();
    RCP_TRY(exclusive_scan_u32(bcnt, bbase, (int64_t)blocks, err + 1));
    unsigned int h_nl = 0;
    uint8_t last = '\n';
    {
        FetchItem it[1] = {{err + 1, &h_nl, 4}};
        RCP_TRY(fetch_and_sync(it, 1));
        if (n_bytes > 0) {
            RCP_CUDA(cudaMemcpyAsync(&last, d_text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, g_ctx.stream));
            RCP_CUDA(cudaStreamSynchronize(g_ctx.stream));
        }
    }
    const int64_t n_nl = (int64_t)h_nl, n_lines = n_nl + ((n_bytes > 0 && last != '\n') ? 1 : 0);
    if (n_lines >= 0x7ffffff0ll) return fail(RCP_ERR_UNSUPPORTED, "more than 2^31 - 17 lines");
    Arena B;
    RCP_TRY(B.reserve(Arena::pad(((size_t)n_nl + 2) * 8) + Arena::pad(((size_t)n_lines + 1) * 4) * 2));
    int64_t* starts = B.take<int64_t>((size_t)n_nl + 2);
    uint32_t* counts = B.take<uint32_t>((size_t)n_lines + 1);
    uint32_t* where = B.take<uint32_t>((size_t)n_lines + 1);
    bed_nl_write_kernel<<<blocks, CTA, 0, g_ctx.stream>>>(reinterpret_cast<const uint4*>(d_text), n_bytes, bbase, starts);
    RCP_LAUNCHED();
    const NameTable nt = {d_blob, d_noff, d_id, n_names};
    const Out none = {nullptr, nullptr, nullptr, nullptr};
    if (n_lines > 0) {
        bed_lines_kernel<false><<<blocks_for(n_lines, CTA), CTA, 0, g_ctx.stream>>>(d_text, n_bytes, starts, n_nl, n_lines,
                                                                                  nt, counts, nullptr, none, err);
        RCP_LAUNCHED();
        RCP_TRY(exclusive_scan_u32(counts, where, n_lines, err + 2));
    }
    unsigned int h_err = 0, h_total = 0;
    FetchItem items[2] = {{err, &h_err, 4}, {err + 2, &h_total, 4}};
    RCP_TRY(fetch_and_sync(items, 2));
    if (h_err & 1u) return fail(RCP_ERR_DATA, "BED: a chromosome name is not among the %d seqlevels", n_names);
    if (h_err & 2u) return fail(RCP_ERR_DATA, "BED: a data line has fewer than three fields or a bad coordinate");
    if (h_err & 4u) return fail(RCP_ERR_DATA, "BED: a strand field is not one of + - . *");
    Decoded* d;
    int h;
    RCP_TRY(new_decoded(&d, &h));
    int rc = alloc_out(*d, (int64_t)h_total);
    if (rc == RCP_OK && n_lines > 0) {
        const Out out = {d->chrom, d->start, d->end, d->strand};
        bed_lines_kernel<true><<<blocks_for(n_lines, CTA), CTA, 0, g_ctx.stream>>>(d_text, n_bytes, starts, n_nl, n_lines, nt,
                                                                                 nullptr, where, out, err);
        g_ctx.launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(RCP_ERR_CUDA, "bed_lines_kernel launch failed");
    }
    if (rc != RCP_OK) {
        decoded_release(*d);
        g_decoded.erase(h);
        return rc;
    }
    *decoded_out = h;
    *n_out = (int64_t)h_total;
    return RCP_OK;
}

int rcp_decoded_fetch(int decoded, int32_t* chrom, int32_t* start, int32_t* end, int8_t* strand, int64_t capacity) {
    RCP_TRY(require_ready());
    auto it = g_decoded.find(decoded);
    if (it == g_decoded.end()) return fail(RCP_ERR_HANDLE, "unknown decoded handle %d", decoded);
    const Decoded& d = *it->second;
    if (capacity < d.n) return fail(RCP_ERR_ARG, "rcp_decoded_fetch: capacity %lld < %lld ranges", (long long)capacity, (long long)d.n);
    const size_t n = (size_t)d.n;
    if (n == 0) return RCP_OK;
    if (chrom) RCP_CUDA(cudaMemcpyAsync(chrom, d.chrom, n * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    if (start) RCP_CUDA(cudaMemcpyAsync(start, d.start, n * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    if (end) RCP_CUDA(cudaMemcpyAsync(end, d.end, n * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    if (strand) RCP_CUDA(cudaMemcpyAsync(strand, d.strand, n, cudaMemcpyDeviceToHost, g_ctx.stream));
    RCP_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return RCP_OK;
}

int rcp_decoded_free(int decoded) {
    auto it = g_decoded.find(decoded);
    if (it == g_decoded.end()) return fail(RCP_ERR_HANDLE, "unknown decoded handle %d", decoded);
    decoded_release(*it->second);
    g_decoded.erase(it);
    return RCP_OK;
}

int rcp_reads_load_decoded(int decoded, int n_chrom, const int64_t* chrom_len, int frag_len, int* reads_out) {
    RCP_TRY(require_ready());
    auto it = g_decoded.find(decoded);
    if (it == g_decoded.end()) return fail(RCP_ERR_HANDLE, "unknown decoded handle %d", decoded);
    const Decoded& d = *it->second;
    return rcp_reads_load(d.n, d.chrom, d.start, d.end, d.strand, n_chrom, chrom_len, frag_len, RCP_MEM_DEVICE,
                          reads_out);
}

int rcp_decoded_width_quantile(int decoded, double prob, double* quantile_out, int64_t* n_le_out) {
    RCP_TRY(require_ready());
    auto it = g_decoded.find(decoded);
    if (it == g_decoded.end()) return fail(RCP_ERR_HANDLE, "unknown decoded handle %d", decoded);
    const Decoded& d = *it->second;
    return rcp_reads_width_quantile(d.n, d.start, d.end, prob, RCP_MEM_DEVICE, quantile_out, n_le_out);
}

int rcp_reads_load_decoded_select(int decoded, double max_width, int64_t k, const int32_t* idx, int n_chrom,
                                  const int64_t* chrom_len, int frag_len, int64_t* n_kept_out, int* reads_out) {
    RCP_TRY(require_ready());
    auto it = g_decoded.find(decoded);
    if (it == g_decoded.end()) return fail(RCP_ERR_HANDLE, "unknown decoded handle %d", decoded);
    if (k < 0 || (k > 0 && idx == nullptr)) return fail(RCP_ERR_ARG, "rcp_reads_load_decoded_select: bad index");
    const Decoded& d = *it->second;
    DevIn<int32_t> d_idx;
    RCP_TRY(d_idx.init(idx, (size_t)k, RCP_MEM_HOST));
    return rcp_reads_load_select(d.n, d.chrom, d.start, d.end, d.strand, max_width, k, d_idx.ptr, n_chrom, chrom_len,
                                 frag_len, RCP_MEM_DEVICE, n_kept_out, reads_out);
}

}  // extern "C"
