// Exclusive prefix sums over int64 / uint32 (region offsets, work-list offsets, cell lists), one or two independent
// sequences per call.  Three small kernels: per-tile reduce -> one-block scan of the tile sums
// -> per-tile scan + carry.  The arrays are region-sized (<= a few million), so this is
// launch-latency work, not bandwidth work.
#include "rcp_internal.cuh"

namespace rcp {

namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class V, int NC>
struct ScanArgs {
    const V* in[NC];
    V* out[NC];
    V* partial[NC];
    V* total[NC];   // device scalars, may be null
    const uint32_t* n_dev = nullptr;   // optional device-side length (<= the host-side n)
};

template <class V, int NC>
__device__ __forceinline__ int64_t scan_length(const ScanArgs<V, NC>& a, int64_t n) {
    if (a.n_dev == nullptr) return n;
    const int64_t m = (int64_t)*a.n_dev;
    return m < n ? m : n;
}

template <class V>
__device__ __forceinline__ V warp_inclusive(V v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        V o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread over the block; *total = block sum
template <class V, int THREADS>
__device__ __forceinline__ V block_exclusive(V v, V* total) {
    __shared__ V wsum[THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    V inc = warp_inclusive<V>(v);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    V pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        V s = wsum[w];
        if (w < (int)warp) pre += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return pre + inc - v;
}

template <class V, int NC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(ScanArgs<V, NC> a, int64_t n) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    n = scan_length(a, n);
    if (base >= n) {                // beyond the device-side length: contributes nothing
        if (threadIdx.x == 0)
            for (int c = 0; c < NC; c++) a.partial[c][blockIdx.x] = 0;
        return;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
        V s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
            if (i < n) s += a.in[c][i];
        }
        V tot;
        block_exclusive<V, SCAN_THREADS>(s, &tot);
        if (threadIdx.x == 0) a.partial[c][blockIdx.x] = tot;
    }
}

template <class V, int NC>
__global__ void __launch_bounds__(1024) scan_partials_kernel(ScanArgs<V, NC> a, int64_t nb) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
        V carry = 0;
        for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
            int64_t i = b0 + threadIdx.x;
            V v = (i < nb) ? a.partial[c][i] : 0;
            V tot;
            V ex = block_exclusive<V, 1024>(v, &tot);
            if (i < nb) a.partial[c][i] = carry + ex;
            carry += tot;
        }
        if (threadIdx.x == 0 && a.total[c]) *a.total[c] = carry;
    }
}

template <class V, int NC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(ScanArgs<V, NC> a, int64_t n) {
    // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    n = scan_length(a, n);
    if ((int64_t)blockIdx.x * SCAN_TILE >= n) return;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        V v[SCAN_ITEMS];
        V s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            v[k] = (i < n) ? a.in[c][i] : 0;
            s += v[k];
        }
        V tot;
        V ex = block_exclusive<V, SCAN_THREADS>(s, &tot) + a.partial[c][blockIdx.x];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            if (i < n) a.out[c][i] = ex;
            ex += v[k];
        }
    }
}

// Very short sequences (the plan of C1 has 100 regions): ONE launch of one CTA, 4 096 elements per
// trip, instead of three launches.  (Longer ones stay with the three kernels: one CTA walking
// 20 000 entries is one SM's worth of memory parallelism, 49 us against 20 us on the C2 plan.)
constexpr int SMALL_THREADS = 1024, SMALL_ITEMS = 4;
constexpr int64_t SMALL_MAX = 2 * SMALL_THREADS * SMALL_ITEMS;
template <class V, int NC>
__global__ void __launch_bounds__(SMALL_THREADS) scan_small_kernel(ScanArgs<V, NC> a, int64_t n) {
    n = scan_length(a, n);
#pragma unroll
    for (int c = 0; c < NC; c++) {
        V carry = 0;
        for (int64_t b0 = 0; b0 < n; b0 += SMALL_THREADS * SMALL_ITEMS) {
            const int64_t base = b0 + (int64_t)threadIdx.x * SMALL_ITEMS;
            V v[SMALL_ITEMS];
            V s = 0;
#pragma unroll
            for (int k = 0; k < SMALL_ITEMS; k++) {
                v[k] = (base + k < n) ? a.in[c][base + k] : 0;
                s += v[k];
            }
            V tot;
            V ex = carry + block_exclusive<V, SMALL_THREADS>(s, &tot);
#pragma unroll
            for (int k = 0; k < SMALL_ITEMS; k++) {
                if (base + k < n) a.out[c][base + k] = ex;
                ex += v[k];
            }
            carry += tot;
        }
        if (threadIdx.x == 0 && a.total[c]) *a.total[c] = carry;
    }
}

template <class V, int NC>
int scan_impl(ScanArgs<V, NC> a, int64_t n) {
    if (n <= 0) {
        for (int c = 0; c < NC; c++)
            if (a.total[c]) RCP_CUDA(cudaMemsetAsync(a.total[c], 0, sizeof(V), g_ctx.stream));
        return RCP_OK;
    }
    if (n <= SMALL_MAX) {
        scan_small_kernel<V, NC><<<1, SMALL_THREADS, 0, g_ctx.stream>>>(a, n);
        RCP_LAUNCHED();
        return RCP_OK;
    }
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    V* partial = nullptr;
    RCP_TRY(dalloc(&partial, (size_t)nb * NC));
    for (int c = 0; c < NC; c++) a.partial[c] = partial + (size_t)c * nb;
    scan_reduce_kernel<V, NC><<<(unsigned)nb, SCAN_THREADS, 0, g_ctx.stream>>>(a, n);
    RCP_LAUNCHED();
    scan_partials_kernel<V, NC><<<1, 1024, 0, g_ctx.stream>>>(a, nb);
    RCP_LAUNCHED();
    scan_apply_kernel<V, NC><<<(unsigned)nb, SCAN_THREADS, 0, g_ctx.stream>>>(a, n);
    RCP_LAUNCHED();
    dfree(partial);
    return RCP_OK;
}
}  // namespace

int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, int64_t* d_total) {
    ScanArgs<int64_t, 1> a;
    a.in[0] = in;
    a.out[0] = out;
    a.partial[0] = nullptr;
    a.total[0] = d_total;
    return scan_impl<int64_t, 1>(a, n);
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* d_total) {
    ScanArgs<uint32_t, 1> a;
    a.in[0] = in;
    a.out[0] = out;
    a.partial[0] = nullptr;
    a.total[0] = d_total;
    return scan_impl<uint32_t, 1>(a, n);
}

int exclusive_scan_u32_bounded(const uint32_t* in, uint32_t* out, int64_t n_upper, const uint32_t* n_dev,
                               uint32_t* d_total) {
    ScanArgs<uint32_t, 1> a;
    a.in[0] = in;
    a.out[0] = out;
    a.partial[0] = nullptr;
    a.total[0] = d_total;
    a.n_dev = n_dev;
    return scan_impl<uint32_t, 1>(a, n_upper);
}

int exclusive_scan2_i64(const int64_t* in0, int64_t* out0, int64_t* d_total0, const int64_t* in1,
                        int64_t* out1, int64_t* d_total1, int64_t n) {
    ScanArgs<int64_t, 2> a;
    a.in[0] = in0;
    a.out[0] = out0;
    a.total[0] = d_total0;
    a.in[1] = in1;
    a.out[1] = out1;
    a.total[1] = d_total1;
    a.partial[0] = a.partial[1] = nullptr;
    return scan_impl<int64_t, 2>(a, n);
}

}  // namespace rcp
