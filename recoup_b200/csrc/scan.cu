// Exclusive prefix sums over int64 (region offsets, work-list offsets), one or two independent
// sequences per call.  Three small kernels: per-tile reduce -> one-block scan of the tile sums
// -> per-tile scan + carry.  The arrays are region-sized (<= a few million), so this is
// launch-latency work, not bandwidth work.
#include "rcp_internal.cuh"

namespace rcp {

namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <int NC>
struct ScanArgs {
    const int64_t* in[NC];
    int64_t* out[NC];
    int64_t* partial[NC];
    int64_t* total[NC];   // device scalars, may be null
};

__device__ __forceinline__ int64_t warp_inclusive(int64_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread over the block; *total = block sum
template <int THREADS>
__device__ __forceinline__ int64_t block_exclusive(int64_t v, int64_t* total) {
    __shared__ int64_t wsum[THREADS / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = warp_inclusive(v);
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int64_t pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        int64_t s = wsum[w];
        if (w < (int)warp) pre += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return pre + inc - v;
}

template <int NC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(ScanArgs<NC> a, int64_t n) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        int64_t s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
            if (i < n) s += a.in[c][i];
        }
        int64_t tot;
        block_exclusive<SCAN_THREADS>(s, &tot);
        if (threadIdx.x == 0) a.partial[c][blockIdx.x] = tot;
    }
}

template <int NC>
__global__ void __launch_bounds__(1024) scan_partials_kernel(ScanArgs<NC> a, int64_t nb) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
        int64_t carry = 0;
        for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
            int64_t i = b0 + threadIdx.x;
            int64_t v = (i < nb) ? a.partial[c][i] : 0;
            int64_t tot;
            int64_t ex = block_exclusive<1024>(v, &tot);
            if (i < nb) a.partial[c][i] = carry + ex;
            carry += tot;
        }
        if (threadIdx.x == 0 && a.total[c]) *a.total[c] = carry;
    }
}

template <int NC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(ScanArgs<NC> a, int64_t n) {
    // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        int64_t v[SCAN_ITEMS];
        int64_t s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            v[k] = (i < n) ? a.in[c][i] : 0;
            s += v[k];
        }
        int64_t tot;
        int64_t ex = block_exclusive<SCAN_THREADS>(s, &tot) + a.partial[c][blockIdx.x];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            if (i < n) a.out[c][i] = ex;
            ex += v[k];
        }
    }
}

template <int NC>
int scan_impl(ScanArgs<NC> a, int64_t n) {
    if (n <= 0) {
        for (int c = 0; c < NC; c++)
            if (a.total[c]) RCP_CUDA(cudaMemsetAsync(a.total[c], 0, sizeof(int64_t), g_ctx.stream));
        return RCP_OK;
    }
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* partial = nullptr;
    RCP_TRY(dalloc(&partial, (size_t)nb * NC));
    for (int c = 0; c < NC; c++) a.partial[c] = partial + (size_t)c * nb;
    scan_reduce_kernel<NC><<<(unsigned)nb, SCAN_THREADS, 0, g_ctx.stream>>>(a, n);
    RCP_LAUNCHED();
    scan_partials_kernel<NC><<<1, 1024, 0, g_ctx.stream>>>(a, nb);
    RCP_LAUNCHED();
    scan_apply_kernel<NC><<<(unsigned)nb, SCAN_THREADS, 0, g_ctx.stream>>>(a, n);
    RCP_LAUNCHED();
    dfree(partial);
    return RCP_OK;
}
}  // namespace

int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, int64_t* d_total) {
    ScanArgs<1> a;
    a.in[0] = in;
    a.out[0] = out;
    a.partial[0] = nullptr;
    a.total[0] = d_total;
    return scan_impl<1>(a, n);
}

int exclusive_scan2_i64(const int64_t* in0, int64_t* out0, int64_t* d_total0, const int64_t* in1,
                        int64_t* out1, int64_t* d_total1, int64_t n) {
    ScanArgs<2> a;
    a.in[0] = in0;
    a.out[0] = out0;
    a.total[0] = d_total0;
    a.in[1] = in1;
    a.out[1] = out1;
    a.total[1] = d_total1;
    a.partial[0] = a.partial[1] = nullptr;
    return scan_impl<2>(a, n);
}

}  // namespace rcp
