// Microbenchmarks that size the design of the one-pass partition (run under gpurun):
//   smem atomics (returning / not), match_any, global atomics on few hot addresses, scattered stores.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int MODE>   // 0 returning ATOMS, 1 non-returning, 2 plain STS, 3 match_any, 4 ballot rank 8 bits
__global__ void __launch_bounds__(512) smem_kernel(int iters, int nbins, uint32_t* out) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t x = blockIdx.x * 9781u + threadIdx.x * 7919u + 1u;
    for (int it = 0; it < iters; it++) {
        x = hash32(x + it);
        const uint32_t b = x % (uint32_t)nbins;
        if (MODE == 0) acc += atomicAdd(&sm[b], 1u);
        else if (MODE == 1) atomicAdd(&sm[b], 1u);
        else if (MODE == 2) sm[b] = x;
        else if (MODE == 3) acc += __match_any_sync(0xffffffffu, b);
        else {
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t m = __ballot_sync(0xffffffffu, (b >> k) & 1u);
                peers &= ((b >> k) & 1u) ? m : ~m;
            }
            acc += __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        }
    }
    __syncthreads();
    if (acc == 0xdeadbeefu || sm[threadIdx.x % nbins] == 0xdeadbeefu) out[0] = acc;
}

template <int MODE>   // 0 returning ATOMG, 1 RED, 2 scattered 4-byte store in a window
__global__ void __launch_bounds__(512) gmem_kernel(int iters, uint32_t nbins, uint32_t* tab, uint32_t* out) {
    uint32_t acc = 0;
    uint32_t x = blockIdx.x * 9781u + threadIdx.x * 7919u + 1u;
    for (int it = 0; it < iters; it++) {
        x = hash32(x + it);
        const uint32_t b = x % nbins;
        if (MODE == 0) acc += atomicAdd(&tab[b], 1u);
        else if (MODE == 1) atomicAdd(&tab[b], 1u);
        else tab[b] = x;
    }
    if (acc == 0xdeadbeefu) out[0] = acc;
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t *out, *tab;
    CK(cudaMalloc(&out, 4));
    CK(cudaMalloc(&tab, 256u << 20));
    CK(cudaMemset(tab, 0, 256u << 20));
    const int iters = 2000;
    const char* names[5] = {"ATOMS returning", "ATOMS no-return", "STS scattered", "match_any", "ballot-rank 8 bit"};
    for (int nbins : {64, 1024, 16384}) {
        for (int ctas : {1, 2, 4}) {
            const dim3 grid(sms * ctas);
            const size_t smem = (size_t)nbins * 4;
            float ms[5];
            ms[0] = timeit([&] { smem_kernel<0><<<grid, 512, smem>>>(iters, nbins, out); });
            ms[1] = timeit([&] { smem_kernel<1><<<grid, 512, smem>>>(iters, nbins, out); });
            ms[2] = timeit([&] { smem_kernel<2><<<grid, 512, smem>>>(iters, nbins, out); });
            ms[3] = timeit([&] { smem_kernel<3><<<grid, 512, smem>>>(iters, nbins, out); });
            ms[4] = timeit([&] { smem_kernel<4><<<grid, 512, smem>>>(iters, nbins, out); });
            const double ops = (double)sms * ctas * 512 * iters;
            for (int m = 0; m < 5; m++)
                printf("smem nbins=%5d ctas/sm=%d %-18s %8.3f ms  %7.1f Gop/s  %.3f cyc/op/SM@1.9GHz\n", nbins, ctas,
                       names[m], ms[m], ops / ms[m] * 1e-6, ms[m] * 1e-3 * 1.9e9 / (ops / sms));
        }
    }
    const char* gn[3] = {"ATOMG returning", "RED", "STG scattered 4B"};
    for (uint32_t nbins : {1024u, 65536u, 1u << 20, 16u << 20, 64u << 20}) {
        const dim3 grid(sms * 4);
        float ms[3];
        ms[0] = timeit([&] { gmem_kernel<0><<<grid, 512>>>(200, nbins, tab, out); });
        ms[1] = timeit([&] { gmem_kernel<1><<<grid, 512>>>(200, nbins, tab, out); });
        ms[2] = timeit([&] { gmem_kernel<2><<<grid, 512>>>(200, nbins, tab, out); });
        const double ops = (double)sms * 4 * 512 * 200;
        for (int m = 0; m < 3; m++)
            printf("gmem nbins=%9u %-18s %8.3f ms  %7.1f Gop/s\n", nbins, gn[m], ms[m], ops / ms[m] * 1e-6);
    }
    return 0;
}
