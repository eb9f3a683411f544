// base-R random numbers needed by recoup's bin layout, usable on host and device.
//
// Product-side implementation (independent of oracle/r_rng.py) of `set.seed(seed)` +
// `sample()` as used by splitVector (/root/reference/R/util.R:25-28,55-58,78-79).  base R is a
// third-party dependency of the reference that is not vendored; the algorithm is R's published
// one: RNG_Init() LCG scrambling (69069*s+1), Mersenne-Twister MT19937 genrand scaled by
// 2.3283064365386963e-10 and clamped into (0,1), R_unif_index() by rejection over
// ceil(log2(n)) bits drawn in 16-bit chunks (R >= 3.6) or floor(n*u) (R < 3.6), and the partial
// Fisher-Yates loop of do_sample().
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define RCP_HD
#else
#define RCP_HD __host__ __device__
#endif

namespace rcp {

struct RRng {
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int mti;
    int kind;   // RCP_SAMPLE_REJECTION (0) / RCP_SAMPLE_ROUNDING (1)

    RCP_HD void seed(uint32_t s, int sample_kind) {
        kind = sample_kind;
        for (int j = 0; j < 50; j++) s = 69069u * s + 1u;
        s = 69069u * s + 1u;              // word 0 = mti, overwritten by FixupSeeds
        for (int j = 0; j < N; j++) {
            s = 69069u * s + 1u;
            mt[j] = s;
        }
        mti = N;
    }

    RCP_HD uint32_t genrand() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
        if (mti >= N) {
            int kk;
            for (kk = 0; kk < N - M; kk++) {
                uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
            }
            for (; kk < N - 1; kk++) {
                uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
            }
            uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }

    RCP_HD double unif() {
        const double i2_32m1 = 2.328306437080797e-10;
        double v = (double)genrand() * 2.3283064365386963e-10;
        if (v <= 0.0) return 0.5 * i2_32m1;
        if (1.0 - v <= 0.0) return 1.0 - 0.5 * i2_32m1;
        return v;
    }

    // R_unif_index(dn) for 1 <= dn < 2^31
    RCP_HD uint32_t index(uint32_t dn) {
        if (kind == 1) return (uint32_t)floor((double)dn * unif());
        if (dn == 0) return 0;
        int bits = 0;
        while (bits < 32 && (1ull << bits) < (unsigned long long)dn) bits++;   // ceil(log2(dn))
        for (;;) {
            unsigned long long v = 0;
            for (int n = 0; n <= bits; n += 16) {
                unsigned long long v1 = (unsigned long long)floor(unif() * 65536.0);
                v = 65536ull * v + v1;
            }
            v &= ((1ull << bits) - 1ull);
            if (v < dn) return (uint32_t)v;
        }
    }

    // `sample.int(n, k)` without replacement; x = scratch of n ints; out = k ints (1-based).
    RCP_HD void sample(int n, int k, int* x, int* out) {
        for (int i = 0; i < n; i++) x[i] = i;
        for (int i = 0; i < k; i++) {
            int j = (int)index((uint32_t)n);
            out[i] = x[j] + 1;
            x[j] = x[--n];
        }
    }
};

}  // namespace rcp
