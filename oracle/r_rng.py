"""ORACLE (test infrastructure, never imported by the product).

Restatement of the part of base R's random number machinery that recoup's bin layout depends on:
`set.seed(42); sample(1:n, dif)` in /root/reference/R/util.R:78-79 (and `sample(3:(n-2), k)` at
util.R:25-28,55-58).  base R is a third-party dependency that is NOT under /root/reference
(DESCRIPTION:7-17 pins no R version), so the published algorithm is restated here:

* `set.seed(seed)`            R src/main/RNG.c  RNG_Init(): 50 rounds of the LCG 69069*s+1, then
                              625 more rounds fill the Mersenne-Twister state words; word 0 (mti)
                              is forced to 624 by FixupSeeds().
* `unif_rand()`               MT19937 genrand (MT_sgenrand is never reached because mti == N only
                              triggers the block regeneration), scaled by 2.3283064365386963e-10
                              and clamped into the open interval (0,1).
* `R_unif_index(dn)`          R >= 3.6.0 default ("Rejection"): draw ceil(log2(dn)) random bits in
                              16-bit chunks, reject values >= dn.  R < 3.6.0 ("Rounding"):
                              floor(dn * unif_rand()).
* `sample.int(n, k)` without  R src/main/random.c do_sample(): partial Fisher-Yates,
  replacement                 y[i] = x[j] + 1; x[j] = x[--n].

Pinned by the known answers in SURVEY.md section 8c (see tests/test_oracle_rng.py):
  set.seed(42); runif(3)        -> 0.9148060 0.9370754 0.2861395
  set.seed(42); sample(1:10)    -> 1 5 10 8 2 4 6 9 7 3        (Rejection)
                                -> 10 9 3 6 4 8 5 1 2 7        (Rounding)
  set.seed(42); sample(1:100,10)-> 49 65 25 74 18 100 47 24 71 89
"""
import math

N, M = 624, 397
MATRIX_A = 0x9908B0DF
UPPER_MASK = 0x80000000
LOWER_MASK = 0x7FFFFFFF
I2_32M1 = 2.328306437080797e-10


class RRandom:
    def __init__(self, seed=42, sample_kind="Rejection"):
        if sample_kind not in ("Rejection", "Rounding"):
            raise ValueError("sample_kind must be 'Rejection' or 'Rounding'")
        self.sample_kind = sample_kind
        self.set_seed(seed)

    def set_seed(self, seed):
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        state = []
        for _ in range(N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            state.append(s)
        # dummy[0] is mti; FixupSeeds sets it to N so the first draw regenerates the block
        self.mt = state[1:]
        self.mti = N

    def _genrand(self):
        mt = self.mt
        if self.mti >= N:
            for kk in range(N - M):
                y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK)
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ (MATRIX_A if y & 1 else 0)
            for kk in range(N - M, N - 1):
                y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK)
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (MATRIX_A if y & 1 else 0)
            y = (mt[N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK)
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (MATRIX_A if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def unif_rand(self):
        v = self._genrand() * 2.3283064365386963e-10
        if v <= 0.0:
            return 0.5 * I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * I2_32M1
        return v

    def _rbits(self, bits):
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn):
        if self.sample_kind == "Rounding":
            return int(math.floor(dn * self.unif_rand()))
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample_int(self, n, k=None):
        """`sample.int(n, k)` without replacement (1-based values)."""
        if k is None:
            k = n
        if k > n:
            raise ValueError("cannot take a sample larger than the population")
        if k < 0:
            raise ValueError("invalid 'size' argument")
        if n > 1e7 and k <= n / 2:
            # sample.int's hash variant (R: .Internal(sample2(n, size)), do_sample2 in
            # src/main/unique.c): draw, redraw on a duplicate, at most 100 tries
            seen = set()
            out = []
            for _ in range(k):
                for _try in range(100):
                    v = self.unif_index(n) + 1
                    if v not in seen:
                        break
                seen.add(v)
                out.append(v)
            return out
        x = list(range(n))
        out = []
        for _ in range(k):
            j = self.unif_index(n)
            out.append(x[j] + 1)
            n -= 1
            x[j] = x[n]
        return out


def r_sample(n, k, seed=42, sample_kind="Rejection"):
    """`set.seed(seed); sample(1:n, k)`."""
    return RRandom(seed, sample_kind).sample_int(n, k)


def rank_table(n, seed=42, sample_kind="Rejection"):
    """rank[i] (1-based value, 0-based bin i) = position of bin i+1 in `set.seed(seed);
    sample(1:n, n)`.  Because `sample(1:n, d)` is a prefix of `sample(1:n, n)` under a fixed
    seed, bin i receives an extra base iff rank[i] <= d  (util.R:74-80)."""
    perm = r_sample(n, n, seed, sample_kind)
    rank = [0] * n
    for pos, b in enumerate(perm):
        rank[b - 1] = pos + 1
    return rank
