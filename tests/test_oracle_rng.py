"""Known-answer tests pinning the R RNG restatements (oracle AND product) -- SURVEY.md 8c."""
import ctypes as C

import pytest

from oracle.r_rng import RRandom, r_sample, rank_table


def test_runif_kat():
    r = RRandom(42)
    got = [r.unif_rand() for _ in range(3)]
    assert [round(v, 7) for v in got] == [0.9148060, 0.9370754, 0.2861395]


def test_sample_kats():
    assert r_sample(10, 10) == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]                       # R >= 3.6
    assert r_sample(10, 10, sample_kind="Rounding") == [10, 9, 3, 6, 4, 8, 5, 1, 2, 7]  # R < 3.6
    assert r_sample(100, 10) == [49, 65, 25, 74, 18, 100, 47, 24, 71, 89]


@pytest.mark.parametrize("n,head", [
    (50, [49, 37, 1, 25, 10, 36, 18, 24, 7, 45, 47, 50]),
    (100, [49, 65, 25, 74, 18, 100, 47, 24, 71, 89, 37, 20]),
    (150, [49, 65, 74, 146, 122, 150, 128, 47, 24, 71, 100, 89]),
    (200, [49, 65, 153, 74, 146, 122, 200, 128, 47, 24, 71, 100]),
])
def test_seed42_permutation_heads(n, head):
    assert r_sample(n, n)[:12] == head


def test_prefix_property():
    # sample(1:n, d) under a fixed seed is a prefix of sample(1:n, n): one rank table per n
    for n in (7, 50, 150):
        full = r_sample(n, n)
        for d in (0, 1, n // 2, n - 1):
            assert r_sample(n, d) == full[:d]
        rank = rank_table(n)
        assert sorted(rank) == list(range(1, n + 1))
        assert [full[r - 1] for r in rank] == list(range(1, n + 1))


@pytest.mark.parametrize("kind,code", [("Rejection", 0), ("Rounding", 1)])
def test_library_rng_matches_oracle(kind, code):
    """The product's C++ RNG (csrc/r_rng.cuh) and the oracle's python RNG are independent
    restatements; they must agree for every n the profile path can meet."""
    from recoup_b200 import _lib
    for n in list(range(1, 70)) + [100, 150, 200, 250, 1000, 4097]:
        out = (C.c_int * n)()
        assert _lib.lib.rcp_r_sample(n, n, 42, code, out) == 0
        assert list(out) == r_sample(n, n, 42, kind), (n, kind)
        rank = (C.c_int * n)()
        assert _lib.lib.rcp_r_rank_table(n, 42, code, rank) == 0
        assert list(rank) == rank_table(n, 42, kind)
    out = (C.c_int * 5)()
    assert _lib.lib.rcp_r_sample(31, 5, 7, code, out) == 0
    assert list(out) == r_sample(31, 5, 7, kind)


def test_library_rng_rejects_bad_sizes():
    from recoup_b200 import _lib
    out = (C.c_int * 4)()
    assert _lib.lib.rcp_r_sample(3, 4, 42, 0, out) == _lib.RCP_ERR_ARG
    assert b"sample" in _lib.lib.rcp_last_error()
