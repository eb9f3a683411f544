import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fixture_data():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "recoup_test_data.npz"))


@pytest.fixture(scope="session", params=["auto", "split", "buckets", "index", "blocks"])
def gpu(request):
    """Binds the CUDA library to cuda:0; fails (never skips) when the GPU path is unusable.
    Every parity test runs once per coverage path (buckets / sorted index / the automatic choice
    between them / the one-pass split / the block partition): all must match the oracle bit for bit."""
    import recoup_b200 as rb
    rb.init(0)
    rb.set_coverage_path(request.param)
    yield rb
    rb.set_coverage_path("auto")
