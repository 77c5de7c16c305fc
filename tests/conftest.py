import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


TEST_MODELS = ROOT / "tests" / "models" / "libdde_test_models.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def test_model_library():
    """The reference's fixture models (seir, the growth family, logistic, fibonacci ...) are an
    out-of-tree model library (tests/models/build.sh), loaded here the way a user's is: the
    product library has none of them compiled in."""
    if TEST_MODELS.exists():
        import dde_b200
        dde_b200.load_model_library(TEST_MODELS)
    return TEST_MODELS


def bits_equal(a, b):
    """Bit-for-bit equality of float64 arrays; NaNs match NaNs of any payload
    (payloads do not survive device arithmetic, see lag_ctx.cuh)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint64)[~na], b.view(np.uint64)[~nb])


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if bits_equal(a, b):
        return
    na, nb = np.isnan(a), np.isnan(b)
    bad = (na != nb) | (~na & ~nb & (a.view(np.uint64) != b.view(np.uint64)))
    idx = np.argwhere(bad)
    first = tuple(idx[0])
    raise AssertionError("%s: %d of %d values differ; first at %s: %r vs %r" %
                         (what, bad.sum(), bad.size, first, a[first], b[first]))


def checker_backends():
    from oracle import pyoracle
    return [b for b in ("ref", "oracle") if pyoracle.available(b)]


@pytest.fixture(params=["ref", "oracle"])
def backend(request):
    from oracle import pyoracle
    if not pyoracle.available(request.param):
        pytest.skip("checker '%s' not built" % request.param)
    return request.param
