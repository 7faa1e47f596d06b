import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("NBREF_QUIET", "1")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_path(name):
    return os.path.join(GOLDEN, name)


def load_golden_npz(tag, precision="f64"):
    return dict(np.load(golden_path("%s_%s.npz" % (tag, precision))))


def rel_err_per_body(a, b, n, floor=0.0):
    """max_i |a_i - b_i| / |b_i| over bodies, for the acceleration rows of two f vectors (6N) or 3xN blocks.
    floor > 0: |b_i| is taken as at least floor * max_j |b_j| (bodies whose attractions cancel by symmetry)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, n)[-3:]
    b = np.asarray(b, dtype=np.float64).reshape(-1, n)[-3:]
    num = np.sqrt(((a - b) ** 2).sum(axis=0))
    den = np.sqrt((b ** 2).sum(axis=0))
    if floor > 0:
        den = np.maximum(den, floor * den.max())
    return float((num / den).max())


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    return Oracle("f64")


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    return Oracle("f32")


@pytest.fixture(scope="session")
def ref64():
    from oracle import refharness as R
    if not R.available("f64"):
        pytest.skip("oracle/_ref/libnbref_f64.so not built (needs /root/reference)")
    return R.load("f64")


@pytest.fixture(scope="session")
def ref32():
    from oracle import refharness as R
    if not R.available("f32"):
        pytest.skip("oracle/_ref/libnbref_f32.so not built (needs /root/reference)")
    return R.load("f32")
