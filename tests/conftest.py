"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def rel_err(a, b, floor=1e-300):
    """max element-wise relative error where |ref| > floor (SURVEY 8c metric for log-pdfs / rho)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    m = np.abs(b) > floor
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def log_err(a, b):
    """max |a - b| / max(|ref|, 1): for logarithms of normalised quantities (``log_rho`` of VB is ln r_nk, which is
    -1e-9 for the dominant component of a row).  Such an entry is the difference of two numbers of size |ln rho~| >= 1,
    so its rounding error is absolute -- relative to an entry near zero it is unbounded in the reference as well."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def exp_err(a, b, floor=1e-300):
    """max |a - b| / (|ref| max(1, |ln ref|)) for quantities that are exponentials of a compared logarithm
    (r_nk = exp(ln r_nk)): d r / r = d(ln r), and ln r itself is only contracted to 1e-10 RELATIVE, so an entry
    r = 1e-235 (|ln r| = 540) can differ by 540 x 1e-10 between two correct float64 evaluations.  Entries with
    |ln r| <= 1 are held to the plain relative tolerance."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    m = np.abs(b) > floor
    if not m.any():
        return 0.0
    scale = np.abs(b[m]) * np.maximum(1.0, np.abs(np.log(np.abs(b[m]))))
    return float(np.max(np.abs(a[m] - b[m]) / scale))


def mat_err(a, b):
    """max|a-b| / max|ref| per trailing matrix (SURVEY 8c metric for covariance-type outputs)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    a2, b2 = a.reshape(-1, a.shape[-2] * a.shape[-1]), b.reshape(-1, b.shape[-2] * b.shape[-1])
    scale = np.maximum(np.abs(b2).max(axis=1), 1e-300)
    return float(np.max(np.abs(a2 - b2).max(axis=1) / scale))
