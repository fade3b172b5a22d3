"""CPU tier: the C-ABI library loads, exports every symbol ``include/pmcb200.h`` declares, its host-only entry
points work without a GPU, and every compute path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "pmcb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmcb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pypmc_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), "libpmcb200.so does not export %s" % name
    # the ctypes binding types exactly the declared set: nothing missing, nothing undeclared
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().pmcb200_version() >= 100


def test_pack_record_layout_on_host():
    from pypmc_b200 import _lib
    for d in (1, 2, 5, 30):
        dp = (d + 1) & ~1
        nt = (dp // 2) * (dp // 2 + 1) * 2
        assert _lib.record_len(d) == nt + dp + _lib.NUM_SCALARS
        rng = np.random.default_rng(d)
        t = np.tril(rng.normal(size=(d, d)))
        c, sc = rng.normal(size=d), rng.normal(size=_lib.NUM_SCALARS)
        rec = _lib.pack_record(t, c, sc)
        for i in range(dp):
            for j in range(i + 1):
                r, p = i // 2, j // 2
                got = rec[2 * r * (r + 1) + 4 * p + 2 * (i % 2) + (j % 2)]
                assert got == (t[i, j] if i < d and j < d else 0.0)
        np.testing.assert_array_equal(rec[nt:nt + d], c)
        assert (rec[nt + d:nt + dp] == 0).all()
        np.testing.assert_array_equal(rec[nt + dp:], sc)
    with pytest.raises(ValueError):
        _lib.record_len(_lib.MAX_DIM + 1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pypmc_b200 import _lib
    from pypmc_b200.density.mixture import create_gaussian_mixture
    assert _lib.device_count() == 0
    h = ctypes.c_void_p()
    assert _lib.load().pmcb200_create(0, ctypes.byref(h)) != 0
    assert b"cuda" in _lib.load().pmcb200_last_error().lower()
    mix = create_gaussian_mixture(np.zeros((2, 3)), np.array([np.eye(3)] * 2))
    with pytest.raises(_lib.PmcB200Error, match="no CPU fallback"):
        mix.multi_evaluate(np.zeros((4, 3)))


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under pypmc_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pypmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
