"""pypmc_b200 -- B200-native (sm_100a) implementation of pypmc's mixture-density / proposal-update hot path.

Same class API as ``pypmc.density`` and ``pypmc.mix_adapt`` for that path; the N-sized loops run in two
hand-written float64 CUDA kernels behind the C ABI of ``include/pmcb200.h``.  There is no CPU fallback.
"""
from . import _lib
from . import density, mix_adapt, sampler, tools  # noqa: F401

__version__ = "0.1.0"

_lib.load()  # fail loudly at import time if the CUDA library is missing
