"""Probability densities with the public API of ``pypmc.density`` (base.py, gauss.pyx, student_t.pyx,
mixture.pyx); every N-sized loop runs in the CUDA kernels behind ``libpmcb200.so``."""
from . import base, gauss, student_t, mixture  # noqa: F401
