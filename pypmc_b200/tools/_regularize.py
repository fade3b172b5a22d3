"""``regularize`` -- host mirror of pypmc/tools/_regularize.pyx:6-17 (K-sized vectors only).

The N-sized log-sum-exp loops of that module (``logsumexp2D``, _regularize.pyx:57-83) run inside the
CUDA kernel K1 (csrc/k1_mixture_eval.cuh); there is no host implementation of them in this package.
"""
import numpy as _np

tiny = float(_np.finfo("d").tiny)


def regularize(x):
    """Replace exact zeros by the smallest positive normal double, in place; return ``x``."""
    x[_np.where(x == 0)] = tiny
    return x
