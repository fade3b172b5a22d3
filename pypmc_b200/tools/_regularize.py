"""``regularize`` and ``logsumexp2D`` -- host mirrors of pypmc/tools/_regularize.pyx:6-17 and :57-83.

The N-sized log-sum-exp of a Gauss / StudentT mixture runs inside the CUDA kernel K1 (csrc/k1_mma_eval.cuh and the
DFMA forms).  ``logsumexp2D`` here serves only mixtures with user-defined components (``MixtureDensity``'s generic
route), whose N-loop is the user's own per-point ``evaluate`` anyway (base.py:42-50).
"""
import numpy as _np

tiny = float(_np.finfo("d").tiny)


def regularize(x):
    """Replace exact zeros by the smallest positive normal double, in place; return ``x``."""
    x[_np.where(x == 0)] = tiny
    return x


def logsumexp(a, weights):
    """log(sum_i weights[i] exp(a[i])) as max + log(sum w exp(a - max)) (_regularize.pyx:19-55), 1-d."""
    a = _np.asarray(a, dtype=float)
    weights = _np.asarray(weights, dtype=float)
    assert a.ndim == 1 and a.shape == weights.shape
    m = a.max() if len(a) else -_np.finfo("d").max
    return float(_np.log((weights * _np.exp(a - m)).sum()) + m)


def logsumexp2D(a, weights):
    """res[n] = max_k a[n, k] + log(sum_k weights[k] exp(a[n, k] - max_k a[n, k])); the maximum runs over ALL columns,
    also those with weight zero (_regularize.pyx:72-81)."""
    a = _np.asarray(a, dtype=float)
    weights = _np.asarray(weights, dtype=float)
    assert a.ndim == 2 and len(weights) == a.shape[1]
    assert (weights >= 0.0).all(), "Found negative weight"
    if a.shape[0] == 0:
        return _np.empty(0)
    m = a.max(axis=1)
    with _np.errstate(divide="ignore", invalid="ignore"):
        return m + _np.log((weights[None, :] * _np.exp(a - m[:, None])).sum(axis=1))
