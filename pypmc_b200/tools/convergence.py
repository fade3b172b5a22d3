"""Quality measures of a weight vector with the API of pypmc/tools/convergence.py (``perp`` :6-39, ``ess`` :42-72).

A numpy array is reduced with numpy like the reference does.  Weights that live on the device -- a torch CUDA tensor, or a
:class:`~pypmc_b200.mix_adapt.pmc.DeviceSamples` whose weights were formed by ``weigh`` -- are reduced by kernel K4
(csrc/k4_weights.cuh): one streaming pass leaves sum w, sum w^2 and sum w log w, from which both measures follow;
for a ``DeviceSamples`` that pass already happened when the weights were formed.
"""
import numpy as _np

from .. import _device as _dev
from .. import _lib


def _device_sums(weights):
    """(sum w, sum w^2, sum w log w, N) of a CUDA weight vector: one launch of K4 on log w."""
    t = _dev.torch()
    w = weights.contiguous().to(t.float64)
    index = w.device.index
    sums = t.empty(5, dtype=t.float64, device=w.device)
    _lib.Context.get(index).importance_weights(None, t.log(w), w.numel(), None, sums, _dev.current_stream_ptr(index))
    s = sums.cpu().numpy()
    return s[0], s[2], s[3], float(w.numel())


def perp(weights):
    r"""Normalised perplexity exp(H)/N with H = -sum w_i log w_i over the normalised weights (zeros contribute 0);
    0 is terrible, 1 is perfect."""
    if hasattr(weights, "weight_sums"):                  # DeviceSamples
        return weights.perp()
    if _dev.is_device_tensor(weights):
        s0, _, s3, n = _device_sums(weights)
        return float(_np.exp(_np.log(s0) - s3 / s0) / n)   # H = log S - (sum w log w) / S
    w = _np.asarray(weights) / _np.sum(weights)
    entr = -_np.sum(w * _np.log(_np.where(w == 0, 1.0, w)))
    return _np.exp(entr) / len(w)


def ess(weights):
    r"""Normalised effective sample size 1 / (1 + C^2), C^2 = mean (N w_i - 1)^2 over the normalised weights [LC95]."""
    if hasattr(weights, "weight_sums"):
        return weights.ess()
    if _dev.is_device_tensor(weights):
        s0, s2, _, n = _device_sums(weights)
        return float(s0 * s0 / (n * s2))                   # 1 + C^2 = N sum w^2 / (sum w)^2
    w = _np.asarray(weights) / _np.sum(weights)
    return 1.0 / (1.0 + _np.sum((len(w) * w - 1) ** 2) / len(w))
