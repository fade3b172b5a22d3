"""Quality measures of a weight vector with the API of pypmc/tools/convergence.py (``perp`` :6-39, ``ess`` :42-72).
O(N) reductions over importance weights: a numpy array is reduced with numpy, a torch CUDA tensor (weights that
never left the device, e.g. ``exp(log_target - proposal.multi_evaluate(x))``) on the device."""
import numpy as _np

from .. import _device as _dev


def perp(weights):
    r"""Normalised perplexity exp(H)/N with H = -sum w_i log w_i over the normalised weights (zeros contribute 0);
    0 is terrible, 1 is perfect."""
    if _dev.is_device_tensor(weights):
        t = _dev.torch()
        w = weights / weights.sum()
        entr = -(w * t.log(t.where(w == 0, t.ones_like(w), w))).sum()
        return float(t.exp(entr) / w.numel())
    w = _np.asarray(weights) / _np.sum(weights)
    entr = -_np.sum(w * _np.log(_np.where(w == 0, 1.0, w)))
    return _np.exp(entr) / len(w)


def ess(weights):
    r"""Normalised effective sample size 1 / (1 + C^2), C^2 = mean (N w_i - 1)^2 over the normalised weights [LC95]."""
    if _dev.is_device_tensor(weights):
        w = weights / weights.sum()
        n = w.numel()
        return float(1.0 / (1.0 + ((n * w - 1) ** 2).sum() / n))
    w = _np.asarray(weights) / _np.sum(weights)
    return 1.0 / (1.0 + _np.sum((len(w) * w - 1) ** 2) / len(w))
