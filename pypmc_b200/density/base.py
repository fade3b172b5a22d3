"""Abstract density protocol -- same surface as pypmc/density/base.py:7-108."""
import numpy as _np


class ProbabilityDensity(object):
    """Base class of a probability density usable as importance-sampling proposal
    (pypmc/density/base.py:7-66): ``evaluate``, ``multi_evaluate`` and ``propose``."""
    dim = 0

    def __init__(self):
        raise NotImplementedError('Do not create instances from this class, use derived classes instead.')

    def evaluate(self, x):
        """log q(x) for one point ``x``."""
        raise NotImplementedError()

    def multi_evaluate(self, x, out=None):
        """log q(x_n) for each row of ``x``; written into ``out`` when given (same object returned).
        Generic per-point loop for user-defined densities that only implement ``evaluate``
        (base.py:42-50); ``Gauss``, ``StudentT`` and ``MixtureDensity`` override it with the CUDA path."""
        if out is None:
            out = _np.empty(len(x))
        else:
            assert len(out) == len(x)
        for i, point in enumerate(x):
            out[i] = self.evaluate(point)
        return out

    def propose(self, N=1, rng=_np.random.mtrand):
        """Draw ``N`` points using ``rng``."""
        raise NotImplementedError()


class LocalDensity(object):
    """Local (Markov-chain) proposal protocol, pypmc/density/base.py:68-108."""
    symmetric = False

    def __init__(self):
        raise NotImplementedError('Do not create instances from this class, use derived classes instead.')

    def evaluate(self, x, y):
        raise NotImplementedError()

    def propose(self, y, rng=_np.random.mtrand):
        raise NotImplementedError()
