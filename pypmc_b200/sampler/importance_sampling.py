"""Importance sampling with the API of pypmc/sampler/importance_sampling.py: ``ImportanceSampler`` (:132-236) and
``combine_weights`` (:238-371).

What differs from the reference is WHERE the N-loops run, not what they compute:

* the reference weights a run with one ``target(x_i) - proposal.evaluate(x_i)`` per sample in Python
  (importance_sampling.py:203-215).  Here the proposal is evaluated for the whole run by ONE launch of kernel K1
  (``proposal.multi_evaluate``), and so is the target when it can be evaluated in batch -- a
  :class:`MixtureDensity` (or its bound ``evaluate``), or any object with ``multi_evaluate``.  Other targets are
  arbitrary Python callables and keep the reference's per-sample loop.
* ``combine_weights`` evaluates every proposal on every run's samples (T^2 launches of K1) and keeps the
  reference's two formulations (log scale when all weights are positive, linear otherwise).
"""
from copy import deepcopy as _cp

import numpy as _np

from ..tools._history import History as _History
from .. import _device as _dev


def _batch_form(target):
    """Batch evaluator ``f(samples[N, D]) -> log-values[N]`` of ``target`` if it has one, else None."""
    owner = getattr(target, "__self__", None)
    if owner is not None and getattr(target, "__name__", "") == "evaluate" and hasattr(owner, "multi_evaluate"):
        return owner.multi_evaluate
    if hasattr(target, "multi_evaluate"):
        return target.multi_evaluate
    return None


class ImportanceSampler(object):
    """Generate weighted samples from ``target`` using ``proposal`` (importance_sampling.py:132-236).

    :param target: callable returning the log of the target at one point, or a density with ``multi_evaluate``.
    :param proposal: density with ``propose`` and ``evaluate`` / ``multi_evaluate`` (deep-copied).
    :param indicator: optional callable; points where it is False get target value -inf.
    :param prealloc: number of samples to reserve memory for.
    :param save_target_values: keep the target's log-values in ``self.target_values``.
    :param rng: numpy-style generator handed to ``proposal.propose``.
    """

    def __init__(self, target, proposal, indicator=None, prealloc=0, save_target_values=False, rng=_np.random.mtrand):
        self.proposal = _cp(proposal)
        self.rng = rng
        self._target_point = target if callable(target) else target.evaluate
        self._target_batch = _batch_form(target)
        self._indicator = indicator
        self.target_values = _History(1, prealloc) if save_target_values else None
        self.weights = _History(1, prealloc)
        self.samples = _History(proposal.dim, prealloc)

    def target(self, x):
        """log target at one point, -inf outside the indicator's support (tools/indicator/_indicator_merge.py)."""
        if self._indicator is not None and not self._indicator(x):
            return -_np.inf
        return self._target_point(x)

    def clear(self):
        """Forget samples, weights and target values; the proposal is untouched."""
        self.samples.clear()
        self.weights.clear()
        if self.target_values is not None:
            self.target_values.clear()

    def run(self, N=1, trace_sort=False):
        """Draw ``N`` samples from the proposal and weight them; with ``trace_sort`` the samples are ordered by
        component and the array of responsible components is returned (importance_sampling.py:158-195)."""
        if N == 0:
            return 0
        this_run = self.samples.append(N)
        origin = None
        if trace_sort:
            this_run[:], origin = self.proposal.propose(N, self.rng, trace=True, shuffle=False)
        else:
            this_run[:] = self.proposal.propose(N, self.rng)
        self._calculate_weights(this_run, N)
        return origin

    def _target_values_of(self, this_samples, N):
        if self._target_batch is not None:
            vals = _np.array(self._target_batch(_np.ascontiguousarray(this_samples)), dtype=float)
            if self._indicator is not None:
                for i in range(N):
                    if not self._indicator(this_samples[i]):
                        vals[i] = -_np.inf
            return vals
        vals = _np.empty(N)
        for i in range(N):                       # arbitrary Python target: the reference's loop (:203-207)
            tmp = self.target(this_samples[i])
            vals[i] = tmp.item() if _np.ndim(tmp) != 0 else tmp
        return vals

    def _calculate_weights(self, this_samples, N):
        """w_i = exp(log target(x_i) - log proposal(x_i)) for the run (importance_sampling.py:197-215)."""
        this_weights = self.weights.append(N)[:, 0]
        target_values = self._target_values_of(this_samples, N)
        if self.target_values is not None:
            self.target_values.append(N)[:, 0] = target_values
        if hasattr(self.proposal, "multi_evaluate"):
            log_q = _np.asarray(self.proposal.multi_evaluate(_np.ascontiguousarray(this_samples)))
        else:
            log_q = _np.array([self.proposal.evaluate(x) for x in this_samples], dtype=float)
        _np.exp(target_values - log_q, out=this_weights)


def combine_weights(samples, weights, proposals):
    """`Deterministic mixture weights` [Cor+12] of importance samples drawn for the same target from different
    proposals (importance_sampling.py:238-371).  Returns a :class:`History` with one run per proposal.

    :param samples: iterable of (N_t x D) arrays, one per step.
    :param weights: iterable of 1-d arrays, the standard weights P(x)/q_t(x) of each step.
    :param proposals: iterable of the densities the samples were drawn from.
    """
    samples = [_np.asarray(s) for s in samples]
    weights = [_np.asarray(w) for w in weights]
    assert len(samples) == len(weights), \
        "Got %i importance-sampling runs but %i weights" % (len(samples), len(weights))
    assert len(samples) == len(proposals), \
        "Got %i importance-sampling runs but %i proposal densities" % (len(samples), len(proposals))
    T = len(proposals)
    N = _np.empty(T)
    for i in range(T):
        assert samples[i].ndim == 2, '``samples[%i]`` is not matrix like.' % i
        dim = samples[0].shape[-1]
        assert samples[i].shape[-1] == dim, \
            "Dimension of samples[0] (%i) does not match the dimension of samples[%i] (%i)" % (dim, i, samples[i].shape[-1])
        N[i] = len(samples[i])
        assert N[i] == len(weights[i]), \
            'Length of weights[%i] (%i) does not match length of samples[%i] (%i)' % (i, N[i], i, len(weights[i]))
    N_total = int(N.sum())
    combined = _History(1, N_total)
    log_scale = all((w > 0.0).all() for w in weights)       # all weights positive => log scale (:300-308)

    t_ = _dev.torch()
    for t in range(T):
        out = combined.append(N[t])[:, 0]
        y = _dev.to_device(_np.ascontiguousarray(samples[t], dtype=float))      # uploaded once, evaluated T times
        q = t_.stack([proposals[l].multi_evaluate(y) for l in range(T)], dim=1)  # K1: log q_l(y_i^t), [N_t, T] on device
        w_t = _dev.to_device(_np.ascontiguousarray(weights[t], dtype=float))
        n_dev = _dev.to_device(N)
        if log_scale:
            # log w = log omega + log q_t + log sum_j N_j - log sum_l N_l q_l(y)   (:333-362); the weighted row-wise
            # log-sum-exp (logsumexp2D(q, N), _regularize.pyx:57-83) as max + log sum N_l exp(q_l - max)
            m = q.max(dim=1).values
            lse = m + t_.log((n_dev[None, :] * t_.exp(q - m[:, None])).sum(dim=1))
            res = t_.exp(t_.log(w_t) + q[:, t] + float(_np.log(N_total)) - lse)
        else:
            # [Cor+12] eq. (3) on linear scale (:314-328)
            denominator = (n_dev[None, :] * t_.exp(q)).sum(dim=1) / N_total
            res = t_.exp(q[:, t]) * w_t / denominator
        out[:] = res.cpu().numpy()
    if log_scale:
        sum_w = combined[:][:, 0].sum()
        assert sum_w > 0, 'Sum of weights <=0 (%g)' % sum_w
    assert _np.isfinite(combined[:][:, 0]).all(), 'Encountered inf or nan mixture weights'
    return combined
