"""Importance sampling with the API of pypmc/sampler/importance_sampling.py: ``ImportanceSampler`` (:132-236) and
``combine_weights`` (:238-371).

What differs from the reference is WHERE the N-loops run, not what they compute:

* the reference weights a run with one ``target(x_i) - proposal.evaluate(x_i)`` per sample in Python
  (importance_sampling.py:203-215).  Here the proposal is evaluated for the whole run by ONE launch of kernel K1
  (``proposal.multi_evaluate``), and so is the target when it can be evaluated in batch -- a
  :class:`MixtureDensity` (or its bound ``evaluate``), or any object with ``multi_evaluate``.  Other targets are
  arbitrary Python callables and keep the reference's per-sample loop.
* ``combine_weights`` needs log sum_l N_l q_l(y) for every sample of every run.  That sum is itself a mixture (all
  components of all proposals, weights N_l w_lk), so ONE launch of K1 over all runs' samples yields it, the
  cross-proposal log-sum-exp included; T more launches give q_t on each run's own samples.  The reference's two
  formulations are kept (log scale when all weights are positive, linear otherwise).
"""
from copy import deepcopy as _cp

import numpy as _np

from ..tools._history import History as _History
from ..density._eval import run_k1 as _run_k1
from .. import _device as _dev


def calculate_expectation(samples, weights, f):
    r"""sum_n wbar_n f(x_n) with the weights normalised to one (importance_sampling.py:13-44); ``f`` is an arbitrary
    Python callable, so this is the reference's per-sample loop."""
    assert len(samples) == len(weights), \
        "The number of samples (got %i) must equal the number of weights (got %i)." % (len(samples), len(weights))
    normalization, out = 0., 0.
    for weight, sample in zip(weights, samples):
        normalization += weight
        out += weight * f(sample)
    return out / normalization


def calculate_mean(samples, weights):
    """Mean of weighted samples (importance_sampling.py:46-60)."""
    assert len(samples) == len(weights), \
        "The number of samples (got %i) must equal the number of weights (got %i)." % (len(samples), len(weights))
    return _np.average(samples, axis=0, weights=weights)


def calculate_covariance(samples, weights):
    """Unbiased covariance of weighted samples (importance_sampling.py:62-84), as one weighted outer-product sum."""
    assert len(samples) == len(weights), \
        "The number of samples (got %i) must equal the number of weights (got %i)." % (len(samples), len(weights))
    samples, weights = _np.asarray(samples, dtype=float), _np.asarray(weights, dtype=float)
    sum_weights_sq = weights.sum() ** 2
    sum_sq_weights = (weights ** 2).sum()
    centred = samples - calculate_mean(samples, weights)
    second = _np.einsum('n,ni,nj->ij', weights, centred, centred) / weights.sum()
    return sum_weights_sq / (sum_weights_sq - sum_sq_weights) * second


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


def _as_device_vector(v, like):
    return v if _dev.is_device_tensor(v) else _dev.torch().from_numpy(_np.ascontiguousarray(v, dtype=float)).to(like.device)


def _pooled_proposal(proposals, N):
    """Packed records of ALL components of all ``proposals`` with weights N_l w_lk, or None if the proposals are not
    mixtures of one CUDA-evaluable kind (all Gauss or all StudentT, same dimension)."""
    from ..density.mixture import MixtureDensity
    from .. import _lib
    if not all(isinstance(p, MixtureDensity) for p in proposals):
        return None
    modes = {p._kernel_mode() for p in proposals}
    if len(modes) != 1 or None in modes or len({p.dim for p in proposals}) != 1 or proposals[0].dim > _lib.MAX_DIM:
        return None
    recs = _np.concatenate([_np.stack([c._packed_record() for c in p.components]) for p in proposals])
    w = _np.concatenate([N[l] * _np.asarray(p.weights, dtype=float) for l, p in enumerate(proposals)])
    assert (w >= 0.0).all(), "Found negative weight"
    return _dev.PackedComponents(recs, list(range(len(w))), weights=w), len(w), modes.pop()


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
    y_dev = [_dev.to_device(_np.ascontiguousarray(s_, dtype=float)) for s_ in samples]     # each run uploaded once
    pooled = _pooled_proposal(proposals, N)
    if pooled is not None:
        # sum_l N_l q_l(y) is itself a mixture: all components of all proposals with weights N_l w_lk.  ONE launch of K1
        # over all runs' samples evaluates its logarithm -- the cross-proposal log-sum-exp (logsumexp2D(q, N),
        # :333-362) happens inside the kernel -- and one launch per run gives q_t on the run's own samples: T + 1
        # launches instead of T^2 and no N x T matrix.
        packed, k_all, mode = pooled
        y_all = t_.cat(y_dev) if T > 1 else y_dev[0]
        lse_all = t_.empty(N_total, dtype=t_.float64, device=y_all.device)
        _run_k1(y_all, packed, k_all, mode, logq=lse_all)
        lse = t_.split(lse_all, [int(n) for n in N])
        q_own = [proposals[t].multi_evaluate(y_dev[t]) for t in range(T)]
    else:
        # proposals of other kinds (user densities, mixed mixtures): every proposal on every run, reference formulation
        n_dev = _dev.to_device(N)
        lse, q_own = [], []
        for t in range(T):
            q = t_.stack([_as_device_vector(proposals[l].multi_evaluate(y_dev[t]), y_dev[t]) for l in range(T)], dim=1)
            m = q.max(dim=1).values
            lse.append(m + t_.log((n_dev[None, :] * t_.exp(q - m[:, None])).sum(dim=1)))
            q_own.append(q[:, t])
    for t in range(T):
        out = combined.append(N[t])[:, 0]
        w_t = _dev.to_device(_np.ascontiguousarray(weights[t], dtype=float))
        if log_scale:
            # log w = log omega + log q_t + log sum_j N_j - log sum_l N_l q_l(y)   (:333-362)
            res = t_.exp(t_.log(w_t) + q_own[t] + float(_np.log(N_total)) - lse[t])
        else:
            # [Cor+12] eq. (3) on linear scale (:314-328)
            denominator = t_.exp(lse[t]) / N_total
            res = t_.exp(q_own[t]) * w_t / denominator
        out[:] = res.cpu().numpy()
    if log_scale:
        sum_w = combined[:][:, 0].sum()
        assert sum_w > 0, 'Sum of weights <=0 (%g)' % sum_w
    assert _np.isfinite(combined[:][:, 0]).all(), 'Encountered inf or nan mixture weights'
    return combined
