"""Mixture densities with the API of pypmc/density/mixture.pyx (``MixtureDensity`` :21-212 and the
create/recover helpers :214-350).  ``multi_evaluate`` is one launch of the fused CUDA kernel K1."""
import numpy as _np
from copy import deepcopy as _deepcopy

from .base import ProbabilityDensity
from .gauss import Gauss
from .student_t import StudentT
from ._eval import run_k1
from .. import _lib
from .. import _device as _dev


class MixtureDensity(ProbabilityDensity):
    """Weighted sum of component densities (mixture.pyx:21-212).

    :param components: iterable of densities (deep-copied).
    :param weights: iterable of floats, normalised on construction; equal weights if omitted.
    """

    def __init__(self, components, weights=None):
        self.components = [_deepcopy(c) for c in components]
        assert self.components, "Must have at least one component!"
        self.dim = self.components[0].dim
        _np.testing.assert_equal([c.dim for c in self.components], [self.dim] * len(self.components))
        if weights is None:
            self.weights = _np.ones(len(self.components))
        else:
            self.weights = _np.array(weights, dtype=float)
            assert len(self.weights) == len(self.components)
        self.normalize()

    def __len__(self):
        k = len(self.components)
        assert k == len(self.weights)
        return k

    def __deepcopy__(self, memo):
        # what copy.deepcopy would build, without its generic walk over the list and the dict (0.4 ms per update at
        # K = 32): components through their own __deepcopy__, arrays copied flat, anything else deep-copied
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for key, val in self.__dict__.items():
            if key == "components":
                new.components = [_deepcopy(c, memo) for c in val]
            elif isinstance(val, _np.ndarray):
                new.__dict__[key] = val.copy()
            else:
                new.__dict__[key] = _deepcopy(val, memo)
        return new

    def normalize(self):
        """Scale the component weights to sum to one."""
        self.weights /= self.weights.sum()

    def normalized(self):
        """True if the component weights sum to one (within ``allclose``)."""
        return bool(_np.allclose(self.weights.sum(), 1.0))

    def prune(self, threshold=0.0):
        """Remove components with weight <= ``threshold``; return ``[(index, component, weight), ...]``
        in descending index order (mixture.pyx:64-94)."""
        removed = []
        for idx in range(len(self.weights) - 1, -1, -1):
            if self.weights[idx] <= threshold:
                removed.append((idx, self.components.pop(idx), self.weights[idx]))
        self.weights = _np.delete(self.weights, [r[0] for r in removed])
        return removed

    # -- CUDA path ------------------------------------------------------------------------------------------
    def _kernel_mode(self):
        """K1 mode if every component is a ``Gauss`` or every component is a ``StudentT``; else None."""
        if all(type(c) is Gauss or isinstance(c, Gauss) for c in self.components):
            return _lib.MODE_GAUSS
        if all(isinstance(c, StudentT) for c in self.components):
            return _lib.MODE_STUDENT_T
        return None

    def _require_mode(self):
        mode = self._kernel_mode()
        if mode is None:
            raise NotImplementedError(
                "this operation (device-side propose / PMC update) needs a mixture whose components are all Gauss or all "
                "StudentT; mixtures of other kinds can only be evaluated (MixtureDensity.multi_evaluate)")
        if self.dim > _lib.MAX_DIM:
            raise NotImplementedError("dimension %d exceeds the CUDA kernels' maximum of %d" % (self.dim, _lib.MAX_DIM))
        return mode

    def _packed(self, components=None, compact=False):
        """Packed records of ``components`` (default all); ``compact`` numbers the output columns 0..len-1."""
        ks = list(range(len(self))) if components is None else [int(k) for k in components]
        recs = _np.stack([self.components[k]._packed_record() for k in ks])
        cols = list(range(len(ks))) if compact else ks
        return _dev.PackedComponents(recs, cols, weights=self.weights[ks])

    def evaluate(self, x, individual=False):
        """log q(x) for one point (mixture.pyx:101-110); with ``individual`` also the component log-pdfs."""
        x = _np.ascontiguousarray(x, dtype=float).reshape(1, -1)
        ind = _np.empty((1, len(self)))
        res = float(self.multi_evaluate(x, individual=ind)[0])
        return (res, ind[0]) if individual else res

    def multi_evaluate(self, x, out=None, individual=None, components=None):
        """Evaluate the mixture at every row of ``x`` (mixture.pyx:112-156).

        Returns log q(x_n) (in ``out`` if given) unless ``components`` is given, in which case only the
        columns ``individual[:, k]``, k in ``components``, are filled and None is returned.  ``x`` may be
        a float64 numpy array (streamed through the GPU) or a float64 torch CUDA tensor (device resident;
        outputs are CUDA tensors).  Assumes normalised weights.
        """
        x = _dev.as_samples(x)
        n, k = x.shape[0], len(self)
        assert x.shape[1] == self.dim, "The points in ``x`` have the wrong dimension (%i instead of %i)" % (x.shape[1], self.dim)
        if individual is not None:
            assert len(x) == len(individual), "For the provided ``x``, ``individual`` must have shape %s" % ((n, k),)
            assert individual.shape[1] == k, "For the provided ``x``, ``individual`` must have shape %s" % ((n, k),)
        mode = self._kernel_mode()
        if mode is None or self.dim > _lib.MAX_DIM:
            return self._multi_evaluate_generic(x, out, individual, components)
        assert (self.weights >= 0.0).all(), "Found negative weight"
        on_device = _dev.is_device_tensor(x)
        t = _dev.torch() if on_device else None

        if components is not None:
            assert out is None, 'If ``components`` is not None, ``out`` must be None.'
            comps = [int(c) for c in components]
            if not comps:
                return None
            if on_device:
                if individual is None:
                    individual = t.empty((n, k), dtype=t.float64, device=x.device)
                run_k1(x, self._packed(comps), k, mode, lp=individual)
            else:
                if individual is None:
                    individual = _np.empty((n, k))
                tmp = _np.empty((n, len(comps)))
                run_k1(x, self._packed(comps, compact=True), len(comps), mode, lp=tmp)
                individual[:, comps] = tmp
            return None

        if out is not None:
            assert len(out) == len(x), '``out`` must have length %i' % (len(x))
        if on_device:
            res = out if out is not None else t.empty(n, dtype=t.float64, device=x.device)
            run_k1(x, self._packed(), k, mode, logq=res, lp=individual)
            return res
        direct_out = out is None or (isinstance(out, _np.ndarray) and out.dtype == _np.float64
                                     and out.flags.c_contiguous and out.ndim == 1)
        res = (out if out is not None else _np.empty(n)) if direct_out else _np.empty(n)
        direct_ind = individual is None or (individual.dtype == _np.float64 and individual.flags.c_contiguous)
        ind = individual if direct_ind else _np.empty((n, k))
        run_k1(x, self._packed(), k, mode, logq=res, lp=ind)
        if not direct_ind:
            individual[:] = ind
        if not direct_out:
            out[:] = res
            return out
        return res

    def _multi_evaluate_generic(self, x, out, individual, components):
        """Mixtures that are not all-Gauss or all-StudentT (mixed kinds, user-defined densities): mixture.pyx:138-156
        as written -- every component fills its column of ``individual`` through its OWN ``multi_evaluate``, then the
        weighted log-sum-exp.  Gauss and StudentT components still go through kernel K1 (one launch per kind); a
        user density runs whatever it implements (the per-point loop of base.py:42-50 if nothing else), which is the
        reference's contract for such components (mixture_test.py:15-23, 82-104)."""
        from ..tools._regularize import logsumexp2D
        n, k = x.shape[0], len(self)
        on_device = _dev.is_device_tensor(x)
        xh = x.cpu().numpy() if on_device else x
        ks = list(range(k)) if components is None else [int(c) for c in components]
        if components is not None:
            assert out is None, 'If ``components`` is not None, ``out`` must be None.'
            if not ks:
                return None
        elif out is not None:
            assert len(out) == n, '``out`` must have length %i' % n
        host_ind = isinstance(individual, _np.ndarray)
        ind = individual if (host_ind and individual.dtype == _np.float64) else _np.empty((n, k))
        if individual is not None and ind is not individual and components is not None:
            ind[:] = individual.cpu().numpy() if _dev.is_device_tensor(individual) else individual   # untouched columns survive
        rest = list(ks)
        if self.dim <= _lib.MAX_DIM:
            for kind, kmode in ((Gauss, _lib.MODE_GAUSS), (StudentT, _lib.MODE_STUDENT_T)):
                idx = [j for j in ks if isinstance(self.components[j], kind)]
                if idx and n > 0:
                    tmp = _np.empty((n, len(idx)))
                    run_k1(_np.ascontiguousarray(xh), self._packed(idx, compact=True), len(idx), kmode, lp=tmp)
                    ind[:, idx] = tmp
                rest = [j for j in rest if j not in idx]
        for j in rest:
            col = _np.empty(n)
            self.components[j].multi_evaluate(xh, col)
            ind[:, j] = col
        if individual is not None and ind is not individual:
            if _dev.is_device_tensor(individual):
                individual.copy_(_dev.torch().from_numpy(ind))
            else:
                individual[:] = ind
        if components is not None:
            return None
        res = logsumexp2D(ind, self.weights)
        if on_device:
            res = _dev.torch().from_numpy(res).to(x.device)
        if out is not None:
            if _dev.is_device_tensor(out):
                out.copy_(res if on_device else _dev.torch().from_numpy(res))
            else:
                out[:] = res.cpu().numpy() if on_device else res
            return out
        return res

    def propose(self, N=1, rng=_np.random.mtrand, trace=False, shuffle=True):
        """Draw ``N`` points (mixture.pyx:159-212): multinomial counts per component, component draws in
        component order, then either the origin array (``trace``) or an in-place shuffle (``shuffle``)."""
        if trace and shuffle:
            raise ValueError('Either ``shuffle`` or ``trace`` must be ``False``!')
        counts = rng.multinomial(N, self.weights)
        samples = _np.empty((N, self.dim))
        start = 0
        for comp, cnt in zip(self.components, counts):
            if cnt != 0:
                samples[start:start + cnt] = comp.propose(cnt)
            start += cnt
        if trace:
            return samples, _np.repeat(_np.arange(len(self.components)), counts)
        if shuffle:
            rng.shuffle(samples)
        return samples


    def propose_device(self, N=1, rng=_np.random.mtrand, trace=False, shuffle=False, seed=None, index0=0, device=None):
        """Draw ``N`` points on the GPU (kernel K3) and return them as a float64 CUDA tensor [N, dim] (with
        ``trace`` also the int64 origin tensor, like :meth:`propose`).

        The per-component counts come from ``rng.multinomial(N, self.weights)`` exactly as in the reference
        (mixture.pyx:193) and the samples are laid out in component order; the normal / chi-square variates are a
        Philox4x32-10 stream keyed by (``seed``, ``index0`` + row) -- ``seed`` defaults to a draw from ``rng``,
        ``index0`` is this rank's global row offset when several ranks draw from one logical stream.
        ``shuffle`` applies a device-side random permutation (off by default: importance sampling and the
        PMC update do not depend on the order)."""
        if trace and shuffle:
            raise ValueError('Either ``shuffle`` or ``trace`` must be ``False``!')
        mode = self._require_mode()
        t = _dev.torch()
        device = _lib.default_device() if device is None else int(device)
        dev = "cuda:%d" % device
        counts = _np.asarray(rng.multinomial(N, self.weights), dtype=_np.int64)
        starts = _np.concatenate([[0], _np.cumsum(counts)]).astype(_np.int64)
        if seed is None:
            seed = int(rng.randint(0, 2 ** 31 - 1))
        k, d = len(self), self.dim
        means = _dev.to_device(_np.array([c.mu for c in self.components], dtype=float).reshape(k, d), device)
        chol_src = [c._local_gauss.cholesky_sigma if mode == _lib.MODE_GAUSS else c._local_t.cholesky_sigma
                    for c in self.components]
        chol = _dev.to_device(_np.array(chol_src, dtype=float).reshape(k, d, d), device)
        dofs = None
        if mode == _lib.MODE_STUDENT_T:
            dofs = _dev.to_device(_np.array([c.dof for c in self.components], dtype=float), device)
        x = t.empty((N, d), dtype=t.float64, device=dev)
        latent = t.empty(N, dtype=t.int32, device=dev) if trace else None
        _lib.Context.get(device).mixture_propose(N, d, k, means, chol, dofs, starts, seed, index0, x, d, latent,
                                                 _dev.current_stream_ptr(device))
        if trace:
            return x, latent.to(t.int64)
        if shuffle:
            g = t.Generator(device=dev).manual_seed(seed)
            x = x[t.randperm(N, device=dev, generator=g)]
        return x


def create_gaussian_mixture(means, covs, weights=None):
    """:class:`MixtureDensity` of :class:`Gauss` components (mixture.pyx:214-246)."""
    assert len(means) == len(covs), \
        'Number of means (%i) does not match number of covariances (%i)' % (len(means), len(covs))
    return MixtureDensity([Gauss(m, c) for m, c in zip(means, covs)], weights)


def recover_gaussian_mixture(mixture):
    """``(means, covs, weights)`` of a Gaussian mixture (mixture.pyx:248-278)."""
    means = _np.array([c.mu for c in mixture.components]).reshape(len(mixture.components), mixture.dim)
    covs = _np.array([c.sigma for c in mixture.components]).reshape(len(means), mixture.dim, mixture.dim)
    return means, covs, _np.array(mixture.weights)


def create_t_mixture(means, covs, dofs, weights=None):
    """:class:`MixtureDensity` of :class:`StudentT` components (mixture.pyx:280-318)."""
    assert (len(means) == len(covs)) and (len(means) == len(dofs)), \
        'Number of ``means`` (%i), ``covs`` (%i) and ``dofs`` (%i) do not match.' % (len(means), len(covs), len(dofs))
    return MixtureDensity([StudentT(m, c, d) for m, c, d in zip(means, covs, dofs)], weights)


def recover_t_mixture(mixture):
    """``(means, covs, dofs, weights)`` of a Student-t mixture (mixture.pyx:320-350)."""
    means, covs, weights = recover_gaussian_mixture(mixture)
    return means, covs, _np.array([c.dof for c in mixture.components], dtype=float), weights
