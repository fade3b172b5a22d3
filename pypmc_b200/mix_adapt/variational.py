"""Variational-Bayes Gaussian mixture inference with the API of ``GaussianInference``
(pypmc/mix_adapt/variational.pyx:27-1033; notation of Bishop, PRML ch. 10.2).

E-step = one launch of kernel K1 in VB mode (expectation of the Gauss exponent, log rho, softmax -> r,
sum r log r) + one launch of kernel K2 (N_k, x_mean_k, S_k) + one all-reduce of the statistics packet when
the data are sharded over GPUs.  M-step, the bound's K-sized terms and pruning are host arithmetic.  The
N x K array ``r`` lives on the device and is downloaded only when the attribute is read; ``log_rho`` and
``expectation_gauss_exponent``, which no step of the algorithm consumes once sum r log r is fused into the launch,
are recomputed by one extra launch when read.  ``VBMerge`` (variational.pyx:1035-1218) loops over input *components*,
not samples, and is outside this package's scope.
"""
from __future__ import division

import logging

import numpy as _np
from scipy.special import digamma as _digamma
from scipy.special import gammaln as _gammaln

from ..density.gauss import Gauss
from ..density.mixture import MixtureDensity, recover_gaussian_mixture
from ..density._eval import run_k1
from ..tools._linalg import chol_inv_det, chol_inv_det_batch, tri_from_precision
from .. import _device as _dev
from .. import _lib
from .. import parallel as _parallel
from ._stats import PacketLayout, moments_from_stats, shift_groups, grouped_suffstats, small_problem, two_pass_suffstats
from .pmc import DeviceSamples

logger = logging.getLogger(__name__)


class GaussianInference(object):
    """Approximate the density behind ``data`` by a Gaussian mixture with variational Bayes.

    :param data: (N x D) or (N,) float64 array (numpy or torch CUDA tensor): this rank's samples.
    :param components: number of mixture components K (taken from ``initial_guess`` if that is a mixture).
    :param weights: N nonnegative finite sample weights, optional.
    :param initial_guess: "first", "random" or a :class:`MixtureDensity` of Gauss components.

    Keyword arguments are handed to :meth:`set_variational_parameters`.
    """

    def __init__(self, data, components=0, weights=None, initial_guess="first", **kwargs):
        if _dev.is_device_tensor(data):
            data2d = data.reshape(data.shape[0], 1) if data.dim() == 1 else data
        else:
            data = _np.asarray(data, dtype=_np.float64)
            data2d = data.reshape(data.shape[0], 1) if data.ndim == 1 else data
        self.data = data2d
        n_local = int(data2d.shape[0])
        self.dim = int(data2d.shape[1])
        self.N = int(round(self._global_sum(float(n_local))))       # all ranks' samples

        self.weights = None
        w_dev = None
        if weights is not None:
            t = _dev.torch()
            w_dev = _dev.to_device(weights if _dev.is_device_tensor(weights) else _np.asarray(weights, dtype=float))
            assert tuple(w_dev.shape) == (n_local,), \
                "The number of samples (%s) does not match the number of weights (%s)" % (n_local, w_dev.shape[0])
            assert bool(t.isfinite(w_dev).all()), 'Some weights are not finite; i.e., inf or nan\n' + str(weights)
            sum_w = self._global_sum(float(w_dev.sum()))
            assert sum_w > 0, 'Sum of weights <= 0 (%g)' % sum_w
            w_dev = self.N * (w_dev / sum_w)                         # normalised to N, variational.pyx:94
            self.weights = w_dev.cpu().numpy()
        self._ds = DeviceSamples(data2d, w_dev)

        self._initialize_K(initial_guess, components, kwargs)
        self.set_variational_parameters(initial_guess=initial_guess, **kwargs)
        if not isinstance(initial_guess, str):
            self._parse_initial_guess(initial_guess)
        self._initialize_intermediate()
        self.E_step()

    # ------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _global_sum(value):
        if not _parallel.enabled():
            return value
        t = _dev.torch()
        buf = t.tensor([value], dtype=t.float64, device="cuda:%d" % _lib.default_device())
        _parallel.allreduce_(buf)
        return float(buf[0])

    def _initialize_K(self, initial_guess, components, kwargs):
        if not isinstance(initial_guess, str):
            self.K = len(initial_guess)
            for name in ('m', 'W', 'alpha', 'beta', 'nu'):
                if name in kwargs:
                    raise ValueError('Specify EITHER ``%s`` OR ``initial_guess``' % name)
        elif components > 0:
            self.K = components
        else:
            raise ValueError('Specify either `components` or a mixture density as `initial_guess` to set the initial values')

    def _initialize_intermediate(self):
        K, D = self.K, self.dim
        self.x_mean_comp = _np.zeros((K, D))
        self.S = _np.empty((K, D, D))
        self.N_comp = _np.zeros(K)
        self.inv_N_comp = _np.zeros(K)
        self.expectation_det_ln_lambda = _np.zeros(K)
        self.expectation_ln_pi = _np.zeros(K)
        self._r_dev = None
        self._host_cache = {}
        self._estep_packed = None

    def _k_vector(self, kwargs, name, default, lower=0.0):
        v = kwargs.pop(name, default)
        v = v * _np.ones(self.K) if not _np.iterable(v) else _np.array(v, dtype=float)
        if v.ndim != 1:
            raise ValueError('%s is not a vector but has shape %s' % (name, v.shape))
        if len(v) != self.K:
            raise ValueError('len(%s)=%d does not match K=%d' % (name, len(v), self.K))
        if not (v > lower).all():
            raise ValueError('All elements of %s must exceed %g. %s=%s' % (name, lower, name, v))
        return v

    def set_variational_parameters(self, *args, **kwargs):
        r"""Reset prior (``alpha0, beta0, nu0, m0, W0``) and initial posterior (``alpha, beta, nu, m, W``)
        hyper-parameters (variational.pyx:361-569).  Scalars are promoted to K-vectors, a single ``m0`` / ``W0``
        to K copies.  Defaults: alpha0 = beta0 = 1e-5, nu0 = D - 1 + 1e-5, m0 = 0, W0 = identity; ``m`` from
        ``initial_guess`` ("first"/"random" data points)."""
        if args:
            raise TypeError('keyword args only')
        K, D = self.K, self.dim
        self.alpha0 = self._k_vector(kwargs, 'alpha0', 1e-5)
        self.alpha = self._k_vector(kwargs, 'alpha', self.alpha0.copy())
        self.beta0 = self._k_vector(kwargs, 'beta0', 1e-5)
        self.beta = self._k_vector(kwargs, 'beta', self.beta0.copy())
        nu_min = D - 1.
        self.nu0 = self._k_vector(kwargs, 'nu0', nu_min + 1e-5, lower=nu_min)
        self.nu = self._k_vector(kwargs, 'nu', self.nu0.copy(), lower=nu_min)

        m0 = _np.array(kwargs.pop('m0', _np.zeros(D)), dtype=float)
        self.m0 = _np.tile(m0, (K, 1)) if m0.shape == (D,) else m0
        initial_guess = kwargs.pop('initial_guess')
        m = kwargs.pop('m', None)
        if m is None:
            if isinstance(initial_guess, str):
                m = self._initialize_m(initial_guess)
            else:
                m = _np.linspace(-1., 1., K * D).reshape((K, D))
        self.m = _np.array(m, dtype=float)
        for name in ('m0', 'm'):
            if getattr(self, name).shape != (K, D):
                raise ValueError('Shape of %s %s does not match (K,d)=%s' % (name, getattr(self, name).shape, (K, D)))

        W0 = kwargs.pop('W0', None)
        if W0 is None:
            W0 = _np.eye(D)
        W0 = _np.array(W0, dtype=float)
        if W0.shape == (D, D):
            inv, log_det = chol_inv_det(W0)[1:]
            self.W0 = _np.array([W0] * K)
            self.inv_W0 = _np.array([inv] * K)
            self.log_det_W0 = _np.array([log_det] * K)
        elif W0.shape == (K, D, D):
            self.W0 = W0
            self.inv_W0 = _np.empty_like(W0)
            self.log_det_W0 = _np.empty(K)
            for k in range(K):
                self.inv_W0[k], self.log_det_W0[k] = chol_inv_det(W0[k])[1:]
        else:
            raise ValueError('W0 is neither None, nor a %s array, nor a %s array.' % ((D, D), (K, D, D)))
        self.W = _np.array(kwargs.pop('W', self.W0.copy()), dtype=float)
        if self.W.shape != (K, D, D):
            raise ValueError('Shape of W %s does not match (K, d, d)=%s' % (self.W.shape, (K, D, D)))
        self.log_det_W = _np.array([chol_inv_det(W)[2] for W in self.W])
        if kwargs:
            raise TypeError('unexpected keyword(s): ' + str(kwargs.keys()))

    def _initialize_m(self, initial_guess):
        if self.K > self.N:
            raise ValueError("Can't auto-initialize ``m`` with more output components than samples."
                             " Specify ``m`` explicitly.")
        if initial_guess == 'first':
            rows = self.data[:self.K]
        elif initial_guess == 'random':
            rows = self.data[_np.random.choice(self.data.shape[0], size=self.K, replace=False)]
        else:
            raise ValueError('Invalid ``initial_guess``: ' + str(initial_guess))
        rows = rows.cpu().numpy() if _dev.is_device_tensor(rows) else _np.array(rows, dtype=float)
        if _parallel.enabled():           # every rank starts from rank 0's points
            import torch.distributed as dist
            t = _dev.torch()
            buf = t.from_numpy(_np.ascontiguousarray(rows)).to("cuda:%d" % _lib.default_device())
            dist.broadcast(buf, src=0)
            rows = buf.cpu().numpy()
        return rows.copy()

    def _parse_initial_guess(self, initial_guess):
        """Hyper-parameters whose posterior mode is the given mixture (variational.pyx:651-673)."""
        means, covs, component_weights = recover_gaussian_mixture(initial_guess)
        N, K, D = self.N, self.K, self.dim
        self.alpha = component_weights * (self.alpha0.sum() + N - K) + 1
        self.beta = self.beta0 + N * component_weights
        self.nu = self.nu0 + N * component_weights
        assert (self.alpha > 0.0).all()
        assert (self.beta > 0.0).all()
        assert (self.nu > D - 1).all()
        self.m = means
        self.W = _np.empty_like(covs)
        for k in range(K):
            self.W[k], log_det = chol_inv_det(covs[k] * (self.nu[k] - D))[1:]
            self.log_det_W[k] = -log_det

    # ------------------------------------------------------------------------------------------ E / M step
    def _vb_records(self):
        D, K = self.dim, self.K
        scalars = _np.zeros((K, _lib.NUM_SCALARS))
        scalars[:, 0] = self.expectation_ln_pi
        scalars[:, 1] = self.expectation_det_ln_lambda
        scalars[:, 2] = D * _np.log(2. * _np.pi)
        scalars[:, 3] = D / self.beta
        scalars[:, 4] = self.nu
        scalars[:, _lib.S_WEIGHT] = 1.0
        # T_k with T_k^T T_k = W_k: left behind by the M-step that produced W (the inverse Cholesky factor of W^-1),
        # recomputed only if W was set from outside since
        cached = getattr(self, "_W_factor", None)
        if cached is not None and cached[0].shape == self.W.shape and _np.array_equal(cached[0], self.W):
            t = cached[1]
        else:
            t = _np.array([tri_from_precision(self.W[k]) for k in range(K)])
            self._W_factor = (self.W.copy(), t)
        return _dev.PackedComponents(_lib.pack_records(t, self.m, scalars), list(range(K)))

    def E_step(self):
        """Expectation values and summary statistics (variational.pyx:116-127)."""
        t = _dev.torch()
        K, D, ds = self.K, self.dim, self._ds
        if D > _lib.MAX_DIM:
            raise NotImplementedError("dimension %d exceeds the CUDA kernels' maximum of %d" % (D, _lib.MAX_DIM))
        # K digammas each (variational.pyx:759-772, 800-804); an invalid W raises here, before the sample loop
        self.expectation_det_ln_lambda = _digamma(0.5 * (self.nu[:, None] + 1. - _np.arange(1, D + 1)[None, :])).sum(axis=1) \
            + D * _np.log(2.) + self.log_det_W
        self.expectation_ln_pi = _digamma(self.alpha) - _digamma(self.alpha.sum())
        packed = self._vb_records()

        n = ds.N
        if self._r_dev is None or tuple(self._r_dev.shape) != (n, K):
            self._r_dev = t.empty((n, K), dtype=t.float64, device=ds.x.device)
        lay = PacketLayout(K, D)
        packet = t.zeros(lay.size, dtype=t.float64, device=ds.x.device)
        # r and sum_n w_n sum_k r log r leave the launch; the normalised log rho (variational.pyx:741,755) is needed by no
        # step of the algorithm once that sum is fused, so it is materialised only when the attribute is read
        # (5 GB and 4 % of the E-step at N = 1e7, K = 64)
        run_k1(ds.x, packed, K, _lib.MODE_VB, resp=self._r_dev, weights=ds.w,
               sums=packet[lay.off_sum_a:lay.off_sum_a + 2])
        # shift vector(s) of the raw moments: one (alpha-weighted centre) unless components lie > 100 sigma apart;
        # component k is N(m_k, (nu_k W_k)^-1) in expectation
        if small_problem(n, K):
            # small problem: the reference's two passes (means first, second moments about them)
            shift = two_pass_suffstats(_lib.Context.get(), ds, lay, packet, self._r_dev, None, range(K))
        else:
            groups = shift_groups(self.m, self.nu[:, None, None] * self.W, self.alpha, range(K))
            shift = grouped_suffstats(_lib.Context.get(), ds, lay, packet, groups, self._r_dev, None, _dev.current_stream_ptr())
        _parallel.allreduce_(packet)
        st = lay.unpack(packet.cpu().numpy())
        self._estep_packed = packed
        self._host_cache = {}

        self.N_comp = st["A"]
        if not _np.isfinite(self.N_comp).any():
            raise _np.linalg.LinAlgError('Encountered inf or nan in update of responsibilities\n' + str(self.N_comp))
        self.N_comp, self.x_mean_comp, self.S = moments_from_stats(st, shift, "B")   # N_comp regularised in place like
        self.inv_N_comp = 1. / self.N_comp                             # variational.pyx:699-709 (exact zeros -> tiny)
        if not _np.isfinite(self.S).any():
            raise _np.linalg.LinAlgError('Encountered inf or nan in update of sample covariance\n' + str(self.S))
        self._expectation_log_q_Z = st["sum_a"]                        # variational.pyx:1003-1013

    def M_step(self):
        """Update the Gauss-Wishart / Dirichlet hyper-parameters (variational.pyx:129-136, 693-697, 934-946)."""
        self.nu = self.nu0 + self.N_comp
        self.alpha = self.alpha0 + self.N_comp
        self.beta = self.beta0 + self.N_comp
        self.m = (self.beta0[:, None] * self.m0 + self.N_comp[:, None] * self.x_mean_comp) / self.beta[:, None]
        diff = self.x_mean_comp - self.m0
        cov = (self.beta0 / (self.beta0 + self.N_comp))[:, None, None] * (diff[:, :, None] * diff[:, None, :])
        cov += self.S
        cov *= self.N_comp[:, None, None]
        cov += self.inv_W0
        try:
            # K inversions in a few batched LAPACK calls; W_k = cov_k^-1 = T_k^T T_k with T_k the inverse Cholesky factor
            _, inv, log_det, t = chol_inv_det_batch(cov)
            self.W = inv
            self.log_det_W = -log_det
            self._W_factor = (self.W.copy(), t)
        except (_np.linalg.LinAlgError, ValueError):
            for k in range(self.K):                                    # per matrix: the reference's error for the offender
                self.W[k], log_det = chol_inv_det(cov[k])[1:]
                self.log_det_W[k] = -log_det

    def update(self):
        """One M-step followed by one E-step."""
        self.M_step()
        self.E_step()

    # ------------------------------------------------------------------------------------------ N x K attributes
    def _download(self, name, tensor):
        if name not in self._host_cache:
            self._host_cache[name] = tensor.cpu().numpy()
        return self._host_cache[name]

    @property
    def r(self):
        """(N x K) responsibilities of the last E-step (downloaded on first access)."""
        return self._download("r", self._r_dev)

    @property
    def log_rho(self):
        """(N x K) normalised log responsibilities of the last E-step (variational.pyx:741,755); recomputed by one
        extra launch of K1 when read (same arithmetic as the E-step's launch), because no step of the algorithm needs
        it materialised."""
        if "log_rho" not in self._host_cache:
            t = _dev.torch()
            ds = self._ds
            lr = t.empty((ds.N, self.K), dtype=t.float64, device=ds.x.device)
            run_k1(ds.x, self._estep_packed, self.K, _lib.MODE_VB, lp=lr)
            self._host_cache["log_rho"] = lr.cpu().numpy()
        return self._host_cache["log_rho"]

    @property
    def expectation_gauss_exponent(self):
        """(N x K) E[(x-mu)^T Lambda (x-mu)] of the last E-step (variational.pyx:774-798); recomputed by one
        extra launch of K1 when read, because no step of the algorithm needs it materialised."""
        if "E" not in self._host_cache:
            t = _dev.torch()
            ds = self._ds
            e = t.empty((ds.N, self.K), dtype=t.float64, device=ds.x.device)
            run_k1(ds.x, self._estep_packed, self.K, _lib.MODE_VB, aux=e)
            self._host_cache["E"] = e.cpu().numpy()
        return self._host_cache["E"]

    # ------------------------------------------------------------------------------------------ results
    def make_mixture(self):
        """Mixture at the mode of the variational posterior (variational.pyx:138-191): weights alpha_k - 1,
        means m_k, covariances ((nu_k - D) W_k)^-1; components without a mode are skipped."""
        components, weights, skipped = [], [], []
        for k in range(self.K):
            pi = self.alpha[k] - 1.
            if pi <= 0:
                logger.warning("Skipped component %i because of zero weight" % k)
                skipped.append(k)
                continue
            if self.nu[k] <= self.dim:
                logger.warning("Gauss-Wishart mode of component %i is not defined" % k)
                skipped.append(k)
                continue
            try:
                cov = chol_inv_det((self.nu[k] - self.dim) * self.W[k])[1]
                components.append(Gauss(self.m[k], cov))
            except Exception as error:
                logger.error("Could not create component %i. The error was: %s" % (k, repr(error)))
                skipped.append(k)
                continue
            weights.append(pi)
        if skipped:
            logger.warning("The following components have been skipped: %s" % skipped)
        return MixtureDensity(components, weights)

    def likelihood_bound(self):
        """Lower bound L(Q) on the log marginal likelihood (PRML 10.70-10.77; variational.pyx:194-209, 948-1033)."""
        K, D = self.K, self.dim
        ln2pi = _np.log(2 * _np.pi)
        e_det, e_pi = self.expectation_det_ln_lambda, self.expectation_ln_pi
        log_p_X = log_p_mu_lambda = 0.
        log_q_mu_lambda = -0.5 * K * D
        for k in range(K):
            W = self.W[k]
            dx = self.x_mean_comp[k] - self.m[k]
            log_p_X += self.N_comp[k] * (e_det[k] - D / self.beta[k]
                                         - self.nu[k] * (_np.trace(self.S[k].dot(W)) + dx.dot(W).dot(dx)) - D * ln2pi)
            dm = self.m[k] - self.m0[k]
            log_p_mu_lambda += D * _np.log(self.beta0[k] / (2. * _np.pi)) + e_det[k] - D * self.beta0[k] / self.beta[k] \
                - self.beta0[k] * self.nu[k] * dm.dot(W).dot(dm) \
                + 2 * Wishart_log_B(D, self.nu0[k], self.log_det_W0[k]) + (self.nu0[k] - D - 1) * e_det[k] \
                - self.nu[k] * _np.trace(self.inv_W0[k].dot(W))
            log_q_mu_lambda += 0.5 * (e_det[k] + D * _np.log(self.beta[k] / (2 * _np.pi))) \
                - Wishart_H(D, self.nu[k], self.log_det_W[k])
        self._expectation_log_p_X = 0.5 * log_p_X
        self._expectation_log_p_Z = float(self.N_comp.dot(e_pi))
        self._expectation_log_p_pi = Dirichlet_log_C(self.alpha0) + float((self.alpha0 - 1).dot(e_pi))
        self._expectation_log_p_mu_lambda = 0.5 * log_p_mu_lambda
        self._expectation_log_q_pi = float((self.alpha - 1).dot(e_pi)) + Dirichlet_log_C(self.alpha)
        self._expectation_log_q_mu_lambda = log_q_mu_lambda
        return self._expectation_log_p_X + self._expectation_log_p_Z + self._expectation_log_p_pi \
            + self._expectation_log_p_mu_lambda - self._expectation_log_q_Z - self._expectation_log_q_pi \
            - self._expectation_log_q_mu_lambda

    def posterior2prior(self):
        """Posterior hyper-parameters as keyword arguments for a new instance that uses them as prior."""
        return dict(alpha0=self.alpha.copy(), beta0=self.beta.copy(), nu0=self.nu.copy(),
                    m0=self.m.copy(), W0=self.W.copy(), components=self.K)

    def prior_posterior(self):
        """Copies of all prior and posterior hyper-parameters."""
        return dict(alpha0=self.alpha0.copy(), beta0=self.beta0.copy(), m0=self.m0.copy(),
                    nu0=self.nu0.copy(), W0=self.W0.copy(), alpha=self.alpha.copy(), beta=self.beta.copy(),
                    m=self.m.copy(), nu=self.nu.copy(), W=self.W.copy(), components=self.K)

    def prune(self, threshold=1.):
        """Delete components whose effective number of samples N_k is below ``threshold``, then redo the
        E-step (variational.pyx:233-281)."""
        if not threshold:
            return
        keep = _np.where(self.N_comp >= threshold)[0]
        if len(keep) == 0:
            raise ValueError("Prune threshold %g too large, would remove all components" % threshold)
        self.K = len(keep)
        for name in ('alpha0', 'alpha', 'beta0', 'beta', 'expectation_det_ln_lambda', 'expectation_ln_pi', 'N_comp',
                     'nu0', 'nu', 'm0', 'm', 'S', 'W0', 'inv_W0', 'W', 'log_det_W', 'log_det_W0', 'x_mean_comp'):
            setattr(self, name, getattr(self, name)[keep])
        self.E_step()       # rebuilds r, log_rho and the statistics for the survivors

    def run(self, iterations=1000, prune=1., rel_tol=1e-10, abs_tol=1e-5, verbose=False):
        """Iterate ``update`` until the bound converges (variational.pyx:283-359).  Returns the iteration of
        convergence or None; no convergence while components are being removed or the bound decreases."""
        old_K = None
        bound = None
        for i in range(1, iterations + 1):
            if self.K == old_K:
                old_bound = bound
            else:
                old_bound = self.likelihood_bound()
                logger.info('New bound=%g, K=%d, N_k=%s' % (old_bound, self.K, self.N_comp))
            self.update()
            bound = self.likelihood_bound()
            logger.info('After update %d: bound=%.15g, K=%d, N_k=%s' % (i, bound, self.K, self.N_comp))
            if bound < old_bound:
                logger.warning('Bound decreased from %g to %g' % (old_bound, bound))
            if bound == old_bound:
                return i
            diff = bound - old_bound
            if diff > 0:
                if abs(bound) < abs_tol:
                    if abs(diff) < abs_tol:
                        return i
                elif abs(diff / bound) < rel_tol:
                    return i
            old_K = self.K
            self.prune(prune)
        return None


def Wishart_log_B(D, nu, log_det):
    """ln B(W, nu), first factor of the Wishart normalisation, PRML (B.79), from ln det W."""
    assert D > 0, 'Invalid dimension: %s' % D
    assert nu > D - 1, 'Invalid degree of freedom: %s' % nu
    assert _np.isfinite(log_det), 'Non-finite log(det): %s' % log_det
    return -0.5 * nu * log_det - 0.5 * nu * D * _np.log(2) - 0.25 * D * (D - 1) * _np.log(_np.pi) \
        - _gammaln(0.5 * (nu + 1 - _np.arange(1, D + 1))).sum()


def Wishart_expect_log_lambda(D, nu, log_det):
    """E[ln det Lambda] under a Wishart, PRML (B.81)."""
    assert D > 0, 'Invalid dimension: %s' % D
    assert nu > D - 1, 'Invalid degree of freedom: %s' % nu
    assert _np.isfinite(log_det), 'Non-finite log(det): %s' % log_det
    return _digamma(0.5 * (nu + 1 - _np.arange(1, D + 1))).sum() + D * _np.log(2.) + log_det


def Wishart_H(D, nu, log_det):
    """Entropy of the Wishart distribution, PRML (B.82)."""
    return -Wishart_log_B(D, nu, log_det) - 0.5 * (nu - D - 1) * Wishart_expect_log_lambda(D, nu, log_det) + 0.5 * nu * D


def Dirichlet_log_C(alpha):
    """ln C(alpha), normalisation of the Dirichlet distribution, PRML (B.23)."""
    return float(_gammaln(_np.sum(alpha)) - _gammaln(alpha).sum())
