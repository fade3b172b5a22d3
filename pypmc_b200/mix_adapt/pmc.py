"""Population Monte Carlo proposal updates with the API of pypmc/mix_adapt/pmc.pyx:
``gaussian_pmc`` (:120-246), ``student_t_pmc`` (:499-739) and the ``PMC`` driver (:248-476).

Per update: ONE launch of kernel K1 (log-pdfs, log-sum-exp, rho_nk [, gamma_nk], sum_n w_n log q_n) and ONE
launch of kernel K2 (A_k, B_k, first and second moments, dof statistic), one all-reduce of the K-row packet
when samples are sharded over GPUs (``pypmc_b200.parallel``), then K-sized host arithmetic and the same
per-component ``update`` / ``LinAlgError`` handling as the reference.  Samples, rho and gamma stay on the
device; ``PMC`` uploads the samples once for all its EM steps (pmc.pyx:362,447 keeps them fixed).
"""
from __future__ import division

import logging
from copy import deepcopy as _cp

import numpy as _np
from scipy.optimize import brentq as _find_root
from scipy.special import digamma as _psi

from ..density.gauss import Gauss, batch_update as _batch_update
from ..density.mixture import MixtureDensity
from ..density.student_t import StudentT
from ..density._eval import run_k1
from .. import _device as _dev
from .. import _lib
from .. import parallel as _parallel
from ._stats import PacketLayout, moments_from_stats, shift_groups, grouped_suffstats, small_problem, two_pass_suffstats

logger = logging.getLogger(__name__)


class DeviceSamples(object):
    """This rank's (samples, importance weights, latent indices), resident on the GPU."""

    def __init__(self, samples, weights=None, latent=None):
        x = _dev.as_samples(samples)
        self.N, self.D = int(x.shape[0]), int(x.shape[1])
        self.x = _dev.to_device(x).contiguous()
        self.w = None if weights is None else _dev.to_device(_np.asarray(weights, dtype=_np.float64)
                                                              if not _dev.is_device_tensor(weights) else weights)
        if latent is None:
            self.latent = None
        else:
            t = _dev.torch()
            self.latent = latent.to(t.int64) if _dev.is_device_tensor(latent) else _dev.to_device(_np.asarray(latent, dtype=_np.int64))
        self._src_weights, self._src_latent = weights, latent     # what the caller handed in (identity check in _pmc_update)
        self.rho = None      # [N, K] responsibilities of the last E-pass (device)
        self.gamma = None    # [N, K] Student-t gamma of the last E-pass (device)
        self.epass = None    # (K, live, mode, record fingerprint, local sums[2], mixture weights) of an E-pass computed ahead
        self.weight_sums = None   # device [5]: sum w, sum w log q, sum w^2, sum w log w, #nonzero (kernel K4, set by weigh())

    def weigh(self, proposal, log_target):
        """Importance weights w_n = exp(log_target_n - log q(x_n)) of these samples under ``proposal`` (extension).

        An importance-sampling step evaluates the proposal at every sample for the weights
        (importance_sampling.py:203-207) and the PMC update that follows evaluates it again, at the same samples,
        for the responsibilities (pmc.pyx:23-43).  Here ONE launch of K1 yields log q and rho; the weights are formed
        from it, stored as this object's weights, and the next ``gaussian_pmc`` / ``student_t_pmc`` call on
        ``(self, proposal)`` re-uses the responsibilities instead of launching K1 again.

        :param proposal: the :class:`MixtureDensity` the samples were drawn from (unchanged until the update).
        :param log_target: N log-values of the target at the samples (CUDA tensor or ndarray).
        :return: the weights as a CUDA tensor (also kept in ``self.w``).
        """
        t = _dev.torch()
        mode = proposal._require_mode()
        K, N = len(proposal), self.N
        live = _live_components(proposal)
        student = (mode == _lib.MODE_STUDENT_T)
        _alloc_e_buffers(self, N, K, len(live), student)
        packed = proposal._packed(live)
        logq = t.empty(N, dtype=t.float64, device=self.x.device)
        run_k1(self.x, packed, K, mode, logq=logq, resp=self.rho, aux=self.gamma if student else None)
        index = self.x.device.index
        lt = log_target if _dev.is_device_tensor(log_target) else _dev.to_device(_np.asarray(log_target, dtype=_np.float64), index)
        # K4: the weights and, in the same pass, sum w, sum w log q, sum w^2, sum w log w (perp / ess / likelihood)
        self.w = t.empty(N, dtype=t.float64, device=self.x.device)
        self.weight_sums = t.empty(5, dtype=t.float64, device=self.x.device)
        _lib.Context.get(index).importance_weights(lt.contiguous(), logq, N, self.w, self.weight_sums, _dev.current_stream_ptr(index))
        self._src_weights = self.w
        sums = t.stack([self.weight_sums[1], self.weight_sums[0]])  # what K1 would have left in the packet
        self.epass = (K, tuple(live), mode, _fingerprint(packed), sums, _np.array(proposal.weights, dtype=float))
        return self.w


def _weight_quality(ds):
    """(perp, ess) of the weights formed by :meth:`DeviceSamples.weigh`, from the sums kernel K4 left -- over all
    ranks' samples when the statistics all-reduce is enabled (tools/convergence.py:6-72)."""
    if ds.weight_sums is None:
        raise ValueError("no weights were formed on the device yet: call DeviceSamples.weigh first")
    s = ds.weight_sums.clone()
    n = _dev.torch().tensor([float(ds.N)], dtype=s.dtype, device=s.device)
    _parallel.allreduce_(s)
    _parallel.allreduce_(n)
    s, n = s.cpu().numpy(), float(n[0])
    perp = float(_np.exp(_np.log(s[0]) - s[3] / s[0]) / n)
    ess = float(s[0] * s[0] / (n * s[2]))
    return perp, ess


DeviceSamples.perp = lambda self: _weight_quality(self)[0]
DeviceSamples.ess = lambda self: _weight_quality(self)[1]
DeviceSamples.perp.__doc__ = "Normalised perplexity of the weights of the last ``weigh`` (convergence.py:6-39); no pass over the samples."
DeviceSamples.ess.__doc__ = "Normalised effective sample size of the weights of the last ``weigh`` (convergence.py:42-72); no pass over the samples."


def _check_arguments(samples, weights, latent, mincount, rb):
    """Argument contradictions, same messages as pmc.pyx:70-83."""
    if weights is not None and not isinstance(samples, DeviceSamples):
        shape = tuple(weights.shape) if hasattr(weights, "shape") else _np.asarray(weights).shape
        assert len(shape) == 1, 'Weights must be one-dimensional.'
        assert shape[0] == len(samples), \
            "Number of weights (%s) does not match the number of samples (%s)." % (shape[0], len(samples))
    if latent is None:
        if mincount > 0:
            raise ValueError('`mincount` must be 0 if `latent` is not provided!')
        if not rb:
            raise ValueError('`rb` must be True if `latent` is not provided!')


def _e_pass_and_stats(ds, density, live, rb, mode):
    """K1 + K2 on this rank's samples, all-reduce, return (layout, unpacked statistics, shift)."""
    t = _dev.torch()
    K, D, N = len(density), density.dim, ds.N
    assert ds.D == D, "The points in ``x`` have the wrong dimension (%i instead of %i)" % (ds.D, D)
    lay = PacketLayout(K, D)
    device = ds.x.device
    packet = t.zeros(lay.size, dtype=t.float64, device=device)
    student = (mode == _lib.MODE_STUDENT_T)
    sums = packet[lay.off_sum_a:lay.off_sum_a + 2]
    packed = density._packed(live) if live else None

    ahead = ds.epass
    ds.epass = None
    # a pass computed ahead serves this update only if it was made with the same components AND the same mixture weights
    # (up to the 1 +- 1e-16 rescaling of normalize(), which is why the weights are not part of the fingerprint)
    reuse = bool(live) and rb and ahead is not None and ahead[:4] == (K, tuple(live), mode, _fingerprint(packed)) \
        and _np.allclose(ahead[5], density.weights, rtol=1e-13, atol=0.0)
    if not reuse:
        _alloc_e_buffers(ds, N, K, len(live), student)
    gamma = ds.gamma if student else None

    if reuse:
        sums.copy_(ahead[4])                                 # rho / gamma / sums were computed by PMC.run's fused bound
    elif live:
        if rb:
            run_k1(ds.x, packed, K, mode, resp=ds.rho, aux=gamma, weights=ds.w, sums=sums)
        else:
            # latent variables known: one-hot responsibilities (pmc.pyx:45-51); K1 only for gamma / sums
            ds.rho.zero_()
            live_mask = t.zeros(K + 1, dtype=t.bool, device=device)
            live_mask[t.tensor(live, device=device)] = True
            lat = ds.latent.clamp(min=-1, max=K)
            ok = (lat >= 0) & (lat < K) & live_mask[lat.clamp(min=0)]
            rows = t.nonzero(ok, as_tuple=True)[0]
            ds.rho[rows, lat[rows]] = 1.0
            if student:
                run_k1(ds.x, packed, K, mode, aux=gamma, weights=ds.w, sums=sums)
            else:
                sums[1] = float(N) if ds.w is None else ds.w.sum()
    if ds.latent is not None:
        lat = ds.latent
        inside = (lat >= 0) & (lat < K)
        packet[lay.off_counts:lay.off_counts + K] = t.bincount(lat[inside], minlength=K)[:K].to(t.float64)

    if live and small_problem(N, K):
        # small problem: the reference's two passes (means first, second moments about them)
        shift = two_pass_suffstats(_lib.Context.get(), ds, lay, packet, ds.rho, gamma, live)
        return lay, lay.unpack(packet.cpu().numpy()), shift
    # shift vector(s) of the raw moments: one for the whole mixture unless components lie > 100 sigma apart
    groups = shift_groups([c.mu for c in density.components], [c.inv_sigma for c in density.components],
                          density.weights, live) if live else []
    if len(groups) > 1:
        logger.info("moment kernel: %d shift groups (components far apart in units of their width)" % len(groups))
    if len(groups) <= 1:
        # ONE launch over all K columns (dead components hold rho = 0 and come out as zero rows)
        centre = groups[0][1] if groups else _np.zeros(D)
        groups = [(list(range(K)), centre)]
    shift = grouped_suffstats(_lib.Context.get(), ds, lay, packet, groups, ds.rho, gamma, _dev.current_stream_ptr())
    _parallel.allreduce_(packet)
    return lay, lay.unpack(packet.cpu().numpy()), shift


def _fingerprint(packed):
    """Hash of the evaluated component records without their mixture-weight slot (``normalize()`` may rescale the
    weights by 1 +- 1e-16 between the pass computed ahead and the update that consumes it)."""
    rec = packed.records.copy()
    rec[:, rec.shape[1] - _lib.NUM_SCALARS + _lib.S_WEIGHT] = 0.0
    return hash(rec.tobytes())


def _alloc_e_buffers(ds, N, K, n_live, student):
    t = _dev.torch()
    alloc = t.empty if n_live == K else t.zeros           # dead columns must read 0 (pmc.pyx:26)
    if ds.rho is None or tuple(ds.rho.shape) != (N, K) or n_live != K:
        ds.rho = alloc((N, K), dtype=t.float64, device=ds.x.device)
    if student and (ds.gamma is None or tuple(ds.gamma.shape) != (N, K) or n_live != K):
        ds.gamma = alloc((N, K), dtype=t.float64, device=ds.x.device)


def e_pass_ahead(ds, density, mode):
    """The Rao-Blackwellised E-pass of the NEXT update, run now: one K1 launch that leaves rho (and gamma) in
    ``ds`` and returns the log-likelihood sum_n wbar_n log q(x_n) of ``density`` over all ranks -- the same launch
    serves ``PMC.log_likelihood`` and ``calculate_rho_rb`` (pmc.pyx:388-391 and :23-43 evaluate the same mixture on
    the same samples twice per EM step)."""
    t = _dev.torch()
    K, N = len(density), ds.N
    live = _live_components(density)
    student = (mode == _lib.MODE_STUDENT_T)
    _alloc_e_buffers(ds, N, K, len(live), student)
    sums = t.zeros(2, dtype=t.float64, device=ds.x.device)
    packed = density._packed(live)
    run_k1(ds.x, packed, K, mode, resp=ds.rho, aux=ds.gamma if student else None, weights=ds.w, sums=sums)
    ds.epass = (K, tuple(live), mode, _fingerprint(packed), sums.clone(), _np.array(density.weights, dtype=float))
    _parallel.allreduce_(sums)
    s = sums.cpu().numpy()
    return float(s[0] / s[1])


def _live_components(density):
    return [k for k in range(len(density)) if density.weights[k] != 0]


def _kill_undersampled(density, live, counts, mincount):
    """Components that proposed fewer than ``mincount`` samples die AFTER rho was computed (pmc.pyx:109-116).
    The reference removes from the list it is iterating, so the element following a removed one is not
    examined; the index arithmetic below reproduces exactly that traversal."""
    died = False
    i = 0
    while i < len(live):
        k = live[i]
        if counts[k] < mincount:
            live.pop(i)
            density.weights[k] = 0.
            died = True
            logger.warning("Component %i died because of too few (%i) samples." % (k, counts[k]))
        i += 1
    return died


def _apply_update(density, live, alpha, mean, cov, new_dof=None):
    """Install the new parameters; a component whose covariance is not positive definite keeps its old
    parameters and gets weight zero (pmc.pyx:227-244, :713-737).  Returns True if that happened."""
    live = list(live)
    try:
        # all K factorisations in a few batched LAPACK calls; an unusable covariance anywhere raises before anything is
        # modified and the per-component loop below -- the reference's, with its per-component verdicts -- takes over
        _batch_update([density.components[k] for k in live], [mean[k] for k in live], [cov[k] for k in live],
                      None if new_dof is None else [new_dof[k] for k in live])
        for k in live:
            density.weights[k] = alpha[k]
        return False
    except (_np.linalg.LinAlgError, ValueError, AssertionError):
        pass
    failed = False
    for k in live:
        comp = density.components[k]
        density.weights[k] = alpha[k]
        old = (comp.mu, comp.sigma) if new_dof is None else (comp.mu, comp.sigma, comp.dof)
        try:
            if new_dof is None:
                comp.update(mean[k], cov[k])
            else:
                comp.update(mean[k], cov[k], new_dof[k])
        except _np.linalg.LinAlgError:
            logger.warning("Could not update component %i --> weight is set to zero." % k)
            comp.update(*old)
            density.weights[k] = 0.
            failed = True
    return failed


def _as_device_samples(samples, weights, latent):
    """``samples`` as :class:`DeviceSamples`.  A DeviceSamples object carries its own weights and latent indices; explicit
    ``weights`` / ``latent`` next to it must be the very objects it was built from (PMC passes them along), anything
    else is a contradiction the caller has to resolve."""
    if isinstance(samples, DeviceSamples):
        for name, given, own in (("weights", weights, samples._src_weights), ("latent", latent, samples._src_latent)):
            if given is not None and given is not own:
                raise ValueError("`%s` was passed next to a DeviceSamples object that carries its own %s; "
                                 "put them into the DeviceSamples instead" % (name, name))
        return samples
    return DeviceSamples(samples, weights, latent)


def _pmc_update(samples, density, weights, latent, rb, mincount, copy, mode, dof_args=None):
    ds = _as_device_samples(samples, weights, latent)
    _check_arguments(samples, weights, ds.latent, mincount, rb)       # the latent indices that will actually be used
    if copy:
        if isinstance(density, MixtureDensity) and density._kernel_mode() == mode and density.dim <= _lib.MAX_DIM:
            # the packed CUDA records are cached per component and shared by copies: form them on the caller's object,
            # so that a mixture handed in again (or its next copy) does not pay the packing again (1 ms at K = 32)
            for k in _live_components(density):
                density.components[k]._packed_record()
        density = _cp(density)
    if not isinstance(density, MixtureDensity) or density._require_mode() != mode:
        raise TypeError("``density`` must be a MixtureDensity with %s components"
                        % ("StudentT" if mode == _lib.MODE_STUDENT_T else "Gauss"))
    live = _live_components(density)
    old_dofs = [c.dof for c in density.components] if mode == _lib.MODE_STUDENT_T else None

    lay, st, shift = _e_pass_and_stats(ds, density, live, rb, mode)
    need_renormalize = False
    if ds.latent is not None:
        need_renormalize = _kill_undersampled(density, live, st["counts"], mincount)

    weight_normalization = st["sumw"]
    A_reg, mean, cov = moments_from_stats(st, shift, "B")
    alpha = A_reg / weight_normalization                       # pmc.pyx:191-193

    new_dof = None
    if mode == _lib.MODE_STUDENT_T:
        new_dof = _solve_dofs(density, live, st, old_dofs, weight_normalization, **dof_args)

    if _apply_update(density, live, alpha, mean, cov, new_dof):
        need_renormalize = True
    if need_renormalize:
        density.normalize()
    density._last_sums = (st["sum_a"], st["sumw"])             # sum_n w_n log q_old(x_n), sum_n w_n
    return density


def gaussian_pmc(samples, density, weights=None, latent=None, rb=True, mincount=0, copy=True):
    """Adapt a Gaussian mixture with the (M-)PMC update [Cap+08, Kil+09]; API of pmc.pyx:120-246.

    :param samples: (N x D) float64 numpy array or torch CUDA tensor -- this rank's samples.
    :param density: :class:`MixtureDensity` of :class:`Gauss` components that proposed them.
    :param weights: N importance weights (unnormalised) or None for equal weights.
    :param latent: N component indices that generated the samples, optional.
    :param rb: Rao-Blackwellised responsibilities (True) or one-hot from ``latent`` (False).
    :param mincount: components with fewer proposed samples are switched off (needs ``latent``).
    :param copy: leave ``density`` untouched and return an updated copy (default) or update in place.
    """
    if samples is None:
        raise TypeError("Argument 'samples' must not be None")
    return _pmc_update(samples, density, weights, latent, rb, mincount, copy, _lib.MODE_GAUSS)


class _DOFCondition(object):
    """First-order condition for a Student-t component's degree of freedom, eq. (16) of [HOD12]
    (pmc.pyx:478-497): root of const + ln(nu/2) - psi(nu/2)."""

    def __init__(self, const):
        self.const = float(const)

    def __call__(self, nu):
        return self.const + _np.log(.5 * nu) - _psi(.5 * nu)


def _solve_dofs(density, live, st, old_dofs, weight_normalization, dof_solver_steps, mindof, maxdof):
    K, D = len(density), density.dim
    if not dof_solver_steps:
        return list(old_dofs)
    new_dof = [-1 for _ in range(K)]
    W = st["sumw"]
    for k in live:
        nu = old_dofs[k]
        A, B, L = st["A"][k], st["B"][k], st["L"][k]
        # sum_n w (xi + delta), pmc.pyx:659-679, with ln((q+nu)/2) = ln((nu+D)/2) - ln(gamma)
        total = (A * _np.log(.5 * (nu + D)) - L) - _psi(.5 * (D + nu)) * A \
            + (W - A) * (_np.log(.5 * nu) - _psi(.5 * nu)) + B + (W - A)
        condition = _DOFCondition(1. - total / weight_normalization)
        try:
            new_dof[k] = _find_root(condition, mindof, maxdof, maxiter=dof_solver_steps)
        except RuntimeError:   # not converged
            logger.warning("``dof`` solver for component %i did not converge." % k)
            new_dof[k] = old_dofs[k]
        except ValueError as error:
            # same sign at both ends; the condition is decreasing in nu (pmc.pyx:700-710)
            if condition(mindof) < 0.:
                new_dof[k] = mindof
            elif condition(maxdof) > 0.:
                new_dof[k] = maxdof
            else:
                raise RuntimeError('``dof`` adaptation for component %i raised an error.' % k, error)
    return new_dof


def student_t_pmc(samples, density, weights=None, latent=None, rb=True, dof_solver_steps=100, mindof=1e-5,
                  maxdof=1e3, mincount=0, copy=True):
    """Adapt a Student-t mixture with the PMC update of [Cap+08] and the dof update of [HOD12]; API of
    pmc.pyx:499-739.  Parameters as :func:`gaussian_pmc`, plus ``dof_solver_steps`` (0 = keep the degrees of
    freedom), ``mindof`` and ``maxdof`` bracketing the root search."""
    if samples is None:
        raise TypeError("Argument 'samples' must not be None")
    return _pmc_update(samples, density, weights, latent, rb, mincount, copy, _lib.MODE_STUDENT_T,
                       dict(dof_solver_steps=dof_solver_steps, mindof=mindof, maxdof=maxdof))


class PMC(object):
    """Run several PMC updates on a fixed set of samples; API of pmc.pyx:248-476.

    ``samples``, ``weights`` and ``latent`` are uploaded to the GPU once (they are not re-read afterwards);
    ``density`` is always copied.  Additional keyword arguments go to the update function.
    """

    def __init__(self, samples, density, weights=None, latent=None, rb=True, mincount=0, **kwargs):
        if samples is None:
            raise TypeError("Argument 'samples' must not be None")
        if weights is not None:
            self.weights = weights if _dev.is_device_tensor(weights) else _np.asarray(weights)
            assert len(self.weights.shape) == 1, 'Weights must be one-dimensional.'
            assert len(self.weights) == len(samples), \
                "Number of weights (%s) does not match the number of samples (%s)." % (len(self.weights), len(samples))
        else:
            self.weights = None
        if latent is None:
            if mincount > 0:
                raise ValueError('`mincount` must be 0 if `latent` is not provided!')
            if not rb:
                raise ValueError('`rb` must be True if `latent` is not provided!')

        error_wrong_mixture = '``density`` must be a ``pypmc.density.mixture.MixtureDensity`` with ' \
            '``pypmc.density.gauss.Gauss`` or ``pypmc.density.student_t.StudentT`` components'
        if not isinstance(density, MixtureDensity):
            raise TypeError(error_wrong_mixture)
        first = type(density.components[0])
        if issubclass(first, Gauss):
            self.pmc, kind = gaussian_pmc, Gauss
        elif issubclass(first, StudentT):
            self.pmc, kind = student_t_pmc, StudentT
        else:
            raise TypeError(error_wrong_mixture)
        if not all(isinstance(c, kind) for c in density.components):
            raise TypeError(error_wrong_mixture)

        self.density = _cp(density)
        self.samples = samples
        self.latent = latent
        self.rb = rb
        self.mincount = mincount
        self.additional_args = kwargs
        self._device_samples = DeviceSamples(samples, weights, latent)

    def log_likelihood(self):
        """sum_n wbar_n log q(x_n), eq. (5) of [Cap+08] (pmc.pyx:371-391), over all ranks' samples."""
        ds = self._device_samples
        t = _dev.torch()
        sums = t.zeros(2, dtype=t.float64, device=ds.x.device)
        run_k1(ds.x, self.density._packed(), len(self.density), self.density._require_mode(), weights=ds.w, sums=sums)
        _parallel.allreduce_(sums)
        s = sums.cpu().numpy()
        return float(s[0] / s[1])

    def run(self, iterations=1000, prune=0., rel_tol=1e-10, abs_tol=1e-5, verbose=False, fuse_likelihood=False):
        """Iterate updates until the log-likelihood converges (pmc.pyx:393-476); returns the number of
        iterations at convergence or None.  Convergence is never declared when the bound decreased.

        ``fuse_likelihood`` (extension, off by default): the reference evaluates the updated mixture once for the
        bound and again, unchanged, for the next update's responsibilities; with this flag one launch of K1 serves
        both, which removes a third of an EM step's device time.  The only difference to the two-launch flow is
        that the responsibilities were computed before ``normalize()`` rescaled the weights by 1 +- 1e-16."""
        old_K = None
        bound = None
        for i in range(1, iterations + 1):
            if old_K == len(self.density):
                old_bound = bound
            else:
                old_bound = self.log_likelihood()
                logger.info('New bound=%g, K=%i' % (old_bound, len(self.density)))

            self.pmc(self._device_samples, self.density, self.weights, self.latent, self.rb, mincount=self.mincount,
                     copy=False, **self.additional_args)
            if fuse_likelihood and self.rb:
                bound = e_pass_ahead(self._device_samples, self.density, self.density._require_mode())
            else:
                bound = self.log_likelihood()
            logger.info('After update %d: bound=%.15g, K=%i, component_weights=%s'
                        % (i, bound, len(self.density), self.density.weights))

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

            old_K = len(self.density)
            if self.density.prune(prune):
                self._device_samples.epass = None      # columns moved: the pass computed ahead no longer applies
            self.density.normalize()
        return None
