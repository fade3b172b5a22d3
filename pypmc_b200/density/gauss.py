"""Gaussian densities with the API of pypmc/density/gauss.pyx (``LocalGauss`` :11-67, ``Gauss`` :69-163)."""
import numpy as _np

from .base import ProbabilityDensity, LocalDensity
from ..tools._linalg import chol_inv_det, chol_inv_det_batch, tri_from_chol, bilinear_sym
from .. import _lib
from .. import _device as _dev


class LocalGauss(LocalDensity):
    """Multivariate local Gaussian with redefinable covariance (gauss.pyx:11-67)."""
    symmetric = True

    def __init__(self, sigma):
        self.update(sigma)

    def update(self, sigma):
        """Install a new covariance.  On ``LinAlgError`` nothing is changed (gauss.pyx:38-47)."""
        sigma = _np.array(sigma, dtype=float, ndmin=2, copy=True)   # scalar -> 1x1
        chol, inv, log_det = chol_inv_det(sigma)                    # may raise: state still untouched
        self.cholesky_sigma, self.inv_sigma, self.log_det_sigma = chol, inv, log_det
        self.sigma = sigma
        self.dim = sigma.shape[0]
        self._compute_norm()

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        for key, val in self.__dict__.items():           # arrays are replaced, never written in place: copy them flat
            new.__dict__[key] = val.copy() if isinstance(val, _np.ndarray) else val
        return new

    def _compute_norm(self):
        # gauss.pyx:54-56
        self.log_normalization = -0.5 * self.dim * _np.log(2.0 * _np.pi) - 0.5 * self.log_det_sigma

    def _get_gauss_sample(self, rng):
        return _np.dot(self.cholesky_sigma, rng.normal(0, 1, self.dim))

    def evaluate(self, x, y):
        return self.log_normalization - 0.5 * bilinear_sym(self.inv_sigma, _np.asarray(x, float) - _np.asarray(y, float))

    def propose(self, y, rng=_np.random.mtrand):
        return y + self._get_gauss_sample(rng)


class Gauss(ProbabilityDensity):
    r"""Gaussian density :math:`N(\mu, \Sigma)`; component type of :class:`MixtureDensity` (gauss.pyx:69-163)."""

    def __init__(self, mu, sigma):
        self.update(mu, sigma)

    def update(self, mu, sigma):
        """Re-initialise with a new mean and covariance; on ``LinAlgError`` the old state stays (gauss.pyx:86-116)."""
        local = LocalGauss(sigma)          # raises before anything is modified
        self._local_gauss = local
        self.mu = _np.array(mu, dtype=float)
        self.dim = len(self.mu)
        self.inv_sigma = local.inv_sigma
        self.log_det_sigma = local.log_det_sigma
        self.sigma = local.sigma
        self._record = None                # packed CUDA record, rebuilt lazily
        assert self.dim == self.sigma.shape[0], \
            "Dimensions of mean (%d) and covariance matrix (%d) do not match!" % (self.dim, self.sigma.shape[0])

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        local = self._local_gauss.__deepcopy__(memo)
        new._local_gauss = local
        new.mu = self.mu.copy()
        new.inv_sigma, new.log_det_sigma, new.sigma = local.inv_sigma, local.log_det_sigma, local.sigma
        return new

    # -- CUDA record -----------------------------------------------------------------------------------
    _mode = _lib.MODE_GAUSS

    def _packed_record(self):
        """Record of this component for kernel K1 (T = L^-1, centre, log-normalisation)."""
        if self._record is None:
            scalars = _np.zeros(_lib.NUM_SCALARS)
            scalars[0] = self._local_gauss.log_normalization
            scalars[_lib.S_WEIGHT] = 1.0
            self._record = _lib.pack_record(tri_from_chol(self._local_gauss.cholesky_sigma), self.mu, scalars)
        return self._record

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_record"] = None
        return state

    # -- evaluation ----------------------------------------------------------------------------------------
    def evaluate(self, x):
        x = _np.asarray(x, dtype=float)
        return float(self.multi_evaluate(x.reshape(1, -1))[0])

    def multi_evaluate(self, x, out=None):
        """log-pdf of every row of ``x`` (gauss.pyx:132-153) -- one launch of kernel K1 with K = 1."""
        from ._eval import run_k1
        x = _dev.as_samples(x)
        n = x.shape[0]
        assert x.shape[1] == self.dim, "The points in ``x`` have the wrong dimension (%i instead of %i)" % (x.shape[1], self.dim)
        packed = _dev.PackedComponents(self._packed_record()[None, :], [0])
        if _dev.is_device_tensor(x):
            t = _dev.torch()
            if out is None:
                out = t.empty(n, dtype=t.float64, device=x.device)
            else:
                assert len(out) == n
            run_k1(x, packed, 1, self._mode, lp=out.view(n, 1))
            return out
        if out is None:
            out = _np.empty(n)
        else:
            assert len(out) == n
        direct = isinstance(out, _np.ndarray) and out.dtype == _np.float64 and out.flags.c_contiguous and out.ndim == 1
        buf = out if direct else _np.empty(n)
        run_k1(x, packed, 1, self._mode, lp=buf.reshape(n, 1))
        if not direct:
            out[:] = buf
        return out

    def propose(self, N=1, rng=_np.random.mtrand):
        """``N`` draws mu + L z, z ~ N(0, 1), one ``rng.normal(0, 1, dim)`` call per draw like gauss.pyx:159-163."""
        output = _np.empty((N, self.dim))
        for i in range(N):
            output[i] = self._local_gauss.propose(self.mu, rng)
        return output


def batch_update(components, means, covs, dofs=None):
    """Install new parameters in all ``components`` (all :class:`Gauss`, or all ``StudentT`` with ``dofs``) at once:
    what ``component.update(mean, cov[, dof])`` does for each of them (gauss.pyx:86-116, student_t.pyx:78-117), with the
    K factorisations batched and the packed CUDA records of the new parameters formed on the way.  All or nothing:
    raises ``LinAlgError`` / ``ValueError`` / ``AssertionError`` BEFORE any component is touched if any covariance is
    unusable -- the caller then runs the reference's per-component loop, which decides component by component."""
    k = len(components)
    if k == 0:
        return
    means = _np.array(means, dtype=float).reshape(k, -1)
    covs = _np.array(covs, dtype=float)
    d = means.shape[1]
    assert covs.shape == (k, d, d), \
        "Dimensions of mean (%d) and covariance matrix (%d) do not match!" % (d, covs.shape[-1])
    low, inv, log_det, t = chol_inv_det_batch(covs)
    student = dofs is not None
    scal = _np.zeros((k, _lib.NUM_SCALARS))
    locals_ = []
    for i, comp in enumerate(components):
        local = comp._local_t.__class__.__new__(comp._local_t.__class__) if student else LocalGauss.__new__(LocalGauss)
        if student:
            dof = float(dofs[i])
            assert dof > 0., "Degree of freedom (``dof``) must be greater than zero (got %g)." % dof
            local.dof, local.symmetric = dof, True
        local.cholesky_sigma, local.inv_sigma, local.log_det_sigma = low[i], inv[i], float(log_det[i])
        local.sigma, local.dim = covs[i], d
        local._compute_norm()
        locals_.append(local)
        scal[i, 0] = local.log_normalization
        scal[i, _lib.S_WEIGHT] = 1.0
        if student:
            scal[i, 1:5] = (-.5 * (dof + d), 1. / dof, dof, dof + float(d))
    records = _lib.pack_records(t, means, scal)
    for i, comp in enumerate(components):                       # nothing can fail from here on
        local = locals_[i]
        if student:
            comp._local_t, comp.dof = local, local.dof
            comp._eval_prefactor, comp._inv_dof = -.5 * (local.dof + d), 1. / local.dof
        else:
            comp._local_gauss = local
        comp.mu, comp.dim = means[i].copy(), d
        comp.inv_sigma, comp.log_det_sigma, comp.sigma = local.inv_sigma, local.log_det_sigma, local.sigma
        comp._record = records[i]
