"""Student's t densities with the API of pypmc/density/student_t.pyx (``LocalStudentT`` :13-55, ``StudentT`` :57-176)."""
import numpy as _np
from scipy.special import gammaln as _gammaln

from .base import ProbabilityDensity
from .gauss import LocalGauss, Gauss
from ..tools._linalg import tri_from_chol, bilinear_sym
from .. import _lib


class LocalStudentT(LocalGauss):
    """Multivariate local Student's t with redefinable covariance (student_t.pyx:13-55)."""

    def __init__(self, sigma, dof):
        dof = float(dof)
        self.symmetric = True
        assert dof > 0., "Degree of freedom (``dof``) must be greater than zero (got %g)." % dof
        self.dof = dof
        self.update(sigma)

    def _compute_norm(self):
        # student_t.pyx:32-34
        self.log_normalization = _gammaln(.5 * (self.dof + self.dim)) - _gammaln(.5 * self.dof) \
            - 0.5 * self.dim * _np.log(self.dof * _np.pi) - 0.5 * self.log_det_sigma

    def evaluate(self, x, y):
        diff = _np.asarray(x, float) - _np.asarray(y, float)
        return self.log_normalization - .5 * (self.dof + self.dim) * _np.log(1. + bilinear_sym(self.inv_sigma, diff) / self.dof)

    def propose(self, y, rng=_np.random.mtrand):
        return y + self._get_gauss_sample(rng) * _np.sqrt(self.dof / rng.chisquare(self.dof))


class StudentT(ProbabilityDensity):
    r"""Student's t density with mean ``mu``, scale matrix ``sigma`` and ``dof`` degrees of freedom
    (student_t.pyx:57-176); component type of :class:`MixtureDensity`."""

    def __init__(self, mu, sigma, dof):
        self.update(mu, sigma, dof)

    def update(self, mu, sigma, dof):
        """Re-initialise; on ``LinAlgError`` the old (mu, sigma, dof) stay in place (student_t.pyx:78-117)."""
        dof = float(dof)
        local = LocalStudentT(sigma, dof)     # raises before anything is modified
        self._local_t = local
        self.mu = _np.array(mu, dtype=float)
        self.dim = len(self.mu)
        self.dof = dof
        self.inv_sigma = local.inv_sigma
        self.log_det_sigma = local.log_det_sigma
        self.sigma = local.sigma
        self._record = None
        assert self.dim == self.sigma.shape[0], \
            "Dimensions of mean (%d) and covariance matrix (%d) do not match!" % (self.dim, self.sigma.shape[0])
        self._eval_prefactor = - .5 * (self.dof + self.dim)
        self._inv_dof = 1. / self.dof

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        local = self._local_t.__deepcopy__(memo)
        new._local_t = local
        new.mu = self.mu.copy()
        new.inv_sigma, new.log_det_sigma, new.sigma = local.inv_sigma, local.log_det_sigma, local.sigma
        return new

    _mode = _lib.MODE_STUDENT_T

    def _packed_record(self):
        if self._record is None:
            scalars = _np.zeros(_lib.NUM_SCALARS)
            scalars[0] = self._local_t.log_normalization
            scalars[1] = self._eval_prefactor
            scalars[2] = self._inv_dof
            scalars[3] = self.dof
            scalars[4] = self.dof + float(self.dim)
            scalars[_lib.S_WEIGHT] = 1.0
            self._record = _lib.pack_record(tri_from_chol(self._local_t.cholesky_sigma), self.mu, scalars)
        return self._record

    __getstate__ = Gauss.__getstate__
    evaluate = Gauss.evaluate
    multi_evaluate = Gauss.multi_evaluate      # same launch, Student-t epilogue (student_t.pyx:135-166)

    def propose(self, N=1, rng=_np.random.mtrand):
        """``N`` draws; per draw one ``rng.normal(0,1,dim)`` and one ``rng.chisquare(dof)`` like student_t.pyx:172-176."""
        output = _np.empty((N, self.dim))
        for i in range(N):
            output[i] = self._local_t.propose(self.mu, rng)
        return output
