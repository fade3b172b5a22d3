"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy-level front end of ``oracle/pmc_oracle.c``: a CPU restatement of the
hot path of pypmc v1.2.6 (commit 9e0ab49).  Each function cites the reference
``file:line`` it follows.  The reference's per-component set-up (Cholesky,
explicit inverse, log-determinant, gammaln/digamma) runs through the same
scipy calls the reference uses, so parity at that boundary is by construction;
the N-loops run through the C file.

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker
or the timed CPU baseline, never as part of the product.  ``pypmc_b200`` does
not import it.

Parity status: PINNED (tests/test_oracle.py: reference golden numbers and
fixtures generated from the compiled reference by tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
from scipy.linalg import cholesky as _cholesky
from scipy.linalg.lapack import get_lapack_funcs as _lapack
from scipy.special import digamma as _digamma
from scipy.special import gammaln as _gammaln

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pmc_oracle.c")
_LIB = os.path.join(_HERE, "libpmc_oracle.so")
TINY = np.finfo("d").tiny


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, no FMA contraction -- see the C header)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, _SRC, "-lm"]
        )
    return _LIB


_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_bilinear_sym.restype = ctypes.c_double
        _lib.orc_logsumexp.restype = ctypes.c_double
        _lib.orc_vb_log_q_z.restype = ctypes.c_double
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _live(live, K):
    arr = np.ascontiguousarray(range(K) if live is None else list(live), dtype=np.intc)
    return arr, arr.ctypes.data_as(_ip), ctypes.c_int(len(arr))


_i64 = ctypes.c_int64
_pd = ctypes.c_ssize_t
_int = ctypes.c_int
_dbl = ctypes.c_double


# --------------------------------------------------------------------------- L1
def bilinear_sym(matrix, vector) -> float:
    """pypmc/tools/_linalg.pyx:10-39."""
    m, v = _c(matrix), _c(vector)
    return lib().orc_bilinear_sym(_d(m), _pd(m.shape[1]), _d(v), _int(len(v)))


def chol_inv_det(m):
    """pypmc/tools/_linalg.pyx:41-95: (L, M^-1, log det M) via scipy cholesky +
    LAPACK potri, raising LinAlgError for non-finite / asymmetric / non-PD input."""
    m = np.asarray_chkfinite(m)
    if not np.allclose(m, m.T):
        raise np.linalg.LinAlgError("matrix not symmetric:\n" + repr(m))
    low = _cholesky(m, True)
    inv = _lapack("potri", (m,))(low, True)[0]
    il = np.tril_indices(len(m), -1)
    inv[il[1], il[0]] = inv[il]
    log_det = 0.0
    for i in range(len(m)):
        log_det += np.log(low[i, i])
    log_det *= 2.0
    if not np.isfinite(log_det):
        raise np.linalg.LinAlgError("Nonpositive eigenvalues lead to invalid determinant " + repr(log_det))
    return low, inv, log_det


def logsumexp(a, weights) -> float:
    """pypmc/tools/_regularize.pyx:19-55."""
    a, w = _c(a), _c(weights)
    return lib().orc_logsumexp(_d(a), _d(w), _i64(len(a)))


def logsumexp2D(a, weights):
    """pypmc/tools/_regularize.pyx:57-83."""
    a, w = _c(a), _c(weights)
    assert (w >= 0.0).all(), "Found negative weight"
    res = np.zeros(len(a))
    lib().orc_logsumexp2D(_d(a), _i64(a.shape[0]), _int(a.shape[1]), _pd(a.shape[1]), _d(w), _d(res))
    return res


# --------------------------------------------------------------------------- L2
def gauss_log_norm(dim, log_det_sigma):
    """pypmc/density/gauss.pyx:54-56."""
    return -0.5 * dim * np.log(2 * np.pi) - 0.5 * log_det_sigma


def student_t_log_norm(dim, dof, log_det_sigma):
    """pypmc/density/student_t.pyx:32-34."""
    return _gammaln(0.5 * (dof + dim)) - _gammaln(0.5 * dof) - 0.5 * dim * np.log(dof * np.pi) - 0.5 * log_det_sigma


class Components:
    """Host-side per-component constants exactly as the reference's ``Gauss`` /
    ``StudentT`` objects hold them (gauss.pyx:86-116, student_t.pyx:78-117)."""

    def __init__(self, means, covs, dofs=None):
        self.mu = _c(np.atleast_2d(means))
        self.K, self.D = self.mu.shape
        covs = np.asarray(covs, dtype=float).reshape(self.K, self.D, self.D)
        self.sigma = _c(covs)
        self.inv_sigma = np.empty_like(self.sigma)
        self.log_det = np.empty(self.K)
        for k in range(self.K):
            _, self.inv_sigma[k], self.log_det[k] = chol_inv_det(self.sigma[k])
        self.dof = None if dofs is None else _c(np.broadcast_to(np.asarray(dofs, float), (self.K,)))
        if self.dof is None:
            self.log_norm = np.array([gauss_log_norm(self.D, ld) for ld in self.log_det])
        else:
            self.log_norm = np.array(
                [student_t_log_norm(self.D, nu, ld) for nu, ld in zip(self.dof, self.log_det)]
            )
            self.prefactor = -0.5 * (self.dof + self.D)
            self.inv_dof = 1.0 / self.dof


def component_multi_evaluate(x, comps: Components, k: int, out, out_stride=1):
    """Gauss.multi_evaluate (gauss.pyx:132-153) / StudentT.multi_evaluate
    (student_t.pyx:135-166) for component ``k`` into a (possibly strided) column."""
    x = _c(x)
    N, D = x.shape
    if comps.dof is None:
        lib().orc_gauss_multi_evaluate(
            _d(x), _i64(N), _pd(D), _int(D), _d(comps.mu[k]), _d(comps.inv_sigma[k]),
            _dbl(comps.log_norm[k]), out, _pd(out_stride))
    else:
        lib().orc_student_t_multi_evaluate(
            _d(x), _i64(N), _pd(D), _int(D), _d(comps.mu[k]), _d(comps.inv_sigma[k]),
            _dbl(comps.log_norm[k]), _dbl(comps.prefactor[k]), _dbl(comps.inv_dof[k]),
            out, _pd(out_stride))


def mixture_multi_evaluate(x, comps: Components, weights, individual=None, components=None):
    """MixtureDensity.multi_evaluate (pypmc/density/mixture.pyx:112-156).

    Returns ``(log_q, individual)``; ``log_q`` is None when ``components`` is given.
    """
    x = _c(x)
    N, K = len(x), comps.K
    if individual is None:
        individual = np.empty((N, K))
    assert individual.flags.c_contiguous and individual.shape == (N, K)
    ks = range(K) if components is None else components
    base = individual.ctypes.data
    for k in ks:
        col = ctypes.cast(base + 8 * int(k), _dp)
        component_multi_evaluate(x, comps, int(k), col, K)
    if components is None:
        return logsumexp2D(individual, weights), individual
    return None, individual


# --------------------------------------------------------------------------- L3: PMC
def calculate_rho_rb(x, comps: Components, weights, live=None):
    """pypmc/mix_adapt/pmc.pyx:23-43. Returns ``(rho, log_denominator)``."""
    x, w = _c(x), _c(weights)
    N, K = len(x), comps.K
    rho = np.zeros((N, K))
    live_arr, lp, ln = _live(live, K)
    mixture_multi_evaluate(x, comps, w, individual=rho, components=list(live_arr))
    log_den = np.zeros(N)
    lib().orc_rho_rb_inplace(_d(rho), _i64(N), _int(K), _d(w), lp, ln, _d(log_den))
    return rho, log_den


def calculate_rho_non_rb(N, K, latent, live=None):
    """pypmc/mix_adapt/pmc.pyx:45-51."""
    rho = np.zeros((N, K))
    latent = np.asarray(latent)
    for k in (range(K) if live is None else live):
        rho[latent == k, k] = 1.0
    return rho


def student_t_gamma(x, comps: Components, live=None):
    """pypmc/mix_adapt/pmc.pyx:602-610 (entries of dead components are left at 0)."""
    x = _c(x)
    N, D = x.shape
    gamma = np.zeros((N, comps.K))
    live_arr, lp, ln = _live(live, comps.K)
    lib().orc_student_t_gamma(_d(x), _i64(N), _pd(D), _int(D), _int(comps.K), _d(comps.mu),
                              _d(comps.inv_sigma), _d(comps.dof), lp, ln, _d(gamma))
    return gamma


def pmc_moments(x, rho, sample_weights=None, gamma=None, live=None):
    """The update equations pypmc/mix_adapt/pmc.pyx:188-222 (Gaussian, ``gamma``
    None) and :612-650 (Student-t): returns ``(alpha, mu, cov)`` with
    ``alpha`` *not yet* divided by the weight normalisation's caller-side use:
    alpha = sum w rho / sum w (pmc.pyx:191-193), mu = sum w rho gamma x /
    regularize(sum w rho gamma), cov_k = sum w rho gamma (x-mu_k)(x-mu_k)^T /
    regularize(sum w rho) for live k (rows of dead k are left uninitialised in
    the reference; here they are zero)."""
    x, rho = _c(x), _c(rho)
    N, D = x.shape
    K = rho.shape[1]
    w = None if sample_weights is None else _c(sample_weights)
    g = None if gamma is None else _c(gamma)
    alpha = np.zeros(K)
    mu_norm = np.zeros(K)
    mu = np.zeros((K, D))
    lib().orc_pmc_first_moments(_d(x), _i64(N), _pd(D), _int(D), _int(K), _d(w), _d(rho), _d(g),
                                _d(alpha), _d(mu_norm), _d(mu))
    alpha[alpha == 0] = TINY          # regularize(alpha) in place, pmc.pyx:192
    inv_alpha = 1.0 / alpha
    mu_norm[mu_norm == 0] = TINY      # pmc.pyx:622
    mu *= (1.0 / mu_norm)[:, None]
    weight_normalization = float(N) if w is None else w.sum()
    cov = np.zeros((K, D, D))
    live_arr, lp, ln = _live(live, K)
    lib().orc_pmc_second_moments(_d(x), _i64(N), _pd(D), _int(D), _int(K), _d(w), _d(rho), _d(g),
                                 _d(mu), lp, ln, _d(cov))
    cov *= inv_alpha[:, None, None]
    return alpha / weight_normalization, mu, cov


def student_t_dof_const(x, comps: Components, rho, sample_weights=None, live=None):
    """pypmc/mix_adapt/pmc.pyx:654-691: ``1 - sum_n [w_n](xi+delta)_nk / sum w``."""
    x, rho = _c(x), _c(rho)
    N, D = x.shape
    K = comps.K
    w = None if sample_weights is None else _c(sample_weights)
    psi1 = _c(_digamma(0.5 * (D + comps.dof)))
    psi2 = _c(_digamma(0.5 * comps.dof))
    out = np.zeros(K)
    live_arr, lp, ln = _live(live, K)
    lib().orc_student_t_dof_stat(_d(x), _i64(N), _pd(D), _int(D), _int(K), _d(comps.mu),
                                 _d(comps.inv_sigma), _d(comps.dof), _d(psi1), _d(psi2),
                                 _d(w), _d(rho), lp, ln, _d(out))
    weight_normalization = float(N) if w is None else w.sum()
    return 1.0 - out / weight_normalization


# --------------------------------------------------------------------------- L3: VB
def vb_expectation_det_ln_lambda(nu, log_det_W, dim):
    """pypmc/mix_adapt/variational.pyx:759-772."""
    res = np.zeros_like(nu)
    for i in range(1, dim + 1):
        res += _digamma(0.5 * (nu + 1.0 - i))
    res += dim * np.log(2.0)
    res += log_det_W
    return res


def vb_expectation_ln_pi(alpha):
    """pypmc/mix_adapt/variational.pyx:800-804."""
    return _digamma(alpha) - _digamma(alpha.sum())


def vb_e_step(x, m, W, beta, nu, alpha, log_det_W, sample_weights=None):
    """GaussianInference.E_step (pypmc/mix_adapt/variational.pyx:116-127).

    ``sample_weights`` must already be normalised to sum N (variational.pyx:94).
    Returns a dict with the public attributes the reference fills."""
    x, m, W = _c(x), _c(m), _c(W)
    beta, nu, alpha = _c(beta), _c(nu), _c(alpha)
    N, D = x.shape
    K = len(m)
    w = None if sample_weights is None else _c(sample_weights)
    e_det = _c(vb_expectation_det_ln_lambda(nu, _c(log_det_W), D))
    e_pi = _c(vb_expectation_ln_pi(alpha))
    E = np.zeros((N, K))
    lib().orc_vb_gauss_exponent(_d(x), _i64(N), _pd(D), _int(D), _int(K), _d(m), _d(W), _d(beta),
                                _d(nu), _d(E))
    log_rho = np.zeros((N, K))
    r = np.zeros((N, K))
    lib().orc_vb_update_r(_d(E), _i64(N), _int(K), _int(D), _d(e_pi), _d(e_det), _d(log_rho), _d(r))
    N_comp = np.zeros(K)
    x_mean = np.zeros((K, D))
    S = np.zeros((K, D, D))
    lib().orc_vb_statistics(_d(x), _i64(N), _pd(D), _int(D), _int(K), _d(w), _d(r),
                            _d(N_comp), _d(x_mean), _d(S))
    log_q_z = lib().orc_vb_log_q_z(_d(r), _d(log_rho), _d(w), _i64(N), _int(K))
    return dict(expectation_det_ln_lambda=e_det, expectation_ln_pi=e_pi,
                expectation_gauss_exponent=E, log_rho=log_rho, r=r, N_comp=N_comp,
                inv_N_comp=1.0 / N_comp, x_mean_comp=x_mean, S=S, expectation_log_q_Z=log_q_z)
