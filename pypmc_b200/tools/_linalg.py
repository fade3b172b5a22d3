"""Per-component linear algebra on the host (K matrices of size D x D per update -- off the N loop).

``chol_inv_det`` has the contract of pypmc/tools/_linalg.pyx:41-95 (same scipy calls, same
``LinAlgError`` conditions) because callers rely on its failure behaviour (gauss.pyx:38-47,
pmc.pyx:234-244).  ``tri_from_chol`` / ``tri_from_precision`` produce the lower-triangular factor T with
T^T T = M^-1 resp. = W that kernel K1 multiplies with (x - mu); they replace the explicit-inverse
bilinear form ``bilinear_sym`` (_linalg.pyx:10-39) of the reference's inner loop.
"""
import functools as _functools

import numpy as _np
from scipy.linalg import cholesky as _cholesky
from scipy.linalg.lapack import get_lapack_funcs as _get_lapack_funcs

# float64 LAPACK routines, looked up once (scipy's wrappers re-validate their input on every call, which costs
# more than the factorisation of a 30 x 30 matrix; the update calls this K times per iteration)
_potri, _trtri = _get_lapack_funcs(("potri", "trtri"), (_np.empty((1, 1)),))


@_functools.lru_cache(maxsize=64)
def _strict_lower(d):
    return _np.tril_indices(d, -1)


def chol_inv_det(m):
    """Return ``(L, M^-1, log det M)`` for a symmetric positive-definite ``m`` (L lower, M = L L^T).

    Raises ``numpy.linalg.LinAlgError`` if ``m`` is not symmetric, not positive definite or has a
    non-finite log-determinant; ``ValueError`` for non-finite input (``asarray_chkfinite``).
    """
    m = _np.asarray_chkfinite(m, dtype=float)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("expected square matrix")
    # numpy.allclose(m, m.T) for finite input (_linalg.pyx:62): |m - m^T| <= atol + rtol |m^T|
    if not (_np.abs(m - m.T) <= 1e-8 + 1e-5 * _np.abs(m.T)).all():
        raise _np.linalg.LinAlgError("matrix not symmetric:\n" + repr(m))
    # the reference's own call (_linalg.pyx:67); LinAlgError when not positive definite.  (Calling LAPACK potrf
    # directly is faster but takes the other triangle's code path for C-ordered input: last-bit differences.)
    try:
        low = _cholesky(m, lower=True, check_finite=False)
    except _np.linalg.LinAlgError as error:
        # scipy >= 1.15 words this "Internal potrf return info = ..."; callers and the reference's own tests
        # (tools/linalg_test.py:45) look for the classic wording
        raise _np.linalg.LinAlgError("matrix not positive definite (%s)" % (error,))
    inv, info = _potri(low, lower=True)
    if info != 0:
        raise _np.linalg.LinAlgError("potri failed with info=%d" % info)
    rows, cols = _strict_lower(len(m))
    inv[cols, rows] = inv[rows, cols]                    # potri fills one triangle only
    log_det = 0.0
    for v in _np.log(low.diagonal()).tolist():           # summed in index order like _linalg.pyx:84-90
        log_det += v
    log_det *= 2.0
    if not _np.isfinite(log_det):
        raise _np.linalg.LinAlgError("Nonpositive eigenvalues lead to invalid determinant " + repr(log_det))
    return low, inv, log_det


def chol_inv_det_batch(ms):
    """``chol_inv_det`` for a stack ``ms`` [K, D, D] in a handful of batched LAPACK calls (an update installs K new
    covariances at once; K separate scipy calls cost 2 ms of host time at K = 32).  Returns ``(L, M^-1, log det M, T)``
    with ``T = L^-1`` (lower triangular, T^T T = M^-1 -- the factor kernel K1 multiplies with).  Same acceptance
    conditions as ``chol_inv_det``; raises if ANY matrix fails (``LinAlgError`` / ``ValueError``) without saying which
    -- callers fall back to the per-matrix routine, whose behaviour is the reference's."""
    ms = _np.asarray_chkfinite(ms, dtype=float)
    if ms.ndim != 3 or ms.shape[1] != ms.shape[2]:
        raise ValueError("expected a stack of square matrices")
    mt = ms.transpose(0, 2, 1)
    if not (_np.abs(ms - mt) <= 1e-8 + 1e-5 * _np.abs(mt)).all():
        raise _np.linalg.LinAlgError("matrix not symmetric")
    low = _np.linalg.cholesky(ms)                       # LinAlgError when a matrix is not positive definite
    # T = L^-1 by LAPACK dtrtri per matrix, the routine behind tri_from_chol (same bits as the per-component path; a
    # batched LU solve against the identity cost 0.45 ms at K = 32, D = 30, K dtrtri calls 0.15 ms)
    t = _np.empty_like(low)
    for i in range(low.shape[0]):
        ti, info = _trtri(low[i], lower=1)
        if info != 0:
            raise _np.linalg.LinAlgError("trtri failed with info=%d" % info)
        t[i] = ti
    t = _np.tril(t)
    inv = t.transpose(0, 2, 1) @ t
    inv = 0.5 * (inv + inv.transpose(0, 2, 1))
    log_det = 2.0 * _np.log(_np.diagonal(low, axis1=1, axis2=2)).sum(axis=1)
    if not _np.isfinite(log_det).all():
        raise _np.linalg.LinAlgError("Nonpositive eigenvalues lead to invalid determinant")
    return low, inv, log_det, t


def tri_from_chol(low):
    """T = L^-1 (lower triangular), so that ||T y||^2 = y^T (L L^T)^-1 y."""
    t, info = _trtri(low, lower=1)
    if info != 0:
        raise _np.linalg.LinAlgError("trtri failed with info=%d" % info)
    return _np.tril(t)


def tri_from_precision(w):
    """Lower-triangular T with T^T T = W: Cholesky of the index-reversed matrix, reversed back."""
    rev = _np.ascontiguousarray(w[::-1, ::-1])
    m = _cholesky(rev, lower=True)
    return _np.ascontiguousarray(m.T[::-1, ::-1])


def bilinear_sym(matrix, vector):
    """x^T M x for symmetric M (host utility, K-sized use only; API of _linalg.pyx:10)."""
    vector = _np.asarray(vector, dtype=float)
    return float(vector @ _np.asarray(matrix, dtype=float) @ vector)
