"""Per-component linear algebra on the host (K matrices of size D x D per update -- off the N loop).

``chol_inv_det`` has the contract of pypmc/tools/_linalg.pyx:41-95 (same scipy calls, same
``LinAlgError`` conditions) because callers rely on its failure behaviour (gauss.pyx:38-47,
pmc.pyx:234-244).  ``tri_from_chol`` / ``tri_from_precision`` produce the lower-triangular factor T with
T^T T = M^-1 resp. = W that kernel K1 multiplies with (x - mu); they replace the explicit-inverse
bilinear form ``bilinear_sym`` (_linalg.pyx:10-39) of the reference's inner loop.
"""
import numpy as _np
from scipy.linalg import cholesky as _cholesky
from scipy.linalg import solve_triangular as _solve_triangular
from scipy.linalg.lapack import get_lapack_funcs as _get_lapack_funcs


def chol_inv_det(m):
    """Return ``(L, M^-1, log det M)`` for a symmetric positive-definite ``m`` (L lower, M = L L^T).

    Raises ``numpy.linalg.LinAlgError`` if ``m`` is not symmetric, not positive definite or has a
    non-finite log-determinant; ``ValueError`` for non-finite input (``asarray_chkfinite``).
    """
    m = _np.asarray_chkfinite(m)
    if not _np.allclose(m, m.T):
        raise _np.linalg.LinAlgError("matrix not symmetric:\n" + repr(m))
    low = _cholesky(m, lower=True)                       # LinAlgError when not positive definite
    potri, = _get_lapack_funcs(("potri",), (m,))
    inv, info = potri(low, lower=True)
    if info != 0:
        raise _np.linalg.LinAlgError("potri failed with info=%d" % info)
    rows, cols = _np.tril_indices(len(m), -1)
    inv[cols, rows] = inv[rows, cols]                    # potri fills one triangle only
    log_det = 2.0 * float(_np.sum(_np.log(_np.diag(low))))
    if not _np.isfinite(log_det):
        raise _np.linalg.LinAlgError("Nonpositive eigenvalues lead to invalid determinant " + repr(log_det))
    return low, inv, log_det


def tri_from_chol(low):
    """T = L^-1 (lower triangular), so that ||T y||^2 = y^T (L L^T)^-1 y."""
    return _solve_triangular(low, _np.eye(len(low)), lower=True)


def tri_from_precision(w):
    """Lower-triangular T with T^T T = W: Cholesky of the index-reversed matrix, reversed back."""
    rev = _np.ascontiguousarray(w[::-1, ::-1])
    m = _cholesky(rev, lower=True)
    return _np.ascontiguousarray(m.T[::-1, ::-1])


def bilinear_sym(matrix, vector):
    """x^T M x for symmetric M (host utility, K-sized use only; API of _linalg.pyx:10)."""
    vector = _np.asarray(vector, dtype=float)
    return float(vector @ _np.asarray(matrix, dtype=float) @ vector)
