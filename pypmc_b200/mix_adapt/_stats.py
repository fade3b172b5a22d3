"""The per-rank statistics packet of one proposal update and its host-side finishing.

One contiguous float64 vector per update is what crosses GPUs (a single all-reduce, ``parallel.py``):

    [ K rows of (A, B, m[D], R[D(D+1)/2], L) | K latent counts | sum_n w_n log q_n (or sum w r log r) | sum_n w_n ]

(A, B, m, R, L) come from kernel K2 (csrc/k2_suffstats.cuh), the two scalars from kernel K1.  Everything in
this module is K-sized host arithmetic -- the part of the reference that runs after its N-loops
(pmc.pyx:191-204 normalisations, variational.pyx:699-709/822-830/876-890 normalisations).
"""
from __future__ import annotations

import numpy as np

from ..tools._regularize import regularize


class PacketLayout:
    def __init__(self, K: int, D: int):
        self.K, self.D = K, D
        self.T = D * (D + 1) // 2
        self.row = 3 + D + self.T           # A, B, m, R, L
        self.stats_len = K * self.row
        self.off_counts = self.stats_len
        self.off_sum_a = self.stats_len + K      # K1 writes (sum_a, sumw) as one pair
        self.off_sumw = self.off_sum_a + 1
        self.size = self.off_sumw + 1
        self._tril = np.tril_indices(D)

    def unpack(self, packet: np.ndarray):
        """-> dict(A[K], B[K], m[K,D], R[K,D,D] symmetric, L[K], counts[K], sumw, sum_a)."""
        K, D = self.K, self.D
        rows = packet[:self.stats_len].reshape(K, self.row)
        R = np.zeros((K, D, D))
        tri = rows[:, 2 + D:2 + D + self.T]
        R[:, self._tril[0], self._tril[1]] = tri
        R[:, self._tril[1], self._tril[0]] = tri
        return dict(A=rows[:, 0].copy(), B=rows[:, 1].copy(), m=rows[:, 2:2 + D].copy(), R=R, L=rows[:, -1].copy(),
                    counts=packet[self.off_counts:self.off_counts + K].copy(),
                    sumw=float(packet[self.off_sumw]), sum_a=float(packet[self.off_sum_a]))


#: largest size of the cancelling terms, sum_ij |P_k,ij| |d_i| |d_j| with d = mu_k - c and P_k = Sigma_k^-1, tolerated
#: between a component and the shift c its raw moments are taken about.  The recovered covariance carries an absolute
#: error of ~eps |d|^2 per entry; measured in units of the component's own width (the smallest eigenvalue sees
#: eps |d|^2 lambda_max(P_k)) that is eps times this bound -- which, unlike the plain Mahalanobis distance
#: d^T P_k d, also grows when the offset lies along a WIDE axis of an ill-conditioned covariance.  <= 1e4 keeps the
#: relative rounding error of every covariance below ~2e-12.
SHIFT_CONDITION_LIMIT = 1.0e4


#: below this many (sample, component) pairs an update is launch-bound anyway and the statistics are taken the
#: reference's way, in two passes (see :func:`two_pass_suffstats`)
SMALL_PROBLEM_PAIRS = 200_000


def small_problem(n_rows, k):
    """True if this update takes the two-pass route: a small single-process problem.  With the statistics all-reduce
    enabled the one-pass route is used on every rank (two passes would need two all-reduces, and ranks hold different
    row counts but must decide alike)."""
    from .. import parallel
    return (not parallel.enabled()) and n_rows * k <= SMALL_PROBLEM_PAIRS


def shift_groups(centers, precisions, weights, live, limit=SHIFT_CONDITION_LIMIT):
    """Partition the live components into groups that share one shift vector for kernel K2.

    K2 accumulates raw second moments about a shift c; the covariance of component k recovered from them loses
    about eps * q_k(c), q_k(c) = |mu_k - c|^T |P_k| |mu_k - c| (see SHIFT_CONDITION_LIMIT), to cancellation (the reference centres every component
    on its own new mean, pmc.pyx:200-204 / variational.pyx:876-890, and has no such term).  Normally ONE group --
    shift = weighted centre of the mixture -- keeps every q_k below ``limit`` and K2 runs once.  Components that
    are far apart in units of their own width are put into separate groups, greedily in component order, and K2
    runs once per group on that group's columns: slower, never less accurate.

    Returns a list of ``(indices, shift)``; deterministic, so every rank forms the same groups.
    """
    live = list(live)
    if not live:
        return []

    def centre(idx):
        w = np.array([weights[k] for k in idx], dtype=float)
        mus = np.array([centers[k] for k in idx], dtype=float)
        if not np.isfinite(w).all() or w.sum() <= 0:
            return mus.mean(axis=0)
        return (w[:, None] * mus).sum(axis=0) / w.sum()

    def worst(idx, c):
        d = np.abs(np.array([centers[k] for k in idx], dtype=float) - c)              # [k, D]
        p = np.abs(np.array([precisions[k] for k in idx], dtype=float))               # [k, D, D]
        q = np.einsum("ki,kij,kj->k", d, p, d)
        return float(np.max(q)) if not np.isnan(q).any() else float("nan")            # (NaN: the caller keeps one group)

    c_all = centre(live)
    if not (worst(live, c_all) > limit):          # also taken when the estimate is NaN: one group, like before
        return [(live, c_all)]
    groups = []
    for k in live:
        for g in groups:
            cand = g[0] + [k]
            c = centre(cand)
            if worst(cand, c) <= limit:
                g[0], g[1] = cand, c
                break
        else:
            groups.append([[k], np.array(centers[k], dtype=float)])
    return [(g[0], g[1]) for g in groups]


def grouped_suffstats(ctx, ds, lay, packet, groups, rho, gamma, stream):
    """Run K2 once per shift group and leave the K statistics rows in ``packet`` (device).  Returns the per-component
    shift array [K, D] the rows refer to (zeros for components in no group)."""
    from .. import _device as dev
    K, D, N = lay.K, lay.D, ds.N
    index = ds.x.device.index                               # the device the samples live on
    ctx = type(ctx).get(index)
    stream = dev.current_stream_ptr(index)
    shifts = np.zeros((K, D))
    ldx = ds.x.stride(0) if N > 1 else D
    if len(groups) == 1 and len(groups[0][0]) > 0 and _is_full(groups[0][0], K):
        idx, c = groups[0]
        shifts[:] = c
        ctx.suffstats(ds.x, N, ldx, D, dev.to_device(c, index), rho, gamma, K, K, ds.w, packet, stream)
        return shifts
    t = dev.torch()
    rows = packet[:lay.stats_len].view(K, lay.row)
    rows.zero_()
    for idx, c in groups:
        cols = t.tensor(idx, device=rho.device)
        rho_g = rho.index_select(1, cols).contiguous()
        gamma_g = None if gamma is None else gamma.index_select(1, cols).contiguous()
        out = t.empty((len(idx), lay.row), dtype=t.float64, device=rho.device)
        ctx.suffstats(ds.x, N, ldx, D, dev.to_device(c, index), rho_g, gamma_g, len(idx), len(idx), ds.w, out, stream)
        rows[cols] = out
        shifts[idx] = c
    return shifts


def two_pass_suffstats(ctx, ds, lay, packet, rho, gamma, live):
    """The reference's own two-pass statistics (pmc.pyx:196-204, variational.pyx:822-830, 876-890) for small
    problems: a first launch of K2 about the origin gives sum v and sum v x, hence the new means exactly as the
    reference forms them (a component whose weighted mean is exactly zero comes out as exactly zero); a second launch
    per live component accumulates the second moments about THAT mean, so the covariance needs no cancelling
    correction and is as accurate as the reference's.  K + 1 launches and one host synchronisation in between -- only
    taken where the whole update costs microseconds (SMALL_PROBLEM_PAIRS); the VB bound then stays monotone to the last
    bit in the reference's unit tests (variational_test.py:16-37 asserts bound >= old_bound exactly).  Returns the
    per-component shift array [K, D] the rows of ``packet`` refer to."""
    from .. import _device as dev
    K, D, N = lay.K, lay.D, ds.N
    index = ds.x.device.index
    ctx = type(ctx).get(index)
    stream = dev.current_stream_ptr(index)
    t = dev.torch()
    ldx = ds.x.stride(0) if N > 1 else D
    rows = packet[:lay.stats_len].view(K, lay.row)
    ctx.suffstats(ds.x, N, ldx, D, dev.to_device(np.zeros(D), index), rho, gamma, K, K, ds.w, packet, stream)
    first = rows.cpu().numpy()
    shifts = np.zeros((K, D))
    for k in live:
        b = first[k, 1]
        if not (b != 0.0 and np.isfinite(b)):
            continue                                           # no mass: the zero-shift row stands (regularised later)
        mean = first[k, 2:2 + D] / b
        if not np.isfinite(mean).all():
            continue
        cols = t.tensor([k], device=rho.device)
        rho_k = rho.index_select(1, cols).contiguous()
        gamma_k = None if gamma is None else gamma.index_select(1, cols).contiguous()
        out = t.empty((1, lay.row), dtype=t.float64, device=rho.device)
        ctx.suffstats(ds.x, N, ldx, D, dev.to_device(mean, index), rho_k, gamma_k, 1, 1, ds.w, out, stream)
        rows[k] = out[0]
        rows[k, 2:2 + D] = 0.0          # the mean IS the first pass's (delta = 0): the second pass only supplies R about it
        shifts[k] = mean
    return shifts


def _is_full(idx, K):
    return len(idx) == K and list(idx) == list(range(K))


def moments_from_stats(st, shift, mean_norm: str):
    """Turn the shifted raw moments into the reference's two-pass quantities.

    mean_norm = 'B': PMC conventions (pmc.pyx:191-204, :615-630) -- mean = sum v x / reg(B), second moment
    centred on that mean and divided by reg(A).  For VB (``gamma`` = 1, A = B) the same formulas give
    N_k, x_mean_comp and S (variational.pyx:699-709, 822-830, 876-890).

    Returns (A_reg, mean[K,D], cov[K,D,D]) with A_reg = regularize(A) (zeros -> tiny), the denominators used.
    """
    A = regularize(st["A"].copy())
    B = regularize(st["B"].copy())
    m, R = st["m"], st["R"]
    delta = m / B[:, None]                                   # mean - shift
    shift = np.asarray(shift, dtype=float)
    mean = (shift[None, :] if shift.ndim == 1 else shift) + delta   # one shift, or one per component (shift_groups)
    # sum v (y - delta)(y - delta)^T = R - delta m^T - m delta^T + B_true delta delta^T
    outer_dm = delta[:, :, None] * m[:, None, :]
    cov = R - outer_dm - np.swapaxes(outer_dm, 1, 2) + st["B"][:, None, None] * (delta[:, :, None] * delta[:, None, :])
    cov = 0.5 * (cov + np.swapaxes(cov, 1, 2)) / A[:, None, None]
    return A, mean, cov
