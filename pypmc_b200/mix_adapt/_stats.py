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
    mean = shift[None, :] + delta
    # sum v (y - delta)(y - delta)^T = R - delta m^T - m delta^T + B_true delta delta^T
    outer_dm = delta[:, :, None] * m[:, None, :]
    cov = R - outer_dm - np.swapaxes(outer_dm, 1, 2) + st["B"][:, None, None] * (delta[:, :, None] * delta[:, None, :])
    cov = 0.5 * (cov + np.swapaxes(cov, 1, 2)) / A[:, None, None]
    return A, mean, cov
