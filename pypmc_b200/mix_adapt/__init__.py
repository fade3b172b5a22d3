"""Mixture adaptation with the public API of ``pypmc.mix_adapt`` for the hot path: PMC updates
(pmc.pyx) and variational-Bayes Gaussian inference (variational.pyx)."""
from . import pmc, variational  # noqa: F401
