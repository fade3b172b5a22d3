"""Callers of the hot path: importance sampling with a batched weight pass (pypmc/sampler/importance_sampling.py)."""
from . import importance_sampling  # noqa: F401
