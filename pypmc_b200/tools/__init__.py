"""Host-side helpers of the hot path (per-component set-up, K-sized arithmetic); ``History`` and ``convergence`` are
exported like ``pypmc.tools`` does (pypmc/tools/__init__.py)."""
from ._history import History  # noqa: F401
from . import convergence  # noqa: F401
