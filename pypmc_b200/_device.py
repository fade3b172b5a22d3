"""Host <-> device plumbing for the density / mix_adapt classes (torch is used for device memory only)."""
from __future__ import annotations

import numpy as np

from . import _lib


_BIG_UPLOAD = 32 << 20     # bytes from which a host array is uploaded through the library's pinned staging


def torch():
    import torch as _t
    return _t


def is_device_tensor(x) -> bool:
    return (not isinstance(x, np.ndarray)) and hasattr(x, "is_cuda") and bool(x.is_cuda)


def as_samples(x):
    """Validate a sample matrix with the reference's contract (``np.ndarray[double, ndim=2] not None``,
    mixture.pyx:112): float64, two-dimensional; numpy (host) or torch CUDA tensor (device resident)."""
    if x is None:
        raise TypeError("Argument 'x' must not be None")
    if is_device_tensor(x):
        if x.dtype != torch().float64 or x.dim() != 2:
            raise ValueError("device samples must be a 2-d float64 tensor")
        return x
    if not isinstance(x, np.ndarray):
        raise TypeError("Argument 'x' has incorrect type (expected numpy.ndarray, got %s)" % type(x).__name__)
    if x.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double' but got %r" % x.dtype)
    if x.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % x.ndim)
    return x


def row_major(x: np.ndarray):
    """(array, ldx) of a host sample matrix usable by the C ABI: unit column stride, any row stride."""
    n, d = x.shape
    if n == 0:
        return np.ascontiguousarray(x), d
    if x.strides[1] == 8 and x.strides[0] >= 8 * d and x.strides[0] % 8 == 0:
        return x, x.strides[0] // 8
    x = np.ascontiguousarray(x)
    return x, d


def to_device(a, device=None, dtype=None):
    t = torch()
    if a is None:
        return None
    if is_device_tensor(a):
        return a
    index = _lib.default_device() if device is None else device
    dev = "cuda:%d" % index
    if (isinstance(a, np.ndarray) and a.dtype == np.float64 and dtype in (None, np.float64) and a.nbytes >= _BIG_UPLOAD
            and a.ndim in (1, 2) and a.strides[-1] == 8 and (a.ndim == 1 or (a.strides[0] % 8 == 0 and a.strides[0] >= 8 * a.shape[1]))):
        # large sample matrices: threaded pinned staging (3x a plain copy from pageable memory), strided rows allowed
        rows, d = (a.shape[0], 1) if a.ndim == 1 else a.shape
        ld = 1 if a.ndim == 1 else a.strides[0] // 8
        out = t.empty(a.shape, dtype=t.float64, device=dev)
        t.cuda.current_stream(index).synchronize()
        _lib.Context.get(index).upload(out, a, rows, d, ld)
        return out
    arr = np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=dtype)
    return t.from_numpy(arr).to(dev)


def current_stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on ``device`` (default: this process's device)."""
    return torch().cuda.current_stream(_lib.default_device() if device is None else device).cuda_stream


#: device copies of packed records keyed by (device, content): an unchanged mixture evaluated again (every EM step of
#: PMC.run, every step of a sampler) re-uses its records instead of paying a blocking pageable upload per call
_RECORD_CACHE = {}
_RECORD_CACHE_MAX = 32


class PackedComponents:
    """Records + output columns of the components to evaluate, on host and (lazily) on device."""

    def __init__(self, records: np.ndarray, cols, weights=None):
        self.records = np.ascontiguousarray(records, dtype=np.float64)       # [kl, RL]
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)               # [kl]
        self.kl = len(self.cols)
        if weights is not None:
            # mixture weights live in scalar slot S_WEIGHT (last 8 doubles of a record)
            self.records[:, self.records.shape[1] - _lib.NUM_SCALARS + _lib.S_WEIGHT] = weights
        self._dev = {}

    def device(self, index=None):
        index = _lib.default_device() if index is None else index
        if index not in self._dev:
            key = (index, self.records.shape, hash(self.records.tobytes()), hash(self.cols.tobytes()))
            hit = _RECORD_CACHE.pop(key, None)
            if hit is None:
                hit = (to_device(self.records, index), to_device(self.cols, index))
            _RECORD_CACHE[key] = hit                                  # most recently used last
            while len(_RECORD_CACHE) > _RECORD_CACHE_MAX:
                _RECORD_CACHE.pop(next(iter(_RECORD_CACHE)))
            self._dev[index] = hit
        return self._dev[index]
