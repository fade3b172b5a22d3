"""ctypes binding of the C ABI in ``include/pmcb200.h`` (``pypmc_b200/csrc/libpmcb200.so``).

There is no CPU fallback: if the library cannot be loaded the import fails, and every compute call
needs a CUDA device (``Context`` raises without one).
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpmcb200.so")

MODE_GAUSS, MODE_STUDENT_T, MODE_VB = 0, 1, 2
NUM_SCALARS = 8
MAX_DIM = 64
S_WEIGHT = 5
DBL_MAX = float(np.finfo("d").max)

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p

#: every symbol ``include/pmcb200.h`` declares, with (restype, argtypes)
SIGNATURES = {
    "pmcb200_version": (ctypes.c_int, []),
    "pmcb200_last_error": (ctypes.c_char_p, []),
    "pmcb200_device_count": (ctypes.c_int, []),
    "pmcb200_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "pmcb200_destroy": (ctypes.c_int, [_vp]),
    "pmcb200_record_len": (ctypes.c_int, [ctypes.c_int]),
    "pmcb200_pack_record": (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _vp]),
    "pmcb200_mixture_eval": (ctypes.c_int, [
        _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_suffstats": (ctypes.c_int, [
        _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp,
        _vp]),
    "pmcb200_mixture_eval_host": (ctypes.c_int, [
        _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int64]),
    "pmcb200_upload": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int64]),
    "pmcb200_mixture_propose": (ctypes.c_int, [
        _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_uint64, ctypes.c_uint64, _vp,
        ctypes.c_int64, _vp, _vp]),
    "pmcb200_importance_weights": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _vp, _vp, _vp]),
    "pmcb200_fp64_peak": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p]),
    "pmcb200_last_k1_kernel": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int]),
    "pmcb200_launch_count": (ctypes.c_int64, [_vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load (building first if the in-tree .so is missing) and type the shared library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.pmcb200_version() < 101:
            raise ImportError("libpmcb200.so is stale; rebuild with python -m pypmc_b200._build --force")
        _lib = lib
        return lib


class PmcB200Error(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise PmcB200Error("%s failed: %s" % (what, load().pmcb200_last_error().decode()))


def record_len(d: int) -> int:
    n = load().pmcb200_record_len(int(d))
    if n < 0:
        raise ValueError("dimension %d not supported by the CUDA kernels (1..%d)" % (d, MAX_DIM))
    return n


def pack_record(t_lower: np.ndarray, center: np.ndarray, scalars: np.ndarray) -> np.ndarray:
    """Pack one component (host side; layout in csrc/pmc_common.cuh)."""
    d = len(center)
    t = np.ascontiguousarray(t_lower, dtype=np.float64)
    c = np.ascontiguousarray(center, dtype=np.float64)
    s = np.ascontiguousarray(scalars, dtype=np.float64)
    assert t.shape == (d, d) and s.shape == (NUM_SCALARS,)
    rec = np.empty(record_len(d))
    _check(load().pmcb200_pack_record(d, t.ctypes.data, c.ctypes.data, s.ctypes.data, rec.ctypes.data), "pack_record")
    return rec


_PACK_INDEX = {}


def pack_records(t_lower: np.ndarray, centers: np.ndarray, scalars: np.ndarray) -> np.ndarray:
    """``pack_record`` for K components at once in numpy (same layout, csrc/pmc_common.cuh): ``t_lower`` [K, d, d],
    ``centers`` [K, d], ``scalars`` [K, NUM_SCALARS] -> [K, record_len(d)]."""
    t = np.asarray(t_lower, dtype=np.float64)
    k, d = t.shape[0], t.shape[1]
    idx = _PACK_INDEX.get(d)
    if idx is None:
        ii, jj = np.tril_indices(d)
        pos = 2 * (ii // 2) * (ii // 2 + 1) + 4 * (jj // 2) + 2 * (ii % 2) + (jj % 2)
        idx = _PACK_INDEX[d] = (ii, jj, pos)
    ii, jj, pos = idx
    dp = (d + 1) & ~1
    nt = (dp // 2) * (dp // 2 + 1) * 2
    rec = np.zeros((k, record_len(d)))
    rec[:, pos] = t[:, ii, jj]
    rec[:, nt:nt + d] = centers
    rec[:, nt + dp:] = scalars
    return rec


def device_count() -> int:
    return load().pmcb200_device_count()


def _ptr(t):
    """Device/host address of a torch tensor, numpy array or None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class Context:
    """One ``pmcb200_ctx`` per CUDA device (scratch buffers + copy streams)."""

    _instances: dict = {}

    def __init__(self, device: int):
        lib = load()
        if lib.pmcb200_device_count() <= device:
            raise PmcB200Error(
                "pypmc_b200 needs a CUDA device (sm_100a); none visible as cuda:%d and there is no CPU fallback" % device)
        h = _vp()
        _check(lib.pmcb200_create(device, ctypes.byref(h)), "pmcb200_create")
        self.handle = h
        self.device = device

    @classmethod
    def get(cls, device=None) -> "Context":
        if device is None:
            device = default_device()
        ctx = cls._instances.get(device)
        if ctx is None:
            ctx = cls._instances[device] = Context(device)
        return ctx

    # -- raw calls ------------------------------------------------------------------------------
    def mixture_eval(self, x, n, ldx, d, records, cols, kl, k_out, mode, max_init, logq=None, lp=None, resp=None,
                     aux=None, weights=None, sums=None, stream=0):
        _check(load().pmcb200_mixture_eval(self.handle, _ptr(x), n, ldx, d, _ptr(records), _ptr(cols), kl, k_out, mode,
                                           max_init, _ptr(logq), _ptr(lp), _ptr(resp), _ptr(aux), _ptr(weights),
                                           _ptr(sums), stream), "pmcb200_mixture_eval")

    def suffstats(self, x, n, ldx, d, shift, rho, gamma, k, ld_rho, weights, out, stream=0):
        _check(load().pmcb200_suffstats(self.handle, _ptr(x), n, ldx, d, _ptr(shift), _ptr(rho), _ptr(gamma), k, ld_rho,
                                        _ptr(weights), _ptr(out), stream), "pmcb200_suffstats")

    def mixture_eval_host(self, x, n, ldx, d, records, cols, kl, k_out, mode, max_init, logq=None, lp=None, resp=None,
                          aux=None, weights=None, sums=None, chunk_rows=0):
        _check(load().pmcb200_mixture_eval_host(self.handle, _ptr(x), n, ldx, d, _ptr(records), _ptr(cols), kl, k_out,
                                                mode, max_init, _ptr(logq), _ptr(lp), _ptr(resp), _ptr(aux),
                                                _ptr(weights), _ptr(sums), chunk_rows), "pmcb200_mixture_eval_host")

    def upload(self, dst, src, rows, d, ld_src):
        _check(load().pmcb200_upload(self.handle, _ptr(dst), _ptr(src), rows, d, ld_src), "pmcb200_upload")

    def mixture_propose(self, n, d, k, means, chol, dofs, starts, seed, index0, x, ldx, latent=None, stream=0):
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        _check(load().pmcb200_mixture_propose(self.handle, n, d, k, _ptr(means), _ptr(chol), _ptr(dofs), starts.ctypes.data,
                                              int(seed), int(index0), _ptr(x), ldx, _ptr(latent), stream),
               "pmcb200_mixture_propose")

    def importance_weights(self, log_target, logq, n, w, sums, stream=0):
        _check(load().pmcb200_importance_weights(self.handle, _ptr(log_target), _ptr(logq), n, _ptr(w), _ptr(sums), stream),
               "pmcb200_importance_weights")

    def fp64_peak(self, which=0, iters=4000):
        g, ms = ctypes.c_double(), ctypes.c_double()
        _check(load().pmcb200_fp64_peak(self.handle, which, iters, ctypes.byref(g), ctypes.byref(ms)), "pmcb200_fp64_peak")
        return g.value, ms.value

    def last_k1_kernel(self) -> str:
        """Name of the K1 kernel that did the work in the last ``mixture_eval`` on this context (reads the device flags)."""
        buf = ctypes.create_string_buffer(128)
        _check(load().pmcb200_last_k1_kernel(self.handle, buf, 128), "pmcb200_last_k1_kernel")
        return buf.value.decode()

    def launch_count(self) -> int:
        return int(load().pmcb200_launch_count(self.handle))


_default_device = None


def default_device() -> int:
    """cuda device index of this process: LOCAL_RANK under torchrun, else torch's current device."""
    global _default_device
    if _default_device is None:
        import torch
        if not torch.cuda.is_available():
            raise PmcB200Error("pypmc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if "LOCAL_RANK" in os.environ:
            torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        _default_device = torch.cuda.current_device()
    return _default_device


def set_default_device(device: int):
    global _default_device
    _default_device = int(device)
