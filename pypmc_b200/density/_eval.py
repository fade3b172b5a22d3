"""Single entry to kernel K1 for the density / mix_adapt classes: host (numpy) or device (torch) samples."""
from __future__ import annotations

import numpy as np

from .. import _device as dev
from .. import _lib


def _host_out(a, shape, name):
    if a is None:
        return None
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == shape):
        raise ValueError("%s must be a C-contiguous float64 array of shape %s" % (name, shape))
    return a


def run_k1(x, packed: dev.PackedComponents, k_out, mode, max_init=-_lib.DBL_MAX, logq=None, lp=None, resp=None,
           aux=None, weights=None, want_sums=False, sums=None):
    """Launch K1 on ``x`` ([N, D] numpy array or torch CUDA tensor).  Outputs that are not None are filled in
    place (numpy arrays for host samples, CUDA tensors for device samples).  Returns ``sums`` (2 floats:
    sum_n w_n log q_n [VB: sum_n w_n sum_k r log r], sum_n w_n) when ``want_sums`` (or into the given
    2-element ``sums`` buffer) else None."""
    n, d = x.shape
    if dev.is_device_tensor(x):
        t = dev.torch()
        index = x.device.index                              # context and stream of the device the samples live on
        ctx = _lib.Context.get(index)
        if n > 0 and x.stride(1) != 1:
            x = x.contiguous()
        ldx = x.stride(0) if n > 1 else d
        rec_d, cols_d = packed.device(index)
        if sums is None and want_sums:
            sums = t.zeros(2, dtype=t.float64, device=x.device)
        w = dev.to_device(weights, index) if weights is not None else None
        for o in (logq, lp, resp, aux, w, sums):
            assert o is None or (dev.is_device_tensor(o) and o.is_contiguous() and o.dtype == t.float64)
            assert o is None or o.device == x.device, "all tensors of one K1 launch must live on the samples' device"
        ctx.mixture_eval(x, n, ldx, d, rec_d, cols_d, packed.kl, k_out, mode, max_init, logq, lp, resp, aux, w, sums,
                         dev.current_stream_ptr(index))
        return None if sums is None else sums
    ctx = _lib.Context.get()
    xh, ldx = dev.row_major(x)
    logq = _host_out(logq, (n,), "out")
    lp = _host_out(lp, (n, k_out), "individual")
    resp = _host_out(resp, (n, k_out), "resp")
    aux = _host_out(aux, (n, k_out), "aux")
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    if sums is None and want_sums:
        sums = np.zeros(2)
    ctx.mixture_eval_host(xh, n, ldx, d, packed.records, packed.cols, packed.kl, k_out, mode, max_init, logq, lp, resp,
                          aux, w, sums)
    return sums
