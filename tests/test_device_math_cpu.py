"""CPU emulation of the elementary-function shortcuts in the CUDA kernels, constant for constant: the table logarithm of
k2_colsums / K3 (`k2_log_pos`, k2_suffstats.cuh; `k3_normal2`, k3_propose.cuh) and K3's own Box-Muller.  The tables are
parsed from csrc/k1_exp_table.cuh, the polynomial coefficients from the kernel sources, so a typo in either fails here
without a GPU.  (FMAs are emulated by separate multiply and add: the bounds below leave room for that.)"""
import os
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pypmc_b200", "csrc")
H = float.fromhex


def _table(name):
    src = open(os.path.join(CSRC, "k1_exp_table.cuh")).read()
    m = re.search(name + r"\[\d+\] = \{(.*?)\};", src, re.S)
    return np.array([H(v.strip()) for v in m.group(1).split(",") if v.strip()])


def _hex_constants(fname, func):
    """Hexadecimal floating-point literals inside device function ``func`` of ``fname``, in source order."""
    src = open(os.path.join(CSRC, fname)).read()
    body = src[src.index(func):]
    body = body[:body.index("\n}\n")]
    return [H(v) for v in re.findall(r"-?0x1\.[0-9a-f]+p[+-]\d+", body)]


def _log_pos(x, invc, lnc, ln2hi, ln2lo):
    bits = x.view(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.int64)
    j = (hi >> 13) & 127
    e = ((hi >> 20) - 1023).astype(float)
    m = ((((hi & 0x000fffff) | 0x3ff00000).astype(np.uint64) << np.uint64(32)) | (bits & np.uint64(0xffffffff))).view(np.float64)
    r = m * invc[j] - 1.0
    u = r * (-1.66666666666666657e-01) + 2.00000000000000011e-01
    u = r * u - 0.25
    u = r * u + 3.33333333333333315e-01
    u = r * u - 0.5
    p = (r * r) * u + r
    return e * ln2hi + ((p + lnc[j]) + e * ln2lo)


def test_table_logarithm_any_positive_normal():
    invc, lnc = _table("kLogInvC"), _table("kLogC")
    assert len(invc) == 128 and len(lnc) == 128
    c = 1.0 + (np.arange(128) + 0.5) / 128.0
    np.testing.assert_allclose(invc, 1.0 / c, rtol=2e-16)
    np.testing.assert_allclose(lnc, np.log(c), rtol=3e-16, atol=1e-18)
    consts = _hex_constants("k2_suffstats.cuh", "double k2_log_pos(")
    ln2hi, ln2lo = consts[-2], consts[-1]
    assert abs((ln2hi + ln2lo) - np.log(2.0)) < 1e-16 and ln2hi == H("0x1.62e42fefa38p-1")   # 44-bit head: e * hi is exact
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(np.log(1e-300), np.log(1e300), size=400_000))
    x = np.concatenate([x, rng.uniform(0.5, 2.0, size=200_000), 1.0 + rng.uniform(-1e-6, 1e-6, size=1000)])
    got, ref = _log_pos(x, invc, lnc, ln2hi, ln2lo), np.log(x)
    assert np.max(np.abs(got - ref) / (1.0 + np.abs(ref))) < 4e-16


def test_k3_box_muller_matches_library_on_the_same_bits():
    invc, lnc = _table("kLogInvC"), _table("kLogC")
    k = _hex_constants("k3_propose.cuh", "double2 k3_normal2(")
    # source order: ln2 head, ln2 tail, pi/2, 7 sine coefficients (highest first), 7 cosine coefficients (highest first; -0.5 is decimal)
    ln2hi, ln2lo, half_pi = k[0], k[1], k[2]
    sin_c, cos_c = k[3:10], k[10:17] + [-0.5]
    assert half_pi == np.pi / 2 and len(k) == 17
    from math import factorial
    np.testing.assert_allclose(sin_c[::-1], [(-1) ** (i + 1) / factorial(2 * i + 3) for i in range(7)], rtol=1e-15)
    np.testing.assert_allclose(cos_c[::-1], [(-1) ** (i + 1) / factorial(2 * i + 2) for i in range(8)], rtol=1e-15)
    rng = np.random.default_rng(1)
    n = 500_000
    w = rng.integers(0, 2 ** 32, size=(n, 4), dtype=np.uint64)

    def unit(hi32, lo32):                                       # [1, 2) from 52 random bits (exponent trick)
        return (((np.uint64(0x3ff00000) | (hi32 >> np.uint64(12))) << np.uint64(32)) | lo32).view(np.float64)

    d1, d2 = unit(w[:, 0], w[:, 1]), unit(w[:, 2], w[:, 3])
    u1 = 2.0 - d1
    assert u1.min() > 0.0 and u1.max() <= 1.0
    rad = np.sqrt(-2.0 * _log_pos(u1, invc, lnc, ln2hi, ln2lo))
    a = d2 * 4.0 - 4.0
    magic = 6755399441055744.0
    t = a + magic
    q = (t.view(np.uint64) & np.uint64(0xffffffff)).astype(np.int64)
    x = (a - (t - magic)) * half_pi
    assert np.abs(x).max() <= np.pi / 4 + 1e-15 and q.min() >= 0 and q.max() <= 4
    x2 = x * x
    sp = x2 * sin_c[0] + sin_c[1]
    for c in sin_c[2:]:
        sp = x2 * sp + c
    sn = (x * x2) * sp + x
    cp = x2 * cos_c[0] + cos_c[1]
    for c in cos_c[2:]:
        cp = x2 * cp + c
    cs = x2 * cp + 1.0
    c0 = np.where(q & 1, sn, cs)
    s0 = np.where(q & 1, cs, sn)
    c0 = np.where((q + 1) & 2, -c0, c0)
    s0 = np.where(q & 2, -s0, s0)
    u2 = d2 - 1.0
    assert np.max(np.abs(c0 - np.cos(2 * np.pi * u2))) < 2e-15 and np.max(np.abs(s0 - np.sin(2 * np.pi * u2))) < 2e-15
    z = np.concatenate([rad * c0, rad * s0])
    ref = np.sqrt(-2.0 * np.log(u1))
    zr = np.concatenate([ref * np.cos(2 * np.pi * u2), ref * np.sin(2 * np.pi * u2)])
    assert np.max(np.abs(z - zr)) < 1e-13
    assert abs(z.mean()) < 5.0 / np.sqrt(len(z)) and abs(z.var() - 1.0) < 6.0 * np.sqrt(2.0 / len(z))


def test_vb_logarithm_next_to_one_is_relative():
    """`log_near1` (k1_mma_eval.cuh): ln(sum) for the VB soft-max, where sum = 1 + eps for a sample with one responsible
    component and ln r of that component is -ln(sum): the series below 2^-7 keeps RELATIVE accuracy (the table form's
    ~1e-16 absolute error, systematic next to 1, was worth 2e-12 in sum_n r ln r at N = 1.2e5)."""
    src = open(os.path.join(CSRC, "k1_mma_eval.cuh")).read()
    body = src[src.index("double log_near1("):]
    body = body[:body.index("return log_any")]
    coef = [float(v) for v in re.findall(r"-?\d\.\d+e[+-]\d+|-0\.25|-0\.5", body)]
    np.testing.assert_allclose(coef, [1 / 7, -1 / 6, 1 / 5, -0.25, 1 / 3, -0.5], rtol=1e-15)
    rng = np.random.default_rng(2)
    s = np.concatenate([np.exp(rng.uniform(np.log(1e-18), np.log(2.0 ** -7), size=200_000)), [0.0, 2.0 ** -7 * (1 - 1e-16)]])
    x = 1.0 + s
    s = x - 1.0                                                  # what the kernel sees (exact for x in [1, 2))
    u = s * coef[0] + coef[1]
    for c in coef[2:]:
        u = s * u + c
    got = (s * s) * u + s
    ref = np.log1p(s)
    ok = ref > 0
    assert np.max(np.abs(got[ok] - ref[ok]) / ref[ok]) < 5e-16
    assert (got[s == 0.0] == 0.0).all()
