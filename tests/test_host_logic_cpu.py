"""CPU tier: the K-sized host arithmetic behind the kernels (statistics packet, moment finishing, component
bookkeeping, dof solver) and the sample-sharded path on two ``gloo`` ranks -- no GPU needed.

The per-rank statistics packet that K1/K2 produce on the device is built here with numpy from the oracle's
responsibilities, so what is tested is exactly what runs on the host after the kernels: the all-reduce, the
packet layout and the finishing formulas, against the oracle evaluated on the unsharded data."""
import os
import socket

import numpy as np
import pytest

from conftest import mat_err
from oracle import oracle as orc


def _synth(K, D, N, seed):
    rng = np.random.default_rng(seed)
    means = rng.normal(0, 3, size=(K, D))
    covs = np.array([(lambda a: a @ a.T + 0.5 * np.eye(D))(rng.normal(0, D ** -0.5, size=(D, D))) for _ in range(K)])
    w = rng.uniform(0.5, 1.5, size=K)
    comp = rng.integers(0, K, size=N)
    x = means[comp] + np.einsum("nij,nj->ni", np.linalg.cholesky(covs)[comp], rng.normal(size=(N, D)))
    return means, covs, w / w.sum(), np.ascontiguousarray(x), rng.uniform(0.5, 1.5, size=N)


def numpy_packet(lay, x, rho, sw, shift, logq, gamma=None):
    """What K1 (sums) + K2 (statistics rows) leave in the packet for one rank's rows."""
    K, D = lay.K, lay.D
    u = rho * sw[:, None]
    v = u if gamma is None else u * gamma
    y = x - shift
    il = np.tril_indices(D)
    pkt = np.zeros(lay.size)
    rows = pkt[:lay.stats_len].reshape(K, lay.row)
    rows[:, 0] = u.sum(0)
    rows[:, 1] = v.sum(0)
    rows[:, 2:2 + D] = v.T @ y
    rows[:, 2 + D:2 + D + lay.T] = np.einsum("nk,ni,nj->kij", v, y, y)[:, il[0], il[1]]
    rows[:, -1] = 0.0 if gamma is None else (u * np.log(gamma)).sum(0)
    pkt[lay.off_sum_a] = (sw * logq).sum()
    pkt[lay.off_sumw] = sw.sum()
    return pkt


def test_packet_roundtrip_and_moments_match_oracle():
    from pypmc_b200.mix_adapt._stats import PacketLayout, moments_from_stats
    K, D, N = 4, 5, 3000
    means, covs, w, x, sw = _synth(K, D, N, seed=3)
    comps = orc.Components(means, covs)
    rho, logq = orc.calculate_rho_rb(x, comps, w)
    lay = PacketLayout(K, D)
    shift = (w[:, None] * means).sum(0)
    st = lay.unpack(numpy_packet(lay, x, rho, sw, shift, logq))
    A, mean, cov = moments_from_stats(st, shift, "B")
    alpha_ref, mu_ref, cov_ref = orc.pmc_moments(x, rho, sw)
    np.testing.assert_allclose(A / st["sumw"], alpha_ref, rtol=1e-12)
    np.testing.assert_allclose(mean, mu_ref, rtol=1e-11, atol=1e-13)
    assert mat_err(cov, cov_ref) < 1e-11
    assert st["sum_a"] == pytest.approx(float((sw * logq).sum()), rel=1e-14)


def test_kill_undersampled_reproduces_reference_traversal():
    # pmc.pyx:110-112 removes from the list it iterates: the element after a removed one is skipped
    from pypmc_b200.mix_adapt.pmc import _kill_undersampled
    from pypmc_b200.density.mixture import create_gaussian_mixture
    mix = create_gaussian_mixture(np.zeros((4, 1)), np.ones((4, 1, 1)))
    live = [0, 1, 2, 3]
    counts = np.array([0, 0, 5, 0])
    assert _kill_undersampled(mix, live, counts, mincount=2)
    # reference traversal: 0 dies, 1 is skipped, 2 survives, 3 dies
    ref_live = [0, 1, 2, 3]
    for k in ref_live:
        if counts[k] < 2:
            ref_live.remove(k)
    assert live == ref_live == [1, 2]
    assert (mix.weights[[0, 3]] == 0).all() and (mix.weights[[1, 2]] != 0).all()


def test_dof_condition_matches_reference_formula():
    # pmc.pyx:478-497: const + ln(nu/2) - psi(nu/2); decreasing in nu
    from scipy.special import digamma
    from pypmc_b200.mix_adapt.pmc import _DOFCondition
    c = _DOFCondition(-0.3)
    for nu in (0.5, 3.0, 40.0):
        assert c(nu) == pytest.approx(-0.3 + np.log(0.5 * nu) - digamma(0.5 * nu), rel=1e-15)
    assert c(1.0) > c(2.0) > c(50.0)


def test_shard_rows_partition():
    from pypmc_b200 import parallel
    for n, world in ((10, 3), (7, 8), (1000, 4), (0, 2)):
        spans = [parallel.shard_rows(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pypmc_b200 import parallel
    from pypmc_b200.mix_adapt._stats import PacketLayout, moments_from_stats
    r, ws = parallel.init_from_env(backend="gloo")
    assert (r, ws) == (rank, world) and parallel.enabled()
    K, D, N = 3, 4, 2001
    means, covs, w, x, sw = _synth(K, D, N, seed=9)          # every rank builds the same data, keeps its shard
    lo, hi = parallel.shard_rows(N)
    comps = orc.Components(means, covs)
    rho, logq = orc.calculate_rho_rb(x[lo:hi], comps, w)
    lay = PacketLayout(K, D)
    shift = (w[:, None] * means).sum(0)
    pkt = torch.from_numpy(numpy_packet(lay, x[lo:hi], rho, sw[lo:hi], shift, logq))
    parallel.allreduce_(pkt)                                  # ONE all-reduce of the K-row packet
    st = lay.unpack(pkt.numpy())
    A, mean, cov = moments_from_stats(st, shift, "B")
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), alpha=A / st["sumw"], mean=mean, cov=cov,
             loglik=st["sum_a"] / st["sumw"], lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_update_matches_unsharded_oracle(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert int(res[0]["lo"]) == 0 and int(res[0]["hi"]) == int(res[1]["lo"]) and int(res[1]["hi"]) == 2001
    for key in ("alpha", "mean", "cov", "loglik"):           # identical bits on every rank: no broadcast needed
        np.testing.assert_array_equal(res[0][key], res[1][key])
    K, D, N = 3, 4, 2001
    means, covs, w, x, sw = _synth(K, D, N, seed=9)
    comps = orc.Components(means, covs)
    rho, logq = orc.calculate_rho_rb(x, comps, w)
    alpha_ref, mu_ref, cov_ref = orc.pmc_moments(x, rho, sw)
    np.testing.assert_allclose(res[0]["alpha"], alpha_ref, rtol=1e-12)
    np.testing.assert_allclose(res[0]["mean"], mu_ref, rtol=1e-11, atol=1e-13)
    assert mat_err(res[0]["cov"], cov_ref) < 1e-11
    assert float(res[0]["loglik"]) == pytest.approx(float((sw * logq).sum() / sw.sum()), rel=1e-13)


def test_history_semantics():
    # pypmc/tools/_history.py:7-116 (docstring example, views, negative indices, growth, clear)
    from pypmc_b200.tools._history import History
    h = History(2)
    for i in range(2):
        a = h.append(i + 1)
        a[:] = i + 1
    np.testing.assert_array_equal(h[0], [[1.0, 1.0]])
    np.testing.assert_array_equal(h[1], [[2.0, 2.0], [2.0, 2.0]])
    np.testing.assert_array_equal(h[:], [[1.0, 1.0], [2.0, 2.0], [2.0, 2.0]])
    np.testing.assert_array_equal(h[-1], h[1])
    assert len(h) == 2
    h[0][0, 0] = 7.0                       # index access returns a reference
    assert h[:][0, 0] == 7.0
    big = h.append(1000)
    big[:] = 3.0
    assert h[:].shape == (1003, 2) and h[0:2].shape == (3, 2) and (h[2] == 3.0).all() and h[:][0, 0] == 7.0
    with pytest.raises(NotImplementedError):
        h[::2]
    with pytest.raises(AssertionError):
        h.append(0)
    h.clear()
    assert len(h) == 0 and h[:].size == 0


def test_combine_weights_argument_checks():
    from pypmc_b200.sampler.importance_sampling import combine_weights
    x = [np.zeros((3, 2)), np.zeros((2, 2))]
    with pytest.raises(AssertionError, match="importance-sampling runs but 1 weights"):
        combine_weights(x, [np.ones(3)], [None, None])
    with pytest.raises(AssertionError, match="proposal densities"):
        combine_weights(x, [np.ones(3), np.ones(2)], [None])
    with pytest.raises(AssertionError, match="Length of weights"):
        combine_weights(x, [np.ones(3), np.ones(3)], [None, None])


def test_shift_groups():
    """One shift for the whole mixture unless components are far apart in units of their own width (then K2 runs
    once per group); deterministic and covering every live component exactly once."""
    from pypmc_b200.mix_adapt._stats import shift_groups, SHIFT_CONDITION_LIMIT
    eye = np.eye(2)
    near = shift_groups([np.zeros(2), np.array([3.0, 1.0])], [eye, eye], [0.5, 0.5], [0, 1])
    assert len(near) == 1 and near[0][0] == [0, 1]
    np.testing.assert_allclose(near[0][1], [1.5, 0.5])
    mus = [np.zeros(2), np.array([3.0e4, 1.0]), np.array([1.0, 0.0]), np.array([3.0e4, 2.0]), np.array([-5e5, 0.0])]
    prec = [eye * 1e2] * 5
    far = shift_groups(mus, prec, [0.2] * 5, [0, 1, 2, 3, 4])
    assert sorted(sum((g[0] for g in far), [])) == [0, 1, 2, 3, 4]
    assert [g[0] for g in far] == [[0, 2], [1, 3], [4]]
    for idx, c in far:
        assert max(float((mus[k] - c) @ prec[k] @ (mus[k] - c)) for k in idx) <= SHIFT_CONDITION_LIMIT
    assert shift_groups(mus, prec, [0.2] * 5, [1, 3]) == [] or len(shift_groups(mus, prec, [0.2] * 5, [1, 3])) == 1
    assert shift_groups(mus, prec, [0.2] * 5, []) == []


def test_perp_and_ess_reference_values():
    # pypmc/tools/convergence.py:6-72: equal weights are perfect, one dominant weight is terrible, zeros are ignored
    from pypmc_b200.tools.convergence import perp, ess
    assert perp(np.ones(10)) == pytest.approx(1.0) and ess(np.ones(10)) == pytest.approx(1.0)
    w = np.array([1.0, 0.0, 0.0, 0.0])
    assert perp(w) == pytest.approx(0.25) and ess(w) == pytest.approx(0.25)
    w = np.array([0.2, 0.5, 0.1, 0.0, 1.7])
    wn = w / w.sum()
    assert perp(w) == pytest.approx(np.exp(-np.sum(wn[wn > 0] * np.log(wn[wn > 0]))) / 5)
    assert ess(w) == pytest.approx(1.0 / (1.0 + np.mean((5 * wn - 1) ** 2)))
