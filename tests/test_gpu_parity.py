"""Parity of the CUDA path (through the C ABI, via the pypmc-compatible classes) against

(a) golden fixtures generated from the compiled, unmodified reference (tests/golden/*.npz),
(b) the reference's own hand-computed golden numbers,
(c) the CPU oracle on seeded inputs incl. edge cases (ragged N, odd D, dead components, strided views).

Tolerance (BASELINE.json north_star): 1e-10 relative, float64.  Covariance-type outputs use the
max-norm metric of SURVEY 8c (``mat_err``).  Needs a B200: ``pytest -m gpu``.
"""
import os
import numpy as np
import pytest

from conftest import exp_err, log_err, mat_err, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def pm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pypmc_b200
    return pypmc_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _mix(pm, g, t=False):
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    if t:
        return create_t_mixture(g["means"], g["covs"], g["dofs"], g["weights"])
    return create_gaussian_mixture(g["means"], g["covs"], g["weights"])


# ------------------------------------------------------------------ (b) reference golden numbers
def test_reference_golden_numbers(pm):
    from pypmc_b200.density.gauss import Gauss
    from pypmc_b200.density.student_t import StudentT
    from pypmc_b200.density.mixture import MixtureDensity
    # density/gauss_test.py:43-55, 111-160
    g = Gauss([4.3, 1.1], [[0.01, 0.003], [0.003, 0.0025]])
    # (the reference prints these golden numbers with 9-10 significant digits: the absolute tolerances below are theirs)
    assert g.evaluate(np.array([4.35, 1.2])) == pytest.approx(1.30077135, abs=1e-8)
    out = np.empty(2)
    res = g.multi_evaluate(np.array([[4.35, 1.2]] * 2), out)
    assert res is out
    np.testing.assert_allclose(out, 1.30077135, atol=1e-8)
    # density/student_t_test.py:168-230, 155-166
    t = StudentT([1.25, 4.3], [[0.0049, 0.0], [0.0, 0.01]], 5.0)
    np.testing.assert_allclose(t.multi_evaluate(np.array([[1.3, 4.4], [1.26, 4.424]])), [2.200202941, 2.174596526], atol=1e-9)
    cauchy = StudentT([0.0], [[1.0]], 1.0)
    assert cauchy.evaluate(np.array([3.2])) == pytest.approx(-3.5642087303149452, abs=1e-12)
    # density/mixture_test.py:29-33: two unit Gaussians in 1-d
    mix = MixtureDensity([Gauss([0.0], [[1.0]]), Gauss([1.0], [[1.0]])], [0.4, 0.6])
    x = np.array([[0.3]])
    expect = np.log(0.4 * np.exp(-0.5 * 0.09) + 0.6 * np.exp(-0.5 * 0.49)) - 0.5 * np.log(2 * np.pi)
    assert mix.multi_evaluate(x)[0] == pytest.approx(expect, rel=1e-14)
    assert mix.evaluate(x[0]) == pytest.approx(expect, rel=1e-14)


def test_gaussian_pmc_reference_tables(pm):
    # mix_adapt/pmc_test.py:93-169 (hand-computed tables, assert_allclose default rtol 1e-7)
    from test_oracle import PMC_MEANS, PMC_COVS, PMC_CW, PMC_SAMPLES, PMC_WEIGHTS
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc
    prop = create_gaussian_mixture(PMC_MEANS, PMC_COVS, PMC_CW)
    before = [c.mu.copy() for c in prop.components]
    new = gaussian_pmc(PMC_SAMPLES, prop, PMC_WEIGHTS)
    for c, b in zip(prop.components, before):      # copy=True leaves the input untouched (pmc_test.py:69-91)
        np.testing.assert_array_equal(c.mu, b)
    np.testing.assert_allclose(new.weights, np.array([154.54358983999998, 159.02061223999999]) / 313.56420207999997)
    np.testing.assert_allclose(new.components[0].mu, np.array([1546.302278, -172.1300429, 1279.34733595]) / 154.54358983999998)
    np.testing.assert_allclose(new.components[1].mu, np.array([-1652.19922509, 1150.52591727, 74.098254]) / 159.02061223999999)
    np.testing.assert_allclose(new.components[0].sigma, np.array([[91.13245238, 62.95055712, 4.96175291],
                                                                  [62.95055712, 51.04895641, -16.59026473],
                                                                  [4.96175291, -16.59026473, 111.63047879]]) / 154.54358983999998)
    new = gaussian_pmc(PMC_SAMPLES, prop)          # unweighted
    np.testing.assert_allclose(new.weights, [0.6, 0.4])
    np.testing.assert_allclose([c.mu for c in new.components],
                               [[10.00569514, -1.11356905, 8.27908553], [-10.38983286, 7.23392486, 0.4632788]])
    with pytest.raises(ValueError, match="mincount"):
        gaussian_pmc(PMC_SAMPLES, prop, mincount=2)
    with pytest.raises(ValueError, match="rb"):
        gaussian_pmc(PMC_SAMPLES, prop, rb=False)


# ------------------------------------------------------------------ (a) fixtures from the compiled reference
@pytest.mark.parametrize("name", ["gauss_small", "gauss_c2", "gauss_c2_stress", "gauss_illcond"])
def test_gauss_mixture_fixture(pm, golden, name):
    import torch
    g = golden(name)
    mix = _mix(pm, g)
    x = g["x"]
    n, k = len(x), len(mix)
    ind = np.empty((n, k))
    logq = mix.multi_evaluate(x, individual=ind)
    rows = len(g["individual"])
    assert rel_err(ind[:rows], g["individual"]) < TOL
    assert rel_err(logq, g["logq"]) < TOL
    # the same bits whichever of out / individual is passed (mixture_test.py:92-96)
    out = np.empty(n)
    assert mix.multi_evaluate(x, out) is out
    np.testing.assert_array_equal(out, logq)
    np.testing.assert_array_equal(mix.multi_evaluate(x), logq)
    # device-resident samples give the same bits as streamed host samples
    xd = torch.from_numpy(x).cuda()
    indd = torch.empty((n, k), dtype=torch.float64, device="cuda")
    lq = mix.multi_evaluate(xd, individual=indd)
    np.testing.assert_array_equal(lq.cpu().numpy(), logq)
    np.testing.assert_array_equal(indd.cpu().numpy(), ind)
    # components subset only touches its columns (mixture.pyx:153-156)
    ind2 = np.full((n, k), -7.0)
    assert mix.multi_evaluate(x, individual=ind2, components=[1, k - 1]) is None
    # (the shift c of the fast K1 form is the weighted centre of the EVALUATED components, so a subset agrees
    # with the full evaluation to rounding, not bit for bit; the reference only pins bitwise equality across the
    # out / individual variants above)
    # (kappa = 1e5 in gauss_illcond: the two shifts differ by ~80 units, T has entries of ~300 -- rounding-level there is 1e-11)
    np.testing.assert_allclose(ind2[:, [1, k - 1]], ind[:, [1, k - 1]], rtol=TOL if name == "gauss_illcond" else 1e-13, atol=0)
    assert (np.delete(ind2, [1, k - 1], axis=1) == -7.0).all()


@pytest.mark.parametrize("name", ["gauss_small", "gauss_c2", "gauss_c2_stress", "gauss_illcond"])
def test_gaussian_pmc_fixture(pm, golden, name):
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, PMC
    g = golden(name)
    mix = _mix(pm, g)
    live = [k for k in range(len(mix)) if g["weights"][k] != 0]
    variants = [("pmc_weighted", dict(weights=g["sample_weights"])), ("pmc_unweighted", dict())]
    if "pmc_latent_rb_weights" in g:
        variants += [("pmc_latent_rb", dict(weights=g["sample_weights"], latent=g["latent"], rb=True, mincount=2)),
                     ("pmc_latent_nonrb", dict(weights=g["sample_weights"], latent=g["latent"], rb=False))]
    for tag, kw in variants:
        new = gaussian_pmc(g["x"], mix, **kw)
        np.testing.assert_allclose(new.weights, g[tag + "_weights"], rtol=TOL, atol=1e-300)
        mu = np.array([c.mu for c in new.components])
        cov = np.array([c.sigma for c in new.components])
        np.testing.assert_allclose(mu[live], g[tag + "_means"][live], rtol=TOL, atol=1e-12)
        assert mat_err(cov[live], g[tag + "_covs"][live]) < TOL
    p = PMC(g["x"], mix, weights=g["sample_weights"])
    assert p.log_likelihood() == pytest.approx(float(g["loglik_weighted"]), rel=TOL)
    assert PMC(g["x"], mix).log_likelihood() == pytest.approx(float(g["loglik_unweighted"]), rel=TOL)
    if "pmc_run3_weights" in g:
        conv = p.run(iterations=3)
        assert (-1 if conv is None else conv) == int(g["pmc_run3_converged"])
        np.testing.assert_allclose(p.density.weights, g["pmc_run3_weights"], rtol=TOL)
        assert mat_err(np.array([c.sigma for c in p.density.components]), g["pmc_run3_covs"]) < TOL
        assert p.log_likelihood() == pytest.approx(float(g["pmc_run3_loglik"]), rel=TOL)


@pytest.mark.parametrize("name", ["student_small", "student_c4"])
def test_student_fixture(pm, golden, name):
    from pypmc_b200.mix_adapt.pmc import student_t_pmc
    g = golden(name)
    mix = _mix(pm, g, t=True)
    x = g["x"]
    ind = np.empty((len(x), len(mix)))
    logq = mix.multi_evaluate(x, individual=ind)
    rows = len(g["individual"])
    assert rel_err(ind[:rows], g["individual"]) < TOL
    assert rel_err(logq, g["logq"]) < TOL
    variants = [("pmc_nodof_weighted", dict(weights=g["sample_weights"], dof_solver_steps=0)),
                ("pmc_dof_weighted", dict(weights=g["sample_weights"]))]
    if "pmc_dof_unweighted_weights" in g:
        variants += [("pmc_nodof_unweighted", dict(dof_solver_steps=0)), ("pmc_dof_unweighted", dict()),
                     ("pmc_dof_latent_nonrb", dict(weights=g["sample_weights"], latent=g["latent"], rb=False))]
    for tag, kw in variants:
        new = student_t_pmc(x, mix, **kw)
        np.testing.assert_allclose(new.weights, g[tag + "_weights"], rtol=TOL)
        np.testing.assert_allclose([c.mu for c in new.components], g[tag + "_means"], rtol=TOL, atol=1e-12)
        assert mat_err(np.array([c.sigma for c in new.components]), g[tag + "_covs"]) < TOL
        np.testing.assert_allclose([c.dof for c in new.components], g[tag + "_dofs"], rtol=TOL)


@pytest.mark.parametrize("name", ["vb_small", "vb_c3"])
def test_vb_fixture(pm, golden, name):
    from pypmc_b200.mix_adapt.variational import GaussianInference
    from pypmc_b200.density.mixture import create_gaussian_mixture
    g = golden(name)
    mix = create_gaussian_mixture(g["means"], g["covs"], g["weights"])
    tags = [("unw", None)] + ([("wgt", g["sample_weights"])] if "wgt_init_r" in g else [])
    for tag, sw in tags:
        vb = GaussianInference(g["x"], initial_guess=mix, weights=sw)

        def check(stage):
            p = lambda a: g["%s_%s_%s" % (tag, stage, a)]
            rows = len(p("r"))
            for a in ("alpha", "beta", "nu", "expectation_det_ln_lambda", "expectation_ln_pi", "N_comp"):
                np.testing.assert_allclose(getattr(vb, a), p(a), rtol=TOL, err_msg=a)
            np.testing.assert_allclose(vb.m, p("m"), rtol=TOL, atol=1e-12)
            assert mat_err(vb.W, p("W")) < TOL
            assert rel_err(vb.expectation_gauss_exponent[:rows], p("expectation_gauss_exponent")) < TOL
            assert rel_err(vb.log_rho[:rows], p("log_rho")) < TOL
            # r = exp(log_rho): after an M-step of our own (W, m agree with the reference to ~1e-14) an entry r = 1e-235
            # inherits |ln r| = 540 times the relative difference of ln r -- bound stated in conftest.exp_err; entries
            # with |ln r| <= 1 (the ones that carry the statistics) are held to 1e-10 relative
            assert exp_err(vb.r[:rows], p("r")) < TOL
            big = p("r") > np.exp(-1.0)
            assert rel_err(vb.r[:rows][big], p("r")[big]) < TOL
            np.testing.assert_allclose(vb.x_mean_comp, p("x_mean_comp"), rtol=TOL, atol=1e-12)
            assert mat_err(vb.S, p("S")) < TOL
            np.testing.assert_allclose(vb.inv_N_comp, p("inv_N_comp"), rtol=TOL)

        check("init")
        assert vb.likelihood_bound() == pytest.approx(float(g[tag + "_init_bound"]), rel=TOL)
        vb.update()
        check("upd1")
        assert vb.likelihood_bound() == pytest.approx(float(g[tag + "_upd1_bound"]), rel=TOL)
        vb.update()
        assert vb.likelihood_bound() == pytest.approx(float(g[tag + "_upd2_bound"]), rel=TOL)
        out = vb.make_mixture()
        np.testing.assert_allclose(out.weights, g[tag + "_upd2_mix_weights"], rtol=TOL)
        assert mat_err(np.array([c.sigma for c in out.components]), g[tag + "_upd2_mix_covs"]) < TOL
    if "first_run5_K" in g:
        vb = GaussianInference(g["x"], components=len(mix) + 2)
        assert vb.likelihood_bound() == pytest.approx(float(g["first_init_bound"]), rel=TOL)
        it = vb.run(iterations=5, prune=1.0)
        assert (-1 if it is None else it) == int(g["first_run5_converged"])
        assert vb.K == int(g["first_run5_K"])
        np.testing.assert_allclose(vb.N_comp, g["first_run5_N_comp"], rtol=TOL)
        assert vb.likelihood_bound() == pytest.approx(float(g["first_run5_bound"]), rel=TOL)


# ------------------------------------------------------------------ (c) oracle on seeded edge cases
def _synth(K, D, N, seed, dof=None):
    rng = np.random.default_rng(seed)
    means = rng.normal(0.0, 3.0, size=(K, D))
    covs = np.empty((K, D, D))
    for k in range(K):
        a = rng.normal(0.0, 1.0 / np.sqrt(D), size=(D, D))
        covs[k] = a @ a.T + 0.5 * np.eye(D)
    w = rng.uniform(0.5, 1.5, size=K)
    comp = rng.integers(0, K, size=N)
    x = means[comp] + np.einsum("nij,nj->ni", np.linalg.cholesky(covs)[comp], rng.normal(size=(N, D)))
    if dof is not None:
        x = means[comp] + (x - means[comp]) / np.sqrt(rng.chisquare(dof, size=N) / dof)[:, None]
    return means, covs, w / w.sum(), np.ascontiguousarray(x), rng.uniform(0.5, 1.5, size=N)


@pytest.mark.parametrize("K,D,N", [(1, 1, 1), (3, 2, 31), (2, 3, 33), (5, 7, 1537), (4, 13, 1000), (7, 20, 2049),
                                   (3, 21, 777), (6, 33, 513), (2, 40, 300), (3, 41, 129), (2, 63, 200), (2, 64, 65)])
def test_k1_vs_oracle_shapes(pm, orc, K, D, N):
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    means, covs, w, x, sw = _synth(K, D, N, seed=100 + D)
    for dofs in (None, np.linspace(2.5, 9.0, K)):
        comps = orc.Components(means, covs, dofs)
        lq_ref, ind_ref = orc.mixture_multi_evaluate(x, comps, w)
        mix = create_gaussian_mixture(means, covs, w) if dofs is None else create_t_mixture(means, covs, dofs, w)
        ind = np.empty((N, K))
        lq = mix.multi_evaluate(x, individual=ind)
        assert rel_err(ind, ind_ref) < TOL
        assert rel_err(lq, lq_ref) < TOL
        # strided sample view (every other row, extra trailing columns)
        big = np.zeros((2 * N, D + 3))
        big[::2, :D] = x
        assert rel_err(mix.multi_evaluate(big[::2, :D]), lq_ref) < TOL
        # component-wise API
        col = np.empty((N, K))
        for k, c in enumerate(mix.components):
            c.multi_evaluate(x, col[:, k])
        np.testing.assert_allclose(col, ind, rtol=1e-13, atol=0)   # K = 1 launch: shift = mu_k, so rounding-level only


def test_rho_gamma_vs_oracle_with_dead_components(pm, orc):
    import torch
    from pypmc_b200.density.mixture import create_t_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200 import _lib
    K, D, N = 6, 9, 1201
    means, covs, w, x, sw = _synth(K, D, N, seed=7, dof=4.0)
    w[[1, 4]] = 0.0
    w /= w.sum()
    dofs = np.linspace(3.0, 8.0, K)
    live = [0, 2, 3, 5]
    comps = orc.Components(means, covs, dofs)
    rho_ref, _ = orc.calculate_rho_rb(x, comps, w, live)
    gamma_ref = orc.student_t_gamma(x, comps, live)
    mix = create_t_mixture(means, covs, dofs, w)
    xd = torch.from_numpy(x).cuda()
    rho = torch.zeros((N, K), dtype=torch.float64, device="cuda")
    gam = torch.zeros((N, K), dtype=torch.float64, device="cuda")
    run_k1(xd, mix._packed(live), K, _lib.MODE_STUDENT_T, resp=rho, aux=gam)
    assert rel_err(rho.cpu().numpy(), rho_ref) < TOL
    assert (rho.cpu().numpy()[:, [1, 4]] == 0).all()
    assert rel_err(gam.cpu().numpy(), gamma_ref) < TOL


def test_k2_vs_numpy_moments(pm):
    """K2 alone against a float64 numpy evaluation of the same shifted raw moments."""
    import torch
    from pypmc_b200 import _lib
    rng = np.random.default_rng(3)
    for (K, D, N, use_g, use_w) in [(1, 1, 5, False, False), (3, 2, 77, True, True), (32, 30, 3001, False, True),
                                    (64, 20, 2500, False, False), (16, 40, 1111, True, True), (5, 47, 400, True, False),
                                    (130, 6, 999, False, True),
                                    # component-block counts 3, 5, 6, 7 and two chunks of 7 (k2_inst.cu)
                                    (20, 30, 1500, False, True), (40, 20, 1300, True, True), (48, 30, 1200, False, False),
                                    (50, 9, 800, True, False), (100, 12, 700, False, True), (21, 3, 300, True, True),
                                    # 65..128 components: the producer warps' second load path
                                    (128, 10, 600, True, True), (96, 33, 500, False, False), (72, 20, 900, True, False),
                                    # D > 62: the producers' generic path into the transposed stage; gamma over many
                                    # binades (the table logarithm of k2_colsums, arguments below and above 1)
                                    (7, 70, 333, True, True), (16, 66, 257, False, False)]:
        x = rng.normal(size=(N, D)) + 2.0
        rho = rng.uniform(size=(N, K))
        gam = rng.uniform(0.5, 2.0, size=(N, K)) if use_g else None
        if use_g and D == 70:
            gam = np.exp(rng.uniform(np.log(1e-9), np.log(1e6), size=(N, K)))
        w = rng.uniform(0.5, 1.5, size=N) if use_w else None
        shift = rng.normal(size=D)
        T = D * (D + 1) // 2
        out = torch.empty((K, 3 + D + T), dtype=torch.float64, device="cuda")
        tod = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
        _lib.Context.get().suffstats(tod(x), N, D, D, tod(shift), tod(rho), tod(gam), K, K, tod(w), out)
        got = out.cpu().numpy()
        u = rho * (1.0 if w is None else w[:, None])
        v = u * (1.0 if gam is None else gam)
        y = x - shift
        il = np.tril_indices(D)
        np.testing.assert_allclose(got[:, 0], u.sum(0), rtol=1e-12)
        np.testing.assert_allclose(got[:, 1], v.sum(0), rtol=1e-12)
        np.testing.assert_allclose(got[:, 2:2 + D], v.T @ y, rtol=1e-11, atol=1e-9)
        R = np.einsum("nk,ni,nj->kij", v, y, y)
        np.testing.assert_allclose(got[:, 2 + D:2 + D + T], R[:, il[0], il[1]], rtol=1e-11, atol=1e-9)
        L = (u * np.log(gam)).sum(0) if use_g else np.zeros(K)
        np.testing.assert_allclose(got[:, -1], L, rtol=1e-11, atol=1e-12)


def test_full_size_properties(pm):
    """BASELINE config 2 size (N=1e7 would take the oracle minutes): size-independent properties at N=2e6 --
    permutation invariance of per-sample outputs, agreement of a shard-wise evaluation with the whole, and the
    fused sum_n w_n log q_n against a float64 torch reduction of the per-sample output."""
    import torch
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200 import _lib
    K, D, N = 32, 30, 2_000_000
    means, covs, w, x_small, _ = _synth(K, D, 1000, seed=11)
    mix = create_gaussian_mixture(means, covs, w)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((N, D), dtype=torch.float64, device="cuda", generator=g) * 2.0
    lq = mix.multi_evaluate(x)
    perm = torch.randperm(N, device="cuda", generator=g)
    lq_perm = mix.multi_evaluate(x[perm].contiguous())
    assert torch.equal(lq_perm, lq[perm])
    half = N // 2 + 17
    assert torch.equal(mix.multi_evaluate(x[:half]), lq[:half])
    assert torch.equal(mix.multi_evaluate(x[half:]), lq[half:])
    sw = torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
    sums = run_k1(x, mix._packed(), K, _lib.MODE_GAUSS, weights=sw, want_sums=True).cpu().numpy()
    assert sums[0] == pytest.approx(float((sw * lq).sum()), rel=1e-12)
    assert sums[1] == pytest.approx(float(sw.sum()), rel=1e-12)
    # linear-domain check of rho: rows sum to one where nothing underflows
    rho = torch.empty((N, K), dtype=torch.float64, device="cuda")
    run_k1(x, mix._packed(), K, _lib.MODE_GAUSS, resp=rho)
    ok = lq > -600
    assert float((rho.sum(1)[ok] - 1.0).abs().max()) < 1e-12


# ------------------------------------------------------------------ K3: device-side propose (statistical parity)
def test_propose_device_statistics(pm):
    """K3 cannot be bit-identical to numpy's Mersenne Twister: check the block structure the
    reference produces (mixture.pyx:193-212) exactly and the distribution statistically."""
    import torch
    from scipy import stats
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    K, D, N = 4, 5, 400_000
    means, covs, w, _, _ = _synth(K, D, 10, seed=21)
    mix = create_gaussian_mixture(means, covs, w)
    rng = np.random.RandomState(5)
    x, lat = mix.propose_device(N, rng, trace=True, seed=1234)
    counts = np.random.RandomState(5).multinomial(N, mix.weights)          # same generator state as the call above
    lat = lat.cpu().numpy()
    np.testing.assert_array_equal(lat, np.repeat(np.arange(K), counts))   # component blocks in component order
    xh = x.cpu().numpy()
    chol = np.linalg.cholesky(covs)
    for k in range(K):
        xs = xh[lat == k]
        z = np.linalg.solve(chol[k], (xs - means[k]).T).T                 # whitened: iid N(0, 1) if K3 is right
        n = len(z)
        assert np.abs(z.mean(0)).max() < 5.0 / np.sqrt(n)
        assert np.abs(np.cov(z.T) - np.eye(D)).max() < 6.0 * np.sqrt(2.0 / n)
        assert stats.kstest(z[:20000, 0], "norm").pvalue > 1e-4
        assert stats.kstest(z[:20000, D - 1], "norm").pvalue > 1e-4
    # reproducible, and independent of how the rows are split over launches / ranks
    x2 = mix.propose_device(N, np.random.RandomState(5), seed=1234)
    assert torch.equal(x, x2)
    # Student-t: squared whitened radius / D is F(D, nu) distributed (student_t.pyx:49-55)
    dofs = np.array([3.0, 4.5, 8.0, 0.7])
    tm = create_t_mixture(means, covs, dofs, w)
    xt, latt = tm.propose_device(N, np.random.RandomState(7), trace=True, seed=99)
    xt, latt = xt.cpu().numpy(), latt.cpu().numpy()
    for k in range(K):
        xs = xt[latt == k][:50000]
        z = np.linalg.solve(chol[k], (xs - means[k]).T).T
        f = (z ** 2).sum(1) / D
        assert stats.kstest(f, stats.f(D, dofs[k]).cdf).pvalue > 1e-4, k
    # end to end: the importance weights of samples drawn from the mixture itself are all one
    lq = mix.multi_evaluate(x)
    assert bool(torch.isfinite(lq).all())


def test_propose_device_forms_agree(pm, tmp_path):
    """K3's register-resident form (D <= 40, the default) and its shared-memory form (any D; PMCB200_K3_FORM=smem)
    consume the Philox words and add the products in the same order: the same samples bit for bit.  Also checks the
    normals of k3_normal2 (own Box-Muller: table logarithm, Taylor sine / cosine) in a dimension where the product
    runs in registers: whitened draws at D = 30 are iid N(0, 1), tails included."""
    import subprocess
    import sys
    import textwrap
    from scipy import stats
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path.insert(0, %r)
        from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
        out = {}
        for (K, D, student) in [(5, 7, False), (8, 30, False), (4, 40, True), (6, 23, True), (3, 16, False), (2, 1, False)]:
            rng = np.random.default_rng(K * 100 + D)
            means = rng.normal(0, 3, size=(K, D))
            a = rng.normal(0, 1 / np.sqrt(D), size=(K, D, D))
            covs = a @ a.transpose(0, 2, 1) + 0.5 * np.eye(D)
            mix = create_t_mixture(means, covs, [4.0] * K) if student else create_gaussian_mixture(means, covs)
            out["x_%%d_%%d" %% (K, D)] = mix.propose_device(30011, np.random.RandomState(1), seed=5).cpu().numpy()
            if D == 30:
                out["means"], out["covs"] = means, covs
                out["counts"] = np.random.RandomState(1).multinomial(30011, mix.weights)
        np.savez(sys.argv[1], **out)
    ''') % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for form in ("regs", "smem"):
        path = str(tmp_path / ("k3_%s.npz" % form))
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, PMCB200_K3_FORM=form))
        res[form] = np.load(path)
    for key in res["regs"].files:
        if key.startswith("x_"):
            assert np.isfinite(res["regs"][key]).all()
            np.testing.assert_array_equal(res["regs"][key], res["smem"][key], err_msg=key)
    x, means, covs, counts = res["regs"]["x_8_30"], res["regs"]["means"], res["regs"]["covs"], res["regs"]["counts"]
    lat = np.repeat(np.arange(8), counts)
    z = np.concatenate([np.linalg.solve(np.linalg.cholesky(covs[k]), (x[lat == k] - means[k]).T).T for k in range(8)])
    n = len(z)
    assert np.abs(z.mean(0)).max() < 5.0 / np.sqrt(n)
    assert np.abs(np.cov(z.T) - np.eye(30)).max() < 6.0 * np.sqrt(2.0 / n)
    flat = z.ravel()
    assert stats.kstest(flat[:200000], "norm").pvalue > 1e-4
    assert abs(stats.kurtosis(flat)) < 0.02                               # 9e5 draws: sd of the estimate 0.005
    assert 3.5 < np.abs(flat).max() < 6.5                                 # the largest of 9e5 normals sits near 4.9


# ------------------------------------------------------------------ config 1: the examples/pmc.py loop end to end
def test_pmc_example_loop_matches_reference(pm, golden):
    """BASELINE config 1: ImportanceSampler.run + gaussian_pmc every 1000 samples, 10 steps, seeded global mtrand,
    replayed with this package's classes against the fixture the compiled reference produced
    (tests/golden/make_golden.py::pmc_example_case).  The proposal draws go through numpy exactly as in the
    reference, so the sample streams coincide and every later quantity can be compared number by number."""
    from copy import deepcopy
    from pypmc_b200.density.gauss import Gauss
    from pypmc_b200.density.mixture import MixtureDensity, create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc
    from pypmc_b200.sampler.importance_sampling import ImportanceSampler, combine_weights
    g = golden("pmc_example")
    steps, n = int(g["steps"]), int(g["n_per_step"])
    target = create_gaussian_mixture(g["t_means"], g["t_covs"], g["t_w"])
    proposal = MixtureDensity([Gauss(m, np.eye(2)) for m in g["p_means"]])
    np.random.seed(int(g["seed"]))
    sampler = ImportanceSampler(target.evaluate, proposal)
    proposals = []
    for i in range(steps):
        proposals.append(deepcopy(sampler.proposal))
        origin = sampler.run(n, trace_sort=True)
        np.testing.assert_array_equal(origin, g["origin_%d" % i])
        gaussian_pmc(sampler.samples[-1], sampler.proposal, sampler.weights[-1][:, 0], origin, mincount=20, rb=True,
                     copy=False)
        np.testing.assert_allclose(sampler.proposal.weights, g["prop_after_%d_weights" % i], rtol=TOL, atol=1e-300)
        live = sampler.proposal.weights != 0
        mu = np.array([c.mu for c in sampler.proposal.components])
        cov = np.array([c.sigma for c in sampler.proposal.components])
        np.testing.assert_allclose(mu[live], g["prop_after_%d_means" % i][live], rtol=TOL, atol=1e-12)
        assert mat_err(cov[live], g["prop_after_%d_covs" % i][live]) < TOL
    np.testing.assert_allclose(sampler.samples[:], g["samples"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sampler.weights[:][:, 0], g["weights"], rtol=TOL, atol=1e-300)
    assert len(sampler.samples) == steps and sampler.samples[-1].shape == (n, 2)
    # deterministic mixture weights over all steps, log-scale and linear-scale branches
    cw = combine_weights([sampler.samples[i] for i in range(steps)], [sampler.weights[i][:, 0] for i in range(steps)],
                         proposals)
    np.testing.assert_allclose(cw[:][:, 0], g["combined_weights"], rtol=TOL, atol=1e-300)
    w_lin = [sampler.weights[i][:, 0].copy() for i in range(3)]
    w_lin[1][5] = 0.0
    cl = combine_weights([sampler.samples[i] for i in range(3)], w_lin, proposals[:3])
    np.testing.assert_allclose(cl[:][:, 0], g["combined_weights_linear3"], rtol=TOL, atol=1e-300)
    # the batched weight pass equals the reference's per-sample formulation
    x5 = sampler.samples[0][:5]
    per_sample = np.array([np.exp(target.evaluate(x) - proposals[0].evaluate(x)) for x in x5])
    np.testing.assert_allclose(sampler.weights[0][:5, 0], per_sample, rtol=1e-10)


# ------------------------------------------------------------------ K1 fallback: exact-difference form
def test_exact_difference_fallback_far_narrow_components(pm, orc):
    """Components that lie > 1e4 standard deviations from the common shift make k1_prepare raise its flag: the
    exact-difference kernel (y = x - mu_k per component, k1_mixture_eval.cuh) must then produce the launch's
    outputs -- log q, individual, rho, gamma and the fused sums -- to the same tolerance."""
    import torch
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc
    from pypmc_b200 import _lib
    rng = np.random.default_rng(17)
    K, D, N = 3, 6, 3000
    means = np.array([[-2.0e3] * D, [0.5] * D, [3.0e3] * D]) + rng.normal(size=(K, D))
    covs = np.array([(lambda a: (a @ a.T + np.eye(D)) * 1e-4)(rng.normal(0, 0.3, size=(D, D))) for _ in range(K)])
    w = np.array([0.3, 0.5, 0.2])
    comp = rng.integers(0, K, size=N)
    x = means[comp] + np.einsum("nij,nj->ni", np.linalg.cholesky(covs)[comp], rng.normal(size=(N, D)))
    # |b| = |T (mu - c)| ~ 2e3 / 1e-2 = 2e5 > 1e4: the flag is up
    for dofs in (None, np.array([3.0, 5.0, 7.0])):
        comps = orc.Components(means, covs, dofs)
        lq_ref, ind_ref = orc.mixture_multi_evaluate(x, comps, w)
        mix = create_gaussian_mixture(means, covs, w) if dofs is None else create_t_mixture(means, covs, dofs, w)
        ind = np.empty((N, K))
        lq = mix.multi_evaluate(x, individual=ind)
        assert rel_err(lq, lq_ref) < TOL
        assert rel_err(ind, ind_ref) < TOL
        rho_ref, _ = orc.calculate_rho_rb(x, comps, w)
        xd = torch.from_numpy(x).cuda()
        rho = torch.empty((N, K), dtype=torch.float64, device="cuda")
        aux = torch.empty((N, K), dtype=torch.float64, device="cuda")
        mode = _lib.MODE_GAUSS if dofs is None else _lib.MODE_STUDENT_T
        sums = run_k1(xd, mix._packed(), K, mode, resp=rho, aux=aux, want_sums=True).cpu().numpy()
        assert rel_err(rho.cpu().numpy(), rho_ref, floor=1e-280) < TOL
        assert sums[0] == pytest.approx(float(lq_ref.sum()), rel=1e-12) and sums[1] == N
        if dofs is not None:
            assert rel_err(aux.cpu().numpy(), orc.student_t_gamma(x, comps)) < TOL
    # and the update on top of it: the moment kernel runs once per shift group here (mix_adapt/_stats.py)
    mix = create_gaussian_mixture(means, covs, w)
    new = gaussian_pmc(x, mix)
    rho_ref, _ = orc.calculate_rho_rb(x, orc.Components(means, covs), w)
    alpha, mu, cov = orc.pmc_moments(x, rho_ref)
    np.testing.assert_allclose(new.weights, alpha, rtol=1e-10)
    np.testing.assert_allclose([c.mu for c in new.components], mu, rtol=1e-10)
    assert mat_err(np.array([c.sigma for c in new.components]), cov) < TOL     # three shift groups: no cancellation


def test_empty_and_single_row_inputs(pm):
    """Edge sizes: N = 0 returns empty outputs without a launch, N = 1 works, device weight diagnostics agree with numpy."""
    import torch
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.tools.convergence import perp, ess
    means, covs, w, x, sw = _synth(3, 4, 50, seed=2)
    mix = create_gaussian_mixture(means, covs, w)
    assert mix.multi_evaluate(np.empty((0, 4))).shape == (0,)
    ind = np.empty((0, 3))
    assert mix.multi_evaluate(np.empty((0, 4)), individual=ind).shape == (0,)
    assert mix.multi_evaluate(torch.empty((0, 4), dtype=torch.float64, device="cuda")).shape == (0,)
    one = mix.multi_evaluate(x[:1])
    assert one.shape == (1,) and one[0] == pytest.approx(mix.multi_evaluate(x)[0], rel=1e-14)
    wd = torch.from_numpy(sw).cuda()
    assert perp(wd) == pytest.approx(perp(sw), rel=1e-12) and ess(wd) == pytest.approx(ess(sw), rel=1e-12)
    # the same through the matrix-instruction form of K1 (K >= 9, D >= 8): N = 1, N = 7 (less than one 8-row block),
    # and bit-identical rows whatever their position in a tile (a row's arithmetic must not depend on its neighbours)
    means, covs, w, x, sw = _synth(16, 10, 1000, seed=3)
    mix = create_gaussian_mixture(means, covs, w)
    full = mix.multi_evaluate(x)
    assert mix.multi_evaluate(np.empty((0, 10))).shape == (0,)
    np.testing.assert_array_equal(mix.multi_evaluate(x[:1]), full[:1])
    np.testing.assert_array_equal(mix.multi_evaluate(x[:7]), full[:7])
    np.testing.assert_array_equal(mix.multi_evaluate(x[333:777]), full[333:777])


def test_full_size_update_invariants(pm):
    """Size-independent properties of the update path at N = 2e6 (the oracle would take minutes):
    K2 is additive over row shards; VB responsibilities sum to one per row, sum_k N_k = N and the N_k-weighted
    component means reproduce the data mean; a PMC update preserves sum alpha = 1 and the weighted data mean."""
    import torch
    from pypmc_b200 import _lib
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, DeviceSamples
    from pypmc_b200.mix_adapt.variational import GaussianInference
    K, D, N = 32, 30, 2_000_000
    means, covs, w, _, _ = _synth(K, D, 10, seed=31)
    mix = create_gaussian_mixture(means, covs, w)
    x = mix.propose_device(N, np.random.RandomState(2), seed=3)
    g = torch.Generator(device="cuda").manual_seed(9)
    sw = torch.rand(N, dtype=torch.float64, device="cuda", generator=g) + 0.5
    rho = torch.rand((N, K), dtype=torch.float64, device="cuda", generator=g)
    ctx = _lib.Context.get()
    F = 3 + D + D * (D + 1) // 2
    shift = torch.zeros(D, dtype=torch.float64, device="cuda")

    def stats(lo, hi):
        out = torch.empty((K, F), dtype=torch.float64, device="cuda")
        ctx.suffstats(x[lo:hi], hi - lo, D, D, shift, rho[lo:hi], None, K, K, sw[lo:hi], out)
        return out
    half = N // 2 + 333
    whole, parts = stats(0, N), stats(0, half) + stats(half, N)
    assert float(((whole - parts).abs() / whole.abs().clamp_min(1e-300)).max()) < 1e-11
    # first-moment column against a plain float64 reduction
    col = ((sw[:, None] * rho)[:, :4].T @ x).cpu().numpy()
    np.testing.assert_allclose(whole[:4, 2:2 + D].cpu().numpy(), col, rtol=1e-11, atol=1e-6)

    vb = GaussianInference(x, initial_guess=mix)
    r = vb._r_dev
    assert float((r.sum(dim=1) - 1.0).abs().max()) < 1e-12
    assert vb.N_comp.sum() == pytest.approx(N, rel=1e-12)
    data_mean = x.mean(dim=0).cpu().numpy()
    np.testing.assert_allclose((vb.N_comp[:, None] * vb.x_mean_comp).sum(0) / N, data_mean, rtol=1e-10, atol=1e-12)

    new = gaussian_pmc(DeviceSamples(x, sw), mix)
    assert new.weights.sum() == pytest.approx(1.0, abs=1e-13)
    wmean = ((sw[:, None] * x).sum(0) / sw.sum()).cpu().numpy()
    mix_mean = (new.weights[:, None] * np.array([c.mu for c in new.components])).sum(0)
    np.testing.assert_allclose(mix_mean, wmean, rtol=1e-10, atol=1e-12)


def test_pmc_run_fused_likelihood_matches_two_launch_flow(pm, golden):
    """PMC.run(fuse_likelihood=True) serves the bound and the next update's responsibilities from one K1 launch;
    it must follow the reference flow (same iteration count on the fixture, same mixture to rounding)."""
    from pypmc_b200.mix_adapt.pmc import PMC, gaussian_pmc
    g = golden("gauss_small")
    mix = _mix(pm, g)
    runs = []
    for fuse in (False, True):
        p = PMC(g["x"], mix, weights=g["sample_weights"])
        conv = p.run(iterations=3, fuse_likelihood=fuse)
        runs.append((conv, p.density.weights.copy(), np.array([c.sigma for c in p.density.components]), p.log_likelihood()))
    assert runs[0][0] == runs[1][0] == (None if int(g["pmc_run3_converged"]) < 0 else int(g["pmc_run3_converged"]))
    np.testing.assert_allclose(runs[1][1], runs[0][1], rtol=1e-12)
    assert mat_err(runs[1][2], runs[0][2]) < 1e-12
    assert runs[1][3] == pytest.approx(runs[0][3], rel=1e-13)
    np.testing.assert_allclose(runs[1][1], g["pmc_run3_weights"], rtol=TOL)
    # a pass computed ahead is never applied to a different mixture
    p = PMC(g["x"], mix, weights=g["sample_weights"])
    p.run(iterations=1, fuse_likelihood=True)
    other = gaussian_pmc(p._device_samples, mix, weights=g["sample_weights"])       # not p.density: must recompute rho
    ref = gaussian_pmc(g["x"], mix, weights=g["sample_weights"])
    np.testing.assert_array_equal(other.weights, ref.weights)


def test_upload_and_pageable_staging(pm):
    """Large host arrays go through the library's pinned staging (pmcb200_upload / the host pipeline): contents
    must arrive bit for bit, for contiguous, row-strided and 1-d inputs, and results through pageable and pinned
    buffers must be the same bits."""
    import torch
    from pypmc_b200 import _device as dev
    from pypmc_b200.density.mixture import create_gaussian_mixture
    rng = np.random.default_rng(8)
    big = rng.normal(size=(700_001, 9))                                 # 50 MB: above the staging threshold
    assert torch.equal(dev.to_device(big).cpu(), torch.from_numpy(big))
    view = big[::2, :7]                                                  # row-strided
    assert view.nbytes < dev._BIG_UPLOAD or torch.equal(dev.to_device(view).cpu(), torch.from_numpy(np.ascontiguousarray(view)))
    wide = rng.normal(size=(900_000, 12))[:, :10]                        # 72 MB, padded rows
    assert torch.equal(dev.to_device(wide).cpu(), torch.from_numpy(np.ascontiguousarray(wide)))
    vec = rng.normal(size=5_000_000)
    assert torch.equal(dev.to_device(vec).cpu(), torch.from_numpy(vec))
    means, covs, w, _, _ = _synth(4, 9, 10, seed=5)
    mix = create_gaussian_mixture(means, covs, w)
    ind_a, ind_b = np.empty((len(big), 4)), torch.empty((len(big), 4), dtype=torch.float64, pin_memory=True)
    xa = big
    xb = torch.empty(big.shape, dtype=torch.float64, pin_memory=True)
    xb.numpy()[:] = big
    la = mix.multi_evaluate(xa, individual=ind_a)                        # pageable in, pageable out (staged both ways)
    lb = mix.multi_evaluate(xb.numpy(), individual=ind_b.numpy())        # pinned in, pinned out (in place)
    np.testing.assert_array_equal(la, lb)
    np.testing.assert_array_equal(ind_a, ind_b.numpy())


def test_k2_dfma_form_agrees_with_default(pm, tmp_path):
    """The DFMA register-tile form of K2 (PMCB200_K2_FORM=dfma, kept for comparison with the default FP64
    matrix-instruction form) must produce the same statistics to rounding.  The form is fixed per process, so the
    DFMA run happens in a child process."""
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    from pypmc_b200 import _lib
    script = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from pypmc_b200 import _lib
rng = np.random.default_rng(12)
res = {}
for (K, D, N, use_g) in [(32, 30, 5003, False), (16, 40, 3001, True), (70, 9, 2000, False), (3, 2, 100, True)]:
    x = rng.normal(size=(N, D)) + 1.0
    rho = rng.uniform(size=(N, K)); gam = rng.uniform(0.5, 2.0, size=(N, K)) if use_g else None
    w = rng.uniform(0.5, 1.5, size=N); shift = rng.normal(size=D)
    out = torch.empty((K, 3 + D + D * (D + 1) // 2), dtype=torch.float64, device="cuda")
    tod = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    _lib.Context.get().suffstats(tod(x), N, D, D, tod(shift), tod(rho), tod(gam), K, K, tod(w), out)
    res["%%d_%%d" %% (K, D)] = out.cpu().numpy()
np.savez(sys.argv[1], **res)
''' % ROOT
    outs = {}
    for form in ("mma", "dfma"):
        path = str(tmp_path / (form + ".npz"))
        env = dict(os.environ, PMCB200_K2_FORM=form)
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env, timeout=300)
        outs[form] = dict(np.load(path))
    for key in outs["mma"]:
        a, b = outs["mma"][key], outs["dfma"][key]
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-9)) < 1e-11, key


def test_weigh_reuses_the_e_pass(pm):
    """DeviceSamples.weigh: one K1 launch gives the importance weights AND the responsibilities of the update that
    follows; the result must be the update of the two-launch flow, bit for bit, and a different mixture must not
    pick the cached pass up."""
    import torch
    from pypmc_b200 import _lib
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, student_t_pmc, DeviceSamples
    K, D, N = 6, 9, 40011                              # (above the two-pass threshold: K2 runs once)
    means, covs, w, _, _ = _synth(K, D, 10, seed=41)
    tmeans = means + 0.1
    for student in (False, True):
        if student:
            prop = create_t_mixture(means, covs, np.linspace(3, 9, K), w)
            target = create_t_mixture(tmeans, covs, np.linspace(3, 9, K), w)
            update = student_t_pmc
        else:
            prop = create_gaussian_mixture(means, covs, w)
            target = create_gaussian_mixture(tmeans, covs, w)
            update = gaussian_pmc
        x = prop.propose_device(N, np.random.RandomState(4), seed=11)
        logp = target.multi_evaluate(x)
        # two-launch flow
        wts = torch.exp(logp - prop.multi_evaluate(x))
        ref = update(DeviceSamples(x, wts), prop)
        # fused flow
        ds = DeviceSamples(x)
        launches0 = _lib.Context.get().launch_count()
        wf = ds.weigh(prop, logp)
        new = update(ds, prop)
        used = _lib.Context.get().launch_count() - launches0
        assert torch.equal(wf, wts)
        # same responsibilities bit for bit; the normalisation sum_n w_n comes from kernel K4 in the fused flow and
        # from K1's epilogue in the two-launch flow (two fixed but different summation orders): last-bit agreement
        np.testing.assert_allclose(new.weights, ref.weights, rtol=4e-15, atol=0)
        for a, b in zip(new.components, ref.components):
            np.testing.assert_array_equal(a.mu, b.mu)
            np.testing.assert_array_equal(a.sigma, b.sigma)
        # the weight-quality measures come out of the same pass (tools/convergence.py:6-72)
        from pypmc_b200.tools.convergence import perp, ess
        wh = wts.cpu().numpy()
        assert ds.perp() == pytest.approx(perp(wh), rel=1e-12) and ds.ess() == pytest.approx(ess(wh), rel=1e-12)
        assert perp(wts) == pytest.approx(perp(wh), rel=1e-12) and ess(wts) == pytest.approx(ess(wh), rel=1e-12)
        assert used <= 10                                   # one K1 (prepare, fast, exact, finish) + K2 launches, no second K1
        # stale pass + different mixture: recomputed, not reused
        ds2 = DeviceSamples(x)
        ds2.weigh(prop, logp)
        other = update(ds2, target)
        ref2 = update(DeviceSamples(x, wts), target)
        np.testing.assert_allclose(other.weights, ref2.weights, rtol=4e-15, atol=0)


@pytest.mark.parametrize("K,D,N,dof", [(32, 30, 5003, None), (64, 20, 3001, None), (16, 40, 2000, 4.0), (14, 9, 1537, None),
                                       (70, 11, 999, 5.0), (35, 13, 700, None), (15, 8, 257, 3.0), (21, 30, 600, None),
                                       (44, 12, 500, 7.0), (52, 10, 400, None), (38, 16, 450, None),
                                       (64, 30, 1200, None), (32, 40, 900, 6.0), (96, 20, 800, None),
                                       (24, 33, 300, None), (16, 8, 100, 4.0), (16, 36, 260, None)])
def test_k1_matrix_instruction_form(pm, orc, K, D, N, dof):
    """The DMMA form of K1 (k1_mma_eval.cuh, the default for K >= 9, D >= 8 when the component count pads to blocks of
    8 within 20 %; every block count 2..8 is covered, (70, 11) and the last three shapes run in component groups
    because theta of all components exceeds shared memory) against the oracle, and against the DFMA form (PMCB200_K1_FORM=dfma, read per call) to rounding: log-pdfs, log q, rho / gamma, a component subset
    (non-contiguous output columns) and weighted sums."""
    import torch
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200 import _lib
    means, covs, w, x, sw = _synth(K, D, N, seed=300 + K + D, dof=dof)
    dofs = None if dof is None else np.linspace(2.5, 9.0, K)
    comps = orc.Components(means, covs, dofs)
    lq_ref, ind_ref = orc.mixture_multi_evaluate(x, comps, w)
    mix = create_gaussian_mixture(means, covs, w) if dofs is None else create_t_mixture(means, covs, dofs, w)
    mode = _lib.MODE_GAUSS if dofs is None else _lib.MODE_STUDENT_T
    res = {}
    old = os.environ.get("PMCB200_K1_FORM")
    try:
        for form in ("mma", "dfma"):
            os.environ["PMCB200_K1_FORM"] = form
            ind = np.empty((N, K))
            lq = mix.multi_evaluate(x, individual=ind)
            assert rel_err(ind, ind_ref) < TOL, form
            assert rel_err(lq, lq_ref) < TOL, form
            sub = [1, 4, K - 1]
            ind2 = np.full((N, K), -7.0)
            mix.multi_evaluate(x, individual=ind2, components=sub)
            np.testing.assert_allclose(ind2[:, sub], ind_ref[:, sub], rtol=TOL, atol=0)
            assert (np.delete(ind2, sub, axis=1) == -7.0).all()
            xd = torch.from_numpy(x).cuda()
            rho = torch.zeros((N, K), dtype=torch.float64, device="cuda")
            aux = torch.zeros((N, K), dtype=torch.float64, device="cuda") if dofs is not None else None
            lqd = torch.empty(N, dtype=torch.float64, device="cuda")
            sums = torch.zeros(2, dtype=torch.float64, device="cuda")
            run_k1(xd, mix._packed(list(range(K))), K, mode, logq=lqd, resp=rho, aux=aux,
                   weights=torch.from_numpy(sw).cuda(), sums=sums)
            rho_ref, _ = orc.calculate_rho_rb(x, comps, w, list(range(K)))
            assert rel_err(rho.cpu().numpy(), rho_ref) < TOL, form
            if dofs is not None:
                assert rel_err(aux.cpu().numpy(), orc.student_t_gamma(x, comps, list(range(K)))) < TOL, form
            np.testing.assert_array_equal(lqd.cpu().numpy(), lq)
            s = sums.cpu().numpy()
            assert s[1] == pytest.approx(sw.sum(), rel=1e-13) and s[0] == pytest.approx((sw * lq_ref).sum(), rel=1e-11)
            res[form] = (lq, ind)
    finally:
        if old is None:
            os.environ.pop("PMCB200_K1_FORM", None)
        else:
            os.environ["PMCB200_K1_FORM"] = old
    np.testing.assert_allclose(res["mma"][0], res["dfma"][0], rtol=1e-12, atol=0)
    np.testing.assert_allclose(res["mma"][1], res["dfma"][1], rtol=1e-12, atol=0)
    assert not np.array_equal(res["mma"][1], res["dfma"][1])         # two different kernels did run


def test_k1_matrix_instruction_form_tails_and_dead_components(pm, orc):
    """DMMA form of K1 where its fused second pass differs most from a literal transcription: samples on a line
    leaving the mixture (component log-pdfs run from ~-10 down to below -745, so exp(lp) passes through the subnormal
    range and to zero) and dead components (weight 0: column stays 0, max_init = 0 path of pmc.pyx:26-27)."""
    import torch
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc
    from pypmc_b200 import _lib
    K, D, N = 12, 9, 4000
    means, covs, w, x, sw = _synth(K, D, N, seed=77)
    w[[2, 7]] = 0.0
    w /= w.sum()
    live = [k for k in range(K) if w[k] != 0]
    direction = np.ones(D) / np.sqrt(D)
    x[: N // 2] = means[0] + np.linspace(0.0, 60.0, N // 2)[:, None] * direction    # q/2 from 0 to beyond 745
    comps = orc.Components(means, covs)
    mix = create_gaussian_mixture(means, covs, w)
    rho_ref, _ = orc.calculate_rho_rb(x, comps, w, live)
    xd = torch.from_numpy(x).cuda()
    rho = torch.zeros((N, K), dtype=torch.float64, device="cuda")
    lq = torch.empty(N, dtype=torch.float64, device="cuda")
    run_k1(xd, mix._packed(live), K, _lib.MODE_GAUSS, max_init=0.0, logq=lq, resp=rho)
    got = rho.cpu().numpy()
    assert np.isfinite(got).all()
    assert (got[:, [2, 7]] == 0).all()
    assert rel_err(got, rho_ref) < TOL
    lq_ref, _ = orc.mixture_multi_evaluate(x, comps, w)
    far = lq_ref < -700
    assert far.any() and rel_err(lq.cpu().numpy()[~far], lq_ref[~far]) < TOL
    # the update built on it agrees with the oracle's (weights, means)
    new = gaussian_pmc(x, mix, weights=sw)
    alpha, mu, cov = orc.pmc_moments(x, rho_ref, sample_weights=sw, live=live)
    np.testing.assert_allclose(new.weights[live], alpha[live], rtol=1e-10, atol=1e-300)
    for k in live:
        np.testing.assert_allclose(new.components[k].mu, mu[k], rtol=TOL)
        assert mat_err(new.components[k].sigma, cov[k]) < TOL


# ------------------------------------------------------------------ multi-tile paths of K1 against the oracle
# k1_mma_eval only prefetches the next tile (cp.async) and staggers its warp groups when a CTA owns >= 4 tiles, i.e.
# N >= 4 * 148 * (8 NB NW) = 1.5e5 rows (7.6e4 with NB = 1).  The fixtures and seeded shapes above stay below that, so
# these cases compare every N x K output of the three BASELINE shapes with the oracle at N = 2e5 (ragged).
def test_multi_tile_c2_gauss_vs_oracle(pm, orc):
    import torch
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, DeviceSamples
    from pypmc_b200 import _lib
    K, D, N = 32, 30, 200_003
    means, covs, w, x, sw = _synth(K, D, N, seed=501)
    # a few samples far outside the mixture (log-pdfs of -1e3 ... -1e8): the significance shortcut of the log-sum-exp
    # must still find the component(s) that matter, and nothing may overflow
    direction = np.ones(D) / np.sqrt(D)
    for i, dist in enumerate((40.0, 300.0, 3e3, 2e4)):
        x[7 + i] = means[i] + dist * direction
        x[N - 9 - i] = means[K - 1 - i] - dist * direction
    comps = orc.Components(means, covs)
    mix = create_gaussian_mixture(means, covs, w)
    xd, swd = torch.from_numpy(x).cuda(), torch.from_numpy(sw).cuda()
    lq = torch.empty(N, dtype=torch.float64, device="cuda")
    ind = torch.empty((N, K), dtype=torch.float64, device="cuda")
    rho = torch.empty((N, K), dtype=torch.float64, device="cuda")
    sums = torch.zeros(2, dtype=torch.float64, device="cuda")
    run_k1(xd, mix._packed(), K, _lib.MODE_GAUSS, logq=lq, lp=ind, resp=rho, weights=swd, sums=sums)   # fused second pass
    lq2 = mix.multi_evaluate(xd)                                                                          # eval-only instantiation
    rho_ref, lq_ref = orc.calculate_rho_rb(x, comps, w)
    _, ind_ref = orc.mixture_multi_evaluate(x, comps, w)
    assert rel_err(lq.cpu().numpy(), lq_ref) < TOL
    assert torch.equal(lq, lq2)
    assert rel_err(ind.cpu().numpy(), ind_ref) < TOL
    assert rel_err(rho.cpu().numpy(), rho_ref) < TOL
    s = sums.cpu().numpy()
    assert s[0] == pytest.approx(float((sw * lq_ref).sum()), rel=1e-12) and s[1] == pytest.approx(float(sw.sum()), rel=1e-13)
    # and the whole update (K1 rho + K2 + host finish) against the oracle's two-pass moments
    new = gaussian_pmc(DeviceSamples(xd, swd), mix)
    alpha, mu, cov = orc.pmc_moments(x, rho_ref, sw)
    np.testing.assert_allclose(new.weights, alpha, rtol=TOL)
    np.testing.assert_allclose([c.mu for c in new.components], mu, rtol=TOL, atol=1e-12)
    assert mat_err(np.array([c.sigma for c in new.components]), cov) < TOL


def test_multi_tile_c4_student_vs_oracle(pm, orc):
    import torch
    from pypmc_b200.density.mixture import create_t_mixture
    from pypmc_b200.density._eval import run_k1
    from pypmc_b200 import _lib
    K, D, N = 16, 40, 200_001
    means, covs, w, x, sw = _synth(K, D, N, seed=502, dof=4.0)
    dofs = np.full(K, 4.0)
    comps = orc.Components(means, covs, dofs)
    mix = create_t_mixture(means, covs, dofs, w)
    xd = torch.from_numpy(x).cuda()
    lq = torch.empty(N, dtype=torch.float64, device="cuda")
    rho = torch.empty((N, K), dtype=torch.float64, device="cuda")
    gam = torch.empty((N, K), dtype=torch.float64, device="cuda")
    run_k1(xd, mix._packed(), K, _lib.MODE_STUDENT_T, logq=lq, resp=rho, aux=gam)
    rho_ref, lq_ref = orc.calculate_rho_rb(x, comps, w)
    assert rel_err(lq.cpu().numpy(), lq_ref) < TOL
    assert rel_err(rho.cpu().numpy(), rho_ref) < TOL
    assert rel_err(gam.cpu().numpy(), orc.student_t_gamma(x, comps)) < TOL


def test_multi_tile_c3_vb_vs_oracle(pm, orc):
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.variational import GaussianInference
    K, D, N = 64, 20, 120_007
    means, covs, w, x, sw = _synth(K, D, N, seed=503)
    mix = create_gaussian_mixture(means, covs, w)
    for weights in (None, sw):
        vb = GaussianInference(x, initial_guess=mix, weights=weights)
        wn = None if weights is None else vb.weights
        ref = orc.vb_e_step(x, vb.m, vb.W, vb.beta, vb.nu, vb.alpha, vb.log_det_W, wn)
        assert rel_err(vb.expectation_gauss_exponent, ref["expectation_gauss_exponent"]) < TOL
        assert log_err(vb.log_rho, ref["log_rho"]) < TOL      # ln r_nk ~ -1e-9 for a row's dominant component: see log_err
        assert rel_err(vb.r, ref["r"]) < TOL
        np.testing.assert_allclose(vb.N_comp, ref["N_comp"], rtol=TOL)
        np.testing.assert_allclose(vb.x_mean_comp, ref["x_mean_comp"], rtol=TOL, atol=1e-12)
        assert mat_err(vb.S, ref["S"]) < TOL
        assert vb._expectation_log_q_Z == pytest.approx(float(ref["expectation_log_q_Z"]), rel=TOL)


# ------------------------------------------------------------------ the one collective, over NCCL
def _nccl_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pypmc_b200 import parallel
    parallel.init_from_env(backend="nccl")
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, DeviceSamples
    from pypmc_b200.mix_adapt.variational import GaussianInference
    K, D, N = 16, 12, 60_000
    means, covs, w, x, sw = _synth(K, D, N, seed=77)             # every rank builds the same data, keeps its shard
    mix = create_gaussian_mixture(means, covs, w)
    lo, hi = parallel.shard_rows(N)
    new = gaussian_pmc(DeviceSamples(np.ascontiguousarray(x[lo:hi]), sw[lo:hi]), mix)
    vb = GaussianInference(np.ascontiguousarray(x[lo:hi]), initial_guess=mix, weights=sw[lo:hi])
    vb.update()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), alpha=new.weights, mu=np.array([c.mu for c in new.components]),
             cov=np.array([c.sigma for c in new.components]), N_comp=vb.N_comp, m=vb.m, W=vb.W, bound=vb.likelihood_bound())
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_sharded_update_matches_unsharded(pm, tmp_path):
    """samples sharded over the GPUs of the box, ONE NCCL all-reduce of the statistics packet per update
    (replaces the gather / update-on-root / bcast of examples/pmc_mpi.py:83-131).  Every rank must end with identical
    bits, and the sharded update must agree with the unsharded one at the contract tolerance."""
    import socket
    import torch
    import torch.multiprocessing as mp
    from pypmc_b200 import parallel
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc
    from pypmc_b200.mix_adapt.variational import GaussianInference
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    for r in res[1:]:
        for key in res[0].files:
            np.testing.assert_array_equal(r[key], res[0][key])
    assert not parallel.enabled()
    K, D, N = 16, 12, 60_000
    means, covs, w, x, sw = _synth(K, D, N, seed=77)
    mix = create_gaussian_mixture(means, covs, w)
    one = gaussian_pmc(x, mix, weights=sw)
    np.testing.assert_allclose(res[0]["alpha"], one.weights, rtol=TOL)
    np.testing.assert_allclose(res[0]["mu"], [c.mu for c in one.components], rtol=TOL, atol=1e-12)
    assert mat_err(res[0]["cov"], np.array([c.sigma for c in one.components])) < TOL
    vb = GaussianInference(x, initial_guess=mix, weights=sw)
    vb.update()
    np.testing.assert_allclose(res[0]["N_comp"], vb.N_comp, rtol=TOL)
    np.testing.assert_allclose(res[0]["m"], vb.m, rtol=TOL, atol=1e-12)
    assert mat_err(res[0]["W"], vb.W) < TOL
    assert float(res[0]["bound"]) == pytest.approx(vb.likelihood_bound(), rel=TOL)
