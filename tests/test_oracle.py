"""Pin ``oracle/`` (the CPU restatement used as the parity checker) against

(a) the hand-computed golden numbers in the reference's own unit tests and
(b) fixtures produced by the compiled, unmodified reference (tests/golden/make_golden.py).

CPU only -- runs in the ``-m "not gpu"`` tier.
"""
import numpy as np
import pytest

from conftest import mat_err, rel_err
from oracle import oracle as orc


# ---------------------------------------------------------------- reference golden numbers
def test_bilinear_sym_reference_golden():
    # pypmc/tools/linalg_test.py:9-15
    v = np.array([2.0, 4.3, 7.0])
    m = np.array([[3.0, 5.0, 1.9], [5.0, 0.8, 2.2], [1.9, 2.2, 4.2]])
    assert orc.bilinear_sym(m, v) == pytest.approx(504.23200000000003, abs=1e-7)


def test_chol_inv_det_reference_cases():
    # pypmc/tools/linalg_test.py:18-45
    for m in (np.array([[1.0, 0.5], [0.5, 1.0]]),
              np.array([[1.0, 0.5, 0.1], [0.5, 2.1, -0.4], [0.1, -0.4, 1.8]])):
        low, inv, log_det = orc.chol_inv_det(m)
        np.testing.assert_allclose(low, np.linalg.cholesky(m))
        np.testing.assert_allclose(inv, np.linalg.inv(m))
        assert log_det == pytest.approx(np.log(np.linalg.det(m)))
    with pytest.raises(np.linalg.LinAlgError, match="not symmetric"):
        orc.chol_inv_det(np.array([[0.01, 0.003], [0.001, 0.0025]]))
    for bad in (np.diag([0.0, 0.0025, 0.6]), -np.eye(13)):
        with pytest.raises(np.linalg.LinAlgError):
            orc.chol_inv_det(bad)


def test_logsumexp_reference_golden():
    # pypmc/tools/regularize_test.py:10-24
    assert orc.logsumexp(np.array([1.0, 2.0, 3.0]), np.array([0.3, 0.4, 0.3])) == pytest.approx(2.28205254, abs=1e-7)
    vals = np.array([[4.0, 8.0, 3.0], [0.3, 0.1, 5.0], [2.3, 5.6, 2.3]])
    np.testing.assert_allclose(orc.logsumexp2D(vals, np.array([1.3, 0.4, 0.3])),
                               [7.14628895, 3.844190158, 4.82132340])


def test_gauss_reference_golden():
    # pypmc/density/gauss_test.py:111-160
    comps = orc.Components([[4.3, 1.1]], [[[0.01, 0.003], [0.003, 0.0025]]])
    x = np.array([[4.35, 1.2]] * 2)
    _, ind = orc.mixture_multi_evaluate(x, comps, np.ones(1))
    np.testing.assert_allclose(ind[:, 0], 1.30077135, atol=1e-8)


def test_student_t_reference_golden():
    # pypmc/density/student_t_test.py:168-230, :155-166 (Cauchy)
    comps = orc.Components([[1.25, 4.3]], [[[0.0049, 0.0], [0.0, 0.01]]], dofs=[5.0])
    x = np.array([[1.3, 4.4], [1.26, 4.424]])
    _, ind = orc.mixture_multi_evaluate(x, comps, np.ones(1))
    np.testing.assert_allclose(ind[:, 0], [2.200202941, 2.174596526], atol=1e-9)
    cauchy = orc.Components([[0.0]], [[[1.0]]], dofs=[1.0])
    _, ind = orc.mixture_multi_evaluate(np.array([[3.2]]), cauchy, np.ones(1))
    assert ind[0, 0] == pytest.approx(-3.5642087303149452, abs=1e-12)


# pypmc/mix_adapt/pmc_test.py:11-57  (TestGaussianPMCNoOverlap tables)
_MU = np.array([[10.0, -1.0, 8.0], [-10.0, 7.4, 0.5]])
_COV = np.array([[[1.15, 0.875, 0.0], [0.875, 0.75, -0.2], [0.0, -0.2, 1.1]],
                 [[1.0, 0.01, 0.1], [0.01, 0.75, 0.0], [0.1, 0.0, 2.1]]])
PMC_MEANS = _MU + np.array([[0.001], [-0.005]])
PMC_COVS = _COV + np.array([0.001, -0.005])[:, None, None]
PMC_CW = np.array([0.7, 0.3])
PMC_LATENT = np.array([0] * 12 + [1] * 8)
PMC_WEIGHTS = np.array([12.89295915, 12.89372694, 12.89781423, 12.79548829, 12.89397248, 12.88642498,
                        12.89875608, 12.8977244, 12.8834032, 12.81344527, 12.8966767, 12.89319812,
                        20.02787201, 19.89550322, 19.81661548, 19.9733172, 19.81867511, 19.81555008,
                        19.83955669, 19.83352245])
PMC_SAMPLES = np.array([[9.7070033, -1.14093259, 7.79492513], [9.56875908, -1.3205348, 7.3705522],
                        [10.53728461, -0.93171182, 8.76279014], [9.80289836, -1.15107748, 9.27682257],
                        [8.91717444, -1.62000575, 7.60676764], [9.55705421, -1.65785994, 9.4330834],
                        [10.90155376, -0.42097835, 7.64481752], [11.06838483, -0.65188323, 8.69936008],
                        [8.50673184, -2.45559049, 8.62152455], [10.8097935, -0.33471831, 8.60497435],
                        [10.46129646, -1.04132199, 9.04460811], [10.23040728, -0.63621386, 6.48880065],
                        [-10.76972316, 8.23669361, 2.06283074], [-11.26019812, 7.03488615, -0.87321151],
                        [-9.99070915, 6.83422119, 0.28846651], [-9.39271812, 7.08741571, 1.91672609],
                        [-10.98814859, 7.55372701, 0.48618477], [-9.60983136, 6.24723833, 1.06241101],
                        [-10.61752466, 7.39052825, 1.17726011], [-10.4898097, 7.48668861, -2.41443733]])


def test_gaussian_pmc_reference_golden_tables():
    # pypmc/mix_adapt/pmc_test.py:93-169
    comps = orc.Components(PMC_MEANS, PMC_COVS)
    rho, _ = orc.calculate_rho_rb(PMC_SAMPLES, comps, PMC_CW)
    alpha, mu, cov = orc.pmc_moments(PMC_SAMPLES, rho, PMC_WEIGHTS)
    np.testing.assert_allclose(alpha, np.array([154.54358983999998, 159.02061223999999]) / 313.56420207999997)
    np.testing.assert_allclose(mu[0], np.array([1546.302278, -172.1300429, 1279.34733595]) / 154.54358983999998)
    np.testing.assert_allclose(mu[1], np.array([-1652.19922509, 1150.52591727, 74.098254]) / 159.02061223999999)
    np.testing.assert_allclose(cov[0], np.array([[91.13245238, 62.95055712, 4.96175291],
                                                 [62.95055712, 51.04895641, -16.59026473],
                                                 [4.96175291, -16.59026473, 111.63047879]]) / 154.54358983999998)
    np.testing.assert_allclose(cov[1], np.array([[61.35426434, -30.51320283, 50.0064872],
                                                 [-30.51320283, 47.59366671, 10.77061072],
                                                 [50.0064872, 10.77061072, 311.06169561]]) / 159.02061223999999)
    alpha, mu, cov = orc.pmc_moments(PMC_SAMPLES, rho)  # unweighted, pmc_test.py:145-169
    np.testing.assert_allclose(alpha, [0.6, 0.4])
    np.testing.assert_allclose(mu, [[10.00569514, -1.11356905, 8.27908553], [-10.38983286, 7.23392486, 0.4632788]])
    np.testing.assert_allclose(cov[1], [[0.38545161, -0.19190136, 0.31422734],
                                        [-0.19190136, 0.29882842, 0.06594038],
                                        [0.31422734, 0.06594038, 1.95479308]], rtol=1e-6)


# ---------------------------------------------------------------- fixtures from the compiled reference
@pytest.mark.parametrize("name", ["gauss_small", "gauss_c2", "gauss_c2_stress", "gauss_illcond"])
def test_gauss_mixture_vs_reference_fixture(golden, name):
    g = golden(name)
    comps = orc.Components(g["means"], g["covs"])
    logq, ind = orc.mixture_multi_evaluate(g["x"], comps, g["weights"])
    rows = len(g["individual"])
    assert rel_err(ind[:rows], g["individual"]) < 1e-13
    assert rel_err(logq, g["logq"]) < 1e-13
    live = [k for k in range(comps.K) if g["weights"][k] != 0]
    rho, logden = orc.calculate_rho_rb(g["x"], comps, g["weights"], live)
    for tag, sw in (("weighted", g["sample_weights"]), ("unweighted", None)):
        alpha, mu, cov = orc.pmc_moments(g["x"], rho, sw, live=live)
        np.testing.assert_allclose(alpha[live], g["pmc_%s_weights" % tag][live], rtol=1e-12)
        np.testing.assert_allclose(mu[live], g["pmc_%s_means" % tag][live], rtol=1e-11, atol=1e-13)
        assert mat_err(cov[live], g["pmc_%s_covs" % tag][live]) < 1e-12
    nw = g["sample_weights"] / g["sample_weights"].sum()
    assert (logq * nw).sum() == pytest.approx(float(g["loglik_weighted"]), rel=1e-13)


@pytest.mark.parametrize("name", ["student_small", "student_c4"])
def test_student_mixture_vs_reference_fixture(golden, name):
    g = golden(name)
    comps = orc.Components(g["means"], g["covs"], g["dofs"])
    logq, ind = orc.mixture_multi_evaluate(g["x"], comps, g["weights"])
    rows = len(g["individual"])
    assert rel_err(ind[:rows], g["individual"]) < 1e-13
    assert rel_err(logq, g["logq"]) < 1e-13
    rho, _ = orc.calculate_rho_rb(g["x"], comps, g["weights"])
    gamma = orc.student_t_gamma(g["x"], comps)
    alpha, mu, cov = orc.pmc_moments(g["x"], rho, g["sample_weights"], gamma)
    np.testing.assert_allclose(alpha, g["pmc_nodof_weighted_weights"], rtol=1e-12)
    np.testing.assert_allclose(mu, g["pmc_nodof_weighted_means"], rtol=1e-11, atol=1e-13)
    assert mat_err(cov, g["pmc_nodof_weighted_covs"]) < 1e-12
    # dof: solve the same first-order condition the reference solves (pmc.pyx:478-497, :696)
    from scipy.optimize import brentq
    from scipy.special import digamma
    const = orc.student_t_dof_const(g["x"], comps, rho, g["sample_weights"])
    new = [brentq(lambda nu, c=c: c + np.log(0.5 * nu) - digamma(0.5 * nu), 1e-5, 1e3, maxiter=100) for c in const]
    np.testing.assert_allclose(new, g["pmc_dof_weighted_dofs"], rtol=1e-9)


@pytest.mark.parametrize("name", ["vb_small", "vb_c3"])
def test_vb_e_step_vs_reference_fixture(golden, name):
    g = golden(name)
    x = g["x"]
    tags = ["unw"] + (["wgt"] if "wgt_init_r" in g else [])
    for tag in tags:
        sw = None
        if tag == "wgt":
            sw = len(x) * (g["sample_weights"] / g["sample_weights"].sum())  # variational.pyx:94
        for stage in ("init", "upd1"):
            p = lambda a: g["%s_%s_%s" % (tag, stage, a)]
            res = orc.vb_e_step(x, p("m"), p("W"), p("beta"), p("nu"), p("alpha"), p("log_det_W"), sw)
            rows = len(p("r"))
            np.testing.assert_allclose(res["expectation_det_ln_lambda"], p("expectation_det_ln_lambda"), rtol=1e-13)
            np.testing.assert_allclose(res["expectation_ln_pi"], p("expectation_ln_pi"), rtol=1e-13)
            assert rel_err(res["expectation_gauss_exponent"][:rows], p("expectation_gauss_exponent")) < 1e-13
            assert rel_err(res["log_rho"][:rows], p("log_rho")) < 1e-12
            assert rel_err(res["r"][:rows], p("r")) < 1e-11
            np.testing.assert_allclose(res["N_comp"], p("N_comp"), rtol=1e-12)
            np.testing.assert_allclose(res["x_mean_comp"], p("x_mean_comp"), rtol=1e-10, atol=1e-12)
            assert mat_err(res["S"], p("S")) < 1e-12
