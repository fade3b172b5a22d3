"""Generate the golden fixtures in this directory from the UNMODIFIED, compiled
reference (pypmc v1.2.6 @ 9e0ab49).

Run in the build container only (the reference does not travel to the GPU box):

    python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref /tmp/refbuild   # copy of /root/reference
    PYTHONPATH=baseline/_ref python tests/golden/make_golden.py
    # or, against an in-place build of a scratch copy (cp -r /root/reference /tmp/refbuild; setup.py build_ext --inplace):
    PYPMC_REF=/tmp/refbuild python tests/golden/make_golden.py [pmc_example]

Every array written here is an output of the reference's own public API
(``MixtureDensity.multi_evaluate``, ``gaussian_pmc``, ``student_t_pmc``,
``PMC``, ``GaussianInference``) on seeded synthetic inputs built as SURVEY.md
section 8(d) prescribes.  The fixtures pin ``oracle/`` (tests/test_oracle.py) and the
CUDA path (tests/test_gpu_*.py).
"""
import os
import sys
import logging

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.environ.get("PYPMC_REF", os.path.join(ROOT, "baseline", "_ref")))   # or an in-place build

import pypmc  # noqa: E402  (the compiled reference)
from pypmc.density.mixture import create_gaussian_mixture, create_t_mixture  # noqa: E402
from pypmc.mix_adapt.pmc import gaussian_pmc, student_t_pmc, PMC  # noqa: E402
from pypmc.mix_adapt.variational import GaussianInference  # noqa: E402

logging.getLogger("pypmc").setLevel(logging.ERROR)
assert "/root/repo/pypmc" not in pypmc.__file__ and ("baseline/_ref" in pypmc.__file__ or os.environ.get("PYPMC_REF")), \
    pypmc.__file__


def synth_mixture(K, D, seed=1, ridge=0.5, spread=3.0):
    """mu_k ~ N(0, spread^2), Sigma_k = A A^T + ridge*I with A_ij ~ N(0, 1/D),
    weights ~ U(0.5, 1.5) normalised."""
    rng = np.random.default_rng(seed)
    means = rng.normal(0.0, spread, size=(K, D))
    covs = np.empty((K, D, D))
    for k in range(K):
        a = rng.normal(0.0, 1.0 / np.sqrt(D), size=(D, D))
        covs[k] = a @ a.T + ridge * np.eye(D)
    w = rng.uniform(0.5, 1.5, size=K)
    return means, covs, w / w.sum()


def synth_samples(N, means, covs, seed=2, dof=None):
    """component ~ U{0..K-1}, x = mu_c + L_c z [ / sqrt(chi2_nu/nu) ]."""
    rng = np.random.default_rng(seed)
    K, D = means.shape
    comp = rng.integers(0, K, size=N)
    chol = np.linalg.cholesky(covs)
    z = rng.normal(size=(N, D))
    x = np.einsum("nij,nj->ni", chol[comp], z)
    if dof is not None:
        x /= np.sqrt(rng.chisquare(dof, size=N) / dof)[:, None]
    x += means[comp]
    sw = rng.uniform(0.5, 1.5, size=N)
    return np.ascontiguousarray(x), comp.astype(np.int64), sw


def recover(mix, t=False):
    out = dict(
        weights=np.array(mix.weights),
        means=np.array([c.mu for c in mix.components]),
        covs=np.array([c.sigma for c in mix.components]),
    )
    if t:
        out["dofs"] = np.array([c.dof for c in mix.components])
    return out


def pack(prefix, d):
    return {prefix + "_" + k: v for k, v in d.items()}


def gauss_case(name, N, K, D, ridge=0.5, dead=(), full=True, keep_rows=None):
    means, covs, w = synth_mixture(K, D, ridge=ridge)
    for k in dead:
        w[k] = 0.0
    w = w / w.sum()
    x, latent, sw = synth_samples(N, means, covs)
    mix = create_gaussian_mixture(means, covs, w)
    individual = np.empty((N, K))
    logq = mix.multi_evaluate(x, individual=individual)
    rows = slice(None) if keep_rows is None else slice(0, keep_rows)  # individual is N x K: keep a prefix
    out = dict(x=x, latent=latent, sample_weights=sw, means=means, covs=covs, weights=np.array(mix.weights),
               individual=individual[rows], logq=logq)
    out.update(pack("pmc_weighted", recover(gaussian_pmc(x, mix, weights=sw))))
    out.update(pack("pmc_unweighted", recover(gaussian_pmc(x, mix))))
    p = PMC(x, mix, weights=sw)
    out["loglik_weighted"] = np.array(p.log_likelihood())
    out["loglik_unweighted"] = np.array(PMC(x, mix).log_likelihood())
    if full:
        out.update(pack("pmc_latent_rb", recover(gaussian_pmc(x, mix, weights=sw, latent=latent, rb=True, mincount=2))))
        out.update(pack("pmc_latent_nonrb", recover(gaussian_pmc(x, mix, weights=sw, latent=latent, rb=False))))
        p3 = PMC(x, mix, weights=sw)
        out["pmc_run3_converged"] = np.array(-1 if (c := p3.run(iterations=3)) is None else c)
        out.update(pack("pmc_run3", recover(p3.density)))
        out["pmc_run3_loglik"] = np.array(p3.log_likelihood())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "cond(cov0) = %.3g" % np.linalg.cond(covs[0]))


def illcond_case(name="gauss_illcond", N=1536, K=16, D=8, narrow=1e-5, spread=30.0):
    """ill-conditioned covariances (one axis 1e5 times narrower in variance than the others, kappa = 1e5)
    whose centres are offset from the mixture's centre along the WIDE axes.  The Mahalanobis distance of the offsets
    stays moderate (|b_k|^2 ~ 6e3) while the terms of a quadratic form expanded about the common centre are ~kappa
    times larger -- the case a |b|^2-based guard of the expanded (matrix-instruction) form cannot see."""
    rng = np.random.default_rng(41)
    q, _ = np.linalg.qr(rng.normal(size=(D, D)))                 # common axes; the last one is the narrow one
    lam = np.ones(D)
    lam[-1] = narrow
    covs = np.array([q @ np.diag(lam * rng.uniform(0.5, 1.5, size=D)) @ q.T for _ in range(K)])
    covs = 0.5 * (covs + np.swapaxes(covs, 1, 2))
    coeff = rng.normal(0.0, spread, size=(K, D))
    coeff[:, -1] = rng.normal(0.0, 3.0 * np.sqrt(narrow), size=K)   # along the narrow axis: a few of ITS sigmas only
    means = coeff @ q.T
    w = rng.uniform(0.5, 1.5, size=K)
    w /= w.sum()
    x, latent, sw = synth_samples(N, means, covs, seed=43)
    mix = create_gaussian_mixture(means, covs, w)
    individual = np.empty((N, K))
    logq = mix.multi_evaluate(x, individual=individual)
    out = dict(x=x, latent=latent, sample_weights=sw, means=means, covs=covs, weights=np.array(mix.weights),
               individual=individual, logq=logq)
    out.update(pack("pmc_weighted", recover(gaussian_pmc(x, mix, weights=sw))))
    out.update(pack("pmc_unweighted", recover(gaussian_pmc(x, mix))))
    out["loglik_weighted"] = np.array(PMC(x, mix, weights=sw).log_likelihood())
    out["loglik_unweighted"] = np.array(PMC(x, mix).log_likelihood())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    c = (w[:, None] * means).sum(0)
    b2 = max(float((m - c) @ np.linalg.inv(s) @ (m - c)) for m, s in zip(means, covs))
    a = max(float(np.abs(m - c) @ np.abs(np.linalg.inv(s)) @ np.abs(m - c)) for m, s in zip(means, covs))
    print(name, "cond(cov0) = %.3g, max |b|^2 = %.3g, max |d|^T |P| |d| = %.3g" % (np.linalg.cond(covs[0]), b2, a))


def student_case(name, N, K, D, dof=4.0, full=True, keep_rows=None):
    means, covs, w = synth_mixture(K, D)
    x, latent, sw = synth_samples(N, means, covs, dof=dof)
    dofs = np.full(K, dof)
    dofs[0] = 2.5  # not all equal
    mix = create_t_mixture(means, covs, dofs, w)
    individual = np.empty((N, K))
    logq = mix.multi_evaluate(x, individual=individual)
    rows = slice(None) if keep_rows is None else slice(0, keep_rows)
    out = dict(x=x, latent=latent, sample_weights=sw, means=means, covs=covs, dofs=dofs,
               weights=np.array(mix.weights), individual=individual[rows], logq=logq)
    out.update(pack("pmc_nodof_weighted", recover(student_t_pmc(x, mix, weights=sw, dof_solver_steps=0), True)))
    out.update(pack("pmc_dof_weighted", recover(student_t_pmc(x, mix, weights=sw), True)))
    if full:
        out.update(pack("pmc_nodof_unweighted", recover(student_t_pmc(x, mix, dof_solver_steps=0), True)))
        out.update(pack("pmc_dof_unweighted", recover(student_t_pmc(x, mix), True)))
        out.update(pack("pmc_dof_latent_nonrb", recover(student_t_pmc(x, mix, weights=sw, latent=latent, rb=False), True)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name)


VB_ATTRS = ("expectation_gauss_exponent", "log_rho", "r", "N_comp", "inv_N_comp", "x_mean_comp", "S",
            "expectation_det_ln_lambda", "expectation_ln_pi", "alpha", "beta", "nu", "m", "W", "log_det_W")


def vb_snapshot(vb, keep_rows=None):
    out = {}
    for a in VB_ATTRS:
        v = np.array(getattr(vb, a))
        if keep_rows is not None and v.ndim == 2 and v.shape[0] == vb.N:
            v = v[:keep_rows]   # N x K attributes: keep a prefix
        out[a] = v
    return out


def vb_case(name, N, K, D, full=True, keep_rows=None):
    means, covs, w = synth_mixture(K, D)
    x, latent, sw = synth_samples(N, means, covs)
    mix = create_gaussian_mixture(means, covs, w)
    out = dict(x=x, sample_weights=sw, means=means, covs=covs, weights=np.array(mix.weights))
    for tag, weights in (("unw", None), ("wgt", sw)):
        vb = GaussianInference(x, initial_guess=mix, weights=weights)
        if tag == "wgt" and not full:
            continue
        out.update(pack(tag + "_init", vb_snapshot(vb, keep_rows)))
        out[tag + "_init_bound"] = np.array(vb.likelihood_bound())
        vb.update()
        out.update(pack(tag + "_upd1", vb_snapshot(vb, keep_rows)))
        out[tag + "_upd1_bound"] = np.array(vb.likelihood_bound())
        vb.update()
        out[tag + "_upd2_bound"] = np.array(vb.likelihood_bound())
        out.update(pack(tag + "_upd2_mix", recover(vb.make_mixture())))
    if not full:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name)
        return
    # default initial guess ("first") with explicit component count and pruning run
    vb = GaussianInference(x, components=K + 2)
    out["first_init_bound"] = np.array(vb.likelihood_bound())
    it = vb.run(iterations=5, prune=1.0)
    out["first_run5_converged"] = np.array(-1 if it is None else it)
    out["first_run5_K"] = np.array(vb.K)
    out["first_run5_bound"] = np.array(vb.likelihood_bound())
    out.update(pack("first_run5", {a: np.array(getattr(vb, a)) for a in ("N_comp", "m", "W", "alpha", "beta", "nu")}))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name)


def pmc_example_case(name="pmc_example", seed=123456, steps=10, n_per_step=1000):
    """BASELINE config 1: the loop of the reference's examples/pmc.py (bimodal 2-d Gaussian target, 3-component
    initial proposal, ImportanceSampler.run + gaussian_pmc every 1000 samples) with a seeded global mtrand, plus
    combine_weights over all steps (sampler/importance_sampling.py:238-371)."""
    from copy import deepcopy
    from pypmc.density.gauss import Gauss
    from pypmc.density.mixture import MixtureDensity
    from pypmc.sampler.importance_sampling import ImportanceSampler, combine_weights
    t_means = [np.array([5.0, 0.01]), np.array([-4.0, 1.0])]
    t_covs = [np.array([[0.01, 0.003], [0.003, 0.0025]]), np.array([[0.1, 0.0], [0.0, 0.02]])]
    t_w = np.array([0.3, 0.7])
    target = create_gaussian_mixture(t_means, t_covs, t_w)
    p_means = [np.array([4.0, 0.0]), np.array([-5.0, 0.0]), np.array([0.0, 0.0])]
    proposal = MixtureDensity([Gauss(m, np.eye(2)) for m in p_means])
    np.random.seed(seed)
    sampler = ImportanceSampler(target.evaluate, proposal)
    out = dict(seed=np.array(seed), steps=np.array(steps), n_per_step=np.array(n_per_step), t_means=np.array(t_means),
               t_covs=np.array(t_covs), t_w=t_w, p_means=np.array(p_means))
    proposals = []
    for i in range(steps):
        proposals.append(deepcopy(sampler.proposal))
        origin = sampler.run(n_per_step, trace_sort=True)
        out["origin_%d" % i] = np.array(origin)
        gaussian_pmc(sampler.samples[-1], sampler.proposal, sampler.weights[-1][:, 0], origin, mincount=20, rb=True,
                     copy=False)
        out.update(pack("prop_after_%d" % i, recover(sampler.proposal)))
    out["samples"] = np.array(sampler.samples[:])
    out["weights"] = np.array(sampler.weights[:][:, 0])
    cw = combine_weights([sampler.samples[i] for i in range(steps)], [sampler.weights[i][:, 0] for i in range(steps)],
                         proposals)
    out["combined_weights"] = np.array(cw[:][:, 0])
    # linear-scale branch: one non-positive weight
    w_lin = [sampler.weights[i][:, 0].copy() for i in range(3)]
    w_lin[1][5] = 0.0
    cl = combine_weights([sampler.samples[i] for i in range(3)], w_lin, proposals[:3])
    out["combined_weights_linear3"] = np.array(cl[:][:, 0])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pmc_example":
        pmc_example_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gauss_illcond":
        illcond_case()
        sys.exit(0)
    pmc_example_case()
    gauss_case("gauss_small", N=257, K=5, D=7, dead=(3,))
    gauss_case("gauss_c2", N=2048, K=32, D=30, full=False, keep_rows=128)  # BASELINE config 2 shape
    gauss_case("gauss_c2_stress", N=2048, K=32, D=30, ridge=1e-4, full=False, keep_rows=128)  # kappa ~ 3e4
    student_case("student_small", N=301, K=4, D=5)
    student_case("student_c4", N=1600, K=16, D=40, full=False, keep_rows=128)  # BASELINE config 4 shape
    illcond_case()
    vb_case("vb_small", N=300, K=4, D=3)
    vb_case("vb_c3", N=2048, K=64, D=20, full=False, keep_rows=64)          # BASELINE config 3 shape
