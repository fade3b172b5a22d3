"""Measured parity of the CUDA path against the fixtures from the compiled reference (tests/golden/*.npz):
prints the maximum errors per quantity, in the metrics of SURVEY 8c.  Run under gpurun."""
import json
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, mat_err, rel_err  # noqa: E402

logging.getLogger("pypmc_b200").setLevel(logging.ERROR)
from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture  # noqa: E402
from pypmc_b200.mix_adapt.pmc import gaussian_pmc, student_t_pmc, PMC  # noqa: E402
from pypmc_b200.mix_adapt.variational import GaussianInference  # noqa: E402

out = {}
for name in ("gauss_small", "gauss_c2", "gauss_c2_stress"):
    g = load_golden(name)
    mix = create_gaussian_mixture(g["means"], g["covs"], g["weights"])
    ind = np.empty((len(g["x"]), len(mix)))
    lq = mix.multi_evaluate(g["x"], individual=ind)
    rows = len(g["individual"])
    new = gaussian_pmc(g["x"], mix, weights=g["sample_weights"])
    live = g["weights"] != 0
    out[name] = {
        "logq_rel": rel_err(lq, g["logq"]), "individual_rel": rel_err(ind[:rows], g["individual"]),
        "pmc_alpha_rel": rel_err(new.weights, g["pmc_weighted_weights"]),
        "pmc_mu_rel": rel_err(np.array([c.mu for c in new.components])[live], g["pmc_weighted_means"][live]),
        "pmc_cov_matnorm": mat_err(np.array([c.sigma for c in new.components])[live], g["pmc_weighted_covs"][live]),
        "loglik_rel": abs(PMC(g["x"], mix, weights=g["sample_weights"]).log_likelihood() / float(g["loglik_weighted"]) - 1),
    }
for name in ("student_small", "student_c4"):
    g = load_golden(name)
    mix = create_t_mixture(g["means"], g["covs"], g["dofs"], g["weights"])
    ind = np.empty((len(g["x"]), len(mix)))
    lq = mix.multi_evaluate(g["x"], individual=ind)
    rows = len(g["individual"])
    new = student_t_pmc(g["x"], mix, weights=g["sample_weights"])
    out[name] = {
        "logq_rel": rel_err(lq, g["logq"]), "individual_rel": rel_err(ind[:rows], g["individual"]),
        "pmc_alpha_rel": rel_err(new.weights, g["pmc_dof_weighted_weights"]),
        "pmc_mu_rel": rel_err(np.array([c.mu for c in new.components]), g["pmc_dof_weighted_means"]),
        "pmc_cov_matnorm": mat_err(np.array([c.sigma for c in new.components]), g["pmc_dof_weighted_covs"]),
        "pmc_dof_rel": rel_err(np.array([c.dof for c in new.components]), g["pmc_dof_weighted_dofs"]),
    }
for name in ("vb_small", "vb_c3"):
    g = load_golden(name)
    keys = [k for k in g if k.startswith("mix_init_")]
    tag = "unw_init"
    mix = create_gaussian_mixture(g["means"], g["covs"], g["weights"])
    vb = GaussianInference(g["x"], initial_guess=mix)
    rows = len(g[tag + "_r"]) if (tag + "_r") in g else 0
    rec = {}
    for attr in ("N_comp", "x_mean_comp"):
        if tag + "_" + attr in g:
            rec[attr + "_rel"] = rel_err(getattr(vb, attr), g[tag + "_" + attr])
    if tag + "_S" in g:
        rec["S_matnorm"] = mat_err(vb.S, g[tag + "_S"])
    if rows:
        rec["r_rel"] = rel_err(vb.r[:rows], g[tag + "_r"], floor=1e-280)
        rec["log_rho_rel"] = rel_err(vb.log_rho[:rows], g[tag + "_log_rho"])
    if tag + "_bound" in g:
        rec["bound_rel"] = abs(vb.likelihood_bound() / float(g[tag + "_bound"]) - 1)
    out[name] = rec
print(json.dumps(out, indent=1))
