"""End-to-end timings of whole updates through the pypmc-compatible classes (kernels + all-reduce-free host finishing),
device-resident samples: gaussian_pmc (C2), GaussianInference.update (C3), student_t_pmc (C4), PMC.run (C2, 3 steps).
One JSON line per case; run under gpurun."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import logging  # noqa: E402

logging.getLogger("pypmc_b200").setLevel(logging.ERROR)
from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture  # noqa: E402
from pypmc_b200.mix_adapt.pmc import gaussian_pmc, student_t_pmc, PMC, DeviceSamples  # noqa: E402
from pypmc_b200.mix_adapt.variational import GaussianInference  # noqa: E402


def synth(K, D, seed=1):
    rng = np.random.default_rng(seed)
    means = rng.normal(0.0, 3.0, size=(K, D))
    covs = np.empty((K, D, D))
    for k in range(K):
        a = rng.normal(0.0, 1.0 / np.sqrt(D), size=(D, D))
        covs[k] = a @ a.T + 0.5 * np.eye(D)
    w = rng.uniform(0.5, 1.5, size=K)
    return means, covs, w / w.sum()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    rs = np.random.RandomState(3)
    # C2: gaussian_pmc
    K, D, N = 32, 30, int(1e7 * scale)
    mix = create_gaussian_mixture(*synth(K, D))
    x = mix.propose_device(N, rs, seed=5)
    sw = torch.rand(N, dtype=torch.float64, device=x.device) + 0.5
    ds = DeviceSamples(x, sw)
    t = timed(lambda: gaussian_pmc(ds, mix))
    print(json.dumps({"case": "C2 gaussian_pmc (K1 rho + K2 + host finishing)", "N": N, "K": K, "D": D, "s": t,
                      "pairs_per_s": N * K / t}), flush=True)
    for rep in range(2):
        p = PMC(x, mix, weights=sw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        conv = p.run(iterations=3)
        torch.cuda.synchronize()
        t3 = time.perf_counter() - t0
        n_it = conv if conv is not None else 3
        print(json.dumps({"case": "C2 PMC.run(iterations=3): update + log-likelihood pass per iteration", "N": N, "s": t3,
                          "iterations_run": n_it, "s_per_iteration": t3 / n_it, "rep": rep}), flush=True)
    del ds, p, x, sw
    torch.cuda.empty_cache()
    # C3: GaussianInference.update
    K, D, N = 64, 20, int(1e7 * scale)
    mix = create_gaussian_mixture(*synth(K, D))
    x = mix.propose_device(N, rs, seed=6)
    t0 = time.perf_counter()
    vb = GaussianInference(x, initial_guess=mix)
    torch.cuda.synchronize()
    tinit = time.perf_counter() - t0
    t = timed(vb.update)
    tb = timed(vb.likelihood_bound)
    print(json.dumps({"case": "C3 GaussianInference.update (M-step + E-step: K1 VB + K2)", "N": N, "K": K, "D": D, "s": t,
                      "pairs_per_s": N * K / t, "init_s": tinit, "likelihood_bound_s": tb}), flush=True)
    del vb, x
    torch.cuda.empty_cache()
    # C4: student_t_pmc
    K, D, N = 16, 40, int(5e6 * scale)
    m, c, w = synth(K, D)
    tmix = create_t_mixture(m, c, [4.0] * K, w)
    x = tmix.propose_device(N, rs, seed=7)
    ds = DeviceSamples(x, None)
    t = timed(lambda: student_t_pmc(ds, tmix))
    t0 = timed(lambda: student_t_pmc(ds, tmix, dof_solver_steps=0))
    print(json.dumps({"case": "C4 student_t_pmc (K1 rho+gamma + K2 + dof solver)", "N": N, "K": K, "D": D, "s": t,
                      "s_without_dof_solver": t0, "pairs_per_s": N * K / t}), flush=True)


if __name__ == "__main__":
    main()
