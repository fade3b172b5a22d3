import os, sys, subprocess, numpy as np
sys.path.insert(0, os.getcwd())
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, "scripts")
    from bench_updates import synth
    from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
    out = {}
    for (K, D, student) in [(5, 7, False), (32, 30, False), (16, 40, True), (6, 23, True), (3, 47, False)]:
        means, covs, w = synth(K, D)
        mix = create_t_mixture(means, covs, [4.0] * K, w) if student else create_gaussian_mixture(means, covs, w)
        x = mix.propose_device(20011, np.random.RandomState(1), seed=5)
        out["%d_%d" % (K, D)] = x.cpu().numpy()
    np.savez(sys.argv[1], **out)
else:
    for form in ("regs", "smem"):
        env = dict(os.environ, PMCB200_K3_FORM=form)
        subprocess.check_call([sys.executable, __file__, "/tmp/k3_%s.npz" % form], env=env)
    a, b = np.load("/tmp/k3_regs.npz"), np.load("/tmp/k3_smem.npz")
    for k in a.files:
        print(k, "max abs diff regs vs smem:", np.max(np.abs(a[k] - b[k])), "finite", np.isfinite(a[k]).all())
