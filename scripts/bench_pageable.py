import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "scripts")
from bench_updates import synth
from pypmc_b200.density.mixture import create_gaussian_mixture
K, D, N = 32, 30, 4_000_000
mix = create_gaussian_mixture(*synth(K, D))
x = np.random.default_rng(0).normal(size=(N, D))
out = np.empty(N)
for _ in range(2): mix.multi_evaluate(x, out=out)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); mix.multi_evaluate(x, out=out); ts.append(time.perf_counter() - t0)
print("threads", os.environ.get("PMCB200_COPY_THREADS", "default"), "pageable: %.2f ms per %d rows -> %.1f ms per 1e7, %.1f GB/s" % (np.median(ts) * 1e3, N, np.median(ts) * 1e3 * 1e7 / N, N * D * 8 / np.median(ts) / 1e9), "cores", os.cpu_count())
