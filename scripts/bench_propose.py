import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from pypmc_b200.density.mixture import create_gaussian_mixture, create_t_mixture
sys.path.insert(0, "/root/repo/scripts")
from bench_updates import synth
for (K, D, N, t) in [(32, 30, 10_000_000, False), (16, 40, 5_000_000, True), (64, 20, 10_000_000, False)]:
    m, c, w = synth(K, D)
    mix = create_t_mixture(m, c, [4.0] * K, w) if t else create_gaussian_mixture(m, c, w)
    rs = np.random.RandomState(1)
    x = mix.propose_device(N, rs, seed=3); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); x = mix.propose_device(N, rs, seed=3); e1.record(); torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    print("K3 K=%d D=%d N=%d student=%s: event ms %.3f wall ms %.3f" % (K, D, N, t, min(a for a, _ in ts), min(b for _, b in ts)), flush=True)
