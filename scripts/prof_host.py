import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import numpy as np, torch, logging
logging.getLogger("pypmc_b200").setLevel(logging.ERROR)
sys.path.insert(0, "scripts")
from bench_updates import synth
from pypmc_b200.density.mixture import create_gaussian_mixture
from pypmc_b200.mix_adapt.pmc import gaussian_pmc, DeviceSamples
from pypmc_b200.mix_adapt.variational import GaussianInference
K, D, N = 32, 30, 20000
mix = create_gaussian_mixture(*synth(K, D))
x = mix.propose_device(N, np.random.RandomState(1), seed=5)
ds = DeviceSamples(x, None)
for _ in range(3): gaussian_pmc(ds, mix)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): gaussian_pmc(ds, mix)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:4500])
K, D = 64, 20
mix = create_gaussian_mixture(*synth(K, D))
x = mix.propose_device(N, np.random.RandomState(1), seed=6)
vb = GaussianInference(x, initial_guess=mix)
for _ in range(3): vb.update()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): vb.update()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3800])
