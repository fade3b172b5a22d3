"""Config 5 (BASELINE.json): one full PMC iteration -- propose on the device (K3), weight (2 x K1), update
(K1 rho + K2 + all-reduce + host finishing) -- with the samples sharded over the GPUs of one node.  Launch with torchrun; prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/pmc_sharded.py --rows 10000000 [--check]

--check: rank 0 also runs the same update unsharded on the concatenated (small) data and compares at 1e-12.
Samples are drawn on each rank's device from the proposal mixture (kernel K3, one Philox stream indexed by the
global row so that ranks never overlap); the target is a second
K=32 Gaussian mixture evaluated with the same kernel K1, importance weight = exp(log p - log q) as in
examples/pmc.py:30-32 of the reference.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def synth_mixture(K, D, seed):
    rng = np.random.default_rng(seed)
    means = rng.normal(0.0, 3.0, size=(K, D))
    covs = np.empty((K, D, D))
    for k in range(K):
        a = rng.normal(0.0, 1.0 / np.sqrt(D), size=(D, D))
        covs[k] = a @ a.T + 0.5 * np.eye(D)
    w = rng.uniform(0.5, 1.5, size=K)
    return means, covs, w / w.sum()


def draw(n, means, covs, w, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    K, D = means.shape
    mu = torch.from_numpy(means).to(device)
    chol = torch.from_numpy(np.linalg.cholesky(covs)).to(device)
    x = torch.empty((n, D), dtype=torch.float64, device=device)
    slab = 1_000_000
    wt = torch.from_numpy(w).to(device)
    for s in range(0, n, slab):
        m = min(slab, n - s)
        comp = torch.multinomial(wt, m, replacement=True, generator=g)
        z = torch.randn((m, D), dtype=torch.float64, device=device, generator=g)
        x[s:s + m] = mu[comp] + torch.einsum("nij,nj->ni", chol[comp], z)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000, help="samples per GPU")
    ap.add_argument("--K", type=int, default=32)
    ap.add_argument("--D", type=int, default=30)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fused", action="store_true",
                    help="weights and responsibilities from ONE proposal evaluation (DeviceSamples.weigh)")
    args = ap.parse_args()

    from pypmc_b200 import parallel
    rank, world = parallel.init_from_env(backend="nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, DeviceSamples

    K, D, n = args.K, args.D, args.rows
    prop = create_gaussian_mixture(*synth_mixture(K, D, seed=1))
    pm, pc, pw = synth_mixture(K, D, seed=1)
    if args.check:
        # a target close to the proposal keeps the importance weights O(1), so the update is well conditioned and
        # sharded vs unsharded agree to rounding (with the seed-3 target a handful of samples carry all the weight)
        tm = pm + 0.05 * np.random.default_rng(4).normal(size=pm.shape)
        target = create_gaussian_mixture(tm, pc, pw)
    else:
        target = create_gaussian_mixture(*synth_mixture(K, D, seed=3))
    rng = np.random.RandomState(100 + rank)
    tsplit = {}

    def iteration():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        x = prop.propose_device(n, rng, seed=777, index0=rank * n)   # K3 (multinomial counts on the host)
        e[1].record()
        logp = target.multi_evaluate(x)                     # K1 (target)
        if args.fused:
            ds = DeviceSamples(x)
            ds.weigh(prop, logp)                            # K1 (proposal): log q -> weights, and rho for the update
            e[2].record()
            new = gaussian_pmc(ds, prop)                    # K2 + all-reduce + host update (rho re-used)
        else:
            logq = prop.multi_evaluate(x)                   # K1 (proposal)
            wts = torch.exp(logp - logq)                    # importance weights
            e[2].record()
            new = gaussian_pmc(DeviceSamples(x, wts), prop)     # K1 (rho) + K2 + all-reduce + host update
        torch.cuda.synchronize()
        tsplit.update(propose_ms=e[0].elapsed_time(e[1]), weight_ms=e[1].elapsed_time(e[2]), x=x)
        return new

    new = iteration()                                        # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        new = iteration()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)

    # every rank must hold the same updated mixture, bit for bit
    flat = np.concatenate([new.weights] + [c.mu for c in new.components] + [c.sigma.ravel() for c in new.components])
    mine = torch.from_numpy(flat).to(dev)
    ref = mine.clone()
    if world > 1:
        dist.broadcast(ref, src=0)
    same = bool(torch.equal(mine, ref))
    ok = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)

    check = None
    if args.check:
        # unsharded run of the same update on the gathered samples (small N only)
        x = tsplit["x"]
        xs = [torch.empty_like(x) for _ in range(world)]
        if world > 1:
            dist.all_gather(xs, x)
        else:
            xs = [x]
        if rank == 0:
            parallel.disable()
            xa = torch.cat(xs)
            wts = torch.exp(target.multi_evaluate(xa) - prop.multi_evaluate(xa))
            one = gaussian_pmc(DeviceSamples(xa, wts), prop)
            errs = [np.max(np.abs(one.weights - new.weights) / one.weights)]
            for c1, c2 in zip(one.components, new.components):   # SURVEY 8c metric: max|diff| / max|ref| per array
                errs.append(np.max(np.abs(c1.mu - c2.mu)) / np.max(np.abs(c1.mu)))
                errs.append(np.max(np.abs(c1.sigma - c2.sigma)) / np.max(np.abs(c1.sigma)))
            check = float(max(errs))
    if rank == 0:
        print(json.dumps({"workload": "PMC iteration: propose + 2x multi_evaluate + gaussian_pmc, N=%d/GPU K=%d D=%d" % (n, K, D),
                          "n_gpus": world, "s_per_iteration": float(t[0]), "pairs_per_s": world * n * K / float(t[0]),
                          "fused": bool(args.fused), "propose_ms": tsplit["propose_ms"], "weight_ms": tsplit["weight_ms"],
                          "ranks_identical": bool(ok.item()), "max_rel_diff_vs_unsharded": check,
                          "times": times}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
