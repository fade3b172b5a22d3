"""Sharded variational-Bayes E/M steps: GaussianInference on each rank's row block with ONE all-reduce of the
statistics packet per E-step, checked against the same inference run unsharded on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        scripts/vb_sharded.py --rows 200000
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000, help="samples per GPU")
    ap.add_argument("--K", type=int, default=8)
    ap.add_argument("--D", type=int, default=6)
    ap.add_argument("--updates", type=int, default=3)
    args = ap.parse_args()
    from pypmc_b200 import parallel
    rank, world = parallel.init_from_env(backend="nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.variational import GaussianInference
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pmc_sharded import synth_mixture

    K, D, n = args.K, args.D, args.rows
    mix = create_gaussian_mixture(*synth_mixture(K, D, seed=1))
    # every rank draws the WHOLE data set from one Philox stream and keeps its block, so rank 0 can redo it unsharded
    x_all = mix.propose_device(world * n, np.random.RandomState(5), seed=99)
    g = torch.Generator(device=x_all.device).manual_seed(3)
    w_all = torch.rand(world * n, dtype=torch.float64, device=x_all.device, generator=g) + 0.5
    lo, hi = parallel.shard_rows(world * n)
    start = create_gaussian_mixture(*synth_mixture(K, D, seed=2))
    vb = GaussianInference(x_all[lo:hi].contiguous(), initial_guess=start, weights=w_all[lo:hi].contiguous())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.updates):
        vb.update()
    bound = vb.likelihood_bound()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.updates
    flat = np.concatenate([vb.N_comp, vb.m.ravel(), vb.W.ravel(), vb.alpha, [bound]])
    mine = torch.from_numpy(flat).cuda()
    ref = mine.clone()
    if world > 1:
        dist.broadcast(ref, src=0)
    same = torch.tensor([1 if torch.equal(mine, ref) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    err = None
    if rank == 0:
        parallel.disable()
        one = GaussianInference(x_all, initial_guess=start, weights=w_all)
        for _ in range(args.updates):
            one.update()
        # SURVEY 8c metrics: element-wise relative for the K-vectors and the bound, max|diff| / max|ref| per component for
        # the vectors m_k and the matrices W_k (an element-wise ratio on a near-zero off-diagonal entry measures nothing)
        K_, D_ = one.m.shape
        errs = [np.max(np.abs(one.N_comp - vb.N_comp) / np.abs(one.N_comp)), np.max(np.abs(one.alpha - vb.alpha) / np.abs(one.alpha)),
                abs(one.likelihood_bound() - bound) / abs(bound),
                np.max(np.abs(one.m - vb.m).max(axis=1) / np.abs(one.m).max(axis=1)),
                np.max(np.abs(one.W - vb.W).reshape(K_, -1).max(axis=1) / np.abs(one.W).reshape(K_, -1).max(axis=1))]
        err = float(max(errs))
        print(json.dumps({"workload": "GaussianInference, %d updates, N=%d/GPU K=%d D=%d" % (args.updates, n, K, D),
                          "n_gpus": world, "s_per_update": dt, "ranks_identical": bool(same.item()),
                          "max_rel_diff_vs_unsharded": err, "bound": bound}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
