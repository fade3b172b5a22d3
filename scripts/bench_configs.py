"""Per-kernel timings of K1 / K2 on the BASELINE.json configurations (run under gpurun; one JSON line per case).

    python scripts/bench_configs.py [--scale 1.0] [--reps 5] [--cases c2,c3,c4]

Samples: ``--data mixture`` (default) draws every row from one of the components, centres ~ N(0, spread^2) with
spread 3 (the synthetic workload of SURVEY 8d: a sample then has one or two components that matter);
``--data normal`` gives standard-normal rows under heavily overlapping components (spread 1) -- the worst case for
K1's log-sum-exp, where every component's term is significant for every sample.  The component records are
well-conditioned random factors.  Times are CUDA-event medians on the launching stream.
Flop model: K1 D^2 + 4D per pair, K2 D^2 + 4D + 2 per pair.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pypmc_b200 import _lib  # noqa: E402

CASES = {
    # name: (N, K, D, mode, outputs of K1)
    "c2_eval": (10_000_000, 32, 30, _lib.MODE_GAUSS, ("logq",)),
    "c2_rho": (10_000_000, 32, 30, _lib.MODE_GAUSS, ("logq", "resp")),
    "c3_vb": (10_000_000, 64, 20, _lib.MODE_VB, ("resp", "lp")),
    "c3_vb_r": (10_000_000, 64, 20, _lib.MODE_VB, ("resp",)),          # E-step without materialising log_rho
    "c3_eval": (10_000_000, 64, 20, _lib.MODE_GAUSS, ("logq",)),
    "c3_vb_eval": (10_000_000, 64, 20, _lib.MODE_VB, ("logq",)),       # VB scalars, no N x K output: the loop with two sample blocks
    "c3_vb_lp": (10_000_000, 64, 20, _lib.MODE_VB, ("lp",)),           # normalised log rho only
    "c4_t_eval": (5_000_000, 16, 40, _lib.MODE_STUDENT_T, ("logq",)),
    "c4_t_rho": (5_000_000, 16, 40, _lib.MODE_STUDENT_T, ("logq", "resp", "aux")),
    # beyond the shared-memory residency of theta: component groups (k1_mma_eval launched per group)
    "k64d30_eval": (5_000_000, 64, 30, _lib.MODE_GAUSS, ("logq",)),
    "k64d30_rho": (5_000_000, 64, 30, _lib.MODE_GAUSS, ("logq", "resp")),
    "k32d40_t_eval": (5_000_000, 32, 40, _lib.MODE_STUDENT_T, ("logq",)),
    # component counts between the block sizes (CB = 3, 5, 6, 7; 48 = 32 + 16 in two groups)
    "k21d30_eval": (5_000_000, 21, 30, _lib.MODE_GAUSS, ("logq",)),
    "k40d20_eval": (5_000_000, 40, 20, _lib.MODE_GAUSS, ("logq",)),
    "k48d30_eval": (5_000_000, 48, 30, _lib.MODE_GAUSS, ("logq",)),
    "k56d20_eval": (5_000_000, 56, 20, _lib.MODE_GAUSS, ("logq",)),
    "k17d30_eval": (5_000_000, 17, 30, _lib.MODE_GAUSS, ("logq",)),
    "k20d30_rho": (5_000_000, 20, 30, _lib.MODE_GAUSS, ("logq", "resp")),
    "k40d20_rho": (5_000_000, 40, 20, _lib.MODE_GAUSS, ("logq", "resp")),
    "k48d30_rho": (5_000_000, 48, 30, _lib.MODE_GAUSS, ("logq", "resp")),
    "k100d20_rho": (5_000_000, 100, 20, _lib.MODE_GAUSS, ("logq", "resp")),
}


def records(K, D, mode, rng, spread=1.0, keep=None):
    recs = []
    for k in range(K):
        t = np.tril(rng.normal(0, 0.3 / np.sqrt(D), size=(D, D))) + np.eye(D)
        sc = np.zeros(_lib.NUM_SCALARS)
        if mode == _lib.MODE_GAUSS:
            sc[0] = -0.5 * D * np.log(2 * np.pi)
        elif mode == _lib.MODE_STUDENT_T:
            nu = 4.0
            sc[:5] = [-3.0, -0.5 * (nu + D), 1.0 / nu, nu, nu + D]
        else:
            sc[:5] = [-np.log(K), 0.0, D * np.log(2 * np.pi), D / 10.0, 1.0]
        sc[_lib.S_WEIGHT] = 1.0 / K if mode != _lib.MODE_VB else 1.0
        c = rng.normal(0, spread, size=D)
        if keep is not None:
            keep.append((t, c))
        recs.append(_lib.pack_record(t, c, sc))
    return np.stack(recs)


def mixture_rows(N, D, comps, device, seed):
    """x = c_k + T_k^-1 z for a uniformly drawn component k, in slabs."""
    g = torch.Generator(device=device).manual_seed(seed)
    tinv = torch.from_numpy(np.stack([np.linalg.inv(t) for t, _ in comps])).to(device)
    cen = torch.from_numpy(np.stack([c for _, c in comps])).to(device)
    x = torch.empty((N, D), dtype=torch.float64, device=device)
    for s in range(0, N, 1_000_000):
        m = min(1_000_000, N - s)
        k = torch.randint(0, len(comps), (m,), device=device, generator=g)
        z = torch.randn((m, D), dtype=torch.float64, device=device, generator=g)
        x[s:s + m] = cen[k] + torch.einsum("nij,nj->ni", tinv[k], z)
    return x


def time_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cases", default="c2_eval,c2_rho,c3_vb,c3_eval,c4_t_eval,c4_t_rho")
    ap.add_argument("--kernels", default="k1,k2")
    ap.add_argument("--data", default="mixture", choices=["mixture", "normal"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = _lib.Context.get(0)
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(5)
    peak, _ = ctx.fp64_peak(0, 3000)
    print(json.dumps({"fp64_dfma_peak_gflops": peak}))
    for name in args.cases.split(","):
        N, K, D, mode, outs = CASES[name]
        N = int(N * args.scale)
        comps = []
        rec = torch.from_numpy(records(K, D, mode, rng, spread=3.0 if args.data == "mixture" else 1.0, keep=comps)).to(dev)
        x = mixture_rows(N, D, comps, dev, seed=7) if args.data == "mixture" else torch.randn((N, D), dtype=torch.float64, device=dev)
        cols = torch.arange(K, dtype=torch.int32, device=dev)
        bufs = {o: torch.empty((N,) if o == "logq" else (N, K), dtype=torch.float64, device=dev) for o in outs}
        sums = torch.zeros(2, dtype=torch.float64, device=dev)

        def k1():
            ctx.mixture_eval(x, N, D, D, rec, cols, K, K, mode, -_lib.DBL_MAX, logq=bufs.get("logq"), lp=bufs.get("lp"),
                             resp=bufs.get("resp"), aux=bufs.get("aux"), sums=sums, stream=stream)

        med, best = time_ms(k1, args.reps)
        flops = float(N) * K * (D * D + 4 * D)
        nbytes = 8.0 * N * (D + (1 if "logq" in outs else 0) + K * sum(o != "logq" for o in outs))
        line = {"case": name, "kernel": "K1", "data": args.data, "N": N, "K": K, "D": D, "outputs": list(outs), "ms_median": med, "ms_best": best,
                "tflops": flops / med * 1e-9, "frac_fp64_peak": flops / med * 1e-6 / peak,
                "alg_GBs": nbytes / med * 1e-6, "pairs_per_s": float(N) * K / med * 1e3}
        print(json.dumps(line), flush=True)
        if "resp" in outs and "k2" in args.kernels:
            F = 1 + D + D * (D + 1) // 2
            out = torch.zeros(K * (F + 2), dtype=torch.float64, device=dev)
            shift = torch.zeros(D, dtype=torch.float64, device=dev)
            gamma = bufs.get("aux") if mode == _lib.MODE_STUDENT_T else None
            w = torch.rand(N, dtype=torch.float64, device=dev)

            def k2():
                ctx.suffstats(x, N, D, D, shift, bufs["resp"], gamma, K, K, w, out, stream)

            med, best = time_ms(k2, args.reps)
            flops2 = float(N) * K * (D * D + 4 * D + 2)
            nbytes2 = 8.0 * N * (D + K * (2 if gamma is not None else 1) + 1)
            print(json.dumps({"case": name, "kernel": "K2", "N": N, "K": K, "D": D, "ms_median": med, "ms_best": best,
                              "tflops": flops2 / med * 1e-9, "frac_fp64_peak": flops2 / med * 1e-6 / peak,
                              "alg_GBs": nbytes2 / med * 1e-6, "pairs_per_s": float(N) * K / med * 1e3}), flush=True)
        del x, bufs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
