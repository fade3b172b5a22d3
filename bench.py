#!/usr/bin/env python
"""bench.py -- throughput of the mixture-density hot path on B200 (contract: see the task statement).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm's CPU restatement on the host cores

Workload (BASELINE.json configs[1]): ``MixtureDensity.multi_evaluate`` of N = 1e7 float64 samples per GPU under a
K = 32 component, D = 30 Gaussian mixture.  A step is one pass over the batch.  Metric: sample-component
evaluations per second (N*K/s), whole job.  With several GPUs the samples are sharded (weak scaling, 1e7 per GPU);
``multi_evaluate`` has no exchange step, so there is no data-path collective -- only the timing barrier.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU, K, D = 10_000_000, 32, 30
FLOP_PER_PAIR = D * D + 4 * D            # D subtract + D(D+1) triangular FMA flops + 2D square-accumulate
BYTES_PER_SAMPLE = 8 * D + 8             # read x, write log q


def synth_mixture(seed=1):
    """mu_k ~ N(0, 3^2), Sigma_k = A A^T + 0.5 I with A_ij ~ N(0, 1/D), weights ~ U(0.5, 1.5)."""
    rng = np.random.default_rng(seed)
    means = rng.normal(0.0, 3.0, size=(K, D))
    covs = np.empty((K, D, D))
    for k in range(K):
        a = rng.normal(0.0, 1.0 / np.sqrt(D), size=(D, D))
        covs[k] = a @ a.T + 0.5 * np.eye(D)
    w = rng.uniform(0.5, 1.5, size=K)
    return means, covs, w / w.sum()


def synth_samples_host(n, means, covs, seed=2):
    rng = np.random.default_rng(seed)
    comp = rng.integers(0, K, size=n)
    chol = np.linalg.cholesky(covs)
    z = rng.normal(size=(n, D))
    return np.ascontiguousarray(means[comp] + np.einsum("nij,nj->ni", chol[comp], z))


def synth_samples_device(n, means, covs, seed, device):
    """x_n = mu_c + L_c z on the device (component c uniform), generated in slabs to bound temporary memory."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    mu = torch.from_numpy(means).to(device)
    chol = torch.from_numpy(np.linalg.cholesky(covs)).to(device)
    x = torch.empty((n, D), dtype=torch.float64, device=device)
    slab = 1_000_000
    for s in range(0, n, slab):
        m = min(slab, n - s)
        comp = torch.randint(0, K, (m,), device=device, generator=g)
        z = torch.randn((m, D), dtype=torch.float64, device=device, generator=g)
        x[s:s + m] = mu[comp] + torch.einsum("nij,nj->ni", chol[comp], z)
    return x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summarise the samples that arrived in [t0, t1] (the timed region); if fewer than 3 did, all samples
        since start() (warm-up + timed region, the same load) are used and ``window`` says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if t0 is not None and t0 <= t <= t1 + 0.03]
        window = "timed region"
        if len(inside) < 3:
            inside, window = [ln for (_, ln) in self.lines], "warm-up + timed region"
        for ln in inside:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_pass(x, comps, weights, threads):
    """One multi_evaluate pass of the oracle (CPU restatement of the reference algorithm, oracle/pmc_oracle.c) over
    ``x``, rows sharded over ``threads`` host threads (ctypes releases the GIL) -- the reference's own scaling
    pattern is one process per rank over sample shards (pypmc/tools/parallel_sampler.py:58-66)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    bounds = np.linspace(0, len(x), threads + 1).astype(int)

    def work(i):
        lq, _ = orc.mixture_multi_evaluate(x[bounds[i]:bounds[i + 1]], comps, weights)
        return lq

    with ThreadPoolExecutor(max_workers=threads) as ex:
        return np.concatenate(list(ex.map(work, range(threads))))


def cpu_arm(rows, steps, warmup, threads, x=None):
    from oracle import oracle as orc
    orc.build()
    means, covs, w = synth_mixture()
    comps = orc.Components(means, covs)
    if x is None:
        x = synth_samples_host(rows, means, covs)
    x = x[:rows]
    for _ in range(warmup):
        cpu_reference_pass(x, comps, w, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(x, comps, w, threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rows * K / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = int(min(N_PER_GPU, 125_000 * threads))          # ~1 s of CPU work per step on this pool's hosts
    value, dt = cpu_arm(rows, args.steps, args.warmup, threads)
    sample = "%d of %d rows per step (bounded sample of the same workload), oracle/pmc_oracle.c over %d threads" % (
        rows, N_PER_GPU, threads)
    print(json.dumps({
        "impl": "reference", "metric": "sample-component evals/sec (N*K/s)", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "MixtureDensity.multi_evaluate N=1e7/GPU K=32 D=30 Gaussian (BASELINE configs[1])",
                   "N_per_gpu": N_PER_GPU, "K": K, "D": D},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=device)

    import pypmc_b200
    from pypmc_b200 import _lib
    from pypmc_b200.density.mixture import create_gaussian_mixture

    ctx = _lib.Context.get(local_rank)
    means, covs, w = synth_mixture()
    mix = create_gaussian_mixture(means, covs, w)
    n = args.rows or N_PER_GPU
    x = synth_samples_device(n, means, covs, seed=2 + rank, device=device)
    logq = torch.empty(n, dtype=torch.float64, device=device)

    def step():
        mix.multi_evaluate(x, out=logq)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)                                    # let nvidia-smi come up before the load starts
    for _ in range(args.warmup):
        step()
    sync_all()
    launches0 = ctx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    t_begin = time.time()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    sync_all()
    t_end = time.time()
    launches = ctx.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clk = clocks.stop(t_begin, t_end) if rank == 0 else None
    t_ms = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms[0])
    ms_per_step = total_ms / args.steps
    value = world * n * K / (ms_per_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned): H2D of x and D2H of log q inside the timed region
    e2e_rows = args.e2e_rows or n
    xh_t = torch.empty((e2e_rows, D), dtype=torch.float64, pin_memory=True)
    xh_t.copy_(x[:e2e_rows])
    outh_t = torch.empty(e2e_rows, dtype=torch.float64, pin_memory=True)
    xh, outh = xh_t.numpy(), outh_t.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    mix.multi_evaluate(xh, out=outh)                      # warm-up (allocates the chunk buffers)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mix.multi_evaluate(xh, out=outh)                  # synchronous: returns when log q is on the host
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_rows * K / float(t_e[0])
    e2e_ok = bool(torch.equal(outh_t[:1000].to(device), logq[:1000]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: the binding roof is the FP64 FMA pipe, measured live; HBM fraction reported beside it.
    # The timed region is short (steps x ~13 ms), so the burst DFMA figure is the denominator; a 1 s DFMA run
    # (sustained, power/thermal steady state) is reported next to it.
    peak_gflops, _ = ctx.fp64_peak(0, 3000)
    sustained_gflops, sustained_ms = ctx.fp64_peak(0, 2500000)   # ~1 s per repetition
    kernel_ms = float(np.median(per_step))                # prepare + K1 (the exact-difference form returns at once)
    flops = FLOP_PER_PAIR * float(n) * K
    achieved_tf = flops / (kernel_ms * 1e-3) * 1e-12
    peaks, peak_src = measured_peaks()
    hbm_bytes = BYTES_PER_SAMPLE * float(n)
    hbm_gbs = hbm_bytes / (kernel_ms * 1e-3) * 1e-9
    traffic, kernel_name = None, "k1_mma_eval<4, 2, 16, false>"
    tpath = os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")   # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("rows") == n and "k1_mma_eval" in tj.get("kernel", ""):
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": peak_gflops * 1e-3, "unit": "TFLOP/s",
        "frac": achieved_tf / (peak_gflops * 1e-3), "traffic": traffic,
        "peak_source": "measured: DFMA microbenchmark run in this process (pmcb200_fp64_peak, register-resident "
                       "chains on every SM, burst); MEASURED_PEAKS.json has no FP64 entry",
        "peak_sustained": sustained_gflops * 1e-3, "peak_sustained_ms": sustained_ms,
        "kernel": kernel_name, "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops,
        "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src},
    }

    # ---- CPU baseline: the oracle (port of the reference algorithm) on a bounded sample, all host threads.
    # A 200k-row probe sizes the sample to ~15 s of CPU work (at most the whole workload).
    threads = os.cpu_count() or 1
    probe_rows = int(min(e2e_rows, 200_000))
    probe_value, _ = cpu_arm(probe_rows, 1, 1, threads, x=xh)
    cpu_rows = int(min(e2e_rows, max(probe_rows, 15.0 * probe_value / K)))
    cpu_value, cpu_dt = cpu_arm(cpu_rows, 1, 0, threads, x=xh)
    one_rows = int(min(e2e_rows, 100_000))                 # the reference as shipped is single-threaded
    one_value, one_dt = cpu_arm(one_rows, 1, 0, 1, x=xh)
    cpu = {"value": cpu_value, "unit": "pairs/s", "cores": threads, "kind": "port",
           "sample": "first %d of %d rows of the same workload, 1 timed pass (%.1f s), oracle/pmc_oracle.c over %d threads"
                     % (cpu_rows, n, cpu_dt, threads),
           "value_1_core": one_value, "sample_1_core": "first %d rows, 1 thread (%.1f s)" % (one_rows, one_dt)}

    print(json.dumps({
        "metric": "sample-component evals/sec (N*K/s)", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "MixtureDensity.multi_evaluate N=1e7/GPU K=32 D=30 Gaussian (BASELINE configs[1])",
                   "N_per_gpu": n, "K": K, "D": D, "l2": "inputs (2.4 GB/GPU) larger than L2, no flush needed",
                   "parallelism": "samples sharded over %d GPU(s), no data-path collective" % world},
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 8 * D * e2e_rows,
                "d2h_bytes_per_step": 8 * e2e_rows, "rows": e2e_rows, "s_per_step": float(t_e[0]),
                "matches_device_result": e2e_ok},
        "gpu_launches": int(launches),
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="samples per GPU (default 1e7, the BASELINE config)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the end-to-end leg (default: all)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
