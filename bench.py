#!/usr/bin/env python
"""bench.py -- throughput of the mixture-density hot path on B200 (contract: see the task statement).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

Workload (BASELINE.json configs[1]): ``MixtureDensity.multi_evaluate`` of N = 1e7 float64 samples per GPU under a
K = 32 component, D = 30 Gaussian mixture.  A step is one pass over the batch.  Metric: sample-component
evaluations per second (N*K/s), whole job.  With several GPUs the samples are sharded (weak scaling, 1e7 per GPU);
``multi_evaluate`` has no exchange step, so the headline has no data-path collective.  The path's ONE collective --
the all-reduce of the per-component statistics packet inside a proposal update (replaces the gather /
update-on-root / bcast of examples/pmc_mpi.py:83-131) -- is measured in the same process and reported in the
``update`` record of every line: K1(rho) + K2 + all-reduce + host finish, the all-reduce alone, and one whole
device-born PMC iteration (BASELINE configs[4]: propose -> weigh -> gaussian_pmc, N = 1e7 per GPU).

Reference arm (``--impl reference``): the compiled, UNMODIFIED reference (pypmc 1.2.6 Cython, installed from
/root/reference into baseline/_ref) evaluating the same N x K workload through its own public API
(``pypmc.density.mixture.MixtureDensity.multi_evaluate``), rows sharded over one process per host core (its own
scaling pattern: one process per rank over sample shards, pypmc/tools/parallel_sampler.py:58-66).  If baseline/_ref
cannot be imported the oracle port (oracle/pmc_oracle.c) takes its place and ``cpu_baseline.kind`` says "port".
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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

N_PER_GPU, K, D = 10_000_000, 32, 30
FLOP_PER_PAIR = D * D + 4 * D            # D subtract + D(D+1) triangular FMA flops + 2D square-accumulate
BYTES_PER_SAMPLE = 8 * D + 8             # read x, write log q
PARITY_ROWS = 1_000_000                  # rows of the bench batch compared with the oracle in every run (task 1)
PARITY_TOL = 1e-10                       # BASELINE.json north_star
METRIC = "sample-component evals/sec (N*K/s)"
WORKLOAD = "MixtureDensity.multi_evaluate N=1e7/GPU K=32 D=30 Gaussian (BASELINE configs[1])"


def config_dict(n_per_gpu, world):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "N_per_gpu": n_per_gpu, "K": K, "D": D,
            "l2": "inputs (2.4 GB/GPU) larger than L2, no flush needed",
            "parallelism": "samples sharded over %d GPU(s); multi_evaluate has no data-path collective "
                           "(the update's one all-reduce is timed in `update`)" % world}


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
    chol = np.linalg.cholesky(covs)
    x = np.empty((n, D))
    slab = 500_000
    for s in range(0, n, slab):
        m = min(slab, n - s)
        comp = rng.integers(0, K, size=m)
        x[s:s + m] = means[comp] + np.einsum("nij,nj->ni", chol[comp], rng.normal(size=(m, D)))
    return x


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


# --------------------------------------------------------------------------------------------- CPU arms
def oracle_pass(x, comps, weights, threads):
    """One multi_evaluate pass of the oracle (CPU restatement of the reference algorithm, oracle/pmc_oracle.c) over
    ``x``, rows sharded over ``threads`` host threads (ctypes releases the GIL).  Returns log q."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    bounds = np.linspace(0, len(x), threads + 1).astype(int)

    def work(i):
        lq, _ = orc.mixture_multi_evaluate(x[bounds[i]:bounds[i + 1]], comps, weights)
        return lq

    with ThreadPoolExecutor(max_workers=threads) as ex:
        return np.concatenate(list(ex.map(work, range(threads))))


def oracle_arm(rows, steps, warmup, threads, x):
    """(pairs/s, s per pass, log q of the last pass) of the oracle port over the first ``rows`` rows of x."""
    from oracle import oracle as orc
    orc.build()
    means, covs, w = synth_mixture()
    comps = orc.Components(means, covs)
    x = x[:rows]
    lq = None
    for _ in range(warmup):
        oracle_pass(x, comps, w, threads)
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        lq = oracle_pass(x, comps, w, threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rows * K / dt, dt, lq


_REF_STATE = {}


def _ref_worker_init(shm_name, n_rows):
    """Pool worker: import the compiled reference from baseline/_ref (never /root/reference), map the shared samples."""
    import logging
    from multiprocessing import shared_memory
    sys.path.insert(0, REF_DIR)
    import pypmc                                        # noqa: F401  (installs a stdout log handler: silence it)
    logging.getLogger("pypmc").handlers.clear()
    from pypmc.density.mixture import create_gaussian_mixture
    means, covs, w = synth_mixture()
    shm = shared_memory.SharedMemory(name=shm_name)
    _REF_STATE["shm"] = shm
    _REF_STATE["x"] = np.ndarray((n_rows, D), dtype=np.float64, buffer=shm.buf)
    _REF_STATE["mix"] = create_gaussian_mixture(means, covs, w)


def _ref_worker_eval(bounds):
    lo, hi = bounds
    x = _REF_STATE["x"][lo:hi]
    return _REF_STATE["mix"].multi_evaluate(x)          # pypmc/density/mixture.pyx:112-156, the reference's own loops


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "pypmc"))


class ReferencePool:
    """The unmodified compiled reference on ``procs`` host processes over row shards of one shared sample matrix."""

    def __init__(self, x, procs):
        import multiprocessing as mp
        from multiprocessing import shared_memory
        self.rows, self.procs = len(x), procs
        self.shm = shared_memory.SharedMemory(create=True, size=max(x.nbytes, 8))
        np.ndarray(x.shape, dtype=np.float64, buffer=self.shm.buf)[:] = x
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_ref_worker_init, initargs=(self.shm.name, self.rows))

    def evaluate(self, rows=None):
        rows = self.rows if rows is None else rows
        chunks = max(self.procs, 1) * 4                  # a few shards per process: the tail of the slowest one stays short
        b = np.linspace(0, rows, chunks + 1).astype(int)
        parts = self.pool.map(_ref_worker_eval, [(int(b[i]), int(b[i + 1])) for i in range(chunks)], chunksize=1)
        return np.concatenate(parts)

    def close(self):
        self.pool.close()
        self.pool.join()
        self.shm.close()
        self.shm.unlink()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.rows or N_PER_GPU
    means, covs, w = synth_mixture()
    use_ref = reference_available() and not args.port
    # size the step: every step evaluates all N rows unless that would push the whole run beyond ~10 minutes on this
    # host (5.5 s per pass on the 32-vCPU boxes of this pool); the probe below decides, the sample says what ran
    probe_rows = min(n, 20_000 * threads)
    x = synth_samples_host(n, means, covs)
    if use_ref:
        pool = ReferencePool(x, threads)
        pool.evaluate(min(1000 * threads, n))            # import + first touch
        t0 = time.perf_counter()
        pool.evaluate(probe_rows)
        rate = probe_rows / (time.perf_counter() - t0)
        evaluate = pool.evaluate
        kind = "reference"
        what = "pypmc 1.2.6 (compiled Cython, baseline/_ref) MixtureDensity.multi_evaluate, %d processes" % threads
    else:
        from oracle import oracle as orc
        orc.build()
        comps = orc.Components(means, covs)
        t0 = time.perf_counter()
        oracle_pass(x[:probe_rows], comps, w, threads)
        rate = probe_rows / (time.perf_counter() - t0)
        evaluate = lambda rows: oracle_pass(x[:rows], comps, w, threads)
        kind = "port"
        what = "oracle/pmc_oracle.c (port of the reference loops), %d threads" % threads
    budget_s = 600.0
    rows = int(min(n, max(probe_rows, rate * budget_s / max(args.steps + args.warmup, 1))))
    for _ in range(args.warmup):
        evaluate(rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        evaluate(rows)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    if use_ref:
        pool.close()
    value = rows * K / dt
    sample = ("all %d rows per step" % n) if rows == n else ("%d of %d rows per step (bounded sample of the same workload)" % (rows, n))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(n, args.gpus),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample + ", " + what},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_cython_one_core(args):
    """Child process of the GPU arm: the compiled reference on ONE core (as shipped it is single-threaded) over
    ``--rows`` rows; prints {"pairs_per_s", "rows", "s"} or {"unavailable": why}."""
    if not reference_available():
        print(json.dumps({"unavailable": "baseline/_ref not present"}))
        return
    import logging
    sys.path.insert(0, REF_DIR)
    try:
        import pypmc  # noqa: F401
        logging.getLogger("pypmc").handlers.clear()
        from pypmc.density.mixture import create_gaussian_mixture
    except Exception as exc:                                 # pragma: no cover
        print(json.dumps({"unavailable": "import failed: %r" % (exc,)}))
        return
    means, covs, w = synth_mixture()
    rows = args.rows or 100_000
    x = synth_samples_host(rows, means, covs)
    mix = create_gaussian_mixture(means, covs, w)
    mix.multi_evaluate(x[:2000])
    t0 = time.perf_counter()
    lq = mix.multi_evaluate(x)
    dt = time.perf_counter() - t0
    from oracle import oracle as orc                          # the oracle against the real reference, same rows
    orc.build()
    lq_o, _ = orc.mixture_multi_evaluate(x, orc.Components(means, covs), w)
    print(json.dumps({"pairs_per_s": rows * K / dt, "rows": rows, "s": dt,
                      "oracle_vs_reference_max_rel": float(np.max(np.abs(lq_o - lq) / np.abs(lq)))}))


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
    from pypmc_b200 import _lib, parallel
    from pypmc_b200.density.mixture import create_gaussian_mixture
    from pypmc_b200.mix_adapt.pmc import gaussian_pmc, DeviceSamples
    from pypmc_b200.mix_adapt._stats import PacketLayout
    if world > 1:
        parallel.enable()

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

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

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
    kernel_name = ctx.last_k1_kernel()                     # which of the three K1 forms did the work (read from the device flags)
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clk = clocks.stop(t_begin, t_end) if rank == 0 else None
    ms_per_step = max_over_ranks(total_ms) / args.steps
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
    host_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    host_value = world * e2e_rows * K / host_s
    e2e_ok = bool(torch.equal(outh_t[:1000].to(device), logq[:1000]))
    host_stream = {"value": host_value, "unit": "pairs/s", "h2d_bytes_per_step": 8 * D * e2e_rows,
                   "d2h_bytes_per_step": 8 * e2e_rows, "rows": e2e_rows, "s_per_step": host_s,
                   "matches_device_result": e2e_ok, "host_buffers": "pinned (torch pin_memory)",
                   "what": "MixtureDensity.multi_evaluate(ndarray) -> ndarray: samples streamed from host memory, log q back"}
    pageable_s = None
    if rank == 0 and world == 1:                          # an ordinary (pageable) ndarray: threaded pinned staging inside the library
        xp = np.array(xh[:min(e2e_rows, 2_000_000)])
        mix.multi_evaluate(xp)
        t0 = time.perf_counter()
        mix.multi_evaluate(xp)
        pageable_s = (time.perf_counter() - t0) * (e2e_rows / len(xp))
        host_stream["pageable_s_per_step_extrapolated"] = pageable_s
        del xp

    # ---- the update and its one collective, same process, same samples ----
    lay = PacketLayout(K, D)
    g = torch.Generator(device=device).manual_seed(11 + rank)
    sw = torch.rand(n, dtype=torch.float64, device=device, generator=g) + 0.5
    ds = DeviceSamples(x, sw)
    upd_reps = max(3, min(args.steps, 5))
    gaussian_pmc(ds, mix)                                  # warm-up
    sync_all()
    upd = []
    for _ in range(upd_reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        new = gaussian_pmc(ds, mix)                        # K1 (rho) + K2 + all-reduce of the packet + host finish
        torch.cuda.synchronize()
        upd.append(time.perf_counter() - t0)
    update_ms = max_over_ranks(float(np.median(upd))) * 1e3
    allreduce_us = None
    if world > 1:                                          # the collective alone: CUDA events around dist.all_reduce of a packet
        pkt = torch.zeros(lay.size, dtype=torch.float64, device=device)
        for _ in range(5):
            dist.all_reduce(pkt)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            a.record()
            dist.all_reduce(pkt)
            b.record()
        torch.cuda.synchronize()
        allreduce_us = max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in evs]))) * 1e3
    # every rank must hold the same updated mixture, bit for bit (no broadcast follows the all-reduce)
    flat = torch.from_numpy(np.concatenate([new.weights] + [c.mu for c in new.components] +
                                           [c.sigma.ravel() for c in new.components])).to(device)
    ref = flat.clone()
    same = torch.ones(1, device=device)
    if world > 1:
        dist.broadcast(ref, src=0)
        same = torch.tensor([1.0 if torch.equal(flat, ref) else 0.0], device=device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)

    # ---- one whole PMC iteration with device-born samples (BASELINE configs[4]): propose -> weigh -> update ----
    target = create_gaussian_mixture(*synth_mixture(seed=3))
    rng = np.random.RandomState(100 + rank)
    param_bytes = [0, 0]

    def pmc_iteration():
        xs = mix.propose_device(n, rng, seed=777, index0=rank * n)    # K3; multinomial counts on the host (mixture.pyx:193)
        logp = target.multi_evaluate(xs)                               # K1 (target)
        dsi = DeviceSamples(xs)
        dsi.weigh(mix, logp)                                           # K1 (proposal): weights AND rho from one evaluation
        return gaussian_pmc(dsi, mix)                                  # K2 + all-reduce + host finish (rho re-used)

    rl = _lib.record_len(D)
    # host -> device per iteration: proposal means + Cholesky factors (K3), proposal and target records + columns (K1), shift (K2);
    # device -> host: the statistics packet (the updated mixture is formed from it on the host)
    param_bytes[0] = 8 * (K * D + K * D * D) + 2 * (8 * K * rl + 4 * K) + 8 * D + 8 * (K + 1)
    param_bytes[1] = 8 * lay.size
    pmc_iteration()
    sync_all()
    its = []
    it_reps = max(3, min(args.steps, 5))
    for _ in range(it_reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pmc_iteration()
        torch.cuda.synchronize()
        its.append(time.perf_counter() - t0)
    iteration_s = max_over_ranks(float(np.median(its)))
    iteration = {"value": world * n * K / iteration_s, "unit": "pairs/s", "h2d_bytes_per_step": param_bytes[0],
                 "d2h_bytes_per_step": param_bytes[1], "rows": n, "s_per_step": iteration_s,
                 "what": "one PMC iteration, samples born on the device: propose_device -> target.multi_evaluate -> "
                         "DeviceSamples.weigh -> gaussian_pmc (K3 + 2 x K1 + K2 + all-reduce + host finish); pairs = N*K per "
                         "iteration; host traffic is the mixture parameters in, the statistics packet out"}
    update = {"update_ms": update_ms, "what": "gaussian_pmc on N=%d device-resident weighted samples per GPU: K1 (rho) + K2 + "
              "all-reduce of the %d-byte packet + host finish; wall clock around the call, median of %d, max over ranks"
              % (n, 8 * lay.size, upd_reps), "allreduce_us": allreduce_us, "packet_bytes": 8 * lay.size,
              "pmc_iteration_ms": iteration_s * 1e3, "ranks_bit_identical": bool(same.item() == 1.0)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: the binding roof is the FP64 pipe, measured live; HBM fraction reported beside it.
    # The kernel issues its FMAs as FP64 matrix instructions (DMMA), so the DMMA probe is the denominator; the DFMA
    # probe (the same units fed through the register file) is reported next to it, and a ~1 s DMMA run (sustained).
    dmma_gflops, _ = ctx.fp64_peak(3, 400)
    dfma_gflops, _ = ctx.fp64_peak(0, 3000)
    sustained_gflops, sustained_ms = ctx.fp64_peak(3, 60000)      # ~1 s per repetition
    kernel_ms = float(np.median(per_step))                # prepare + K1 (the forms that did not get the launch return at once)
    flops = FLOP_PER_PAIR * float(n) * K
    achieved_tf = flops / (kernel_ms * 1e-3) * 1e-12
    peaks, peak_src = measured_peaks()
    hbm_bytes = BYTES_PER_SAMPLE * float(n)
    hbm_gbs = hbm_bytes / (kernel_ms * 1e-3) * 1e-9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")   # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        if tj.get("rows") == n and kernel_name.split("<")[0] in tj.get("kernel", ""):
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = "committed ncu --set full capture (profiles/k1_ncu_traffic.json: %s), not re-measured in this run" % tj.get("capture")
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": dmma_gflops * 1e-3, "unit": "TFLOP/s",
        "frac": achieved_tf / (dmma_gflops * 1e-3), "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": "measured in this process: FP64 matrix-instruction (DMMA) microbenchmark on every SM "
                       "(pmcb200_fp64_peak which=3, burst); MEASURED_PEAKS.json has no FP64 entry",
        "peak_dfma": dfma_gflops * 1e-3, "frac_of_dfma_peak": achieved_tf / (dfma_gflops * 1e-3),
        "peak_sustained": sustained_gflops * 1e-3, "peak_sustained_ms": sustained_ms,
        "kernel": kernel_name, "kernel_source": "device flags of the last launch (pmcb200_last_k1_kernel)",
        "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops,
        "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src},
    }

    # ---- parity, in every run: the oracle over the first PARITY_ROWS rows of the same batch vs the GPU's log q
    threads = os.cpu_count() or 1
    par_rows = int(min(e2e_rows, args.parity_rows))
    cpu_value, cpu_dt, lq_oracle = oracle_arm(par_rows, 1, 0, threads, xh)
    rel = np.abs(outh[:par_rows] - lq_oracle) / np.abs(lq_oracle)
    parity = {"rows": par_rows, "max_rel_logq": float(rel.max()), "tolerance": PARITY_TOL,
              "against": "oracle/pmc_oracle.c (pinned to the compiled reference, tests/test_oracle.py) on the first rows of the bench batch",
              "ok": bool(rel.max() <= PARITY_TOL)}

    # ---- CPU baselines: the oracle port on all host threads (the pass above), and the compiled reference on one core
    one_value, one_dt, _ = oracle_arm(int(min(e2e_rows, 100_000)), 1, 0, 1, xh)
    cpu = {"value": cpu_value, "unit": "pairs/s", "cores": threads, "kind": "port",
           "sample": "first %d of %d rows of the same workload, 1 timed pass (%.1f s), oracle/pmc_oracle.c over %d threads"
                     % (par_rows, n, cpu_dt, threads),
           "value_1_core": one_value, "sample_1_core": "first %d rows, 1 thread (%.1f s)" % (min(e2e_rows, 100_000), one_dt)}
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cython1", "--rows", "100000"],
                             capture_output=True, text=True, timeout=600, env={k: v for k, v in os.environ.items()
                                                                               if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        cy = json.loads(line[-1]) if line else {"unavailable": (out.stderr or "no output")[-300:]}
    except Exception as exc:                                  # pragma: no cover
        cy = {"unavailable": repr(exc)}
    if "pairs_per_s" in cy:
        cpu["reference_cython"] = {"value": cy["pairs_per_s"], "unit": "pairs/s", "cores": 1, "kind": "reference",
                                   "sample": "pypmc 1.2.6 (compiled Cython, baseline/_ref) MixtureDensity.multi_evaluate, "
                                             "%d rows of the same workload, 1 process (%.1f s)" % (cy["rows"], cy["s"]),
                                   "oracle_vs_reference_max_rel": cy["oracle_vs_reference_max_rel"]}
    else:
        cpu["reference_cython"] = cy

    e2e = dict(host_stream) if world == 1 else dict(iteration)
    e2e["kind"] = "host_stream" if world == 1 else "device_born_iteration"
    e2e["note"] = ("N = 1: the public call on HOST arrays (what the reference arm evaluates); N > 1: the device-born PMC iteration "
                   "(host streaming of 2.4 GB per rank cannot scale on one host).  Both records are on every line as "
                   "e2e_host_stream / e2e_iteration: follow ONE of them across N, not `e2e`, for a scaling figure.")
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(n, world),
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "e2e": e2e, "e2e_host_stream": host_stream, "e2e_iteration": iteration, "update": update,
        "gpu_launches": int(launches),
    }))
    if world > 1:
        dist.destroy_process_group()
    if not parity["ok"]:
        sys.stderr.write("bench.py: parity violated: max rel diff of log q %.3e > %.1e on %d rows\n"
                         % (parity["max_rel_logq"], PARITY_TOL, par_rows))
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cython1"])
    ap.add_argument("--port", action="store_true", help="reference arm: time the oracle port even if baseline/_ref is present")
    ap.add_argument("--rows", type=int, default=0, help="samples per GPU (default 1e7, the BASELINE config)")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the host-buffer end-to-end leg (default: all)")
    ap.add_argument("--parity-rows", type=int, default=PARITY_ROWS, help="rows compared with the oracle (default 1e6)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "cython1":
        run_cython_one_core(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
