"""Sample-sharded multi-GPU execution: one process per GPU, each rank holds N/G rows of the samples.

The per-sample outputs (log q, rho / r) stay with their shard; the only data that crosses GPUs is the
statistics packet of ``mix_adapt/_stats.py`` (~127 KB at K=32, D=30) -- ONE all-reduce (sum) per update,
NCCL over NVLink when the packet is a CUDA tensor, gloo for the CPU tests.  After it every rank holds
identical bits and applies the identical host update, so no broadcast follows.  This replaces the
gather-to-root / update-on-root / broadcast pattern of the reference's MPI example
(pypmc/tools/parallel_sampler.py:58-66, examples/pmc_mpi.py:92-131).
"""
from __future__ import annotations

import os

_group = None
_enabled = False


def enable(group=None):
    """Turn on the all-reduce of update statistics over ``group`` (default: the world group).
    ``torch.distributed`` must be initialised (see ``init_from_env``)."""
    import torch.distributed as dist
    global _group, _enabled
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable():
    global _group, _enabled
    _group, _enabled = None, False


def enabled() -> bool:
    return _enabled


def init_from_env(backend: str = "nccl"):
    """Initialise ``torch.distributed`` from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT), bind this process to its GPU and enable the statistics all-reduce."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        else:
            dist.init_process_group(backend=backend)
    if world > 1:
        enable()
    return rank(), world_size()


def rank() -> int:
    import torch.distributed as dist
    return dist.get_rank(_group) if dist.is_available() and dist.is_initialized() else 0


def world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size(_group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_(packet):
    """In-place sum of the statistics packet over all ranks (no-op unless ``enable`` was called).
    ``packet`` is a torch tensor (CUDA -> NCCL, CPU -> gloo)."""
    if not _enabled:
        return packet
    import torch.distributed as dist
    dist.all_reduce(packet, op=dist.ReduceOp.SUM, group=_group)
    return packet


def shard_rows(n_total: int, r: int = None, world: int = None):
    """Contiguous row block [start, stop) of rank ``r`` out of ``world`` (SURVEY 8e partitioning)."""
    r = rank() if r is None else r
    world = world_size() if world is None else world
    base, rem = divmod(n_total, world)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)
