"""Multi-GPU plumbing for the render path: one process per GPU, object instances sharded across ranks.

The reference's only parallelism is DDP over object instances (scripts/train.py:50-56,157-158; tu/ddp.py): every
rank renders its own instances with a full weight replica, and the forward path has NO collective (rays and
instances are independent).  The only exchange is the gradient all-reduce that DDP fires during backward, which
stays torch's NCCL all-reduce.  This module holds the host-side logic that is independent of CUDA so that it can
be covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_instances(n_instances_total: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced shard of instance indices for `rank` (the first `rem` ranks get one extra)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(n_instances_total, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def rank_seed(base_seed: int, rank: int) -> int:
    """Per-rank RNG seed, as the reference does (`set_seed_benchmark(seed + rank)`, scripts/train.py:136)."""
    return base_seed + rank


def aggregate_throughput(units_local: float, seconds_local: float, group=None) -> Tuple[float, float, float]:
    """Whole-job throughput = (sum of units over ranks) / (max of elapsed time over ranks).
    Returns (throughput, total_units, max_seconds).  Works on any backend (gloo on CPU, nccl on GPU)."""
    if not (dist.is_available() and dist.is_initialized()):
        return units_local / seconds_local, units_local, seconds_local
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([seconds_local], dtype=torch.float64, device=dev)
    u = torch.tensor([units_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u) / float(t), float(u), float(t)
