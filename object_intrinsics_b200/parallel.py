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


# ---------------------------------------------------------------------------------------------------------------
# Training-side data parallelism: the reference wraps its Generator in
# DistributedDataParallel (scripts/train.py:157-158) and the gradient all-reduce fires inside
# `loss_final.backward()` (src/trainers/gan_pose_trainer.py:141).  `RenderModule` is the part of that Generator the
# render path owns -- the three networks registered as children under the reference's names, plus the renderer --
# so that wrapping it in DDP exercises exactly that exchange: per rank one grad-mode render of its own instances,
# `oi_render_backward` writes the parameter gradients, DDP's reducer hooks all-reduce them over NCCL/NVLink.
# ---------------------------------------------------------------------------------------------------------------
def nccl_nvlink_env() -> None:
    """Single-node NVLink/NVSwitch only: no InfiniBand / socket transports for the gradient all-reduce."""
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


class RenderModule(torch.nn.Module):
    """`Generator`-shaped container of the render path (children named as in src/models/generator.py:45-47).

    forward(rays_o, rays_d, near, far, z, ...) = `w = sdf_network.style(z)` (generator.py:237) followed by
    `renderer.render(...)` (generator.py:245-252); returns the renderer's dict.  `render_fn(module, rays_o, rays_d,
    near, far, w, cos_anneal_ratio, perturb_overwrite)` replaces the CUDA renderer when given (CPU gloo tests use the
    differentiable torch formulation)."""

    def __init__(self, sdf_network, color_network, deviation_network, n_samples=64, n_importance=0, render_fn=None,
                 **renderer_kwargs):
        super().__init__()
        self.sdf_network, self.color_network, self.deviation_network = sdf_network, color_network, deviation_network
        self.n_samples, self.n_importance, self.up_sample_steps = n_samples, n_importance, 1
        self.render_fn = render_fn
        self.renderer = None
        if render_fn is None:
            from .renderer import NeuSRenderer
            self.renderer = NeuSRenderer(nerf=None, sdf_network=sdf_network, deviation_network=deviation_network,
                                         color_network=color_network, n_samples=n_samples,
                                         n_importance=n_importance, n_outside=0, up_sample_steps=1, perturb=0,
                                         **renderer_kwargs)

    def forward(self, rays_o, rays_d, near, far, z, cos_anneal_ratio=1.0, perturb_overwrite=0):
        w = self.sdf_network.style(z)
        if self.render_fn is not None:
            return self.render_fn(self, rays_o, rays_d, near, far, w, cos_anneal_ratio, perturb_overwrite)
        return self.renderer.render(rays_o, rays_d, near, far, cos_anneal_ratio=cos_anneal_ratio,
                                    perturb_overwrite=perturb_overwrite, z=z, w=w)


def training_loss(out):
    """A generator-shaped scalar of the render outputs the trainer consumes (image, mask, eikonal term)."""
    img = out["color_fine"] + (1.0 - out["weight_sum"])
    return (img ** 2).mean() + 0.5 * out["weight_sum"].mean() + 0.1 * out["gradient_error"]


def flat_grads(module: torch.nn.Module) -> torch.Tensor:
    """All parameter gradients of `module` as one flat fp64 vector (parameters() order; missing grads count as 0)."""
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().double().reshape(-1)
                      for p in module.parameters()])


def check_ddp_gradients(ddp_module, local_module, inputs, group=None) -> Tuple[float, int]:
    """Runs one backward through `ddp_module` (all-reduced gradients) and one through the un-wrapped replica
    `local_module` (this rank's own gradients), all-gathers the local ones and returns
    (max |g_ddp - mean_r g_r| / max |mean_r g_r|, number of gradient elements).  Both modules must hold the same
    parameter values."""
    for m in (ddp_module, local_module):
        for p in m.parameters():
            p.grad = None
    training_loss(ddp_module(*inputs)).backward()
    training_loss(local_module(*inputs)).backward()
    g_ddp = flat_grads(ddp_module.module if hasattr(ddp_module, "module") else ddp_module)
    g_loc = flat_grads(local_module)
    world = dist.get_world_size(group)
    parts = [torch.empty_like(g_loc) for _ in range(world)]
    dist.all_gather(parts, g_loc, group=group)
    mean = torch.stack(parts).mean(0)
    scale = float(mean.abs().max()) + 1e-300
    return float((g_ddp - mean).abs().max()) / scale, g_ddp.numel()
