"""Row e2: the path's only collective is DDP's gradient all-reduce during backward
(scripts/train.py:157-158, src/trainers/gan_pose_trainer.py:141).  `parallel.RenderModule` is the Generator-shaped
container of the render path; wrapped in DistributedDataParallel its post-backward `.grad` must equal the MEAN over
ranks of the per-rank gradients.

CPU: world_size-2 gloo group, the differentiable torch formulation as the renderer (host logic + reducer hooks).
GPU (-m gpu, needs >= 2 devices): world_size-2 NCCL group, the CUDA renderer and `oi_render_backward` writing the
gradients from inside a custom autograd.Function.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torch_render_fn(mod, rays_o, rays_d, near, far, w, cos_anneal_ratio, perturb_overwrite):
    from object_intrinsics_b200 import torch_graph
    return torch_graph.render_differentiable(mod, rays_o, rays_d, near, far, w, cos_anneal_ratio, None)


def _cpu_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import load_params
    from oracle import neus_oracle as O
    from object_intrinsics_b200.parallel import RenderModule, check_ddp_gradients, rank_seed
    from test_torch_graph import _nets
    torch.manual_seed(0)                                   # same replica on every rank
    P = load_params("params_D4.npz")

    def build():
        sdf, col, dev = _nets(P, 4, torch.float32)
        # CPU stand-in for the style MLP (the real one is a CUDA op): still reaches the FiLM linears through w
        sdf.style = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.LeakyReLU(0.2))
        return RenderModule(sdf, col, dev, n_samples=16, n_importance=0, render_fn=_torch_render_fn)
    local = build()
    wrapped = build()
    wrapped.load_state_dict(local.state_dict())
    ddp = torch.nn.parallel.DistributedDataParallel(wrapped)
    seed = rank_seed(1234, rank)                           # every rank renders ITS OWN instance (scripts/train.py:136)
    ro, rd, near, far = O.synthetic_rays(1, 4, seed=seed)
    z = torch.randn(1, 64, generator=torch.Generator().manual_seed(seed))
    err, n = check_ddp_gradients(ddp, local, (ro, rd, near, far, z))
    n_params = sum(p.numel() for p in local.parameters())
    q.put((rank, err, n, n_params, float(z.sum())))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(worker, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23000 + os.getpid() % 4000
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


def test_ddp_gradients_are_the_rank_mean_gloo_world_size_2():
    res = _spawn(_cpu_worker, 2)
    assert res[0][4] != res[1][4]                          # the two ranks really rendered different instances
    for _, err, n, n_params, _ in res:
        assert n == n_params > 100_000
        assert err < 1e-5, err


def _gpu_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.parallel import RenderModule, check_ddp_gradients, nccl_nvlink_env, rank_seed
    nccl_nvlink_env()
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from helpers import load_params
    from oracle import neus_oracle as O
    P = load_params("params_D8.npz")

    def build():
        sdf, col, devn = fields.build_networks(D=8, device=dev)
        fields.load_flat_params(sdf, col, devn, P)
        return RenderModule(sdf, col, devn, n_samples=64, n_importance=0)
    local, wrapped = build(), build()
    ddp = torch.nn.parallel.DistributedDataParallel(wrapped, device_ids=[rank])   # as scripts/train.py:157-158
    seed = rank_seed(1234, rank)
    ro, rd, near, far = (t.to(dev) for t in O.synthetic_rays(1, 16, seed=seed))
    z = torch.randn(1, 64, generator=torch.Generator().manual_seed(seed)).to(dev)
    err, n = check_ddp_gradients(ddp, local, (ro, rd, near, far, z))
    # a second step: gradients are re-reduced (bucket views are reused) and still the rank mean
    err2, _ = check_ddp_gradients(ddp, local, (ro, rd, near, far, z))
    torch.cuda.synchronize()
    q.put((rank, max(err, err2), n, sum(p.numel() for p in local.parameters()), float(z.sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_ddp_gradients_are_the_rank_mean_nccl_world_size_2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); the world_size-2 gloo test covers the host logic")
    res = _spawn(_gpu_worker, 2)
    for _, err, n, n_params, _ in res:
        assert n == n_params
        # both passes run the same kernels on the same inputs; only the atomics' summation order and the all-reduce
        # rounding differ
        assert err < 2e-3, err
