"""Synthetic inputs of the benchmark legs (bench.py, tools_*.py).  Bench-side helper: imports nothing from oracle/.

SURVEY.md 8(d) "Synthetic inputs": per object instance a pinhole camera at `cam_dist = 1/tan(5 deg)` looking at the
origin from a seeded random direction, a P x P grid of unit ray directions spanning the 10-degree frustum, `rays_o`
constant within the instance, `near/far = -(o.d)/(d.d) -+ 1` (src/models/generator.py:255-279, 317-342); rays of
instance b occupy rows [b P^2, (b+1) P^2).  SDF weights: checkpoints/sphere_init.pt as committed in
tests/golden/params_D8.npz (data written from the reference's state_dict by oracle/gen_golden.py); ~20 % of the
rays hit the radius-0.5 sphere.  tests/test_oracle.py checks that this generator and the oracle's agree bit for bit.
"""
import math
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
CAM_DIST = 11.430052


def load_flat_params(fname="params_D8.npz"):
    with np.load(os.path.join(ROOT, "tests", "golden", fname)) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def camera_frame(gen):
    """Orthonormal (forward, right, up) of a camera looking at the origin from a random direction."""
    v = torch.randn(3, generator=gen, dtype=torch.float64)
    fwd = -v / v.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    if abs(float(fwd @ up)) > 0.95:
        up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    return fwd, right, torch.linalg.cross(right, fwd)


def synthetic_rays(bs, patch, seed=1234, cam_dist=CAM_DIST, fov_deg=10.0):
    """rays_o, rays_d [bs*patch^2, 3], near, far [bs*patch^2, 1] (fp32, CPU)."""
    gen = torch.Generator().manual_seed(seed)
    half = math.tan(math.radians(fov_deg / 2))
    lin = torch.linspace(-half, half, patch, dtype=torch.float64)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    origins, dirs = [], []
    for _ in range(bs):
        fwd, right, up = camera_frame(gen)
        d = fwd[None, None] + xx[..., None] * right[None, None] + yy[..., None] * up[None, None]
        dirs.append((d / d.norm(dim=-1, keepdim=True)).reshape(-1, 3))
        origins.append((-fwd * cam_dist).expand(patch * patch, 3))
    rays_o = torch.cat(origins).float().contiguous()
    rays_d = torch.cat(dirs).float().contiguous()
    mid = -(rays_o * rays_d).sum(-1, keepdim=True) / (rays_d * rays_d).sum(-1, keepdim=True)   # generator.py:336-342
    return rays_o, rays_d, mid - 1.0, mid + 1.0


def latent(bs, seed):
    return torch.randn(bs, 64, generator=torch.Generator().manual_seed(seed))
