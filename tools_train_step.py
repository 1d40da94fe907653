"""A training-shaped composition of the path's kernels -- no trainer code, no discriminator networks.

Mirrors the launch sequence of one `GANPoseTrainer.train_step` (src/trainers/gan_pose_trainer.py:77-101) as far
as the render path and its two neighbours reach:

    G step   gen_rays_at -> grad-mode render -> render_maps (shading + compositing) -> loss -> backward (:104-141)
    D step   no-grad gen_rays_at + render + render_maps, AugmentPipe on fake and real image            (:85-87)
    mask-D   no-grad gen_rays_at + render + render_maps, AugmentPipe on fake and real mask             (:89-91)
    (the G step itself pushes its image and its mask through both pipes: 6 AugmentPipe forwards per step)

The discriminator convolutions, optimisers and the pose prior are out of scope; a quadratic loss on
the augmented images stands in for the discriminator logits so that the adjoint of the augmentation and the whole
render backward run.  Shapes: `cfg5` = 4 instances x 64x64 patch x 64 samples, `cfg3` = the shipped training
config, 1 instance x 128x128 patch x (16 + 4 hierarchical) samples.

    python tools_train_step.py [--shape cfg5|cfg3] [--steps 10]      # prints one JSON object
"""
import json
import math
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_inputs as BI  # noqa: E402

SHAPES = {"cfg5": dict(bs=4, patch=64, n_samples=64, n_importance=0),
          "cfg3": dict(bs=1, patch=128, n_samples=16, n_importance=4)}


def _pose(bs, seed, device):
    """b2w / c2b [bs,4,4]: identity object pose, camera on a seeded random viewing direction."""
    gen = torch.Generator().manual_seed(seed)
    c2b = torch.eye(4).repeat(bs, 1, 1)
    for b in range(bs):
        fwd, right, up = BI.camera_frame(gen)
        c2b[b, :3, 0], c2b[b, :3, 1], c2b[b, :3, 2] = right.float(), up.float(), fwd.float()   # camera looks along +z
        c2b[b, :3, 3] = (-fwd * BI.CAM_DIST).float()
    return torch.eye(4).repeat(bs, 1, 1).to(device), c2b.to(device)


def _generator_stand_in(patch, device):
    """The attributes generator_ops reads from the reference Generator (generator.py:28-44, camera.py)."""
    fov = math.radians(10.0)
    f = 0.5 * patch / math.tan(fov / 2)
    K = torch.tensor([[f, 0, patch / 2, 0], [0, f, patch / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
    w2c = torch.eye(4)
    w2c[2, 3] = BI.CAM_DIST
    cam = types.SimpleNamespace(cam_dist=BI.CAM_DIST, w2c=w2c.to(device), c2w=torch.linalg.inv(w2c).to(device),
                                intrinsics_inv=torch.linalg.inv(K).to(device))
    g = types.SimpleNamespace(camera=cam, resolution=patch, scene_resolution=patch)
    g.bg_color = lambda n: torch.ones(n, 3, patch, patch, device=device)
    return g


def _light(bs, device):
    base = types.SimpleNamespace(param_direction=torch.tensor([0.3, 0.2, -1.0], device=device),
                                 ambient_color=torch.full((3,), 0.4, device=device),
                                 diffuse_color=torch.full((3,), 0.6, device=device),
                                 specular_color=torch.full((3,), 0.2, device=device),
                                 shininess=torch.tensor(8.0, device=device))
    return types.SimpleNamespace(light=base, w2b=torch.eye(4, device=device).repeat(bs, 1, 1))


def _shade_torch(out, rays_d, light, bs, patch):
    """Differentiable stand-in of `Generator.render_maps` (generator.py:80-174, lighting.py:126-225) for the G
    step: per-sample Phong shading composited with the render weights -> image [bs,3,P,P], mask [bs,1,P,P]."""
    base = light.light
    ldir = base.param_direction / base.param_direction.norm()
    n = torch.nn.functional.normalize(out["gradients"], dim=-1)
    ndl = (n * (-ldir)).sum(-1, keepdim=True).clamp(min=0)
    view = -torch.nn.functional.normalize(rays_d, dim=-1)[:, None, :]
    half = torch.nn.functional.normalize(view - ldir, dim=-1)
    spec = (n * half).sum(-1, keepdim=True).clamp(min=0) ** base.shininess
    shaded = out["raw_color"] * (base.ambient_color + base.diffuse_color * ndl) + base.specular_color * spec
    img = (out["weights"][..., None] * shaded).sum(1) + (1.0 - out["weight_sum"])
    return (img.reshape(bs, patch, patch, 3).permute(0, 3, 1, 2).contiguous(),
            out["weight_sum"].reshape(bs, patch, patch, 1).permute(0, 3, 1, 2).contiguous())


def measure(dev, P, kernel="auto", flush=None, shape="cfg5", steps=8, shading="kernel"):
    from object_intrinsics_b200 import fields, generator_ops
    from object_intrinsics_b200.augment import AugmentPipe
    from object_intrinsics_b200.renderer import NeuSRenderer
    cfg = SHAPES[shape]
    bs, patch = cfg["bs"], cfg["patch"]
    sdf, col, devn = fields.build_networks(D=8, device=dev)
    fields.load_flat_params(sdf, col, devn, P)
    renderer = NeuSRenderer(nerf=None, sdf_network=sdf, deviation_network=devn, color_network=col,
                            n_samples=cfg["n_samples"], n_importance=cfg["n_importance"], n_outside=0,
                            up_sample_steps=1, perturb=1, impl=kernel)
    gparams = list(sdf.parameters()) + list(col.parameters()) + list(devn.parameters())
    gen, light = _generator_stand_in(patch, dev), _light(bs, dev)
    aug_img = AugmentPipe(xint=1, scale=1).to(dev)      # configs/train.yaml:80-100
    aug_mask = AugmentPipe(xint=1, scale=1).to(dev)
    real_img = torch.rand(bs, 3, patch, patch, device=dev)
    real_mask = (torch.rand(bs, 1, patch, patch, device=dev) > 0.5).float()
    b2w, c2b = _pose(bs, 7, dev)
    z = torch.randn(bs, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(3))

    def rays():
        r = generator_ops.gen_rays_at(gen, {}, {"b2w": b2w, "c2b": c2b}, with_near_far=True)
        return r["rays_o"].reshape(-1, 3), r["rays_d"].reshape(-1, 3), r["near"], r["far"]

    def render(ro, rd, near, far):
        return renderer.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=sdf.style(z))

    def g_step():
        for p in gparams:
            p.grad = None
        ro, rd, near, far = rays()
        out = render(ro, rd, near, far)
        if shading == "torch":
            img, mask = _shade_torch(out, rd, light, bs, patch)
        else:   # the differentiable oi_render_maps (forward + oi_render_maps_backward inside autograd)
            maps = generator_ops.render_maps(gen, bs, out, {"rays_o": ro}, {"light": light}, False)
            img, mask = maps["image"], maps["mask"]
        loss = (aug_img(img) ** 2).mean() + (aug_mask(mask) ** 2).mean() + 0.1 * out["gradient_error"]
        loss.backward()

    def d_step(pipe, real, key):
        with torch.no_grad():
            ro, rd, near, far = rays()
            # contract B: render + render_maps in one library call, no per-point tensor materialised for the caller
            _, maps = generator_ops.render_and_maps(gen, renderer, bs, {"rays_o": ro, "rays_d": rd, "near": near,
                                                                        "far": far}, {"light": light}, sdf.style(z),
                                                    return_raw=False, cos_anneal_ratio=1.0)
            return pipe(maps[key]), pipe(real)

    def step():
        g_step()
        d_step(aug_img, real_img, "image")
        d_step(aug_mask, real_mask, "mask")

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    parts = {"g_step": [], "d_step_x2": [], "total": []}
    for i in range(steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        g_step()
        ev[1].record()
        d_step(aug_img, real_img, "image")
        d_step(aug_mask, real_mask, "mask")
        ev[2].record()
        ev[2].synchronize()
        parts["g_step"].append(ev[0].elapsed_time(ev[1]))
        parts["d_step_x2"].append(ev[1].elapsed_time(ev[2]))
        parts["total"].append(ev[0].elapsed_time(ev[2]))
    med = {k: sorted(v)[len(v) // 2] for k, v in parts.items()}
    # free-running: `steps` steps back to back as the trainer issues them (it has no host read-back between steps),
    # L2 flush fills in stream and their measured duration subtracted
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize(dev)
    a.record()
    if flush is not None:
        for i in range(steps):
            flush.fill_(i & 0xFF)
    b.record()
    import time
    t0 = time.perf_counter()
    for i in range(steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        step()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps      # host time to ENQUEUE a step
    c.record()
    c.synchronize()
    free_ms = (b.elapsed_time(c) - a.elapsed_time(b)) / steps
    rays_per_step = 3 * bs * patch * patch
    return {"shape": shape, **cfg, "ms_per_step": med["total"], "ms_per_step_free_running": free_ms,
            "host_enqueue_ms_per_step": host_ms,
            "g_step_ms": med["g_step"],
            "two_no_grad_renders_ms": med["d_step_x2"], "rays_per_step": rays_per_step,
            "rays_per_sec": rays_per_step / (med["total"] * 1e-3),
            "shading": shading,
            "what": "gen_rays -> grad render -> Phong shading maps (oi_render_maps + oi_render_maps_backward, or a torch "
                    "stand-in with shading='torch') -> 2 AugmentPipe -> quadratic loss -> "
                    "backward (oi_render_backward + augment adjoint); then 2 x (gen_rays -> no-grad render -> "
                    "oi_render_maps -> 2 AugmentPipe); mirrors gan_pose_trainer.py:77-101 without the "
                    "discriminator networks / optimisers"}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="both", choices=["cfg5", "cfg3", "both"])
    ap.add_argument("--steps", type=int, default=8)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    P = BI.load_flat_params("params_D8.npz")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {s: measure(dev, P, "auto", flush, s, a.steps) for s in (("cfg5", "cfg3") if a.shape == "both" else (a.shape,))}
    print(json.dumps(res))
