"""CUDA replacements for the two `Generator` methods either side of the render call.

Both functions take the reference `Generator` instance (or anything with the same attributes) as their first
argument and keep the reference methods' signatures and return dicts, so that they can be bound in place:

    import object_intrinsics_b200.generator_ops as G
    Generator.gen_rays_at = G.gen_rays_at          # src/models/generator.py:255-279 (+ build_rays :317-333)
    Generator.render_maps = G.render_maps          # src/models/generator.py:80-174 (+ lighting.py:126-225)

They are forward-only kernels: under autograd (the generator step) the reference's own torch code must be used
-- `render_maps` raises if any input requires grad while grad mode is on.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def gen_rays_at(generator, data, prior_info: Dict, with_near_far: bool = False) -> Dict[str, torch.Tensor]:
    """rays_o, rays_d [bs,P,P,3], x_offset, y_offset [bs] (and near/far [bs*P*P,1] with `with_near_far`)."""
    cam = generator.camera
    b2w, c2b = prior_info["b2w"], prior_info["c2b"]
    if not b2w.is_cuda:
        raise RuntimeError("gen_rays_at: CUDA tensors only (object_intrinsics_b200 has no CPU path)")
    dev = b2w.device
    bs, P = b2w.shape[0], int(generator.resolution)
    rays_o = torch.empty(bs, P, P, 3, device=dev)
    rays_d = torch.empty(bs, P, P, 3, device=dev)
    xo, yo = torch.empty(bs, device=dev), torch.empty(bs, device=dev)
    d = _lib.OiGenRaysDesc()
    d.n_instances, d.resolution, d.scene_resolution = bs, P, int(generator.scene_resolution)
    d.cam_dist = float(cam.cam_dist)
    keep = [_f32c(b2w, dev), _f32c(c2b, dev), _f32c(cam.w2c, dev), _f32c(cam.intrinsics_inv, dev)]
    d.b2w, d.c2b, d.w2c, d.intrinsics_inv = [t.data_ptr() for t in keep]
    d.rays_o, d.rays_d, d.x_offset, d.y_offset = rays_o.data_ptr(), rays_d.data_ptr(), xo.data_ptr(), yo.data_ptr()
    out = {"rays_o": rays_o, "rays_d": rays_d, "x_offset": xo, "y_offset": yo}
    if with_near_far:
        out["near"] = torch.empty(bs * P * P, 1, device=dev)
        out["far"] = torch.empty(bs * P * P, 1, device=dev)
        d.near, d.far = out["near"].data_ptr(), out["far"].data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().oi_gen_rays(C.byref(d), _lib.current_stream_ptr(dev)), "oi_gen_rays")
    return out


_RAW_KEYS = ("amb_shading_map", "diff_shading_map", "normal_map", "no_specular_map", "specular_map")


def render_maps(generator, bs, render_out, rays_info, prior_info, return_raw):
    """Phong shading of every sample + per-ray compositing of all maps in one kernel."""
    light = prior_info["light"]            # BatchDirectionalLightWithSpecularFixInit (lighting.py:79-119)
    weights = render_out["weights"]
    if not weights.is_cuda:
        raise RuntimeError("render_maps: CUDA tensors only (object_intrinsics_b200 has no CPU path)")
    if torch.is_grad_enabled() and any(render_out[k].requires_grad for k in ("weights", "gradients", "raw_color")):
        raise RuntimeError("generator_ops.render_maps is forward-only; use the reference render_maps under autograd")
    dev = weights.device
    P = int(generator.resolution)
    R, S = weights.shape
    base = light.light
    w2b = light.w2b
    direction = (base.param_direction / torch.linalg.norm(base.param_direction)).detach()
    light_dir = _f32c(torch.einsum("bij,j->bi", w2b[:, :3, :3].detach(), direction), dev)   # lighting.py:115-119
    bg = generator.bg_color(bs)            # [bs,3,h,w] expand of a per-instance colour (utils/prior.py:12-29)
    bg_color = _f32c(bg[:, :, 0, 0], dev)
    rays_o = _f32c(rays_info["rays_o"].reshape(-1, 3), dev)
    ins = {k: _f32c(render_out[k], dev) for k in ("weights", "gradients", "raw_color", "pts", "weight_sum",
                                                  "color_fine")}
    d = _lib.OiRenderMapsDesc()
    d.n_rays, d.rays_per_instance, d.n_samples = R, R // bs, S
    # ONE device->host read for the ten light scalars (ten float(tensor[i]) calls would be ten blocking syncs)
    lp = torch.cat([base.ambient_color.detach().reshape(3).float(), base.diffuse_color.detach().reshape(3).float(),
                    base.specular_color.detach().reshape(3).float(),
                    torch.as_tensor(base.shininess, dtype=torch.float32, device=base.ambient_color.device).reshape(1)
                    ]).tolist()
    d.shininess = lp[9]
    for i in range(3):
        d.ambient_color[i], d.diffuse_color[i], d.specular_color[i] = lp[i], lp[3 + i], lp[6 + i]
    for k, t in ins.items():
        setattr(d, k, t.data_ptr())
    d.rays_o, d.light_dir, d.bg_color = rays_o.data_ptr(), light_dir.data_ptr(), bg_color.data_ptr()
    ret = {}

    def new(name, c):
        ret[name] = torch.empty(bs, c, P, P, device=dev)
        setattr(d, name, ret[name].data_ptr())

    new("weight_sum_map", 1)
    new("color_map", 3)
    if return_raw:
        new("amb_shading_map", 3)
        new("diff_shading_map", 3)
    new("shading_map", 3)
    if return_raw:
        new("normal_map", 3)
        new("no_specular_map", 3)
        new("specular_map", 3)
    new("image_no_bg", 3)
    new("image", 3)
    new("mask", 1)
    zmin = None
    if return_raw:
        mid = _f32c(render_out["mid_z_vals"], dev)
        d.mid_z_vals = mid.data_ptr()
        new("z_map", 1)
        zmin = torch.empty(R, device=dev)
        d.z_min_per_ray = zmin.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().oi_render_maps(C.byref(d), _lib.current_stream_ptr(dev)), "oi_render_maps")
    if return_raw:
        ret["z_min"] = zmin.reshape(bs, -1).min(-1).values
    render_out.pop("gradients", None)      # the reference deletes these two from the dict (generator.py:126,129)
    render_out.pop("pts", None)
    return ret
