"""CUDA replacements for the two `Generator` methods either side of the render call.

Both functions take the reference `Generator` instance (or anything with the same attributes) as their first
argument and keep the reference methods' signatures and return dicts, so that they can be bound in place:

    import object_intrinsics_b200.generator_ops as G
    Generator.gen_rays_at = G.gen_rays_at          # src/models/generator.py:255-279 (+ build_rays :317-333)
    Generator.render_maps = G.render_maps          # src/models/generator.py:80-174 (+ lighting.py:126-225)

`gen_rays_at` is forward-only (poses are sampled, rays are constants of the training path); `render_maps` is
differentiable (oi_render_maps_backward) with respect to the render outputs and the light's parameters.
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


_MAPS_BASE = ("weight_sum_map", "color_map", "shading_map", "image_no_bg", "image", "mask")
_MAPS_RAW = ("amb_shading_map", "diff_shading_map", "normal_map", "no_specular_map", "specular_map", "z_map")
_MAP_CHANNELS = {"weight_sum_map": 1, "mask": 1, "z_map": 1}


def _fill_maps_desc(d, n_rays, rays_per_instance, n_samples, tensors):
    d.n_rays, d.rays_per_instance, d.n_samples = n_rays, rays_per_instance, n_samples
    for k, t in tensors.items():
        setattr(d, k, None if t is None else t.data_ptr())


class _RenderMapsFunction(torch.autograd.Function):
    """oi_render_maps / oi_render_maps_backward behind autograd.  Differentiable inputs: weights, gradients,
    raw_color, weight_sum, color_fine, light_params [10], light_dir [bs,3]; the rest are constants of the training
    path (sample positions, camera position, background colour)."""

    @staticmethod
    def forward(ctx, weights, gradients, raw_color, weight_sum, color_fine, light_params, light_dir, pts, mid_z,
                rays_o, bg_color, bs, P, return_raw):
        dev = weights.device
        R, S = weights.shape
        ins = {"weights": _f32c(weights, dev), "gradients": _f32c(gradients, dev), "raw_color": _f32c(raw_color, dev),
               "weight_sum": _f32c(weight_sum, dev), "color_fine": _f32c(color_fine, dev),
               "light_params": _f32c(light_params, dev), "light_dir": _f32c(light_dir, dev), "pts": _f32c(pts, dev),
               "mid_z_vals": None if mid_z is None else _f32c(mid_z, dev), "rays_o": _f32c(rays_o, dev),
               "bg_color": _f32c(bg_color, dev)}
        names = _MAPS_BASE + (_MAPS_RAW if return_raw else ())
        outs = {n: torch.empty(bs, _MAP_CHANNELS.get(n, 3), P, P, device=dev) for n in names}
        zmin = torch.empty(R, device=dev) if return_raw else None
        d = _lib.OiRenderMapsDesc()
        _fill_maps_desc(d, R, R // bs, S, {**ins, **outs, "z_min_per_ray": zmin})
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().oi_render_maps(C.byref(d), _lib.current_stream_ptr(dev)), "oi_render_maps")
        ctx.save_for_backward(*[t for t in ins.values() if t is not None])
        ctx.keys = [k for k, t in ins.items() if t is not None]
        ctx.names, ctx.bs, ctx.P = names, bs, P
        ctx.set_materialize_grads(False)
        ret = tuple(outs[n] for n in names)
        if return_raw:
            ctx.mark_non_differentiable(zmin)
            ret = ret + (zmin,)
        return ret

    @staticmethod
    def backward(ctx, *gouts):
        ins = dict(zip(ctx.keys, ctx.saved_tensors))
        dev = ins["weights"].device
        R, S = ins["weights"].shape
        bd = _lib.OiRenderMapsBwdDesc()
        _fill_maps_desc(bd.fwd, R, R // ctx.bs, S, ins)
        keep = []
        for n, g in zip(ctx.names, gouts):
            if g is not None:
                g = _f32c(g, dev)
                keep.append(g)
                setattr(bd, "g_" + n, g.data_ptr())
        grads = {"d_weights": torch.empty(R, S, device=dev), "d_gradients": torch.empty(R, S, 3, device=dev),
                 "d_raw_color": torch.empty(R, S, 3, device=dev), "d_weight_sum": torch.empty(R, 1, device=dev),
                 "d_color_fine": torch.empty(R, 3, device=dev), "d_light_params": torch.empty(10, device=dev),
                 "d_light_dir": torch.empty(ctx.bs, 3, device=dev)}
        for k, t in grads.items():
            setattr(bd, k, t.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().oi_render_maps_backward(C.byref(bd), _lib.current_stream_ptr(dev)),
                       "oi_render_maps_backward")
        return (grads["d_weights"], grads["d_gradients"], grads["d_raw_color"], grads["d_weight_sum"],
                grads["d_color_fine"], grads["d_light_params"], grads["d_light_dir"], None, None, None, None, None,
                None, None)


def render_maps(generator, bs, render_out, rays_info, prior_info, return_raw):
    """Phong shading of every sample + per-ray compositing of all maps in one kernel; differentiable with respect to
    the render outputs and the light (module parameters stay on the device: no host read-back)."""
    light = prior_info["light"]            # BatchDirectionalLightWithSpecularFixInit (lighting.py:79-119)
    weights = render_out["weights"]
    if not weights.is_cuda:
        raise RuntimeError("render_maps: CUDA tensors only (object_intrinsics_b200 has no CPU path)")
    dev = weights.device
    P = int(generator.resolution)
    base = light.light
    direction = base.param_direction / torch.linalg.norm(base.param_direction)
    light_dir = torch.einsum("bij,j->bi", light.w2b[:, :3, :3].to(dev), direction.to(dev))   # lighting.py:115-119
    light_params = torch.cat([torch.as_tensor(v, dtype=torch.float32, device=dev).reshape(-1).expand(n)
                              for v, n in ((base.ambient_color, 3), (base.diffuse_color, 3),
                                           (base.specular_color, 3), (base.shininess, 1))])
    bg = generator.bg_color(bs)            # [bs,3,h,w] expand of a per-instance colour (utils/prior.py:12-29)
    outs = _RenderMapsFunction.apply(
        weights, render_out["gradients"], render_out["raw_color"], render_out["weight_sum"], render_out["color_fine"],
        light_params, light_dir, render_out["pts"].detach(),
        render_out["mid_z_vals"].detach() if return_raw else None, rays_info["rays_o"].reshape(-1, 3).detach(),
        bg[:, :, 0, 0].detach(), bs, P, bool(return_raw))
    names = _MAPS_BASE + (_MAPS_RAW if return_raw else ())
    ret = dict(zip(names, outs))
    if return_raw:
        ret["z_min"] = outs[-1].reshape(bs, -1).min(-1).values
    render_out.pop("gradients", None)      # the reference deletes these two from the dict (generator.py:126,129)
    render_out.pop("pts", None)
    return ret


def render_and_maps(generator, renderer, bs, rays_info, prior_info, w, return_raw=False, cos_anneal_ratio=1.0,
                    perturb_overwrite=-1):
    """Contract B for the no-grad renders of a training step and for inference: what `Generator.forward` does between
    generator.py:245 (`self.renderer.render(...)`) and :174 (`render_maps`) in ONE library call -- the per-point
    tensors never leave the chip when the render kernel can composite in its tile tail.  Returns (render_out with the
    per-ray keys and the two scalars, maps) like the two reference calls would; not differentiable."""
    light = prior_info["light"]
    base = light.light
    dev = rays_info["rays_o"].device
    direction = base.param_direction / torch.linalg.norm(base.param_direction)
    light_dir = torch.einsum("bij,j->bi", light.w2b[:, :3, :3].to(dev), direction.to(dev))
    light_params = torch.cat([torch.as_tensor(v, dtype=torch.float32, device=dev).reshape(-1).expand(n)
                              for v, n in ((base.ambient_color, 3), (base.diffuse_color, 3),
                                           (base.specular_color, 3), (base.shininess, 1))])
    bg = generator.bg_color(bs)
    ro = rays_info["rays_o"].reshape(-1, 3)
    rd = rays_info["rays_d"].reshape(-1, 3)
    near, far = rays_info.get("near"), rays_info.get("far")
    if near is None:                                                    # near_far_from_sphere, generator.py:336-342
        mid = -(ro * rd).sum(-1, keepdim=True) / (rd * rd).sum(-1, keepdim=True)
        near, far = mid - 1.0, mid + 1.0
    return renderer.render_with_maps(ro, rd, near, far, w=w, light_params=light_params, light_dir=light_dir,
                                     bg_color=bg[:, :, 0, 0], resolution=int(generator.resolution),
                                     return_raw=return_raw, cos_anneal_ratio=cos_anneal_ratio,
                                     perturb_overwrite=perturb_overwrite)
