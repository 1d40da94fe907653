"""Differentiable (autograd) formulation of the render path, used ONLY for grad-mode calls.

Render #1 of every training step (`gan_pose_trainer.py:110`) needs dL/dtheta through colour, alpha AND the SDF
normal, i.e. second-order terms (`fields.py:104-122` builds them with `create_graph=True`).  Until the
hand-written backward kernel exists grad-mode calls run this module: plain torch ops on
the GPU over the same `nn.Parameter`s, so gradients land on the original parameters (DDP / optimiser / EMA
unchanged).  The normal is formed by an explicit reverse sweep -- the same recurrence the CUDA kernel
implements -- which is itself made of differentiable ops, so no nested `autograd.grad` is needed.

This is not a fallback for the forward-only path: no-grad calls always go through the CUDA library and fail
loudly without it.  Sampling (coarse pass + inverse-CDF up-sampling) is non-differentiable in the reference
(`renderer.py:390` runs it under `no_grad`) and is done here under `no_grad` as well.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .renderer import collect_params


def _film(P, prefix, w):
    g = 15.0 * F.linear(w, P[prefix + ".gamma.weight"], P[prefix + ".gamma.bias"]) + 30.0
    b = 0.25 * F.linear(w, P[prefix + ".beta.weight"], P[prefix + ".beta.bias"])
    return g[:, None, :], b[:, None, :]


def _sdf_forward(P, D, x, w, want_normal):
    """x [bs, n, 3] -> sdf [bs,n,1], features [bs,n,W], normal [bs,n,3] (or None)."""
    h, cs = x, []
    for l in range(D):
        pre = f"sdf_network.pts_linears.{l}"
        g, b = _film(P, pre, w)
        arg = g * F.linear(h, P[pre + ".weight"], P[pre + ".bias"]) + b
        h = torch.sin(arg)
        if want_normal:
            cs.append(g * torch.cos(arg))
    sdf = F.linear(h, P["sdf_network.sigma_linear.weight"], P["sdf_network.sigma_linear.bias"])
    if not want_normal:
        return sdf, h, None
    gvec = P["sdf_network.sigma_linear.weight"].expand(x.shape[0], x.shape[1], -1)
    for l in reversed(range(D)):
        gvec = (gvec * cs[l]) @ P[f"sdf_network.pts_linears.{l}.weight"]
    return sdf, h, gvec


def _excl_cumprod(x):
    ones = torch.ones_like(x[:, :1])
    return torch.cumprod(torch.cat([ones, x], -1), -1)[:, :-1]


def _sample_fine(P, D, rays_o, rays_d, z_vals, w, n, m, inv_s=64.0):
    """One up-sampling step: renderer.py:137-181 (`up_sample` with the given inv_s) + sample_pdf :44-74 + the sort of
    cat_z_vals :183-197, under no_grad.  The SDF is evaluated at all n current z (the reference evaluates the new
    points only and gathers; per point the value is the same)."""
    R = rays_o.shape[0]
    bs = w.shape[0]
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    sdf, _, _ = _sdf_forward(P, D, pts.reshape(bs, -1, 3), w, False)
    sdf = sdf.reshape(R, n)
    radius = pts.norm(dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    dz = z_vals[:, 1:] - z_vals[:, :-1]
    mid_sdf = (sdf[:, :-1] + sdf[:, 1:]) * 0.5
    cos_val = (sdf[:, 1:] - sdf[:, :-1]) / (dz + 1e-5)
    prev_cos = torch.cat([torch.zeros_like(cos_val[:, :1]), cos_val[:, :-1]], -1)
    cos_val = torch.minimum(prev_cos, cos_val).clip(-1e3, 0.0) * inside
    prev_cdf = torch.sigmoid((mid_sdf - cos_val * dz * 0.5) * inv_s)
    next_cdf = torch.sigmoid((mid_sdf + cos_val * dz * 0.5) * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    weights = alpha * _excl_cumprod(1.0 - alpha + 1e-7) + 1e-5
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / m, 1.0 - 0.5 / m, m, device=z_vals.device, dtype=z_vals.dtype).expand(R, m).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=n - 1)
    cb, ca = cdf.gather(1, below), cdf.gather(1, above)
    zb, za = z_vals.gather(1, below), z_vals.gather(1, above)
    denom = ca - cb
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    new_z = zb + (u - cb) / denom * (za - zb)
    return torch.sort(torch.cat([z_vals, new_z], -1), dim=-1)[0]


def render_differentiable(renderer, rays_o, rays_d, near, far, w, cos_anneal_ratio, t_rand=None, z_vals=None):
    named = dict(collect_params(renderer.sdf_network, renderer.color_network, renderer.deviation_network,
                                with_style=False))
    D = len(renderer.sdf_network.pts_linears)
    n, m = renderer.n_samples, renderer.n_importance
    R, bs = rays_o.shape[0], w.shape[0]
    sample_dist = 2.0 / n
    if z_vals is None:
        lin = torch.linspace(0.0, 1.0, n, device=rays_o.device)
        z_vals = near + (far - near) * lin[None, :]
        if t_rand is not None:
            z_vals = z_vals + t_rand * 2.0 / n
        if m > 0:
            steps = max(int(renderer.up_sample_steps), 1)
            if m % steps != 0:
                raise ValueError(f"n_importance ({m}) must be a multiple of up_sample_steps ({steps})")
            with torch.no_grad():                                           # renderer.py:400-413
                Pd = {k: v.detach() for k, v in named.items()}
                z_vals = z_vals.detach()
                for i in range(steps):
                    z_vals = _sample_fine(Pd, D, rays_o.detach(), rays_d.detach(), z_vals, w.detach(),
                                          n + i * (m // steps), m // steps, 64.0 * 2 ** i)
    S = z_vals.shape[1]
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., None]
    sdf, feat, normal = _sdf_forward(named, D, pts.reshape(bs, -1, 3), w, True)
    g, b = _film(named, "color_network.views_linears", w)
    hc = torch.sin(g * F.linear(torch.cat([feat, normal], -1), named["color_network.views_linears.weight"],
                                named["color_network.views_linears.bias"]) + b)
    rgb = torch.sigmoid(F.linear(hc, named["color_network.rgb_linear.weight"], named["color_network.rgb_linear.bias"]))
    sdf = sdf.reshape(R, S)
    normal = normal.reshape(R, S, 3)
    rgb = rgb.reshape(R, S, 3)

    inv_s = torch.exp(named["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)
    true_cos = (rays_d[:, None, :] * normal).sum(-1)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    prev_cdf = torch.sigmoid((sdf - iter_cos * dists * 0.5) * inv_s)
    next_cdf = torch.sigmoid((sdf + iter_cos * dists * 0.5) * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)
    pts_norm = pts.norm(dim=-1)
    relax = (pts_norm < 1.2).to(pts.dtype).detach()
    weights = alpha * _excl_cumprod(1.0 - alpha + 1e-7)
    weight_sum = weights.sum(-1, keepdim=True)
    ge = (normal.norm(dim=-1) - 1.0) ** 2
    return {
        "s_val": (1.0 / inv_s).expand(R, S).mean(-1, keepdim=True),
        "cdf_fine": prev_cdf,
        "weight_sum": weight_sum,
        "weight_max": weights.max(-1, keepdim=True)[0],
        "gradients": normal,
        "weights": weights,
        "gradient_error": (relax * ge).sum() / (relax.sum() + 1e-5),
        "inside_sphere": (pts_norm < 1.0).to(pts.dtype).detach(),
        "mid_z_vals": mid_z,
        "surface_loss": torch.exp(-1e2 * sdf.abs()).mean(),
        "sdf": sdf,
        "pts_norm": pts_norm,
        "pts": pts,
        "color_fine": (rgb * weights[..., None]).sum(1),
        "raw_color": rgb,
    }
