"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-torch (CPU, fp32 or fp64) restatement of the reference's NeuS-style SDF volume-rendering hot
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / the timed CPU baseline.  The product path
(`object_intrinsics_b200/`) never imports anything from `oracle/`.

Every function cites the reference file:line it restates (paths relative to /root/reference).  The
arithmetic lives in torch ops (F.linear, sin, sigmoid, cumprod, sort, searchsorted, autograd.grad)
exactly as in the reference, which has no third-party numerical dependency besides PyTorch
(2.11.0+cu128 in this image; the authors pinned 2.0.0, environment.yml:82).

Pinning: the reference ships NO golden vectors or numerical tests for this
path, so the oracle is pinned against (1) the known-answer vectors computed from the reference's own
`checkpoints/sphere_init.pt` in the survey (tests/golden/kav_sphere_init.json) and (2) outputs of the
unmodified reference run in the authoring container through oracle/ref_harness.py and committed as
fixtures by oracle/gen_golden.py (tests/golden/*.npz).  tests/test_oracle.py checks both.

Parameters are passed as a flat dict keyed like the reference Generator's state_dict
(`sdf_network.pts_linears.0.weight`, `color_network.views_linears.gamma.bias`,
`deviation_network.variance`, ...), see `extract_params`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# parameter plumbing
# ----------------------------------------------------------------------------------------------
def extract_params(sdf_network, color_network, deviation_network, dtype=None) -> Params:
    """Flat {state_dict key: tensor} for the three nets the renderer holds (renderer.py:87-90)."""
    out = {}
    for prefix, mod in (("sdf_network", sdf_network), ("color_network", color_network),
                        ("deviation_network", deviation_network)):
        for k, v in mod.state_dict().items():
            v = v.detach().cpu()
            out[f"{prefix}.{k}"] = v.to(dtype) if (dtype is not None and v.is_floating_point()) else v.clone()
    return out


def cast_params(P: Params, dtype) -> Params:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in P.items()}


def num_sdf_layers(P: Params) -> int:
    d = 0
    while f"sdf_network.pts_linears.{d}.weight" in P:
        d += 1
    return d


def init_params(D=8, W=128, style_dim=64, seed=0, dtype=torch.float32) -> Params:
    """Random parameters with the reference's init distributions, for machines without the reference.

    FiLMSiren init: stylesdf/volume_renderer.py:33-48; LinearLayer init: :12-25; MappingLinear init:
    stylesdf/model.py:32-47; SingleVarianceNetwork init 0.3: configs/train.yaml:49-52.
    (Only the distributions are followed; bitwise RNG-stream equality with the reference is not needed --
    fixtures carry the actual reference parameters.)
    """
    g = torch.Generator().manual_seed(seed)
    P: Params = {}

    def U(shape, a):
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * a

    def kaiming(out_dim, in_dim, scale):
        gain = math.sqrt(2.0 / (1 + 0.2 ** 2))
        return torch.randn(out_dim, in_dim, generator=g, dtype=torch.float64) * (gain / math.sqrt(in_dim)) * scale

    def film(prefix, cin, cout, first):
        a = 1.0 / 3 if first else math.sqrt(6.0 / cin) / 25
        P[prefix + ".weight"] = U((cout, cin), a)
        P[prefix + ".bias"] = U((cout,), math.sqrt(1.0 / cin))
        for nm in ("gamma", "beta"):
            P[f"{prefix}.{nm}.weight"] = kaiming(cout, style_dim, 0.25)
            P[f"{prefix}.{nm}.bias"] = U((cout,), math.sqrt(1.0 / style_dim))

    for i in range(3):
        P[f"sdf_network.style.{i}.weight"] = kaiming(style_dim, style_dim, 1.0)
        P[f"sdf_network.style.{i}.bias"] = U((style_dim,), math.sqrt(1.0 / style_dim))
    for i in range(D):
        film(f"sdf_network.pts_linears.{i}", 3 if i == 0 else W, W, i == 0)
    P["sdf_network.sigma_linear.weight"] = U((1, W), math.sqrt(6.0 / W) / 25)
    P["sdf_network.sigma_linear.bias"] = U((1,), math.sqrt(1.0 / W))
    film("color_network.views_linears", W + 3, W, False)
    P["color_network.rgb_linear.weight"] = U((3, W), math.sqrt(6.0 / W) / 25)
    P["color_network.rgb_linear.bias"] = U((3,), math.sqrt(1.0 / W))
    P["deviation_network.variance"] = torch.tensor(0.3, dtype=torch.float64)
    return cast_params(P, dtype)


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------
def fused_leaky_relu(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """stylesdf/op/fused_act.py:104-119 (CPU branch) == fused_bias_act_kernel.cu:19-52 act=3,grad=0."""
    return F.leaky_relu(x + bias.view(1, -1), negative_slope=negative_slope) * scale


def style_mlp(P: Params, z: torch.Tensor) -> torch.Tensor:
    """ShapeNetwork.style = 3 x MappingLinear(64,64,'fused_lrelu'): src/models/fields.py:14-19;
    MappingLinear.forward: stylesdf/model.py:49-56 (bias added inside fused_leaky_relu, scale=1)."""
    w = z
    for i in range(3):
        w = F.linear(w, P[f"sdf_network.style.{i}.weight"])
        w = fused_leaky_relu(w, P[f"sdf_network.style.{i}.bias"], scale=1)
    return w


def _linear_layer(x, weight, bias, std_init=1.0, bias_init=0.0):
    """LinearLayer.forward: stylesdf/volume_renderer.py:27-30."""
    return std_init * F.linear(x, weight, bias=bias) + bias_init


def film_gamma_beta(P: Params, prefix: str, w: torch.Tensor):
    """gamma = 15*lin+30, beta = 0.25*lin: stylesdf/volume_renderer.py:47-48,56-57."""
    gamma = _linear_layer(w, P[prefix + ".gamma.weight"], P[prefix + ".gamma.bias"], 15.0, 30.0)
    beta = _linear_layer(w, P[prefix + ".beta.weight"], P[prefix + ".beta.bias"], 0.25, 0.0)
    return gamma, beta


def _film_siren(P: Params, prefix: str, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """FiLMSiren.forward: stylesdf/volume_renderer.py:50-61.  x: [bs, n, C_in]; w: [bs, style]."""
    out = F.linear(x, P[prefix + ".weight"], bias=P[prefix + ".bias"])
    gamma, beta = film_gamma_beta(P, prefix, w)
    return torch.sin(gamma[:, None, :] * out + beta[:, None, :])


def sdf_network_forward(P: Params, pts: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """ShapeNetwork.forward: src/models/fields.py:49-70.  pts [N,3], w [bs,style] -> [N, 1+W].
    Instance of point i is i // (N/bs) (reshape at fields.py:55)."""
    n, bs = pts.shape[0], w.shape[0]
    assert n % bs == 0
    h = pts.reshape(bs, n // bs, 3)
    for i in range(num_sdf_layers(P)):
        h = _film_siren(P, f"sdf_network.pts_linears.{i}", h, w)
    sdf = _linear_layer(h, P["sdf_network.sigma_linear.weight"], P["sdf_network.sigma_linear.bias"])
    return torch.cat([sdf, h], -1).flatten(0, 1)


def sdf_only(P: Params, pts: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """ShapeNetwork.sdf: src/models/fields.py:72-73."""
    return sdf_network_forward(P, pts, w)[:, :1]


def sdf_gradient(P: Params, pts: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """ShapeNetwork.gradient -> gradient(): src/models/fields.py:75-77,104-122.
    A fresh forward + autograd.grad w.r.t. the points, as the reference executes it."""
    with torch.enable_grad():
        x = pts.detach().clone().requires_grad_(True)
        y = sdf_only(P, x, w).squeeze(-1)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=False)
    return g.detach()


def sdf_gradient_analytic(P: Params, pts: torch.Tensor, w: torch.Tensor):
    """Same quantity as sdf_gradient by an explicit reverse sweep (used to cross-check autograd, and to
    document the recurrence the CUDA kernel implements).  Returns (sdf [N,1], feat [N,W], grad [N,3])."""
    n, bs = pts.shape[0], w.shape[0]
    D = num_sdf_layers(P)
    h = pts.reshape(bs, n // bs, 3)
    cs = []
    for i in range(D):
        pre = f"sdf_network.pts_linears.{i}"
        gamma, beta = film_gamma_beta(P, pre, w)
        arg = gamma[:, None, :] * F.linear(h, P[pre + ".weight"], P[pre + ".bias"]) + beta[:, None, :]
        h = torch.sin(arg)
        cs.append(gamma[:, None, :] * torch.cos(arg))
    sdf = F.linear(h, P["sdf_network.sigma_linear.weight"], P["sdf_network.sigma_linear.bias"])
    g = P["sdf_network.sigma_linear.weight"].expand(bs, n // bs, -1)
    for i in reversed(range(D)):
        g = (g * cs[i]) @ P[f"sdf_network.pts_linears.{i}.weight"]
    return sdf.flatten(0, 1), h.flatten(0, 1), g.flatten(0, 1)


def color_network_forward(P: Params, normals: torch.Tensor, feats: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """ColorNetwork.forward: src/models/fields.py:89-101 (ignores points and view_dirs; input order
    [features, normals]; sigmoid output)."""
    n, bs = feats.shape[0], w.shape[0]
    x = torch.cat([feats.reshape(bs, n // bs, -1), normals.reshape(bs, n // bs, 3)], -1)
    h = _film_siren(P, "color_network.views_linears", x, w)
    rgb = _linear_layer(h, P["color_network.rgb_linear.weight"], P["color_network.rgb_linear.bias"])
    return torch.sigmoid(rgb).flatten(0, 1)


def inv_s_value(P: Params) -> torch.Tensor:
    """SingleVarianceNetwork.forward + clip: neus/models/fields.py:267-268; renderer.py:266."""
    return torch.exp(P["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)


# ----------------------------------------------------------------------------------------------
# caller-side helper (generator.py:336-342)
# ----------------------------------------------------------------------------------------------
def near_far_from_sphere(rays_o, rays_d):
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return mid - 1.0, mid + 1.0


# ----------------------------------------------------------------------------------------------
# hierarchical sampling
# ----------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples, det=True):
    """renderer.py:44-74 (det=True is the only mode the renderer uses, :180)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    assert det
    u = torch.linspace(0. + 0.5 / n_samples, 1. - 0.5 / n_samples, steps=n_samples, dtype=bins.dtype)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bins_b, bins_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bins_b + t * (bins_a - bins_b)


def up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
    """renderer.py:137-181."""
    batch_size, n_samples = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=False)
    inside_sphere = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    sdf = sdf.reshape(batch_size, n_samples)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos_val = torch.cat([torch.zeros([batch_size, 1], dtype=z_vals.dtype), cos_val[:, :-1]], dim=-1)
    cos_val = torch.stack([prev_cos_val, cos_val], dim=-1)
    cos_val, _ = torch.min(cos_val, dim=-1, keepdim=False)
    cos_val = cos_val.clip(-1e3, 0.0) * inside_sphere
    dist = next_z - prev_z
    prev_esti_sdf = mid_sdf - cos_val * dist * 0.5
    next_esti_sdf = mid_sdf + cos_val * dist * 0.5
    prev_cdf = torch.sigmoid(prev_esti_sdf * inv_s)
    next_cdf = torch.sigmoid(next_esti_sdf * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    weights = alpha * torch.cumprod(
        torch.cat([torch.ones([batch_size, 1], dtype=z_vals.dtype), 1. - alpha + 1e-7], -1), -1)[:, :-1]
    return sample_pdf(z_vals, weights, n_importance, det=True)


# ----------------------------------------------------------------------------------------------
# render_core / render
# ----------------------------------------------------------------------------------------------
def render_core(P: Params, rays_o, rays_d, z_vals, sample_dist, w, cos_anneal_ratio=0.0,
                analytic_gradient=False):
    """renderer.py:199-349 restricted to the live configuration (n_outside=0, no background blend,
    siren_network=None, second_order=None, compute_color=True)."""
    batch_size, n_samples = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)      # :219-222
    mid_z_vals = z_vals + dists * 0.5                                                 # :225
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z_vals[..., :, None]           # :228
    dirs = rays_d[:, None, :].expand(pts.shape)
    pts = pts.reshape(-1, 3)
    dirs = dirs.reshape(-1, 3)

    if analytic_gradient:
        sdf, feature_vector, gradients = sdf_gradient_analytic(P, pts, w)
    else:
        sdf_nn_output = sdf_network_forward(P, pts, w)                                # :241
        sdf = sdf_nn_output[:, :1]
        feature_vector = sdf_nn_output[:, 1:]
        gradients = sdf_gradient(P, pts, w)                                           # :253
    sampled_color = color_network_forward(P, gradients, feature_vector, w)           # :261
    sampled_color = sampled_color.unflatten(0, (batch_size, n_samples))

    inv_s = inv_s_value(P).reshape(1, 1).expand(batch_size * n_samples, 1)            # :266-267
    true_cos = (dirs * gradients).sum(-1, keepdim=True)                               # :269
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) +
                 F.relu(-true_cos) * cos_anneal_ratio)                                # :273-274
    estimated_next_sdf = sdf + iter_cos * dists.reshape(-1, 1) * 0.5                  # :277
    estimated_prev_sdf = sdf - iter_cos * dists.reshape(-1, 1) * 0.5                  # :278
    prev_cdf = torch.sigmoid(estimated_prev_sdf * inv_s)
    next_cdf = torch.sigmoid(estimated_next_sdf * inv_s)
    p = prev_cdf - next_cdf
    c = prev_cdf
    alpha = ((p + 1e-5) / (c + 1e-5)).reshape(batch_size, n_samples).clip(0.0, 1.0)   # :286

    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(batch_size, n_samples)
    inside_sphere = (pts_norm < 1.0).to(z_vals.dtype)
    relax_inside_sphere = (pts_norm < 1.2).to(z_vals.dtype)

    weights = alpha * torch.cumprod(
        torch.cat([torch.ones([batch_size, 1], dtype=z_vals.dtype), 1. - alpha + 1e-7], -1), -1)[:, :-1]  # :300
    weights_sum = weights.sum(dim=-1, keepdim=True)
    color = (sampled_color * weights[:, :, None]).sum(dim=1)                           # :304

    gradient_error = (torch.linalg.norm(gradients.reshape(batch_size, n_samples, 3), ord=2, dim=-1) - 1.0) ** 2
    gradient_error = (relax_inside_sphere * gradient_error).sum() / (relax_inside_sphere.sum() + 1e-5)  # :309-311

    return {
        'sdf': sdf.reshape(batch_size, n_samples),
        'dists': dists,
        'gradients': gradients.reshape(batch_size, n_samples, 3),
        's_val': 1.0 / inv_s,
        'mid_z_vals': mid_z_vals,
        'weights': weights,
        'cdf': c.reshape(batch_size, n_samples),
        'gradient_error': gradient_error,
        'surface_loss': torch.exp(-1e2 * sdf.abs()).mean(),                            # :338
        'inside_sphere': inside_sphere,
        'pts_norm': pts_norm,
        'pts': pts.reshape(batch_size, n_samples, 3),
        'alpha': alpha,
        'color': color,
        'raw_color': sampled_color,
        'weights_sum': weights_sum,
    }


def coarse_z_vals(near, far, n_samples, t_rand: Optional[torch.Tensor] = None):
    """renderer.py:359-360 and the per-ray jitter :371-373 (t_rand = U[0,1)-0.5, one per ray)."""
    z = torch.linspace(0.0, 1.0, n_samples, dtype=near.dtype)
    z = near + (far - near) * z[None, :]
    if t_rand is not None:
        z = z + t_rand * 2.0 / n_samples
    return z


def fine_z_vals(P: Params, rays_o, rays_d, z_vals, w, n_samples, n_importance, up_sample_steps=1):
    """The `if self.n_importance > 0` block of render(): renderer.py:389-413 with cat_z_vals :183-197."""
    batch_size = rays_o.shape[0]
    with torch.no_grad():
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
        sdf = sdf_only(P, pts.flatten(0, 1), w).reshape(batch_size, n_samples)
        for i in range(up_sample_steps):
            new_z = up_sample(rays_o, rays_d, z_vals, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
            last = (i + 1 == up_sample_steps)
            n_new = new_z.shape[1]
            new_pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z[..., :, None]
            z_vals = torch.cat([z_vals, new_z], dim=-1)
            z_vals, index = torch.sort(z_vals, dim=-1)
            if not last:
                new_sdf = sdf_only(P, new_pts.reshape(-1, 3), w).reshape(batch_size, n_new)
                sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), 1, index)
    return z_vals


def render(P: Params, rays_o, rays_d, near, far, z=None, w=None, *, n_samples, n_importance=0,
           up_sample_steps=1, cos_anneal_ratio=0.0, t_rand: Optional[torch.Tensor] = None,
           z_vals_override: Optional[torch.Tensor] = None, analytic_gradient=False):
    """NeuSRenderer.render: renderer.py:351-473 (live configuration).  `t_rand` [R,1] replaces the
    torch.rand draw of :372 (None == perturb off).  `z_vals_override` feeds given fine z-values to
    render_core (used for matched-z parity of per-point tensors, SURVEY 8c tolerance plan)."""
    if w is None:
        w = style_mlp(P, z)
    batch_size = len(rays_o)
    sample_dist = 2.0 / n_samples                                                     # :356
    z_vals = coarse_z_vals(near, far, n_samples, t_rand)
    n_total = n_samples
    if n_importance > 0:
        z_vals = fine_z_vals(P, rays_o, rays_d, z_vals, w, n_samples, n_importance, up_sample_steps)
        n_total = n_samples + n_importance
    if z_vals_override is not None:
        z_vals = z_vals_override
        n_total = z_vals.shape[1]
    rf = render_core(P, rays_o, rays_d, z_vals, sample_dist, w, cos_anneal_ratio, analytic_gradient)
    weights = rf['weights']
    return {
        's_val': rf['s_val'].reshape(batch_size, n_total).mean(dim=-1, keepdim=True),  # :449
        'cdf_fine': rf['cdf'],
        'weight_sum': weights.sum(dim=-1, keepdim=True),
        'weight_max': torch.max(weights, dim=-1, keepdim=True)[0],
        'gradients': rf['gradients'],
        'weights': weights,
        'gradient_error': rf['gradient_error'],
        'inside_sphere': rf['inside_sphere'],
        'mid_z_vals': rf['mid_z_vals'],
        'surface_loss': rf['surface_loss'],
        'sdf': rf['sdf'],
        'pts_norm': rf['pts_norm'],
        'pts': rf['pts'],
        'color_fine': rf['color'],
        'raw_color': rf['raw_color'],
        'z_vals': z_vals,     # extra (commented out in the reference dict, :458); handy for matched-z tests
    }


# ----------------------------------------------------------------------------------------------
# deterministic synthetic inputs (SURVEY 8d "Synthetic inputs"): closed-form equivalent of gen_rays_at
# ----------------------------------------------------------------------------------------------
def synthetic_rays(bs: int, patch: int, seed: int = 1234, dtype=torch.float32, cam_dist=11.430052,
                   fov_deg=10.0):
    """Per instance: a camera at distance cam_dist looking at the origin from a random direction; a
    patch x patch pinhole grid spanning fov_deg.  Mirrors what generator.py:255-279,317-342 produces
    (unit rays_d, rays_o constant per instance, near/far = mid -+ 1).  Returns rays_o, rays_d [R,3],
    near, far [R,1] with rays of instance b in [b*P^2, (b+1)*P^2)."""
    g = torch.Generator().manual_seed(seed)
    ro, rd = [], []
    for _ in range(bs):
        v = torch.randn(3, generator=g, dtype=torch.float64)
        fwd = -v / v.norm()
        up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
        if abs(float(fwd @ up)) > 0.95:
            up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        upv = torch.linalg.cross(right, fwd)
        o = -fwd * cam_dist
        t = math.tan(math.radians(fov_deg / 2))
        lin = torch.linspace(-t, t, patch, dtype=torch.float64)
        yy, xx = torch.meshgrid(lin, lin, indexing="ij")
        d = fwd[None, None] + xx[..., None] * right[None, None] + yy[..., None] * upv[None, None]
        d = d / d.norm(dim=-1, keepdim=True)
        ro.append(o.expand(patch, patch, 3).reshape(-1, 3))
        rd.append(d.reshape(-1, 3))
    rays_o = torch.cat(ro).to(dtype).contiguous()
    rays_d = torch.cat(rd).to(dtype).contiguous()
    near, far = near_far_from_sphere(rays_o, rays_d)
    return rays_o, rays_d, near, far
