"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Hand-derived reverse-mode sweep of the render path (the algorithm `csrc/oi_render_bwd.cu` implements),
written with plain torch ops and NO autograd, so that every formula of the CUDA backward has a CPU
restatement that can be checked against `torch.autograd` in fp64 (tests/test_backward_math.py).

What is differentiated (paths relative to /root/reference): `render_core`, renderer.py:199-349, with
`ShapeNetwork.forward` (src/models/fields.py:49-73), the normal `ShapeNetwork.gradient`
(fields.py:75-77,104-122 -- itself an autograd.grad with create_graph=True, hence second-order terms) and
`ColorNetwork.forward` (fields.py:89-101).  Sampling (renderer.py:389-413) runs under no_grad in the
reference and is not differentiated: z_vals are an input here.

Notation per sample point (FiLM gamma_l, beta_l per instance and channel):
  forward   u_l = W_l h_l + b_l,  a_l = gamma_l u_l + beta_l,  h_{l+1} = sin a_l,  c_l = gamma_l cos a_l
            sdf = w_s . h_D + b_s
  reverse   t_{D-1} = w_s * c_{D-1};  g_l = W_l^T t_l;  t_{l-1} = g_l * c_{l-1};  normal = g_0
  colour    u_c = W_cf h_D + W_cg normal + b_c,  h_c = sin(gamma_c u_c + beta_c),  rgb = sigmoid(W_rgb h_c + b_rgb)
  tail      alpha(sdf, normal) (renderer.py:266-286), weights = alpha * exclusive-cumprod(1 - alpha + 1e-7)
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

Params = Dict[str, torch.Tensor]


def film_tables(P: Params, D: int, w: torch.Tensor):
    """gamma, beta [bs, 9, 128] (slot 8 = colour layer); volume_renderer.py:27-30,47-48,56-57."""
    bs = w.shape[0]
    g = w.new_zeros(bs, 9, 128)
    b = w.new_zeros(bs, 9, 128)
    for slot, pre in [(l, f"sdf_network.pts_linears.{l}") for l in range(D)] + [(8, "color_network.views_linears")]:
        g[:, slot] = 15.0 * (w @ P[pre + ".gamma.weight"].T + P[pre + ".gamma.bias"]) + 30.0
        b[:, slot] = 0.25 * (w @ P[pre + ".beta.weight"].T + P[pre + ".beta.bias"])
    return g, b


def tail_backward(sdf, normal, rgb, rays_d, z_vals, inv_s, cos_anneal, sample_dist, pts_norm, adj):
    """Adjoints of the per-ray compositing and the NeuS alpha (renderer.py:261-305).

    sdf [R,S], normal/rgb [R,S,3]; adj: dict of incoming output adjoints (missing = zero).
    Returns sdf_bar [R,S], normal_bar [R,S,3], zrgb_bar [R,S,3] (adjoint of the rgb pre-activation),
    inv_s_bar []."""
    R, S = sdf.shape
    z = lambda *s: sdf.new_zeros(*s)
    g = lambda k, *s: adj[k] if adj.get(k) is not None else z(*s)
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], sample_dist)], -1)
    tc = (rays_d[:, None, :] * normal).sum(-1)
    r = cos_anneal
    ic = -(torch.relu(0.5 - 0.5 * tc) * (1.0 - r) + torch.relu(-tc) * r)
    hs = ic * dists * 0.5
    A, B = (sdf - hs) * inv_s, (sdf + hs) * inv_s
    p, q = torch.sigmoid(A), torch.sigmoid(B)
    raw = (p - q + 1e-5) / (p + 1e-5)
    alpha = raw.clip(0.0, 1.0)
    f = 1.0 - alpha + 1e-7
    T = torch.cumprod(torch.cat([torch.ones_like(f[:, :1]), f], -1), -1)[:, :-1]
    weights = alpha * T

    # ---- adjoint of weights from every consumer
    wbar = g("weights", R, S) + g("weight_sum", R, 1) + (g("color_fine", R, 3)[:, None, :] * rgb).sum(-1)
    if adj.get("weight_max") is not None:
        idx = weights.argmax(-1, keepdim=True)
        wbar = wbar + torch.zeros_like(wbar).scatter(1, idx, adj["weight_max"])
    rgb_bar = g("raw_color", R, S, 3) + weights[..., None] * g("color_fine", R, 3)[:, None, :]
    # ---- exclusive cumprod: alpha_bar_i = wbar_i T_i - (sum_{k>i} wbar_k w_k) / f_i
    ww = wbar * weights
    suffix = ww.flip(-1).cumsum(-1).flip(-1) - ww
    alpha_bar = wbar * T - suffix / f
    raw_bar = alpha_bar * ((raw >= 0.0) & (raw <= 1.0)).to(sdf.dtype)
    p_bar = raw_bar * q / (p + 1e-5) ** 2 + g("cdf_fine", R, S)
    q_bar = -raw_bar / (p + 1e-5)
    A_bar, B_bar = p_bar * p * (1 - p), q_bar * q * (1 - q)
    sdf_bar = (A_bar + B_bar) * inv_s + g("sdf", R, S)
    hs_bar = (B_bar - A_bar) * inv_s
    inv_s_bar = (A_bar * (sdf - hs) + B_bar * (sdf + hs)).sum()
    ic_bar = hs_bar * dists * 0.5
    tc_bar = ic_bar * (0.5 * (1.0 - r) * (0.5 - 0.5 * tc > 0).to(sdf.dtype) + r * (-tc > 0).to(sdf.dtype))
    normal_bar = g("gradients", R, S, 3) + tc_bar[..., None] * rays_d[:, None, :]
    # ---- the two global scalars (renderer.py:295-299) and s_val
    if adj.get("gradient_error") is not None:
        relax = (pts_norm < 1.2).to(sdf.dtype)
        nn = normal.norm(dim=-1)
        coef = adj["gradient_error"] / (relax.sum() + 1e-5)
        normal_bar = normal_bar + (coef * relax * 2.0 * (nn - 1.0) / nn)[..., None] * normal
    if adj.get("surface_loss") is not None:
        sdf_bar = sdf_bar + adj["surface_loss"] / (R * S) * (-100.0 * torch.sign(sdf)) * torch.exp(-100.0 * sdf.abs())
    if adj.get("s_val") is not None:
        inv_s_bar = inv_s_bar - adj["s_val"].sum() / inv_s ** 2
    zrgb_bar = rgb_bar * rgb * (1.0 - rgb)
    return sdf_bar, normal_bar, zrgb_bar, inv_s_bar


def manual_backward(P: Params, D: int, rays_o, rays_d, z_vals, w, cos_anneal, n_samples,
                    adj: Dict[str, Optional[torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Parameter gradients (keys as in P, plus 'w') of sum_k <adj[k], out[k]> for the render outputs."""
    R, S = z_vals.shape
    bs = w.shape[0]
    n = R * S // bs
    sample_dist = 2.0 / n_samples
    gam, bet = film_tables(P, D, w)
    Wl = [P[f"sdf_network.pts_linears.{l}.weight"] for l in range(D)]
    bl = [P[f"sdf_network.pts_linears.{l}.bias"] for l in range(D)]
    ws, bsig = P["sdf_network.sigma_linear.weight"][0], P["sdf_network.sigma_linear.bias"]
    Wc, bc = P["color_network.views_linears.weight"], P["color_network.views_linears.bias"]
    Wcf, Wcg = Wc[:, :128], Wc[:, 128:]
    Wrgb, brgb = P["color_network.rgb_linear.weight"], P["color_network.rgb_linear.bias"]
    var = P["deviation_network.variance"]
    inv_s_raw = torch.exp(var * 10.0)
    inv_s = inv_s_raw.clip(1e-6, 1e6)

    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], sample_dist)], -1)
    mid = z_vals + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid[..., None]
    x = pts.reshape(bs, n, 3)

    # ---------------- recompute: forward sweep (keeps u_l) and reverse sweep (keeps g_l)
    G = lambda l: gam[:, l][:, None, :]
    Bt = lambda l: bet[:, l][:, None, :]
    h, u = [x], []
    for l in range(D):
        u.append(h[l] @ Wl[l].T + bl[l])
        h.append(torch.sin(G(l) * u[l] + Bt(l)))
    c = [G(l) * torch.cos(G(l) * u[l] + Bt(l)) for l in range(D)]
    sdf = h[D] @ ws + bsig
    t = [None] * D
    g = [None] * (D + 1)
    t[D - 1] = ws * c[D - 1]
    for l in range(D - 1, -1, -1):
        g[l] = t[l] @ Wl[l]
        if l > 0:
            t[l - 1] = g[l] * c[l - 1]
    normal = g[0]
    u_c = h[D] @ Wcf.T + normal @ Wcg.T + bc
    a_c = G(8) * u_c + Bt(8)
    h_c = torch.sin(a_c)
    rgb = torch.sigmoid(h_c @ Wrgb.T + brgb)

    # ---------------- per-ray tail
    sdf_bar, n_bar, z_bar, inv_s_bar = tail_backward(
        sdf.reshape(R, S), normal.reshape(R, S, 3), rgb.reshape(R, S, 3), rays_d, z_vals, inv_s, cos_anneal,
        sample_dist, pts.norm(dim=-1), adj)
    sdf_bar, n_bar, z_bar = sdf_bar.reshape(bs, n), n_bar.reshape(bs, n, 3), z_bar.reshape(bs, n, 3)

    out = {}
    dgam, dbet = torch.zeros_like(gam), torch.zeros_like(bet)
    sum_pts = lambda a, b: torch.einsum("bni,bnj->ij", a, b)      # weight-gradient contraction over points

    # ---------------- colour head
    out["color_network.rgb_linear.weight"] = sum_pts(z_bar, h_c)
    out["color_network.rgb_linear.bias"] = z_bar.sum((0, 1))
    hc_bar = z_bar @ Wrgb
    ac_bar = hc_bar * torch.cos(a_c)
    dbet[:, 8] = ac_bar.sum(1)
    dgam[:, 8] = (ac_bar * u_c).sum(1)
    uc_bar = ac_bar * G(8)
    out["color_network.views_linears.weight"] = torch.cat([sum_pts(uc_bar, h[D]), sum_pts(uc_bar, normal)], 1)
    out["color_network.views_linears.bias"] = uc_bar.sum((0, 1))
    hD_bar = uc_bar @ Wcf + sdf_bar[..., None] * ws
    n_bar = n_bar + uc_bar @ Wcg
    dws = (sdf_bar[..., None] * h[D]).sum((0, 1))
    out["sdf_network.sigma_linear.bias"] = sdf_bar.sum().reshape(1)

    # ---------------- backward of the reverse sweep (ascending l): g_bar_0 = normal_bar
    dW = [torch.zeros_like(Wl[l]) for l in range(D)]
    c_bar = [None] * D
    g_bar = n_bar
    for l in range(D):
        t_bar = g_bar @ Wl[l].T                      # t_bar_l = W_l g_bar_l
        dW[l] = dW[l] + sum_pts(t[l], g_bar)         # dW_l += t_l (x) g_bar_l
        if l < D - 1:
            c_bar[l] = t_bar * g[l + 1]
            g_bar = t_bar * c[l]
        else:
            c_bar[l] = t_bar * ws
            dws = dws + (t_bar * c[l]).sum((0, 1))
    out["sdf_network.sigma_linear.weight"] = dws.reshape(1, -1)

    # ---------------- backward of the forward sweep (descending l)
    h_bar = hD_bar
    for l in range(D - 1, -1, -1):
        a = G(l) * u[l] + Bt(l)
        a_bar = h_bar * torch.cos(a) - c_bar[l] * G(l) * torch.sin(a)
        dbet[:, l] = a_bar.sum(1)
        dgam[:, l] = (a_bar * u[l] + c_bar[l] * torch.cos(a)).sum(1)
        u_bar = a_bar * G(l)
        dW[l] = dW[l] + sum_pts(u_bar, h[l])
        out[f"sdf_network.pts_linears.{l}.bias"] = u_bar.sum((0, 1))
        h_bar = u_bar @ Wl[l]
    for l in range(D):
        out[f"sdf_network.pts_linears.{l}.weight"] = dW[l]

    # ---------------- variance and the FiLM linears (volume_renderer.py:27-30) / latent w
    inside = ((inv_s_raw >= 1e-6) & (inv_s_raw <= 1e6)).to(var.dtype)
    out["deviation_network.variance"] = inv_s_bar * 10.0 * inv_s_raw * inside
    w_bar = torch.zeros_like(w)
    for slot, pre in [(l, f"sdf_network.pts_linears.{l}") for l in range(D)] + [(8, "color_network.views_linears")]:
        dg, db = 15.0 * dgam[:, slot], 0.25 * dbet[:, slot]
        out[pre + ".gamma.weight"] = dg.T @ w
        out[pre + ".gamma.bias"] = dg.sum(0)
        out[pre + ".beta.weight"] = db.T @ w
        out[pre + ".beta.bias"] = db.sum(0)
        w_bar = w_bar + dg @ P[pre + ".gamma.weight"] + db @ P[pre + ".beta.weight"]
    out["w"] = w_bar
    out["film_gamma"], out["film_beta"] = dgam, dbet
    return out
