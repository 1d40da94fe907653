"""CPU check of the hand-derived backward sweep (oracle/backward_oracle.py = the algorithm of the CUDA backward,
csrc/oi_render_bwd.cu) against torch.autograd through the differentiable formulation, in fp64."""
import types

import pytest
import torch

from helpers import load_case, load_params
from oracle import backward_oracle as B
from object_intrinsics_b200 import torch_graph
from test_torch_graph import _nets

ADJ_KEYS = ["weights", "weight_sum", "color_fine", "weight_max", "raw_color", "gradients", "sdf", "cdf_fine",
            "gradient_error", "surface_loss", "s_val"]


def _setup(name, n_rays, cos_anneal, seed=0):
    meta, inp, _, r64 = load_case(name)
    D = meta["D"]
    P = load_params(meta["params"], torch.float64)
    sdf, col, dev = _nets(P, D, torch.float64)
    a = {k: v.double()[:n_rays] for k, v in inp.items() if k not in ("z", "w")}
    w = r64["w"].double()[:1].clone().requires_grad_(True)
    n = meta["n_samples"]
    r = types.SimpleNamespace(sdf_network=sdf, color_network=col, deviation_network=dev, n_samples=n,
                              n_importance=0, up_sample_steps=1)
    lin = torch.linspace(0.0, 1.0, n, dtype=torch.float64)
    z_vals = a["near"] + (a["far"] - a["near"]) * lin[None, :]
    return meta, P, r, a, w, z_vals


@pytest.mark.parametrize("name,cos_anneal", [("cfg1_n16_m0", 0.3), ("cfgd_n16_m4_D8", 1.0)])
def test_manual_backward_matches_autograd(name, cos_anneal):
    torch.manual_seed(0)
    meta, P, r, a, w, z_vals = _setup(name, 24, cos_anneal)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        out = torch_graph.render_differentiable(r, a["rays_o"], a["rays_d"], a["near"], a["far"], w, cos_anneal,
                                                z_vals=z_vals)
        adj = {k: torch.randn_like(out[k]) for k in ADJ_KEYS}
        loss = sum((adj[k] * out[k]).sum() for k in ADJ_KEYS)
        named = dict(torch_graph.collect_params(r.sdf_network, r.color_network, r.deviation_network, with_style=False))
        keys = list(named)
        g_auto = torch.autograd.grad(loss, [named[k] for k in keys] + [w])
        Pd = {k: v.detach() for k, v in named.items()}
        g_man = B.manual_backward(Pd, meta["D"], a["rays_o"], a["rays_d"], z_vals, w.detach(), cos_anneal,
                                  meta["n_samples"], adj)
    finally:
        torch.set_default_dtype(old)
    for k, ga in zip(keys + ["w"], g_auto):
        gm = g_man[k].reshape(ga.shape)
        scale = float(ga.abs().max()) + 1e-30
        err = float((gm - ga).abs().max()) / scale
        assert err < 1e-9, (k, err, scale)


def test_manual_backward_two_instances():
    """Per-instance FiLM tables: rays of instance b are rows [b R/bs, (b+1) R/bs) (fields.py:55)."""
    torch.manual_seed(1)
    meta, P, r, a, w1, z_vals = _setup("cfg1_n16_m0", 24, 1.0)
    w = torch.cat([w1.detach(), w1.detach() + 0.3 * torch.randn(1, 64, dtype=torch.float64)]).requires_grad_(True)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        out = torch_graph.render_differentiable(r, a["rays_o"], a["rays_d"], a["near"], a["far"], w, 1.0, z_vals=z_vals)
        keys_adj = ["color_fine", "weight_sum", "gradient_error", "weights"]
        adj = {k: torch.randn_like(out[k]) for k in keys_adj}
        loss = sum((adj[k] * out[k]).sum() for k in keys_adj)
        named = dict(torch_graph.collect_params(r.sdf_network, r.color_network, r.deviation_network, with_style=False))
        keys = list(named)
        g_auto = torch.autograd.grad(loss, [named[k] for k in keys] + [w])
        g_man = B.manual_backward({k: v.detach() for k, v in named.items()}, meta["D"], a["rays_o"], a["rays_d"], z_vals,
                                  w.detach(), 1.0, meta["n_samples"], adj)
    finally:
        torch.set_default_dtype(old)
    for k, ga in zip(keys + ["w"], g_auto):
        scale = float(ga.abs().max()) + 1e-30
        assert float((g_man[k].reshape(ga.shape) - ga).abs().max()) / scale < 1e-9, k


def test_dgamma_identity_used_by_the_tensor_core_backward():
    """csrc/oi_render_bwd_tc.cu takes dL/dgamma from the (per-instance) weight gradients instead of per-point column
    sums:  dgamma_l[j] = ( sum_k W_l[j][k] dW_l[j][k] + b_l[j] db_l[j] ) / gamma_l[j]   (same for the colour layer).
    Checked here in fp64 against the reverse sweep's own dgamma (which test_manual_backward_matches_autograd pins to
    torch.autograd)."""
    torch.manual_seed(3)
    meta, P, r, a, w, z_vals = _setup("cfgd_n16_m4_D8", 24, 1.0)
    D = meta["D"]
    named = dict(torch_graph.collect_params(r.sdf_network, r.color_network, r.deviation_network, with_style=False))
    Pd = {k: v.detach() for k, v in named.items()}
    R, S = z_vals.shape
    adj = {"color_fine": torch.randn(R, 3, dtype=torch.float64), "weight_sum": torch.randn(R, 1, dtype=torch.float64),
           "gradients": torch.randn(R, S, 3, dtype=torch.float64), "weights": torch.randn(R, S, dtype=torch.float64),
           "gradient_error": torch.tensor(3.0, dtype=torch.float64)}
    g = B.manual_backward(Pd, D, a["rays_o"], a["rays_d"], z_vals, w.detach(), 1.0, meta["n_samples"], adj)
    gam, _ = B.film_tables(Pd, D, w.detach())
    layers = [(l, f"sdf_network.pts_linears.{l}") for l in range(D)] + [(8, "color_network.views_linears")]
    for slot, pre in layers:
        W, b = Pd[pre + ".weight"], Pd[pre + ".bias"]
        ident = ((W * g[pre + ".weight"]).sum(-1) + b * g[pre + ".bias"]) / gam[0, slot]     # one instance
        ref = g["film_gamma"][0, slot]
        assert float((ident - ref).abs().max()) <= 1e-11 * float(ref.abs().max()), pre
