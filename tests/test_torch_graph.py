"""CPU tests of the grad-mode formulation (object_intrinsics_b200/torch_graph.py): forward values against the
oracle, parameter gradients against a nested-autograd formulation (what the reference does, fields.py:104-122)."""
import types

import torch

from helpers import OUT_KEYS, linf, load_case, load_params
from oracle import neus_oracle as O
from object_intrinsics_b200 import fields, torch_graph


def _nets(P, D, dtype):
    sdf = fields.ShapeNetwork.__new__(fields.ShapeNetwork)
    torch.nn.Module.__init__(sdf)
    sdf.style = torch.nn.Sequential()   # style MLP needs the CUDA op; not exercised here
    sdf.pts_linears = torch.nn.ModuleList([fields.FiLMSiren(3, 128, 64, True)] +
                                          [fields.FiLMSiren(128, 128, 64) for _ in range(D - 1)])
    sdf.sigma_linear = fields.LinearLayer(128, 1)
    col = fields.ColorNetwork(D=D)
    dev = fields.SingleVarianceNetwork(0.3)
    sd = {k[len("sdf_network."):]: v for k, v in P.items() if k.startswith("sdf_network.") and ".style." not in k}
    sdf.load_state_dict(sd)
    col.load_state_dict({k[len("color_network."):]: v for k, v in P.items() if k.startswith("color_network.")})
    dev.load_state_dict({"variance": P["deviation_network.variance"]})
    return sdf.to(dtype), col.to(dtype), dev.to(dtype)


def _renderer(sdf, col, dev, n, m, steps=1):
    return types.SimpleNamespace(sdf_network=sdf, color_network=col, deviation_network=dev, n_samples=n,
                                 n_importance=m, up_sample_steps=steps)


def test_forward_matches_oracle_fp64():
    # the last two: up_sample_steps = 2 / 4 (renderer.py:400-413), fixtures of the unmodified reference
    for name in ("cfg1_n16_m0", "cfg1_n16_m4_jit", "cfgd_n16_m4_D8", "cfgs_n16_m8_s2", "cfgs_n16_m12_s4_D8"):
        meta, inp, r32, r64 = load_case(name)
        P = load_params(meta["params"], torch.float64)
        sdf, col, dev = _nets(P, meta["D"], torch.float64)
        a = {k: v.double() for k, v in inp.items()}
        a["w"] = r64["w"]      # the fp64 reference run derived w from z in fp64
        r = _renderer(sdf, col, dev, meta["n_samples"], meta["n_importance"], meta.get("up_sample_steps", 1))
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            out = torch_graph.render_differentiable(r, a["rays_o"], a["rays_d"], a["near"], a["far"], a["w"],
                                                    meta["cos_anneal_ratio"], a.get("t_rand"))
        finally:
            torch.set_default_dtype(old)
        for k in OUT_KEYS:
            assert linf(out[k].detach(), r64[k]) < 1e-7, (name, k, linf(out[k].detach(), r64[k]))


def test_parameter_gradients_match_nested_autograd():
    meta, inp, _, _ = load_case("cfg1_n16_m0")
    P = load_params(meta["params"], torch.float64)
    sdf, col, dev = _nets(P, meta["D"], torch.float64)
    a = {k: v.double()[:24] for k, v in inp.items() if k not in ("z", "w")}
    w = inp["w"].double()[:1].clone().requires_grad_(True)
    r = _renderer(sdf, col, dev, meta["n_samples"], 0)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        out = torch_graph.render_differentiable(r, a["rays_o"], a["rays_d"], a["near"], a["far"], w, 0.3)
        loss = out["color_fine"].sum() + 0.5 * out["weight_sum"].sum() + 10.0 * out["gradient_error"]
        params = [p for p in list(sdf.parameters()) + list(col.parameters()) + list(dev.parameters())]
        g1 = torch.autograd.grad(loss, params + [w], allow_unused=True)

        # nested-autograd formulation: normal = autograd.grad(sdf, x, create_graph=True)
        z_vals = O.coarse_z_vals(a["near"], a["far"], meta["n_samples"])
        dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full_like(z_vals[:, :1], 2.0 / 16)], -1)
        mid = z_vals + dists * 0.5
        pts = (a["rays_o"][:, None] + a["rays_d"][:, None] * mid[..., None]).reshape(-1, 3).requires_grad_(True)
        o = sdf(pts, w=w)
        s, feat = o[:, :1], o[:, 1:]
        (nrm,) = torch.autograd.grad(s.sum(), pts, create_graph=True)
        rgb = col(None, nrm, None, feat, w=w).reshape(24, 16, 3)
        inv_s = torch.exp(dev.variance * 10.0)
        dirs = a["rays_d"][:, None].expand(24, 16, 3).reshape(-1, 3)
        tc = (dirs * nrm).sum(-1, keepdim=True)
        ic = -(torch.relu(-tc * 0.5 + 0.5) * 0.7 + torch.relu(-tc) * 0.3)
        pc = torch.sigmoid((s - ic * dists.reshape(-1, 1) * 0.5) * inv_s)
        nc = torch.sigmoid((s + ic * dists.reshape(-1, 1) * 0.5) * inv_s)
        alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).reshape(24, 16).clip(0, 1)
        wts = alpha * torch.cumprod(torch.cat([torch.ones(24, 1), 1 - alpha + 1e-7], -1), -1)[:, :-1]
        relax = (pts.detach().norm(dim=-1).reshape(24, 16) < 1.2).double()
        ge = (relax * (nrm.norm(dim=-1).reshape(24, 16) - 1) ** 2).sum() / (relax.sum() + 1e-5)
        loss2 = (rgb * wts[..., None]).sum() + 0.5 * wts.sum() + 10.0 * ge
        g2 = torch.autograd.grad(loss2, params + [w], allow_unused=True)
    finally:
        torch.set_default_dtype(old)
    assert abs(float(loss) - float(loss2)) < 1e-9
    n_nonzero = 0
    for x, y in zip(g1, g2):
        assert (x is None) == (y is None)
        if x is not None:
            assert linf(x, y) <= 1e-8 * max(1.0, float(y.abs().max())), linf(x, y)
            n_nonzero += int(x.abs().max() > 0)
    assert n_nonzero >= len(g1) - 2
