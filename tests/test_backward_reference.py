"""Backward parity pinned to the REFERENCE ITSELF: tests/golden/grad_*.npz hold the parameter gradients of the
unmodified reference render (autograd through NeuSRenderer.render with the create_graph normal of
src/models/fields.py:104-122), written by oracle/gen_golden_grad.py in fp32 and fp64, for

    loss = color_fine.sum() + weight_sum.sum() + 10 * gradient_error

on a fixed-z D=8 2-instance case and a 16+4 hierarchical D=8 3-instance case.

CPU (always): the hand-derived reverse sweep oracle/backward_oracle.py (the algorithm of the CUDA backward) and the
style-MLP chain reproduce the reference's fp64 gradients of all 66 tensors (+ w) to 1e-8 relative.
GPU (-m gpu): the CUDA backward through the drop-in `NeuSRenderer.render` in grad mode, at the z-values the reference
rendered, within  ||g - g64||_inf / ||g64||_inf <= max(1e-3, 3 x the same ratio for the reference's own fp32 run).
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_case, load_params

CASES = ["cfg2_n64_m0", "cfgd_n16_m4_D8"]


def load_grad_case(name):
    with np.load(os.path.join(GOLDEN, f"grad_{name}.npz")) as f:
        meta = json.loads(str(f["meta"]))
        g32 = {k[4:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("g32/")}
        g64 = {k[4:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("g64/")}
        z32, z64 = torch.from_numpy(f["z_vals32"]), torch.from_numpy(f["z_vals64"])
    return meta, g32, g64, z32, z64


def _style_chain(P, z, w_bar):
    """d/d(style layers) of <w_bar, style(z)> by autograd through the oracle's 3-layer style MLP."""
    from oracle import neus_oracle as O
    keys = [f"sdf_network.style.{i}.{n}" for i in range(3) for n in ("weight", "bias")]
    Q = dict(P)
    for k in keys:
        Q[k] = P[k].detach().clone().requires_grad_(True)
    w = O.style_mlp(Q, z)
    gs = torch.autograd.grad((w * w_bar).sum(), [Q[k] for k in keys])
    return dict(zip(keys, gs)), w.detach()


@pytest.mark.parametrize("name", CASES)
def test_backward_oracle_matches_reference_gradients_fp64(name):
    from oracle import backward_oracle as B
    from oracle import neus_oracle as O
    meta, inp, _, _ = load_case(name)
    _, _, g64, _, z64 = load_grad_case(name)
    P = load_params(meta["params"], torch.float64)
    z = inp["z"].double()
    w = O.style_mlp(P, z)
    R = inp["rays_o"].shape[0]
    adj = {"color_fine": torch.ones(R, 3, dtype=torch.float64), "weight_sum": torch.ones(R, 1, dtype=torch.float64),
           "gradient_error": torch.tensor(10.0, dtype=torch.float64)}
    g = B.manual_backward(P, meta["D"], inp["rays_o"].double(), inp["rays_d"].double(), z64, w,
                          meta["cos_anneal_ratio"], meta["n_samples"], adj)
    gs, _ = _style_chain(P, z, g["w"])
    g.update(gs)
    assert len(g64) == 66
    for k, ref in g64.items():
        scale = float(ref.abs().max()) + 1e-30
        err = float((g[k].reshape(ref.shape) - ref).abs().max()) / scale
        assert err < 1e-8, (k, err, scale)


@pytest.mark.gpu
@pytest.mark.parametrize("impl,flags", [("ffma", 0), ("tcgen05", 32), ("tcgen05", 64)],
                         ids=["ffma", "tcgen05-tf32-operands", "tcgen05-fp16-operands"])
@pytest.mark.parametrize("name", CASES)
def test_cuda_backward_matches_reference_gradients(name, impl, flags):
    """flags: the operand format of the tensor-core backward's weight-gradient contraction, forced (bit 5 TF32, bit 6
    scaled fp16 -- what the device-side rule picks after the first, probing call on a workspace)."""
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import NeuSRenderer
    meta, inp, _, _ = load_case(name)
    _, g32, g64, z32, _ = load_grad_case(name)
    P = load_params(meta["params"])
    sdf, col, dev = fields.build_networks(D=meta["D"], device="cuda")
    fields.load_flat_params(sdf, col, dev, P)
    r = NeuSRenderer(nerf=None, sdf_network=sdf, deviation_network=dev, color_network=col,
                     n_samples=meta["n_samples"], n_importance=meta["n_importance"], n_outside=0, up_sample_steps=1,
                     perturb=0, impl=impl)
    r.flags |= flags
    c = {k: inp[k].cuda() for k in ("rays_o", "rays_d", "near", "far", "z")}
    w = sdf.style(c["z"])                     # grad mode: the style layers are reached through w, as in the reference
    w.retain_grad()
    out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=meta["cos_anneal_ratio"],
                   perturb_overwrite=0, z=c["z"], w=w, z_vals=z32.cuda())
    loss = out["color_fine"].sum() + out["weight_sum"].sum() + 10.0 * out["gradient_error"]
    loss.backward()
    torch.cuda.synchronize()
    got = {}
    for prefix, mod in (("sdf_network.", sdf), ("color_network.", col), ("deviation_network.", dev)):
        for k, p in mod.named_parameters():
            assert p.grad is not None, prefix + k
            got[prefix + k] = p.grad.detach().cpu().double()
    got["w"] = w.grad.detach().cpu().double()
    assert set(got) == set(g64)
    worst = (0.0, None)
    for k, ref in g64.items():
        scale = float(ref.abs().max()) + 1e-30
        err = float((got[k].reshape(ref.shape) - ref).abs().max()) / scale
        floor = float((g32[k].double() - ref).abs().max()) / scale
        tol = max(1e-3, 3.0 * floor)
        assert err <= tol, f"{k}: rel Linf {err:.3e} > tol {tol:.3e} (reference fp32 floor {floor:.3e}, scale {scale:.3e})"
        if err / tol > worst[0]:
            worst = (err / tol, k)
    print(f"{name}/{impl}/{flags}: worst err/tol {worst[0]:.2f} ({worst[1]})")
