"""GPU parity tests of the CUDA backward (oi_render_backward through NeuSRenderer.render in grad mode) against
torch.autograd through the differentiable torch formulation (torch_graph.render_differentiable) in fp64.

Tolerance rule (same shape as the forward's): for every parameter gradient g,
    ||g_kernel - g_fp64||_inf / ||g_fp64||_inf  <=  max(1e-3, 3 * floor),
floor = the same ratio for fp32 autograd on the same inputs (the noise torch's own fp32 backward has).
"""
import copy

import pytest
import torch

from helpers import load_case, load_params

pytestmark = pytest.mark.gpu

ADJ_KEYS = ["weights", "weight_sum", "color_fine", "weight_max", "raw_color", "gradients", "sdf", "cdf_fine",
            "gradient_error", "surface_loss", "s_val"]


def _build(meta, dtype=torch.float32, n_importance=None, grad_impl="cuda", impl="auto"):
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import NeuSRenderer
    P = load_params(meta["params"])
    sdf, col, dev = fields.build_networks(D=meta["D"], device="cuda")
    fields.load_flat_params(sdf, col, dev, P)
    sdf, col, dev = sdf.to(dtype), col.to(dtype), dev.to(dtype)
    m = meta["n_importance"] if n_importance is None else n_importance
    return NeuSRenderer(nerf=None, sdf_network=sdf, deviation_network=dev, color_network=col,
                        n_samples=meta["n_samples"], n_importance=m, n_outside=0, up_sample_steps=1, perturb=0,
                        impl=impl, grad_impl=grad_impl)


def _params(r):
    from object_intrinsics_b200.renderer import collect_params
    return collect_params(r.sdf_network, r.color_network, r.deviation_network, with_style=False)


def _loss(out, adj):
    return sum((adj[k].to(out[k].dtype) * out[k]).sum() for k in adj)


def _reference_grads(meta, c, w, z_vals, adj, cos_anneal, dtype, m):
    """autograd through the torch formulation at the given z-values, in `dtype`."""
    r = _build(meta, dtype, n_importance=m, grad_impl="torch")
    wd = w.detach().to(dtype).requires_grad_(True)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        out = r.render(c["rays_o"].to(dtype), c["rays_d"].to(dtype), c["near"].to(dtype), c["far"].to(dtype),
                       cos_anneal_ratio=cos_anneal, perturb_overwrite=0, w=wd, z_vals=z_vals.to(dtype))
        named = _params(r)
        g = torch.autograd.grad(_loss(out, adj), [t for _, t in named] + [wd], allow_unused=True)
    finally:
        torch.set_default_dtype(old)
    return {k: (v.double() if v is not None else None) for k, v in zip([k for k, _ in named] + ["w"], g)}


def _check(meta, c, w, adj_keys, cos_anneal, n_importance=0, t_rand=None, impl="auto", seed=0, flags=0, expect=None):
    torch.manual_seed(seed)
    r = _build(meta, n_importance=n_importance, impl=impl)
    r.flags |= flags
    if expect is not None:
        # the first backward on a workspace keeps TF32 and probes the range of the fp16 operands (overflow guard)
        o = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=cos_anneal, perturb_overwrite=0,
                     w=w, t_rand=t_rand)
        torch.autograd.grad(o["color_fine"].sum(), [t for _, t in _params(r)][:1])
    wk = w.detach().clone().requires_grad_(True)
    out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=cos_anneal, perturb_overwrite=0,
                   w=wk, t_rand=t_rand, return_z_vals=True)
    assert out["color_fine"].requires_grad and out["weights"].requires_grad
    adj = {k: torch.randn_like(out[k]) for k in adj_keys}
    named = _params(r)
    gk = torch.autograd.grad(_loss(out, adj), [t for _, t in named] + [wk], allow_unused=True)
    torch.cuda.synchronize()
    if expect is not None:
        assert r.last_backward_operand_format() == expect
    gk = dict(zip([k for k, _ in named] + ["w"], gk))
    z_vals = out["z_vals"].detach()
    g64 = _reference_grads(meta, c, w, z_vals, adj, cos_anneal, torch.float64, n_importance)
    g32 = _reference_grads(meta, c, w, z_vals, adj, cos_anneal, torch.float32, n_importance)
    worst = (0.0, None)
    for k, ref in g64.items():
        assert gk[k] is not None, k
        assert torch.isfinite(gk[k]).all(), k
        scale = float(ref.abs().max()) + 1e-30
        err = float((gk[k].double().reshape(ref.shape) - ref).abs().max()) / scale
        floor = float((g32[k].reshape(ref.shape) - ref).abs().max()) / scale
        tol = max(1e-3, 3.0 * floor)
        assert err <= tol, f"{k}: rel Linf {err:.3e} > tol {tol:.3e} (fp32-autograd floor {floor:.3e}, scale {scale:.3e})"
        if err / tol > worst[0]:
            worst = (err / tol, k)
    return worst


def _inputs(name, n_rays, n_inst=1):
    meta, inp, _, _ = load_case(name)
    c = {k: v[:n_rays * n_inst].cuda().contiguous() for k, v in inp.items() if k not in ("z", "w")}
    w = inp["w"][:1].cuda().repeat(n_inst, 1)
    if n_inst > 1:
        w = w + 0.3 * torch.randn(n_inst, 64, generator=torch.Generator().manual_seed(5)).cuda()
    return meta, c, w


@pytest.mark.parametrize("impl", ["ffma", "tcgen05"])
@pytest.mark.parametrize("name", ["cfg1_n16_m0", "cfgd_n16_m4_D8"])
def test_all_adjoints_small(name, impl):
    meta, c, w = _inputs(name, 64)
    _check(meta, c, w, ADJ_KEYS, 0.3, impl=impl, expect="fp16" if impl == "tcgen05" else None)


def test_adjoint_scale_invariance_tcgen05():
    """bf16-split adjoint sweeps and per-point power-of-two scales of the fp16 operands: gradients of 1e-6 * loss are
    1e-6 * gradients (no range loss)."""
    meta, c, w = _inputs("cfgd_n16_m4_D8", 64)
    r = _build(meta, n_importance=0, impl="tcgen05")
    named = _params(r)
    grads = []
    for scale in (1.0, 1e-6):
        out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
        loss = scale * (out["color_fine"].sum() + out["weight_sum"].sum() + out["gradient_error"])
        grads.append(torch.autograd.grad(loss, [t for _, t in named]))
    for (k, _), g1, g2 in zip(named, *grads):
        s = float(g1.abs().max()) + 1e-30
        assert float((g2 * 1e6 - g1).abs().max()) / s < 2e-3, k   # TF32 operand rounding differs between the two scales


@pytest.mark.parametrize("flags,expect", [(32, "tf32"), (64, "fp16"), (0, "fp16")])
def test_operand_formats_of_the_contraction(flags, expect):
    """The per-point operands of the weight-gradient contraction as TF32 (flags bit 5) and as scaled fp16 (bit 6; what
    the device-side rule picks for ordinary adjoints): both inside the tolerance of the fp64 reference gradients."""
    meta, c, w = _inputs("cfgd_n16_m4_D8", 64)
    _check(meta, c, w, ADJ_KEYS, 0.3, impl="tcgen05", flags=flags, expect=expect)
    meta, c, w = _inputs("cfgd_n16_m4_D8", 50, n_inst=2)
    _check(meta, c, w, ["color_fine", "weight_sum", "gradient_error", "weights", "gradients", "raw_color"], 1.0,
           impl="tcgen05", flags=flags, expect=expect)


def test_operand_format_follows_the_adjoint_statistics():
    """fp16 operands need the adjoint mass to sit within 2^18 of the largest adjoint.  One point with an adjoint 2^20
    above all others (which then carry 1e-3 of the mass, > 2^-12): the device-side rule keeps TF32, and the gradients
    agree with the forced-TF32 run; ordinary adjoints: fp16."""
    meta, c, w = _inputs("cfgd_n16_m4_D8", 64)
    r = _build(meta, n_importance=0, impl="tcgen05")
    with pytest.raises(RuntimeError):
        r.last_backward_operand_format()
    named = _params(r)

    def grads(adj_sdf, flags):
        r.flags = (r.flags & ~96) | flags
        out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
        g = torch.autograd.grad((out["sdf"] * adj_sdf).sum(), [t for _, t in named], allow_unused=True)
        return g, r.last_backward_operand_format()

    ones = torch.ones(64, meta["n_samples"], device="cuda")
    _, fmt = grads(ones, 0)
    assert fmt == "fp16"
    heavy = ones.clone()
    heavy[3, 5] = 2.0 ** 20
    g_auto, fmt = grads(heavy, 0)
    assert fmt == "tf32"
    g_tf32, fmt = grads(heavy, 32)
    assert fmt == "tf32"
    g_f16, fmt = grads(heavy, 64)
    assert fmt == "fp16"
    for (k, _), a, b, h in zip(named, g_auto, g_tf32, g_f16):
        if a is None:
            continue
        s = float(b.abs().max()) + 1e-30
        assert float((a - b).abs().max()) / s < 1e-5, k      # same kernels (atomics reorder the sums)
        assert float((h - b).abs().max()) / s < 2e-3, k      # the forced fp16 path degrades gracefully here


def test_fp16_overflow_guard_of_the_workspace():
    """First backward on a workspace: TF32 + sampled range probe -> marked safe, fp16 from then on.  An SDF head blown
    up 3000x pushes the forward-type operands past half of fp16's range: the fp16 sweep trips the guard and every
    later call on the workspace keeps TF32 (and matches the forced-TF32 gradients)."""
    meta, c, w = _inputs("cfgd_n16_m4_D8", 64)
    r = _build(meta, n_importance=0, impl="tcgen05")
    named = _params(r)

    def grads():
        out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
        return torch.autograd.grad(out["color_fine"].sum() + out["weight_sum"].sum(), [t for _, t in named],
                                   allow_unused=True)

    grads()
    assert r.last_backward_control_words()[8] == 1 and r.last_backward_operand_format() == "fp16"
    grads()
    assert r.last_backward_control_words()[8] == 1
    g_trip = None
    for _ in range(16):          # the SDF head grows 2x per step, as a diverging training run would
        with torch.no_grad():
            r.sdf_network.sigma_linear.weight.mul_(2.0)
        g_trip = grads()         # fp16 sweep AND fp16 contraction (one snapshot of the guard per call)
        if r.last_backward_control_words()[8] == 2:
            break
    assert r.last_backward_control_words()[8] == 2 and r.last_backward_operand_format() == "tf32"
    g_auto = grads()             # TF32 from now on
    r.flags |= 32
    g_tf32 = grads()
    for (k, _), a, b, t in zip(named, g_auto, g_tf32, g_trip):
        if a is not None:
            s = float(b.abs().max()) + 1e-30
            assert torch.isfinite(a).all(), k
            assert float((a - b).abs().max()) <= 1e-5 * s, k
            # the tripping call: the sampled operands crossed 32768 for the first time, nothing has saturated yet
            assert float((t - b).abs().max()) <= 2e-3 * s, k


@pytest.mark.parametrize("flags", [0, 64])
@pytest.mark.parametrize("bad", [float("inf"), float("nan")])
def test_non_finite_adjoints_propagate(flags, bad):
    """An inf / NaN upstream adjoint must surface as non-finite gradients (not as silently clamped or skipped work), also
    with the fp16 operand format forced: inf keeps the TF32 operands, NaN survives the fp16 conversion."""
    meta, c, w = _inputs("cfgd_n16_m4_D8", 64)
    r = _build(meta, n_importance=0, impl="tcgen05")
    r.flags |= flags
    named = _params(r)
    out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
    adj = torch.ones_like(out["sdf"])
    adj[7, 3] = bad
    g = torch.autograd.grad((out["sdf"] * adj).sum(), [t for _, t in named], allow_unused=True)
    assert any(x is not None and not bool(torch.isfinite(x).all()) for x in g)
    if bad == float("inf"):
        assert r.last_backward_operand_format() == "tf32"


@pytest.mark.parametrize("impl", ["ffma", "tcgen05"])
def test_training_loss_two_instances_ragged_tiles(impl):
    # 2 instances x 50 rays x 16 samples = 800 points per instance: 7 tiles each, the last one 32 points
    meta, c, w = _inputs("cfgd_n16_m4_D8", 50, n_inst=2)
    _check(meta, c, w, ["color_fine", "weight_sum", "gradient_error", "weights", "gradients", "raw_color"], 1.0,
           impl=impl, expect="fp16" if impl == "tcgen05" else None)


def test_hierarchical_and_jitter():
    meta, c, w = _inputs("cfgd_n16_m4_D8", 96)
    t_rand = torch.rand(96, 1, device="cuda") - 0.5
    _check(meta, c, w, ["color_fine", "weight_sum", "gradient_error"], 0.5, n_importance=4, t_rand=t_rand, expect="fp16")


def test_headline_size_backward_and_optimizer_step():
    """cfg2 (64x64 rays x 64 samples, D=8): loss.backward() through the drop-in, grads land on the nn.Parameters,
    match fp32 autograd of the torch formulation, and an optimiser step invalidates the packed blob."""
    meta, inp, _, _ = load_case("cfg2_n64_m0")
    from oracle import neus_oracle as O
    c = dict(zip(("rays_o", "rays_d", "near", "far"), (t.cuda() for t in O.synthetic_rays(1, 64, seed=3))))
    w = inp["w"][:1].cuda()
    r = _build(meta)
    named = _params(r)
    for it in range(2):     # the first backward on a workspace is the TF32 range probe; the second runs the fp16 operands
        for _, p in named:
            p.grad = None
        out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
        img = out["color_fine"] + (1.0 - out["weight_sum"])
        loss = (img ** 2).mean() + 0.1 * out["gradient_error"] + (out["weights"] * out["mid_z_vals"]).sum(-1).mean()
        loss.backward()
    torch.cuda.synchronize()
    assert r.last_backward_operand_format() == "fp16"
    rt = _build(meta, grad_impl="torch")
    out_t = rt.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
    img_t = out_t["color_fine"] + (1.0 - out_t["weight_sum"])
    loss_t = (img_t ** 2).mean() + 0.1 * out_t["gradient_error"] + \
        (out_t["weights"] * out_t["mid_z_vals"]).sum(-1).mean()
    loss_t.backward()
    assert abs(float(loss) - float(loss_t)) < 1e-4
    for (k, p), (_, pt) in zip(named, _params(rt)):
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        scale = float(pt.grad.abs().max()) + 1e-30
        err = float((p.grad - pt.grad).abs().max()) / scale
        assert err < 5e-3, f"{k}: rel Linf {err:.3e} (scale {scale:.3e})"
    repacks = r._packed.repacks
    torch.optim.SGD([p for _, p in named], lr=1e-4).step()
    with torch.no_grad():
        r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
    assert r._packed.repacks == repacks + 1


def test_rays_requiring_grad_raise():
    meta, c, w = _inputs("cfg1_n16_m0", 32)
    r = _build(meta)
    with pytest.raises(NotImplementedError):
        r.render(c["rays_o"].requires_grad_(True), c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0,
                 perturb_overwrite=0, w=w)


def test_tcgen05_matches_ffma_across_chunks_and_instances():
    """2 instances x 4096 rays x 40 samples = 2560 tiles: two slab chunks (2048 + 512 tiles) with the instance
    boundary inside the first chunk's contraction ranges; the tensor-core backward must agree with the FP32 one."""
    meta, inp, _, _ = load_case("cfg2_n64_m0")
    meta = dict(meta, n_samples=40, n_importance=0)
    from oracle import neus_oracle as O
    c = dict(zip(("rays_o", "rays_d", "near", "far"), (t.cuda() for t in O.synthetic_rays(2, 64, seed=11))))
    w = inp["w"][:1].cuda().repeat(2, 1)
    w[1] += 0.2
    grads = {}
    for impl in ("ffma", "tcgen05"):
        r = _build(meta, impl=impl)
        named = _params(r)
        wk = w.clone().requires_grad_(True)
        out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=wk)
        img = out["color_fine"] + (1.0 - out["weight_sum"])
        loss = (img ** 2).mean() + 0.1 * out["gradient_error"] + (out["weights"][..., None] * out["gradients"]).sum() * 1e-3
        grads[impl] = torch.autograd.grad(loss, [t for _, t in named] + [wk])
    torch.cuda.synchronize()
    for (k, _), a, b in zip(_params(r) + [("w", None)], grads["ffma"], grads["tcgen05"]):
        s = float(a.abs().max()) + 1e-30
        assert float((a - b).abs().max()) / s < 2e-3, (k, float((a - b).abs().max()) / s)


def test_cfg4_full_patch_backward_tcgen05_vs_ffma():
    """BASELINE config 4 shape: 128x128 patch, 64 + 64 hierarchical samples (16 384 tiles = 8 slab chunks).  Both
    backward cores on the same rendered z-values; finite, and in agreement."""
    meta, inp, _, _ = load_case("cfg4_n64_m64")
    from oracle import neus_oracle as O
    c = dict(zip(("rays_o", "rays_d", "near", "far"), (t.cuda() for t in O.synthetic_rays(1, 128, seed=5))))
    w = inp["w"][:1].cuda()
    grads, z_vals = {}, None
    for impl in ("tcgen05", "ffma"):
        r = _build(meta, impl=impl)
        named = _params(r)
        out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w,
                       z_vals=z_vals, return_z_vals=True)
        if z_vals is None:
            z_vals = out["z_vals"].detach()          # the second core renders the same (hierarchically sampled) z
        img = out["color_fine"] + (1.0 - out["weight_sum"])
        loss = (img ** 2).mean() + 0.1 * out["gradient_error"]
        grads[impl] = torch.autograd.grad(loss, [t for _, t in named])
    torch.cuda.synchronize()
    for (k, _), a, b in zip(_params(r), grads["ffma"], grads["tcgen05"]):
        assert torch.isfinite(b).all(), k
        s = float(a.abs().max()) + 1e-30
        assert float((a - b).abs().max()) / s < 2e-3, (k, float((a - b).abs().max()) / s)


def test_backward_under_ddp_single_rank_nccl():
    """The reference trains under DDP (tu/ddp.py): gradients produced by oi_render_backward must land on the wrapped
    nn.Parameters through DDP's reducer hooks (NCCL, world_size 1 here; the all-reduce itself is torch's)."""
    import os
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from object_intrinsics_b200.renderer import NeuSRenderer
    meta, c, w = _inputs("cfgd_n16_m4_D8", 64)
    r0 = _build(meta, n_importance=0)

    class Gen(torch.nn.Module):      # the three networks the reference Generator registers (generator.py:51-57)
        def __init__(self, r):
            super().__init__()
            self.sdf_network, self.color_network, self.deviation_network = r.sdf_network, r.color_network, r.deviation_network
            self.sdf_network.style = torch.nn.Sequential()      # w is given: the style MLP takes no part
            self.renderer = NeuSRenderer(None, self.sdf_network, self.deviation_network, self.color_network,
                                         n_samples=meta["n_samples"], n_importance=0, n_outside=0, up_sample_steps=1,
                                         perturb=0)

        def forward(self, c, w):
            out = self.renderer.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0,
                                       perturb_overwrite=0, w=w)
            return out["color_fine"].sum() + out["weight_sum"].sum() + 10.0 * out["gradient_error"]

    gen = Gen(r0).cuda()
    gen(c, w).backward()          # first backward on the workspace: TF32 + range probe (overflow guard); not compared
    gen.zero_grad(set_to_none=True)
    gen(c, w).backward()
    ref = {k: p.grad.clone() for k, p in gen.named_parameters() if p.grad is not None}
    gen.zero_grad(set_to_none=True)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        ddp = DDP(gen, device_ids=[0])
        ddp(c, w).backward()
        torch.cuda.synchronize()
        got = {k: p.grad for k, p in gen.named_parameters() if p.grad is not None}
        assert set(got) == set(ref) and len(got) >= 50
        for k in ref:
            s = float(ref[k].abs().max()) + 1e-30
            assert float((got[k] - ref[k]).abs().max()) / s < 1e-4, k   # atomics: not bit-wise reproducible
    finally:
        dist.destroy_process_group()
