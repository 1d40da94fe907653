"""GPU parity tests of the generator-side kernels (oi_gen_rays, oi_render_maps) against the oracle and the golden
vectors produced by the reference's own Generator.gen_rays_at / render_maps."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import GOLDEN, linf
from oracle import generator_oracle as GO

pytestmark = pytest.mark.gpu


def load():
    with np.load(os.path.join(GOLDEN, "generator_golden.npz")) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def _fake_generator(G, bg=None):
    cam_dist, res, scene_res = [float(v) for v in G["rays/scalars"]]
    cam = types.SimpleNamespace(cam_dist=cam_dist, w2c=G["rays/w2c"].cuda(), c2w=G["rays/c2w"].cuda(),
                                intrinsics_inv=G["rays/intrinsics_inv"].cuda())
    g = types.SimpleNamespace(camera=cam, resolution=int(res), scene_resolution=int(scene_res))
    if bg is not None:
        g.bg_color = lambda n: bg[:, :, None, None].expand(n, 3, int(res), int(res))
    return g


def test_gen_rays_matches_reference_and_near_far():
    from object_intrinsics_b200 import generator_ops
    G = load()
    gen = _fake_generator(G)
    out = generator_ops.gen_rays_at(gen, {}, {"b2w": G["rays/b2w"].cuda(), "c2b": G["rays/c2b"].cuda()},
                                    with_near_far=True)
    assert linf(out["rays_o"].cpu(), G["rays/rays_o"]) == 0.0
    assert linf(out["rays_d"].cpu(), G["rays/rays_d"]) < 2e-6
    assert linf(out["x_offset"].cpu(), G["rays/x_offset"]) < 2e-4
    assert linf(out["y_offset"].cpu(), G["rays/y_offset"]) < 2e-4
    from oracle import neus_oracle as O
    near, far = O.near_far_from_sphere(G["rays/rays_o"].reshape(-1, 3), G["rays/rays_d"].reshape(-1, 3))
    assert linf(out["near"].cpu(), near) < 2e-5 and linf(out["far"].cpu(), far) < 2e-5


@pytest.mark.parametrize("return_raw", [True, False])
def test_render_maps_matches_reference(return_raw):
    from object_intrinsics_b200 import generator_ops
    G = load()
    amb, dif, spec, shin = [float(v) for v in G["maps/light"]]
    bs = G["maps/bg"].shape[0]
    gen = _fake_generator(G, bg=G["maps/bg"])
    # stand-in for the reference light objects: exactly the attributes render_maps reads
    base = types.SimpleNamespace(param_direction=torch.tensor([0.0, 0.0, -1.0]).cuda(),
                                 ambient_color=torch.full((3,), amb), diffuse_color=torch.full((3,), dif),
                                 specular_color=torch.full((3,), spec), shininess=torch.tensor(shin))
    light = types.SimpleNamespace(light=base, w2b=G["maps/w2b"].cuda())
    render_out = {k[len("maps/in/"):]: v.cuda() for k, v in G.items() if k.startswith("maps/in/")}
    with torch.no_grad():
        out = generator_ops.render_maps(gen, bs, render_out, {"rays_o": G["rays/rays_o"].cuda()}, {"light": light},
                                        return_raw)
    torch.cuda.synchronize()
    expect = {k[len("maps/out/"):] for k in G if k.startswith("maps/out/")}
    if not return_raw:
        expect -= {"amb_shading_map", "diff_shading_map", "normal_map", "no_specular_map", "specular_map", "z_map",
                   "z_min"}
    assert set(out) == expect
    for k in out:
        ref = G[f"maps/out/{k}"]
        assert out[k].shape == ref.shape, k
        assert linf(out[k].cpu(), ref) < 5e-6, (k, linf(out[k].cpu(), ref))
    assert "gradients" not in render_out and "pts" not in render_out


def test_render_maps_on_kernel_output_full_patch():
    """End of the chain at a 64x64 patch: fused renderer -> render_maps kernel vs oracle render -> oracle maps."""
    from helpers import load_params
    from oracle import neus_oracle as O
    from object_intrinsics_b200 import fields, generator_ops
    from object_intrinsics_b200.renderer import NeuSRenderer
    P = load_params("params_D8.npz")
    sdf, col, dev = fields.build_networks(D=8, device="cuda")
    fields.load_flat_params(sdf, col, dev, P)
    r = NeuSRenderer(None, sdf, dev, col, n_samples=32, n_importance=0, n_outside=0, up_sample_steps=1, perturb=0)
    bs, res = 2, 32
    ro, rd, near, far = O.synthetic_rays(bs, res, seed=3)
    z = torch.randn(bs, 64, generator=torch.Generator().manual_seed(3))
    w = O.style_mlp(P, z)
    with torch.no_grad():
        out = r.render(ro.cuda(), rd.cuda(), near.cuda(), far.cuda(), cos_anneal_ratio=1.0, perturb_overwrite=0,
                       z=z.cuda(), w=w.cuda())
    ref_out = O.render(P, ro, rd, near, far, w=w, n_samples=32, cos_anneal_ratio=1.0)
    w2b = torch.eye(4).repeat(bs, 1, 1)
    direction = torch.tensor([0.3, -0.5, -0.8])
    direction = direction / direction.norm()
    bg = torch.tensor([[0.1, 0.5, 0.9], [0.7, 0.2, 0.4]])
    base = types.SimpleNamespace(param_direction=direction.cuda(), ambient_color=torch.full((3,), 0.4),
                                 diffuse_color=torch.full((3,), 0.6), specular_color=torch.full((3,), 0.3),
                                 shininess=torch.tensor(7.0))
    gen = types.SimpleNamespace(resolution=res, bg_color=lambda n: bg[:, :, None, None].expand(n, 3, res, res))
    maps = generator_ops.render_maps(gen, bs, dict(out), {"rays_o": ro.cuda().reshape(bs, res, res, 3)},
                                     {"light": types.SimpleNamespace(light=base, w2b=w2b.cuda())}, True)
    ref = GO.render_maps(bs, res, ref_out, ro, GO.light_batch_direction(w2b, direction), torch.full((3,), 0.4),
                         torch.full((3,), 0.6), torch.full((3,), 0.3), torch.tensor(7.0),
                         bg[:, :, None, None].expand(bs, 3, res, res), True)
    for k in ref:
        assert linf(maps[k].cpu(), ref[k]) < 1e-4, (k, linf(maps[k].cpu(), ref[k]))


@pytest.mark.parametrize("return_raw", [True, False])
def test_render_maps_backward_matches_fp64_autograd(return_raw):
    """oi_render_maps_backward vs torch.autograd through the oracle restatement of Generator.render_maps in fp64:
    gradients of a random linear functional of every map w.r.t. weights / normals / albedo / weight_sum /
    color_fine and the light's four parameters (lighting.py:13-52) and direction."""
    from object_intrinsics_b200 import generator_ops
    G = load()
    amb, dif, spec, shin = [float(v) for v in G["maps/light"]]
    bs, res = G["maps/bg"].shape[0], int(G["rays/scalars"][1])
    gen = _fake_generator(G, bg=G["maps/bg"].cuda())
    keys = ("weights", "gradients", "raw_color", "weight_sum", "color_fine")
    render_out = {k[len("maps/in/"):]: v.cuda() for k, v in G.items() if k.startswith("maps/in/")}
    for k in keys:
        render_out[k] = render_out[k].clone().requires_grad_(True)
    # the light as the reference parametrises it: ambient = sigmoid(p), diffuse = 1 - sigmoid(p), specular, shininess
    p_amb = torch.tensor(amb / (amb + dif)).logit().cuda().requires_grad_(True)
    p_spec = torch.tensor(max(spec, 0.05)).cuda().requires_grad_(True)
    p_shin = torch.tensor(shin).cuda().requires_grad_(True)
    p_dir = torch.tensor([0.2, -0.3, -1.0]).cuda().requires_grad_(True)
    base = types.SimpleNamespace(param_direction=p_dir, ambient_color=torch.sigmoid(p_amb).expand(3),
                                 diffuse_color=(1 - torch.sigmoid(p_amb)).expand(3),
                                 specular_color=p_spec.expand(3).clamp(min=0), shininess=p_shin)
    light = types.SimpleNamespace(light=base, w2b=G["maps/w2b"].cuda())
    leaves = [render_out[k] for k in keys] + [p_amb, p_spec, p_shin, p_dir]
    rays_o = G["rays/rays_o"].cuda()
    out = generator_ops.render_maps(gen, bs, dict(render_out), {"rays_o": rays_o}, {"light": light}, return_raw)
    gen_w = torch.Generator().manual_seed(5)
    coef = {k: torch.randn(v.shape, generator=gen_w) for k, v in out.items() if k != "z_min"}
    loss = sum((out[k] * coef[k].cuda()).sum() for k in coef)
    grads = torch.autograd.grad(loss, leaves)
    # fp64 oracle
    ro64 = {k: v.detach().cpu().double() for k, v in render_out.items()}
    for k in keys:
        ro64[k].requires_grad_(True)
    q = [t.detach().cpu().double().requires_grad_(True) for t in (p_amb, p_spec, p_shin, p_dir)]
    direction = q[3] / torch.linalg.norm(q[3])
    ref = GO.render_maps(bs, res, ro64, rays_o.cpu().double().reshape(-1, 3),
                         GO.light_batch_direction(G["maps/w2b"].double(), direction),
                         torch.sigmoid(q[0]).expand(3), (1 - torch.sigmoid(q[0])).expand(3),
                         q[1].expand(3).clamp(min=0), q[2],
                         G["maps/bg"].double()[:, :, None, None].expand(bs, 3, res, res), return_raw)
    loss64 = sum((ref[k] * coef[k].double()).sum() for k in coef)
    grads64 = torch.autograd.grad(loss64, [ro64[k] for k in keys] + q)
    names = list(keys) + ["param_ambient", "param_specular", "param_shininess", "param_direction"]
    for n, g, g64 in zip(names, grads, grads64):
        scale = float(g64.abs().max()) + 1e-12
        err = float((g.detach().cpu().double().reshape(g64.shape) - g64).abs().max()) / scale
        assert err < 2e-4, (n, err, scale)


@pytest.mark.parametrize("n,m,impl,raw,in_kernel", [(64, 0, "tcgen05", True, True), (32, 0, "tcgen05", False, True),
                                                    (64, 0, "tcgen05", True, False), (16, 4, "tcgen05", True, True),
                                                    (16, 0, "ffma", False, False)])
def test_contract_b_single_call_matches_render_then_maps(n, m, impl, raw, in_kernel):
    """`render_and_maps` (OiRenderDesc.maps: maps composited in the tail of the render kernel when 128 % S == 0,
    else render -> composite -> maps kernel inside the one call) vs the two-call path `render` + `render_maps`,
    which the tests above pin to the reference's fixtures."""
    from helpers import load_params
    from oracle import neus_oracle as O
    from object_intrinsics_b200 import fields, generator_ops
    from object_intrinsics_b200.renderer import NeuSRenderer
    P = load_params("params_D8.npz")
    sdf, col, dev = fields.build_networks(D=8, device="cuda")
    fields.load_flat_params(sdf, col, dev, P)
    r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=0,
                     impl=impl)
    if in_kernel:
        r.flags |= 16                                             # maps composited in the render kernel's tile tail
    bs, res = 3, 10                                               # 300 rays: ragged last tile of every instance
    ro, rd, near, far = [t.cuda() for t in O.synthetic_rays(bs, res, seed=4)]
    w = O.style_mlp(P, torch.randn(bs, 64, generator=torch.Generator().manual_seed(4))).cuda()
    direction = torch.tensor([0.3, -0.5, -0.8])
    base = types.SimpleNamespace(param_direction=(direction / direction.norm()).cuda(),
                                 ambient_color=torch.full((3,), 0.4), diffuse_color=torch.full((3,), 0.6),
                                 specular_color=torch.full((3,), 0.3), shininess=torch.tensor(7.0))
    bg = torch.tensor([[0.1, 0.5, 0.9], [0.7, 0.2, 0.4], [0.3, 0.3, 0.8]]).cuda()
    gen = types.SimpleNamespace(resolution=res, bg_color=lambda k: bg[:, :, None, None].expand(k, 3, res, res))
    prior = {"light": types.SimpleNamespace(light=base, w2b=torch.eye(4).repeat(bs, 1, 1).cuda())}
    rays_info = {"rays_o": ro.reshape(bs, res, res, 3), "rays_d": rd.reshape(bs, res, res, 3), "near": near, "far": far}
    with torch.no_grad():
        out2 = r.render(ro, rd, near, far, cos_anneal_ratio=0.8, perturb_overwrite=0, w=w)
        maps2 = generator_ops.render_maps(gen, bs, dict(out2), rays_info, prior, raw)
        launches2 = r.last_launches + 1
        out1, maps1 = generator_ops.render_and_maps(gen, r, bs, rays_info, prior, w, return_raw=raw,
                                                    cos_anneal_ratio=0.8, perturb_overwrite=0)
    torch.cuda.synchronize()
    assert r.last_launches <= launches2 - (1 if in_kernel and (n + m) in (16, 32, 64, 128) and impl == "tcgen05" else 0)
    assert set(maps1) == set(maps2)
    for k in maps2:
        assert linf(maps1[k].cpu(), maps2[k].cpu()) < 5e-6, (k, linf(maps1[k].cpu(), maps2[k].cpu()))
    for k in ("weight_sum", "weight_max", "color_fine", "s_val", "gradient_error", "surface_loss"):
        assert linf(out1[k].cpu(), out2[k].cpu()) < 5e-6, k
    assert "weights" not in out1 and "pts" not in out1            # nothing per-point is materialised
