"""GPU parity tests of the StyleGAN2 ops (C-ABI oi_upfirdn2d / oi_bias_act / oi_fused_bias_act through the
reference-shaped Python wrappers) against oracle/ops_oracle.py and the reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, linf
from oracle import ops_oracle as OO

pytestmark = pytest.mark.gpu


def _g():
    with np.load(os.path.join(GOLDEN, "ops_golden.npz")) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def test_upfirdn2d_golden_and_layouts():
    from object_intrinsics_b200.ops import upfirdn2d as U
    G = _g()
    for name in ("u0", "u1", "u2", "u3", "u4", "u5"):
        upx, upy, dx, dy, px0, px1, py0, py1, flip = [int(v) for v in G[f"{name}/cfg"]]
        x, f = G[f"{name}/x"].cuda(), G[f"{name}/f"].cuda()
        y = U.upfirdn2d(x, f, up=[upx, upy], down=[dx, dy], padding=[px0, px1, py0, py1], flip_filter=bool(flip),
                        gain=float(G[f"{name}/gain"]))
        assert linf(y.cpu(), G[f"{name}/y"]) < 2e-5, name
        if x.shape[1] > 1:   # channels-last input -> channels-last output, same values
            ycl = U.upfirdn2d(x.contiguous(memory_format=torch.channels_last), f, up=[upx, upy], down=[dx, dy],
                              padding=[px0, px1, py0, py1], flip_filter=bool(flip), gain=float(G[f"{name}/gain"]))
            assert ycl.stride(1) == 1 and linf(ycl.cpu(), G[f"{name}/y"]) < 2e-5, name


def test_upfirdn2d_augment_pipe_calls_and_dtypes():
    from object_intrinsics_b200.ops import upfirdn2d as U
    G = _g()
    x, f = G["aug/x"].cuda(), G["aug/f"].cuda()
    assert linf(U.upsample2d(x, f, up=2).cpu(), G["aug/up2"]) < 2e-5
    assert linf(U.downsample2d(x, f, down=2, padding=-6, flip_filter=True).cpu(), G["aug/down2"]) < 2e-5
    assert linf(U.filter2d(x, f).cpu(), G["aug/filter2d"]) < 2e-5
    assert linf(U.upsample2d(x.double(), f, up=2).cpu(), G["aug/up2"]) < 1e-6
    assert linf(U.upsample2d(x.half(), f, up=2).float().cpu(), G["aug/up2"]) < 2e-2


def test_upfirdn2d_stylesdf_flavour_and_errors():
    from object_intrinsics_b200.ops import upfirdn2d as U
    G = _g()
    x, k = G["sdf/x"].cuda(), G["sdf/k"].cuda()
    assert linf(U.upfirdn2d_native_layout(x, k * 4, up=2, down=1, pad=(2, 1)).cpu(), G["sdf/up2"]) < 2e-5
    assert linf(U.upfirdn2d_native_layout(x, k, up=1, down=2, pad=(1, 1)).cpu(), G["sdf/down2"]) < 2e-5
    assert linf(U.upfirdn2d_native_layout(x, k, up=1, down=1, pad=(2, 1)).cpu(), G["sdf/blur"]) < 2e-5
    with pytest.raises(RuntimeError):
        U.upfirdn2d_raw(x.cpu(), k.cpu(), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)        # no CPU path
    with pytest.raises(RuntimeError):
        U.upfirdn2d_raw(x, k, 1, 1, 1, 1, -4, -4, 0, 0, False, 1.0)                  # output < 1x1


def test_upfirdn2d_gradients_first_and_second_order():
    from object_intrinsics_b200.ops import upfirdn2d as U
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 9, 8, generator=g, dtype=torch.float64)
    f = torch.randn(4, 3, generator=g)
    cfg = dict(up=[2, 1], down=[1, 2], padding=[2, 1, 0, 3], flip_filter=False, gain=1.3)
    xc = x.cuda().requires_grad_(True)
    y = U.upfirdn2d(xc, f.cuda(), **cfg)
    xo = x.clone().requires_grad_(True)
    yo = OO.upfirdn2d(xo, f.double(), 2, 1, 1, 2, 2, 1, 0, 3, False, 1.3)
    assert linf(y.detach().cpu(), yo.detach()) < 1e-6   # filter is fp32 inside the kernel
    w = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (gx,) = torch.autograd.grad((y * w.cuda()).sum(), xc, create_graph=True)
    (gxo,) = torch.autograd.grad((yo * w).sum(), xo, create_graph=True)
    assert linf(gx.detach().cpu(), gxo.detach()) < 1e-6
    # the op is linear: the double-backward w.r.t. the incoming gradient is the forward op again
    v = torch.randn(x.shape, generator=g, dtype=torch.float64)
    wc = w.cuda().requires_grad_(True)
    (gx2,) = torch.autograd.grad((U.upfirdn2d(xc, f.cuda(), **cfg) * wc).sum(), xc, create_graph=True)
    (gw,) = torch.autograd.grad((gx2 * v.cuda()).sum(), wc)
    assert linf(gw.cpu(), OO.upfirdn2d(v, f.double(), 2, 1, 1, 2, 2, 1, 0, 3, False, 1.3)) < 1e-6


def test_bias_act_forward_golden():
    from object_intrinsics_b200.ops.bias_act import bias_act
    G = _g()
    x, b = G["ba/x"].cuda(), G["ba/b"].cuda()
    for act in OO.ACTS:
        assert linf(bias_act(x, b, dim=1, act=act).cpu(), G[f"ba/{act}"]) < 2e-6, act
        assert linf(bias_act(x, b, dim=1, act=act, alpha=0.3, gain=1.7, clamp=0.9).cpu(),
                    G[f"ba/{act}_clamp"]) < 2e-6, act
    assert linf(bias_act(x, torch.arange(6.0).cuda(), dim=3, act="lrelu").cpu(), G["ba/nobias_dim3"]) < 2e-6
    xcl = x.contiguous(memory_format=torch.channels_last)
    ycl = bias_act(xcl, b, dim=1, act="swish")
    assert ycl.stride() == xcl.stride() and linf(ycl.cpu(), G["ba/swish"]) < 2e-6


@pytest.mark.parametrize("act", list(OO.ACTS))
def test_bias_act_first_and_second_order_gradients(act):
    from object_intrinsics_b200.ops.bias_act import bias_act
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(3, 4, 5, generator=g, dtype=torch.float64) * 1.5)
    x = torch.where(x.abs() < 0.05, x + 0.2, x)       # keep away from the relu/lrelu kinks
    b = torch.randn(4, generator=g, dtype=torch.float64)
    w = torch.randn(3, 4, 5, generator=g, dtype=torch.float64)
    v = torch.randn(3, 4, 5, generator=g, dtype=torch.float64)

    def run(fn, dev):
        xx = x.to(dev).requires_grad_(True)
        bb = b.to(dev).requires_grad_(True)
        y = fn(xx, bb, dim=1, act=act, gain=1.25)   # alpha/gain/clamp cross the C-ABI as fp32 (like the reference's)
        gx, gb = torch.autograd.grad((y * w.to(dev)).sum(), [xx, bb], create_graph=True)
        tot = (gx * v.to(dev)).sum() + gb.sum()
        ggx, ggb = (torch.autograd.grad(tot, [xx, bb], allow_unused=True) if tot.requires_grad else (None, None))
        z = lambda t, like: torch.zeros_like(like) if t is None else t
        return [y.detach().cpu(), gx.detach().cpu(), gb.detach().cpu(), z(ggx, xx).cpu(), z(ggb, bb).cpu()]

    ours = run(bias_act, "cuda")
    ref = run(OO.bias_act, "cpu")
    for name, a, r in zip(("y", "dx", "db", "d2x", "d2b"), ours, ref):
        assert linf(a, r) < 1e-6 * max(1.0, float(r.abs().max())), (act, name, linf(a, r))


def test_fused_leaky_relu_golden_and_gradients():
    from object_intrinsics_b200.ops.fused_act import fused_leaky_relu, fused_bias_act
    G = _g()
    x, b = G["flr/x"].cuda(), G["flr/b"].cuda()
    assert linf(fused_leaky_relu(x, b, scale=1).cpu(), G["flr/y_scale1"]) < 1e-6
    assert linf(fused_leaky_relu(x, b).cpu(), G["flr/y_default"]) < 1e-6
    assert linf(fused_leaky_relu(G["flr/x4"].cuda(), G["flr/b4"].cuda()).cpu(), G["flr/y4"]) < 1e-6
    # raw op modes (fused_bias_act_kernel.cu:19-52)
    for act, grad in ((1, 0), (1, 1), (1, 2), (3, 0), (3, 1), (3, 2)):
        ref_t = torch.randn(x.shape, generator=torch.Generator().manual_seed(act * 10 + grad)).cuda()
        y = fused_bias_act(x, b, ref_t, act, grad, 0.2, 1.5)
        assert linf(y.cpu(), OO.fused_bias_act(x.cpu(), b.cpu(), ref_t.cpu(), act, grad, 0.2, 1.5)) < 1e-6
    # gradients incl. double backward against autograd through the oracle
    xd = G["flr/x4"].double()
    bd = G["flr/b4"].double()

    def run(fn, dev):
        xx, bb = xd.to(dev).requires_grad_(True), bd.to(dev).requires_grad_(True)
        y = fn(xx, bb, 0.25, 1.75)
        gx, gb = torch.autograd.grad((y * y).sum(), [xx, bb], create_graph=True)
        ggx, ggb = torch.autograd.grad(gx.pow(2).sum() + gb.pow(2).sum(), [xx, bb])
        return [t.detach().cpu() for t in (y, gx, gb, ggx, ggb)]

    for a, r in zip(run(fused_leaky_relu, "cuda"), run(OO.fused_leaky_relu, "cpu")):
        assert linf(a, r) < 1e-9 * max(1.0, float(r.abs().max()))
    with pytest.raises(RuntimeError):
        fused_leaky_relu(G["flr/x"], G["flr/b"])      # CPU tensors: no CPU path


def test_style_mlp_matches_oracle():
    """ShapeNetwork.style through MappingLinear + the fused op, and through the single-kernel oi_style_mlp."""
    import ctypes as C
    from helpers import load_params
    from oracle import neus_oracle as O
    from object_intrinsics_b200 import _lib, fields
    from object_intrinsics_b200.renderer import collect_params, fill_net_params
    P = load_params("params_D8.npz")
    sdf, col, dev = fields.build_networks(D=8, device="cuda")
    fields.load_flat_params(sdf, col, dev, P)
    z = torch.randn(5, 64, generator=torch.Generator().manual_seed(2))
    ref = O.style_mlp(P, z)
    with torch.no_grad():
        assert linf(sdf.style(z.cuda()).cpu(), ref) < 2e-6
    p = fill_net_params({k: t.detach() for k, t in collect_params(sdf, col, dev)})
    w = torch.empty(5, 64, device="cuda")
    _lib.check(_lib.lib().oi_style_mlp(C.byref(p), z.cuda().data_ptr(), w.data_ptr(), 5, _lib.current_stream_ptr()),
               "oi_style_mlp")
    torch.cuda.synchronize()
    assert linf(w.cpu(), ref) < 2e-6
