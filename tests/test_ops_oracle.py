"""CPU tests: oracle/ops_oracle.py against golden vectors produced by the reference's own CPU reference
implementations of upfirdn2d / bias_act / fused_leaky_relu (tests/golden/ops_golden.npz)."""
import os

import numpy as np
import torch

from helpers import GOLDEN, linf
from oracle import ops_oracle as OO


def _g():
    with np.load(os.path.join(GOLDEN, "ops_golden.npz")) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def test_upfirdn2d_generic():
    G = _g()
    for name in ("u0", "u1", "u2", "u3", "u4", "u5"):
        upx, upy, dx, dy, px0, px1, py0, py1, flip = [int(v) for v in G[f"{name}/cfg"]]
        y = OO.upfirdn2d(G[f"{name}/x"], G[f"{name}/f"], upx, upy, dx, dy, px0, px1, py0, py1, bool(flip),
                         float(G[f"{name}/gain"]))
        assert y.shape == G[f"{name}/y"].shape, name
        assert linf(y, G[f"{name}/y"]) < 1e-5, (name, linf(y, G[f"{name}/y"]))


def test_upfirdn2d_augment_pipe_calls():
    G = _g()
    x, f = G["aug/x"], G["aug/f"]
    up = OO.upfirdn2d_separable(x, f, up=2, padding=OO.upsample2d_padding(12, 12, 2), gain=4.0)
    assert linf(up, G["aug/up2"]) < 1e-5
    dn = OO.upfirdn2d_separable(x, f, down=2, padding=OO.downsample2d_padding(12, 12, 2, (-6, -6, -6, -6)),
                                flip_filter=True)
    assert dn.shape == G["aug/down2"].shape and linf(dn, G["aug/down2"]) < 1e-5
    fl = OO.upfirdn2d_separable(x, f, padding=(6, 5, 6, 5))
    assert linf(fl, G["aug/filter2d"]) < 1e-5


def test_upfirdn2d_stylesdf_flavour():
    G = _g()
    x, k = G["sdf/x"], G["sdf/k"]
    assert linf(OO.upfirdn2d(x, k * 4, 2, 2, 1, 1, 2, 1, 2, 1), G["sdf/up2"]) < 1e-5
    assert linf(OO.upfirdn2d(x, k, 1, 1, 2, 2, 1, 1, 1, 1), G["sdf/down2"]) < 1e-5
    assert linf(OO.upfirdn2d(x, k, 1, 1, 1, 1, 2, 1, 2, 1), G["sdf/blur"]) < 1e-5


def test_bias_act_forward_all_activations():
    G = _g()
    x, b = G["ba/x"], G["ba/b"]
    for act in OO.ACTS:
        assert linf(OO.bias_act(x, b, 1, act), G[f"ba/{act}"]) < 1e-6, act
        assert linf(OO.bias_act(x, b, 1, act, alpha=0.3, gain=1.7, clamp=0.9), G[f"ba/{act}_clamp"]) < 1e-6, act
    assert linf(OO.bias_act(x, torch.arange(6.0), 3, "lrelu"), G["ba/nobias_dim3"]) < 1e-6


def test_fused_leaky_relu():
    G = _g()
    assert linf(OO.fused_leaky_relu(G["flr/x"], G["flr/b"], scale=1), G["flr/y_scale1"]) < 1e-6
    assert linf(OO.fused_leaky_relu(G["flr/x"], G["flr/b"]), G["flr/y_default"]) < 1e-6
    assert linf(OO.fused_leaky_relu(G["flr/x4"], G["flr/b4"]), G["flr/y4"]) < 1e-6
