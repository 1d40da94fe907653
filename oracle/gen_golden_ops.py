"""TEST INFRASTRUCTURE -- generates tests/golden/ops_*.npz from the reference's own CPU reference
implementations of the StyleGAN2 ops (run in the authoring container only; needs /root/reference):

    python -m oracle.gen_golden_ops

* ADA upfirdn2d: `upfirdn2d.upfirdn2d(..., impl='ref')`, `upsample2d`, `downsample2d` with the sym6 filter the
  AugmentPipe uses (ada/augment.py:24,116; ada/torch_utils/ops/upfirdn2d.py:120-384)
* stylesdf upfirdn2d: `upfirdn2d_native` (stylesdf/op/upfirdn2d.py:160-201)
* ADA bias_act: `_bias_act_ref` (ada/torch_utils/ops/bias_act.py:93-123); the module imports the un-vendored
  `dnnlib` only for `EasyDict`, which is shimmed here with a 3-line attribute dict
* stylesdf fused_leaky_relu: the CPU branch (stylesdf/op/fused_act.py:104-115)
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as RH  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148]   # ada/augment.py:24 wavelets['sym6']


def main():
    RH.import_reference()
    class EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    sys.modules.setdefault("dnnlib", types.SimpleNamespace(EasyDict=EasyDict))
    from src.third_party.ada.torch_utils.ops import upfirdn2d as U
    from src.third_party.ada.torch_utils.ops import bias_act as B
    from src.third_party.stylesdf.op.upfirdn2d import upfirdn2d_native
    from src.third_party.stylesdf.op.fused_act import fused_leaky_relu

    g = torch.Generator().manual_seed(0)
    blob = {}
    # ---- ADA upfirdn2d, generic 2-D filters
    cases = [
        # name, x shape, f shape, up, down, padding, flip, gain
        ("u0", (2, 3, 9, 11), (3, 3), (1, 1), (1, 1), (1, 1, 1, 1), False, 1.0),
        ("u1", (1, 2, 8, 8), (4, 4), (2, 2), (1, 1), (2, 1, 2, 1), False, 4.0),
        ("u2", (2, 1, 13, 10), (4, 4), (1, 1), (2, 2), (1, 1, 1, 1), True, 1.0),
        ("u3", (1, 3, 7, 9), (5, 3), (2, 3), (3, 2), (3, -1, 0, 2), False, 0.5),
        ("u4", (1, 1, 6, 6), (1, 1), (1, 1), (1, 1), (0, 0, 0, 0), False, 2.0),
        ("u5", (2, 2, 10, 12), (2, 6), (1, 2), (2, 1), (-2, 3, 1, -1), True, 1.5),
    ]
    for name, xs, fs, up, down, pad, flip, gain in cases:
        x = torch.randn(xs, generator=g)
        f = torch.randn(fs, generator=g)
        y = U.upfirdn2d(x, f, up=list(up), down=list(down), padding=list(pad), flip_filter=flip, gain=gain, impl="ref")
        blob[f"{name}/x"], blob[f"{name}/f"], blob[f"{name}/y"] = x.numpy(), f.numpy(), y.numpy()
        blob[f"{name}/cfg"] = np.array([*up, *down, *pad, int(flip)], dtype=np.int64)
        blob[f"{name}/gain"] = np.array(gain, dtype=np.float64)
    # ---- the AugmentPipe calls: separable sym6, up 2 / down 2 (augment.py:290,301)
    f6 = U.setup_filter(SYM6)
    assert f6.ndim == 1 and f6.numel() == 12
    x = torch.randn(2, 3, 20, 23, generator=g)
    blob["aug/x"], blob["aug/f"] = x.numpy(), f6.numpy()
    blob["aug/up2"] = U.upsample2d(x=x, f=f6, up=2, impl="ref").numpy()
    blob["aug/down2"] = U.downsample2d(x=x, f=f6, down=2, padding=-3 * 2, flip_filter=True, impl="ref").numpy()
    blob["aug/filter2d"] = U.filter2d(x=x, f=f6, impl="ref").numpy()
    # ---- stylesdf upfirdn2d_native (blur kernel [1,3,3,1] outer product, up 2 / down 2 as in stylesdf/model.py:75-133)
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k1[None, :] * k1[:, None]
    k = k / k.sum()
    x = torch.randn(2, 3, 8, 8, generator=g)
    blob["sdf/x"], blob["sdf/k"] = x.numpy(), k.numpy()
    blob["sdf/up2"] = upfirdn2d_native(x, k * 4, 2, 2, 1, 1, 2, 1, 2, 1).numpy()
    blob["sdf/down2"] = upfirdn2d_native(x, k, 1, 1, 2, 2, 1, 1, 1, 1).numpy()
    blob["sdf/blur"] = upfirdn2d_native(x, k, 1, 1, 1, 1, 2, 1, 2, 1).numpy()
    # ---- bias_act forward, all activations
    x = torch.randn(3, 5, 4, 6, generator=g) * 2
    b = torch.randn(5, generator=g)
    blob["ba/x"], blob["ba/b"] = x.numpy(), b.numpy()
    for act in B.activation_funcs:
        blob[f"ba/{act}"] = B._bias_act_ref(x, b, dim=1, act=act).numpy()
        blob[f"ba/{act}_clamp"] = B._bias_act_ref(x, b, dim=1, act=act, alpha=0.3, gain=1.7, clamp=0.9).numpy()
    blob["ba/nobias_dim3"] = B._bias_act_ref(x, torch.arange(6.0), dim=3, act="lrelu").numpy()
    # ---- stylesdf fused_leaky_relu (CPU branch)
    x = torch.randn(4, 64, generator=g)
    b = torch.randn(64, generator=g)
    blob["flr/x"], blob["flr/b"] = x.numpy(), b.numpy()
    blob["flr/y_scale1"] = fused_leaky_relu(x, b, scale=1).numpy()
    blob["flr/y_default"] = fused_leaky_relu(x, b).numpy()
    x4 = torch.randn(2, 6, 3, 3, generator=g)
    b4 = torch.randn(6, generator=g)
    blob["flr/x4"], blob["flr/b4"], blob["flr/y4"] = x4.numpy(), b4.numpy(), fused_leaky_relu(x4, b4).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "ops_golden.npz"), **blob)
    print("wrote", len(blob), "arrays")


if __name__ == "__main__":
    main()
