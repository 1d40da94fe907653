"""CPU ORACLE for the StyleGAN2 ops -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-torch restatements of what the reference's native ops compute, each citing the reference file:line.
Pinned by tests/golden/ops_*.npz, generated from the reference's own CPU reference implementations
(`_upfirdn2d_ref`, `upfirdn2d_native`, `_bias_act_ref`, the CPU branch of `fused_leaky_relu`) by
oracle/gen_golden_ops.py.  The first/second-order gradient modes of bias_act / fused_bias_act have no CPU
reference in the tree (they exist only in bias_act.cu:23-147 / fused_bias_act_kernel.cu:19-52); their oracle is
autograd through the forward restatement, which is what those kernels are designed to equal.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# upfirdn2d
# ------------------------------------------------------------------------------------------------
def upfirdn2d(x, f, upx=1, upy=1, downx=1, downy=1, padx0=0, padx1=0, pady0=0, pady1=0, flip_filter=False, gain=1.0):
    """ada/torch_utils/ops/upfirdn2d.py:168-208 (`_upfirdn2d_ref`) for a rank-2 filter f [fh, fw]:
    zero-insert upsample, pad/crop, correlate with the (flipped unless flip_filter) filter, decimate."""
    n, c, h, w = x.shape
    y = x.reshape(n, c, h, 1, w, 1)
    y = F.pad(y, [0, upx - 1, 0, 0, 0, upy - 1])
    y = y.reshape(n, c, h * upy, w * upx)
    y = F.pad(y, [max(padx0, 0), max(padx1, 0), max(pady0, 0), max(pady1, 0)])
    y = y[:, :, max(-pady0, 0): y.shape[2] - max(-pady1, 0), max(-padx0, 0): y.shape[3] - max(-padx1, 0)]
    k = (f * gain).to(x.dtype)
    if not flip_filter:
        k = k.flip([0, 1])
    k = k[None, None].repeat(c, 1, 1, 1)
    y = F.conv2d(y, k, groups=c)
    return y[:, :, ::downy, ::downx]


def upfirdn2d_separable(x, f1, up=1, down=1, padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """Separable 1-D filter path: a row pass then a column pass with sqrt(gain) each
    (upfirdn2d.py:200-204 reference, :233-240 CUDA path)."""
    px0, px1, py0, py1 = padding
    g = math.sqrt(gain)
    y = upfirdn2d(x, f1[None, :], up, 1, down, 1, px0, px1, 0, 0, flip_filter, g)
    return upfirdn2d(y, f1[:, None], 1, up, 1, down, 0, 0, py0, py1, flip_filter, g)


def upsample2d_padding(fw, fh, up, padding=(0, 0, 0, 0)):
    """upfirdn2d.py:330-339."""
    px0, px1, py0, py1 = padding
    return (px0 + (fw + up - 1) // 2, px1 + (fw - up) // 2, py0 + (fh + up - 1) // 2, py1 + (fh - up) // 2)


def downsample2d_padding(fw, fh, down, padding=(0, 0, 0, 0)):
    """upfirdn2d.py:369-378."""
    px0, px1, py0, py1 = padding
    return (px0 + (fw - down + 1) // 2, px1 + (fw - down) // 2, py0 + (fh - down + 1) // 2, py1 + (fh - down) // 2)


# ------------------------------------------------------------------------------------------------
# bias_act
# ------------------------------------------------------------------------------------------------
ACTS = {
    # name: (function, def_alpha, def_gain, cuda_idx)            ada/torch_utils/ops/bias_act.py:23-33
    "linear": (lambda x, a: x, 0.0, 1.0, 1),
    "relu": (lambda x, a: F.relu(x), 0.0, math.sqrt(2), 2),
    "lrelu": (lambda x, a: F.leaky_relu(x, a), 0.2, math.sqrt(2), 3),
    "tanh": (lambda x, a: torch.tanh(x), 0.0, 1.0, 4),
    "sigmoid": (lambda x, a: torch.sigmoid(x), 0.0, 1.0, 5),
    "elu": (lambda x, a: F.elu(x), 0.0, 1.0, 6),
    "selu": (lambda x, a: F.selu(x), 0.0, 1.0, 7),
    "softplus": (lambda x, a: F.softplus(x), 0.0, 1.0, 8),
    "swish": (lambda x, a: torch.sigmoid(x) * x, 0.0, math.sqrt(2), 9),
}


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None):
    """ada/torch_utils/ops/bias_act.py:93-123 (`_bias_act_ref`)."""
    fn, def_alpha, def_gain, _ = ACTS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if b is not None:
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


# ------------------------------------------------------------------------------------------------
# stylesdf fused_bias_act
# ------------------------------------------------------------------------------------------------
def fused_bias_act(x, bias, ref, act, grad, alpha, scale):
    """stylesdf/op/fused_bias_act_kernel.cu:19-52: y = f(x + bias[(i / step) % size]) * scale with
    act*10+grad in {10, 11: identity; 30: lrelu; 31: slope chosen by the sign of `ref`; 12, 32: zero}."""
    v = x
    if bias is not None and bias.numel():
        v = v + bias.view(1, -1, *([1] * (x.ndim - 2)))
    mode = act * 10 + grad
    if mode in (12, 32):
        out = torch.zeros_like(v)
    elif mode == 30:
        out = torch.where(v > 0, v, v * alpha)
    elif mode == 31:
        out = torch.where(ref > 0, v, v * alpha)
    else:
        out = v
    return out * scale


def fused_leaky_relu(x, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """stylesdf/op/fused_act.py:104-119 with the CUDA kernel's semantics (slope honoured)."""
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)
