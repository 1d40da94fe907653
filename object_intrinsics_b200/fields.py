"""Parameter containers with the reference's module structure and state_dict keys.

When this package is used as a drop-in inside the reference tree, the reference's own `ShapeNetwork`,
`ColorNetwork` and `SingleVarianceNetwork` (src/models/fields.py:10-101, neus/models/fields.py:262-268) are
kept and only the renderer is swapped.  On machines without the reference (the GPU test box, bench.py) these
classes provide the same attribute tree -- `sdf_network.style[i]`, `.pts_linears[l].{weight,bias,gamma,beta}`,
`.sigma_linear`, `color_network.views_linears`, `.rgb_linear`, `deviation_network.variance` -- so that
checkpoints / fixtures written from the reference's `state_dict()` load unchanged and the renderer finds the
tensors under the same names.  The per-point evaluation of these nets is NOT done here: it happens inside the
fused CUDA kernel.  The module `forward`s below are plain differentiable torch formulations used by the
grad-mode path (`torch_graph.py`) and for small utility queries.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops.fused_act import fused_leaky_relu


class LinearLayer(nn.Module):
    """y = std_init * (x W^T + b) + bias_init   (stylesdf/volume_renderer.py:12-30)."""

    def __init__(self, in_dim, out_dim, bias_init=0.0, std_init=1.0, weight_bound=None):
        super().__init__()
        bound = weight_bound if weight_bound is not None else math.sqrt(6.0 / in_dim) / 25
        self.weight = nn.Parameter(torch.empty(out_dim, in_dim).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(out_dim).uniform_(-math.sqrt(1 / in_dim), math.sqrt(1 / in_dim)))
        self.bias_init, self.std_init = bias_init, std_init

    def forward(self, x):
        return self.std_init * F.linear(x, self.weight, self.bias) + self.bias_init


class _StyleLinear(LinearLayer):
    """gamma/beta heads: kaiming-normal(0.2) * 0.25 weights (volume_renderer.py:21)."""

    def __init__(self, in_dim, out_dim, bias_init, std_init):
        super().__init__(in_dim, out_dim, bias_init, std_init)
        with torch.no_grad():
            nn.init.kaiming_normal_(self.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
            self.weight.mul_(0.25)


class FiLMSiren(nn.Module):
    """h = sin(gamma(w) * (x W^T + b) + beta(w))   (volume_renderer.py:33-61)."""

    def __init__(self, in_channel, out_channel, style_dim, is_first=False):
        super().__init__()
        bound = 1.0 / 3 if is_first else math.sqrt(6.0 / in_channel) / 25
        self.weight = nn.Parameter(torch.empty(out_channel, in_channel).uniform_(-bound, bound))
        self.bias = nn.Parameter(
            torch.empty(out_channel).uniform_(-math.sqrt(1 / in_channel), math.sqrt(1 / in_channel)))
        self.gamma = _StyleLinear(style_dim, out_channel, bias_init=30.0, std_init=15.0)
        self.beta = _StyleLinear(style_dim, out_channel, bias_init=0.0, std_init=0.25)

    def film(self, w):
        return self.gamma(w), self.beta(w)

    def forward(self, x, w):
        """x: [bs, n, C_in], w: [bs, style]."""
        g, b = self.film(w)
        return torch.sin(g[:, None, :] * F.linear(x, self.weight, self.bias) + b[:, None, :])


class MappingLinear(nn.Module):
    """Style-MLP layer: linear without bias, then fused bias + leaky_relu(0.2) * 1  (stylesdf/model.py:32-56)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        w = torch.empty(out_dim, in_dim)
        nn.init.kaiming_normal_(w, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.empty(out_dim).uniform_(-math.sqrt(1 / in_dim), math.sqrt(1 / in_dim)))

    def forward(self, x):
        return fused_leaky_relu(F.linear(x, self.weight), self.bias, negative_slope=0.2, scale=1)


class StyleMLP(nn.Sequential):
    """ShapeNetwork.style (fields.py:14-19).  Same children / state_dict keys as the reference's nn.Sequential of
    three MappingLinear; when no gradient is needed the whole MLP is ONE kernel (C-ABI `oi_style_mlp`) instead of
    three GEMV + three fused bias/activation launches."""

    def forward(self, z):
        needs_grad = torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad or not z.is_cuda or z.dtype != torch.float32 or len(self) != 3:
            return super().forward(z)
        import ctypes as C
        from . import _lib
        # the one-kernel path is written for style_dim = 64 square layers of contiguous fp32 tensors on z's device
        # (the reference's ShapeNetwork); anything else (e.g. StyleSDF's default 256) takes the per-layer path
        ok = z.dim() == 2 and z.shape[-1] == _lib.OI_STYLE_DIM and all(
            tuple(l.weight.shape) == (_lib.OI_STYLE_DIM, _lib.OI_STYLE_DIM) and l.weight.is_contiguous() and
            l.bias.is_contiguous() and l.weight.dtype == torch.float32 and l.weight.device == z.device
            for l in self)
        if not ok:
            return super().forward(z)
        p = _lib.OiNetParams()
        p.depth, p.width, p.style_dim = 1, _lib.OI_WIDTH, z.shape[-1]
        for i, layer in enumerate(self):
            p.style_weight[i], p.style_bias[i] = layer.weight.data_ptr(), layer.bias.data_ptr()
        zz = z.contiguous()
        w = torch.empty_like(zz)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().oi_style_mlp(C.byref(p), zz.data_ptr(), w.data_ptr(), zz.shape[0],
                                               _lib.current_stream_ptr(z.device)), "oi_style_mlp")
        return w


class ShapeNetwork(nn.Module):
    """FiLM-SIREN SDF network + 3-layer style MLP  (src/models/fields.py:10-77)."""

    def __init__(self, D=8, W=128, style_dim=64, input_ch=3, input_ch_views=3, checkpoint_path=None):
        super().__init__()
        if checkpoint_path is not None:
            raise NotImplementedError("load weights with load_state_dict / load_flat_params instead")
        self.style = StyleMLP(*[MappingLinear(style_dim, style_dim) for _ in range(3)])
        self.pts_linears = nn.ModuleList(
            [FiLMSiren(input_ch, W, style_dim, is_first=True)] + [FiLMSiren(W, W, style_dim) for _ in range(D - 1)])
        self.sigma_linear = LinearLayer(W, 1)

    def forward(self, x, z=None, w=None):
        """[N,3] -> [N, 1+W]; instance of point i is i // (N/bs)  (fields.py:49-70)."""
        latent = self.style(z) if w is None else w
        bs = latent.shape[0]
        h = x.reshape(bs, x.shape[0] // bs, x.shape[-1])
        for layer in self.pts_linears:
            h = layer(h, latent)
        return torch.cat([self.sigma_linear(h), h], -1).flatten(0, 1)

    def sdf(self, x, z=None, w=None):
        return self.forward(x, z=z, w=w)[:, :1]


class ColorNetwork(nn.Module):
    """rgb = sigmoid(W_rgb FiLMSiren([features, normal]) + b)   (src/models/fields.py:80-101)."""

    def __init__(self, D=8, W=128, style_dim=64, input_ch=3, input_ch_views=3):
        super().__init__()
        self.views_linears = FiLMSiren(input_ch_views + W, W, style_dim)
        self.rgb_linear = LinearLayer(W, 3)
        self.style_dim, self.w_dim = style_dim, W

    def forward(self, points, normals, view_dirs, feature_vectors, z=None, w=None):
        bs = w.shape[0]
        x = torch.cat([feature_vectors, normals], -1)
        x = x.reshape(bs, x.shape[0] // bs, x.shape[-1])
        return torch.sigmoid(self.rgb_linear(self.views_linears(x, w))).flatten(0, 1)


class SingleVarianceNetwork(nn.Module):
    """inv_s = exp(10 * variance)   (neus/models/fields.py:262-268)."""

    def __init__(self, init_val=0.3):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(float(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)


def build_networks(D=8, W=128, style_dim=64, device="cuda", seed=0):
    torch.manual_seed(seed)
    sdf = ShapeNetwork(D=D, W=W, style_dim=style_dim)
    col = ColorNetwork(D=D, W=W, style_dim=style_dim)
    dev = SingleVarianceNetwork(0.3)
    return sdf.to(device), col.to(device), dev.to(device)


def load_flat_params(sdf_network, color_network, deviation_network, flat: Dict[str, torch.Tensor]):
    """Loads a flat {`sdf_network.x`: tensor} dict (reference Generator state_dict naming)."""
    for prefix, mod in (("sdf_network.", sdf_network), ("color_network.", color_network),
                        ("deviation_network.", deviation_network)):
        sd = {k[len(prefix):]: v for k, v in flat.items() if k.startswith(prefix)}
        mod.load_state_dict(sd, strict=True)
