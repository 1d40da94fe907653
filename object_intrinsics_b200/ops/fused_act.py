"""`fused_leaky_relu` / `FusedLeakyReLU` with the reference's Python surface
(src/third_party/stylesdf/op/fused_act.py:21-119) on top of the C-ABI `oi_fused_bias_act`
(replaces `fused.fused_bias_act`, stylesdf/op/fused_bias_act.cpp:11-20).

Unlike the reference there is no CPU branch (fused_act.py:105-115): CPU tensors raise.  Note the reference's
CPU branch hard-codes slope 0.2 and ignores `negative_slope`; the CUDA kernel honours it, as does this one.
"""
import ctypes as C

import torch
from torch import nn

from .. import _lib


def fused_bias_act(x, bias, ref, act: int, grad: int, alpha: float, scale: float):
    """Same argument meaning as the reference's pybind op; empty tensors / None mean "absent"."""
    if not x.is_cuda:
        raise RuntimeError("fused_bias_act: CUDA tensors only (object_intrinsics_b200 has no CPU path)")
    if x.dtype not in _lib.DTYPE_CODE:
        raise TypeError(f"fused_bias_act: unsupported dtype {x.dtype}")
    x = x.contiguous()
    bias = None if (bias is None or bias.numel() == 0) else bias.contiguous().to(x.dtype)
    ref = None if (ref is None or ref.numel() == 0) else ref.contiguous().to(x.dtype)
    y = torch.empty_like(x)
    d = _lib.OiFusedBiasActDesc()
    d.x, d.bias, d.ref, d.y = x.data_ptr(), _lib.ptr(bias), _lib.ptr(ref), y.data_ptr()
    d.dtype, d.act, d.grad = _lib.DTYPE_CODE[x.dtype], act, grad
    d.size_x = x.numel()
    d.size_b = bias.numel() if bias is not None else 0
    step = 1
    for i in range(2, x.dim()):
        step *= x.shape[i]
    d.step_b = step
    d.alpha, d.scale = alpha, scale
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().oi_fused_bias_act(C.byref(d), _lib.current_stream_ptr(x.device)), "oi_fused_bias_act")
    return y


class _FusedLeakyReLUBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = fused_bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        dims = [0] + list(range(2, grad_input.ndim))
        grad_bias = grad_input.sum(dims).detach() if has_bias else grad_input.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        (out,) = ctx.saved_tensors
        gradgrad_out = fused_bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None


class _FusedLeakyReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.has_bias = bias is not None
        out = fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        grad_input, grad_bias = _FusedLeakyReLUBackward.apply(grad_output, out, ctx.has_bias, ctx.negative_slope,
                                                              ctx.scale)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """y = leaky_relu(input + bias[None, :, None...], negative_slope) * scale  (fused_act.py:104-119)."""
    return _FusedLeakyReLU.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope, self.scale = negative_slope, scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
