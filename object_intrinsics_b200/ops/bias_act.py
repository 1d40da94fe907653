"""`bias_act` with the reference's Python surface (src/third_party/ada/torch_utils/ops/bias_act.py:55-210) on
top of the C-ABI `oi_bias_act` (replaces `_plugin.bias_act`, bias_act.cpp:32-90).

The reference module cannot even be imported (it needs the un-vendored `dnnlib`, bias_act.py:15); the
activation table below restates bias_act.py:23-33.  First and second order gradients are supported exactly
like the reference (BiasActCuda / BiasActCudaGrad, bias_act.py:131-210).  CUDA tensors only.
"""
import ctypes as C
import math
from types import SimpleNamespace

import torch

from .. import _lib

activation_funcs = {
    "linear": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=1, ref="", has_2nd_grad=False),
    "relu": SimpleNamespace(def_alpha=0, def_gain=math.sqrt(2), cuda_idx=2, ref="y", has_2nd_grad=False),
    "lrelu": SimpleNamespace(def_alpha=0.2, def_gain=math.sqrt(2), cuda_idx=3, ref="y", has_2nd_grad=False),
    "tanh": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=4, ref="y", has_2nd_grad=True),
    "sigmoid": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=5, ref="y", has_2nd_grad=True),
    "elu": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=6, ref="y", has_2nd_grad=True),
    "selu": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=7, ref="y", has_2nd_grad=True),
    "softplus": SimpleNamespace(def_alpha=0, def_gain=1, cuda_idx=8, ref="y", has_2nd_grad=True),
    "swish": SimpleNamespace(def_alpha=0, def_gain=math.sqrt(2), cuda_idx=9, ref="x", has_2nd_grad=True),
}


def _dense_layout_ok(x, like):
    return x.shape == like.shape and x.stride() == like.stride()


def bias_act_raw(x, b, xref, yref, dy, grad: int, dim: int, act: int, alpha: float, gain: float, clamp: float):
    """Same argument meaning as the reference's `_plugin.bias_act` (bias_act.cpp:32); None/empty = absent."""
    if not x.is_cuda:
        raise RuntimeError("bias_act: CUDA tensors only (object_intrinsics_b200 has no CPU path)")
    if x.dtype not in _lib.DTYPE_CODE:
        raise TypeError(f"bias_act: unsupported dtype {x.dtype}")

    def opt(t):
        return None if (t is None or t.numel() == 0) else t

    b, xref, yref, dy = opt(b), opt(xref), opt(yref), opt(dy)
    dense = x.is_contiguous() or (x.ndim == 4 and x.is_contiguous(memory_format=torch.channels_last)) or \
        (x.ndim == 5 and x.is_contiguous(memory_format=torch.channels_last_3d))
    if not dense:
        raise RuntimeError("x must be non-overlapping and dense")
    for name, t in (("xref", xref), ("yref", yref), ("dy", dy)):
        if t is not None and not (_dense_layout_ok(t, x) and t.dtype == x.dtype and t.device == x.device):
            raise RuntimeError(f"{name} must have the same shape, dtype, layout and device as x")
    if b is not None:
        if b.dim() != 1 or not b.is_contiguous() or b.dtype != x.dtype or b.device != x.device:
            raise RuntimeError("b must be a contiguous rank-1 tensor with the dtype and device of x")
        if not (0 <= dim < x.dim()) or b.numel() != x.shape[dim]:
            raise RuntimeError("b has wrong number of elements / dim is out of bounds")
    if grad < 0:
        raise RuntimeError("grad must be non-negative")
    y = torch.empty_like(x)   # preserves the dense layout of x (bias_act.cpp:56-58)
    d = _lib.OiBiasActDesc()
    d.x, d.b, d.xref, d.yref, d.dy, d.y = x.data_ptr(), _lib.ptr(b), _lib.ptr(xref), _lib.ptr(yref), _lib.ptr(dy), \
        y.data_ptr()
    d.dtype, d.grad, d.act = _lib.DTYPE_CODE[x.dtype], grad, act
    d.alpha, d.gain, d.clamp = alpha, gain, clamp
    d.size_x = x.numel()
    d.size_b = b.numel() if b is not None else 0
    d.step_b = x.stride(dim) if b is not None else 1
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().oi_bias_act(C.byref(d), _lib.current_stream_ptr(x.device)), "oi_bias_act")
    return y


_cache = {}


def _bias_act_cuda(dim, act, alpha, gain, clamp):
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]

    class BiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.memory_format = torch.channels_last if x.ndim > 2 and x.stride()[1] == 1 else torch.contiguous_format
            x = x.contiguous(memory_format=ctx.memory_format)
            b = b.contiguous() if b is not None else None
            y = x
            if act != "linear" or gain != 1 or clamp >= 0 or b is not None:
                y = bias_act_raw(x, b, None, None, None, 0, dim, spec.cuda_idx, alpha, gain, clamp)
            empty = x.new_empty(0)
            ctx.save_for_backward(x if "x" in spec.ref or spec.has_2nd_grad else empty,
                                  b if (b is not None and ("x" in spec.ref or spec.has_2nd_grad)) else empty,
                                  y if "y" in spec.ref else empty)
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = dy.contiguous(memory_format=ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy
                if act != "linear" or gain != 1 or clamp >= 0:
                    dx = BiasActGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1]:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ctx.memory_format = torch.channels_last if dy.ndim > 2 and dy.stride()[1] == 1 else torch.contiguous_format
            dx = bias_act_raw(dy, b, x, y, None, 1, dim, spec.cuda_idx, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else dy.new_empty(0), x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = d_dx.contiguous(memory_format=ctx.memory_format)
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = bias_act_raw(d_dx, b, x, y, dy, 2, dim, spec.cuda_idx, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _cache[key] = BiasAct
    return BiasAct


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None, impl="cuda"):
    """y = clamp(act(x + b) * gain)  (bias_act.py:55-91).  `impl` is accepted for signature compatibility;
    there is only the CUDA implementation."""
    assert isinstance(x, torch.Tensor)
    assert impl in ("ref", "cuda")
    assert clamp is None or clamp >= 0
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)
