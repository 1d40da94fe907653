"""`upfirdn2d` with the reference's Python surface (src/third_party/ada/torch_utils/ops/upfirdn2d.py:72-384)
on top of the C-ABI `oi_upfirdn2d` (replaces `_plugin.upfirdn2d`, ada/.../upfirdn2d.cpp:16-94, and the stylesdf
`upfirdn2d_op.upfirdn2d`, stylesdf/op/upfirdn2d.cpp:12-22).

Differences from the reference wrapper, on purpose: CUDA tensors only (no `_upfirdn2d_ref` fallback,
upfirdn2d.py:164-208), and a separable filter is applied as two launches exactly like the reference's CUDA
path (upfirdn2d.py:233-240).  Gradients of any order come from the same op with up/down swapped
(upfirdn2d.py:245-262).
"""
import ctypes as C
import math

import torch

from .. import _lib


def _pair(v):
    if isinstance(v, int):
        return v, v
    assert isinstance(v, (list, tuple)) and len(v) == 2 and all(isinstance(a, int) for a in v)
    return int(v[0]), int(v[1])


def _parse_scaling(scaling):
    sx, sy = _pair(scaling)
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(a, int) for a in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device("cpu"), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR setup helper (upfirdn2d.py:72-116): 1-D taps of length >= 8 stay separable by default."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in (0, 1, 2) and f.numel() > 0
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def upfirdn2d_raw(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    """Same argument meaning as `_plugin.upfirdn2d` (ada/.../upfirdn2d.cpp:16): x [N,C,H,W] (any strides),
    f [fh,fw] fp32."""
    if not x.is_cuda:
        raise RuntimeError("upfirdn2d: CUDA tensors only (object_intrinsics_b200 has no CPU path)")
    if x.dtype not in _lib.DTYPE_CODE:
        raise TypeError(f"upfirdn2d: unsupported dtype {x.dtype}")
    if x.dim() != 4:
        raise RuntimeError("x must be rank 4")
    if f.dim() != 2 or f.dtype != torch.float32 or f.device != x.device:
        raise RuntimeError("f must be a rank-2 float32 tensor on the device of x")
    if f.shape[0] < 1 or f.shape[1] < 1:
        raise RuntimeError("f must be at least 1x1")
    if upx < 1 or upy < 1 or downx < 1 or downy < 1:
        raise RuntimeError("up/down sampling factors must be at least 1")
    N, Cn, H, W = x.shape
    out_w = (W * upx + padx0 + padx1 - f.shape[1] + downx) // downx
    out_h = (H * upy + pady0 + pady1 - f.shape[0] + downy) // downy
    if out_w < 1 or out_h < 1:
        raise RuntimeError("output must be at least 1x1")
    channels_last = x.dim() == 4 and x.stride(1) == 1 and Cn > 1
    y = torch.empty((N, Cn, out_h, out_w), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    d = _lib.OiUpfirdnDesc()
    d.x, d.f, d.y, d.dtype = x.data_ptr(), f.data_ptr(), y.data_ptr(), _lib.DTYPE_CODE[x.dtype]
    d.batch, d.channels, d.in_h, d.in_w = N, Cn, H, W
    d.x_stride_n, d.x_stride_c, d.x_stride_h, d.x_stride_w = x.stride()
    d.out_h, d.out_w = out_h, out_w
    d.y_stride_n, d.y_stride_c, d.y_stride_h, d.y_stride_w = y.stride()
    d.filter_h, d.filter_w = f.shape
    d.f_stride_h, d.f_stride_w = f.stride()
    d.up_x, d.up_y, d.down_x, d.down_y = upx, upy, downx, downy
    d.pad_x0, d.pad_x1, d.pad_y0, d.pad_y1 = padx0, padx1, pady0, pady1
    d.flip, d.gain = int(bool(flip)), float(gain)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().oi_upfirdn2d(C.byref(d), _lib.current_stream_ptr(x.device)), "oi_upfirdn2d")
    return y


_cache = {}


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _cache:
        return _cache[key]

    class Upfirdn2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
            if f.ndim == 2:
                y = upfirdn2d_raw(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:  # separable: a row pass then a column pass, sqrt(gain) each
                y = upfirdn2d_raw(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, math.sqrt(gain))
                y = upfirdn2d_raw(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, math.sqrt(gain))
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            (f,) = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1,
                 fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _upfirdn2d_cuda(up=down, down=up, padding=p, flip_filter=(not flip_filter), gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, None

    _cache[key] = Upfirdn2d
    return Upfirdn2d


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """Pad, upsample, filter and downsample a batch of 2-D images (upfirdn2d.py:120-160)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ("ref", "cuda")
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """Same-size FIR filtering (upfirdn2d.py:270-301)."""
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """FIR upsampling; output size is a multiple of the input (upfirdn2d.py:305-340)."""
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """FIR downsampling; output size is a fraction of the input (upfirdn2d.py:344-379)."""
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2,
         pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


# ---------------------------------------------------------------------------------------------------
# stylesdf flavour: `upfirdn2d(input[N,C,H,W], kernel[kh,kw], up, down, pad=(p0,p1))`
# (stylesdf/op/upfirdn2d.py:146-157): always convolution (kernel flipped), gain 1, same pad on x and y.
# ---------------------------------------------------------------------------------------------------
class _UpFirDn2dStyleSDF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, up, down, pad):
        ctx.cfg = (up, down, pad)
        ctx.in_size = x.shape
        ctx.save_for_backward(kernel)
        return upfirdn2d_raw(x, kernel, up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3], False, 1.0)

    @staticmethod
    def backward(ctx, gy):
        (kernel,) = ctx.saved_tensors
        (upx, upy), (downx, downy), (px0, px1, py0, py1) = ctx.cfg
        _, _, ih, iw = ctx.in_size
        _, _, oh, ow = gy.shape
        kh, kw = kernel.shape
        # adjoint: swap up/down, correlate (flip=True); paddings as in stylesdf/op/upfirdn2d.py:97-107
        g0x, g0y = kw - px0 - 1, kh - py0 - 1
        g1x = iw * upx - ow * downx + px0 - upx + 1
        g1y = ih * upy - oh * downy + py0 - upy + 1
        gx = _UpFirDn2dStyleSDFBackward.apply(gy, kernel, (downx, downy), (upx, upy), (g0x, g1x, g0y, g1y),
                                              ctx.cfg)
        return gx, None, None, None, None


class _UpFirDn2dStyleSDFBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, kernel, up, down, pad, fwd_cfg):
        ctx.fwd_cfg = fwd_cfg
        ctx.save_for_backward(kernel)
        return upfirdn2d_raw(gy, kernel, up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3], True, 1.0)

    @staticmethod
    def backward(ctx, ggx):
        (kernel,) = ctx.saved_tensors
        up, down, pad = ctx.fwd_cfg
        ggy = upfirdn2d_raw(ggx, kernel, up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3], False, 1.0)
        return ggy, None, None, None, None, None


def upfirdn2d_native_layout(input, kernel, up=1, down=1, pad=(0, 0)):
    """The stylesdf entry point `upfirdn2d(input, kernel, up, down, pad)` (stylesdf/op/upfirdn2d.py:146-157)."""
    return _UpFirDn2dStyleSDF.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
