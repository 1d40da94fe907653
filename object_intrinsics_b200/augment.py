"""Drop-in for the reference's `AugmentPipe` (src/third_party/ada/augment.py:113-421) for the options its configs
enable (configs/train.yaml:80-100: `scale`, `xint`; all pixel-blitting / general geometric options are supported).

    model.discriminator.kwargs.aug.__target__=object_intrinsics_b200.augment.AugmentPipe

Same constructor keywords, the `p` buffer (overall probability multiplier, updated by ADA controllers), the same
order of random draws (augment.py:196-264), and `forward(images)` returns the same image as the reference under the
same torch seed (tests/test_augment_gpu.py).  The geometric execution (augment.py:270-301: reflect-pad, 2x
up-sample, affine resample, 2x down-sample) runs as two CUDA kernels through `oi_augment_geom_forward` with the
padding margins kept on the device -- the reference's `margin.ceil().to(torch.int32)` + `F.pad(pad=[...])`
(augment.py:283,286) is a device->host sync on every discriminator forward.  The map is linear in the images:
backward = `oi_augment_geom_backward` (the adjoint), double backward (R1, src/loss/gan.py:5-14) = the forward.
Colour / image-space filtering / noise / cutout (augment.py:303-421) are disabled in the reference's configs and
raise NotImplementedError when enabled.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib

SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148]   # augment.py:24 wavelets['sym6']

SYM2 = [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025]   # wavelets['sym2']


def image_filter_bank(lowpass=SYM2, bands=4):
    """The `Hz_fbank` buffer of the reference pipe (augment.py:169-179): row i is the zero-phase band-pass of
    octave i built from the sym2 low-pass H(z).  With L = H(z)H(1/z)/2 and B = H(-z)H(-1/z)/2, every step
    stretches the bank by 2 (zero stuffing), smooths it with L and drops B into the centre of the next row.
    The image-space filtering that consumes it is disabled in the reference's configs (and raises here); the
    buffer is registered only so that `state_dict()` keys / shapes match the reference's checkpoints."""
    import numpy as np
    h = np.asarray(lowpass, dtype=np.float64)
    g = h * np.where(np.arange(h.size) % 2 == 0, 1.0, -1.0)
    low2 = np.convolve(h, h[::-1]) * 0.5
    high2 = np.convolve(g, g[::-1]) * 0.5
    bank = np.zeros((bands, 1))
    bank[0, 0] = 1.0
    for i in range(1, bands):
        stretched = np.zeros((bands, 2 * bank.shape[1] - 1))
        stretched[:, ::2] = bank
        bank = np.stack([np.convolve(row, low2) for row in stretched])
        c0 = (bank.shape[1] - high2.size) // 2
        bank[i, c0:c0 + high2.size] += high2
    return torch.as_tensor(bank, dtype=torch.float32)


_WS = {}    # one scratch buffer per device, reused by every call (calls on one stream, as in the reference's trainer)


def _workspace(dev, nbytes):
    t = _WS.get(dev)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WS[dev] = t
    return t


def _geom_call(x, theta, margins, taps, backward):
    if not x.is_cuda:
        raise RuntimeError("object_intrinsics_b200.augment has no CPU path: images must be CUDA tensors")
    L = _lib.lib()
    x = x.to(torch.float32).contiguous()
    y = torch.empty_like(x)
    d = _lib.OiAugmentGeomDesc()
    d.batch, d.channels, d.height, d.width = x.shape
    d.filter_taps = len(taps)
    farr = (C.c_float * len(taps))(*taps)
    d.filter = C.cast(farr, C.POINTER(C.c_float))
    d.theta, d.margins, d.x, d.y = theta.data_ptr(), margins.data_ptr(), x.data_ptr(), y.data_ptr()
    nbytes = C.c_size_t(0)
    with torch.cuda.device(x.device):
        _lib.check(L.oi_augment_geom_workspace_bytes(C.byref(d), C.byref(nbytes)), "oi_augment_geom_workspace_bytes")
        ws = _workspace(x.device, nbytes.value)
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        fn = L.oi_augment_geom_backward if backward else L.oi_augment_geom_forward
        _lib.check(fn(C.byref(d), _lib.current_stream_ptr(x.device)), "oi_augment_geom")
    return y


class _GeomForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, margins, taps):
        ctx.save_for_backward(theta, margins)
        ctx.taps = taps
        return _geom_call(x, theta, margins, taps, False)

    @staticmethod
    def backward(ctx, gy):
        theta, margins = ctx.saved_tensors
        return _GeomAdjoint.apply(gy, theta, margins, ctx.taps), None, None, None


class _GeomAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, theta, margins, taps):
        ctx.save_for_backward(theta, margins)
        ctx.taps = taps
        return _geom_call(gy, theta, margins, taps, True)

    @staticmethod
    def backward(ctx, ggx):
        theta, margins = ctx.saved_tensors
        return _GeomForward.apply(ggx, theta, margins, ctx.taps), None, None, None


def _mat(rows, like):
    elems = [x if isinstance(x, torch.Tensor) else torch.full_like(like, float(x)) for row in rows for x in row]
    return torch.stack(elems, dim=-1).reshape(like.shape + (3, 3))


def _translate(tx, ty, like):
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], like)


def _scale(sx, sy, like):
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], like)


def _rotate(theta):
    return _mat([[torch.cos(theta), torch.sin(-theta), 0], [torch.sin(theta), torch.cos(theta), 0], [0, 0, 1]], theta)


def geometric_setup(G_inv, height, width, hz_pad):
    """Padding margins [4] int32 (augment.py:274-283) and affine_grid matrices theta [B,2,3] (augment.py:287-297)
    from the inverse transform, as device tensors: a handful of [B,3,3] torch ops, no host read-back."""
    H, W = height, width
    dev, dt = G_inv.device, torch.float32
    G_inv = G_inv.to(dt)
    B = G_inv.shape[0]
    like = torch.ones([B], device=dev, dtype=dt)
    cx, cy = (W - 1) / 2, (H - 1) / 2
    cp = torch.tensor([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]], device=dev, dtype=dt)
    cp = G_inv @ cp.t()
    m = cp[:, :2, :].permute(1, 0, 2).flatten(1)
    m = torch.cat([-m, m]).max(dim=1).values
    m = m + torch.tensor([hz_pad * 2 - cx, hz_pad * 2 - cy] * 2, device=dev, dtype=dt)
    m = m.max(torch.zeros(4, device=dev, dtype=dt)).min(torch.tensor([W - 1, H - 1] * 2, device=dev, dtype=dt))
    margins = m.ceil().to(torch.int32)                      # [mx0, my0, mx1, my1], stays on the device
    mf = margins.to(dt)
    Wp, Hp = W + mf[0] + mf[2], H + mf[1] + mf[3]           # padded size, device scalars
    G = _translate((mf[0] - mf[2]) / 2 * like, (mf[1] - mf[3]) / 2 * like, like) @ G_inv
    G = _scale(2, 2, like) @ G @ _scale(1 / 2, 1 / 2, like)
    G = _translate(-0.5, -0.5, like) @ G @ _translate(0.5, 0.5, like)
    Wr, Hr = (W + hz_pad * 2) * 2, (H + hz_pad * 2) * 2
    G = _scale(2 / (2 * Wp) * like, 2 / (2 * Hp) * like, like) @ G @ _scale(1 / (2 / Wr), 1 / (2 / Hr), like)
    return G[:, :2, :].contiguous(), margins


def geometric_setup_cuda(G_inv, height, width, n_taps):
    """`geometric_setup` as one tiny kernel (oi_augment_geom_setup) instead of ~40 torch launches."""
    L = _lib.lib()
    G_inv = G_inv.detach().to(torch.float32).contiguous()
    B = G_inv.shape[0]
    theta = torch.empty((B, 2, 3), dtype=torch.float32, device=G_inv.device)
    margins = torch.empty(4, dtype=torch.int32, device=G_inv.device)
    with torch.cuda.device(G_inv.device):
        _lib.check(L.oi_augment_geom_setup(G_inv.data_ptr(), B, height, width, n_taps, theta.data_ptr(),
                                           margins.data_ptr(), _lib.current_stream_ptr(G_inv.device)),
                   "oi_augment_geom_setup")
    return theta, margins


def geometric_setup_ops_cuda(ops, batch, height, width, n_taps, device):
    """`geometric_setup` for a transform given as its elementary factors (`AugmentPipe.sample_ops`): ONE kernel
    composes G_inv = prod M_i and derives margins + affine matrices (oi_augment_geom_setup_ops) instead of the ~15
    torch launches per factor of the reference's matrix helpers.  Returns (theta, margins, G_inv)."""
    L = _lib.lib()
    arr = (_lib.OiAugmentOp * len(ops))()
    keep = []
    for i, (kind, p0, p1) in enumerate(ops):
        p0 = p0.detach().to(device=device, dtype=torch.float32).contiguous()
        keep.append(p0)
        arr[i].kind, arr[i].p0 = kind, p0.data_ptr()
        if p1 is not None:
            p1 = p1.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(p1)
            arr[i].p1 = p1.data_ptr()
    theta = torch.empty((batch, 2, 3), dtype=torch.float32, device=device)
    g_inv = torch.empty((batch, 3, 3), dtype=torch.float32, device=device)
    margins = torch.empty(4, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(L.oi_augment_geom_setup_ops(arr, len(ops), batch, height, width, n_taps, g_inv.data_ptr(),
                                               theta.data_ptr(), margins.data_ptr(), _lib.current_stream_ptr(device)),
                   "oi_augment_geom_setup_ops")
    return theta, margins, g_inv


def geometric_setup_raw_cuda(raw, p, batch, height, width, n_taps, device):
    """`geometric_setup` from the RAW random draws of `AugmentPipe.sample_raw`: gating, parameter arithmetic,
    composition of G_inv, margins and affine matrices in ONE kernel (oi_augment_geom_setup_raw).  Returns
    (theta, margins, G_inv)."""
    L = _lib.lib()
    arr = (_lib.OiAugmentRawOp * len(raw))()
    for i, (form, draw, gate, prob, param) in enumerate(raw):
        arr[i].form, arr[i].draw, arr[i].gate, arr[i].prob, arr[i].param = form, draw.data_ptr(), gate.data_ptr(), prob, param
    out = torch.empty(batch * 15 + 4, dtype=torch.float32, device=device)    # theta [B,2,3] | G_inv [B,3,3] | margins
    theta, g_inv = out[:batch * 6].view(batch, 2, 3), out[batch * 6:batch * 15].view(batch, 3, 3)
    margins = out[batch * 15:].view(torch.int32)
    with torch.cuda.device(device):
        _lib.check(L.oi_augment_geom_setup_raw(arr, len(raw), p.data_ptr(), batch, height, width, n_taps,
                                               g_inv.data_ptr(), theta.data_ptr(), margins.data_ptr(),
                                               _lib.current_stream_ptr(device)), "oi_augment_geom_setup_raw")
    return theta, margins, g_inv


def geometric_transform(images, G_inv, taps=None):
    """augment.py:270-301 for a given inverse transform G_inv [B,3,3] (pixel_out -> pixel_in)."""
    if not images.is_cuda:
        raise RuntimeError("object_intrinsics_b200.augment has no CPU path: images must be CUDA tensors")
    taps = tuple(taps) if taps is not None else tuple(t / sum(SYM6) for t in SYM6)
    theta, margins = geometric_setup_cuda(G_inv.to(images.device), images.shape[2], images.shape[3], len(taps))
    return _GeomForward.apply(images, theta, margins, taps)


class AugmentPipe(torch.nn.Module):
    """Constructor keywords of the reference class (augment.py:114-121)."""

    def __init__(self, xflip=0, rotate90=0, xint=0, xint_max=0.125, scale=0, rotate=0, aniso=0, xfrac=0,
                 scale_std=0.2, rotate_max=1, aniso_std=0.2, xfrac_std=0.125, brightness=0, contrast=0, lumaflip=0,
                 hue=0, saturation=0, brightness_std=0.2, contrast_std=0.5, hue_max=1, saturation_std=1, imgfilter=0,
                 imgfilter_bands=(1, 1, 1, 1), imgfilter_std=1, noise=0, cutout=0, noise_std=0.1, cutout_size=0.5):
        super().__init__()
        if any(float(v) > 0 for v in (brightness, contrast, lumaflip, hue, saturation, imgfilter, noise, cutout)):
            raise NotImplementedError("colour / filtering / noise / cutout augmentations (augment.py:303-421) are "
                                      "disabled in the reference's configs and are not implemented")
        self.register_buffer("p", torch.ones([]))
        self.xflip, self.rotate90, self.xint, self.xint_max = float(xflip), float(rotate90), float(xint), float(xint_max)
        self.scale, self.rotate, self.aniso, self.xfrac = float(scale), float(rotate), float(aniso), float(xfrac)
        self.scale_std, self.rotate_max = float(scale_std), float(rotate_max)
        self.aniso_std, self.xfrac_std = float(aniso_std), float(xfrac_std)
        f = torch.tensor(SYM6, dtype=torch.float32)
        self.register_buffer("Hz_geom", f / f.sum())          # upfirdn2d.setup_filter(wavelets['sym6']), augment.py:116
        self._taps = tuple(float(v) for v in (f / f.sum()))
        self.register_buffer("Hz_fbank", image_filter_bank())  # augment.py:179 (unused: imgfilter is disabled)

    def sample_ops(self, batch, width, height, device):
        """The elementary factors of G_inv (pixel_out -> pixel_in) as a list of (kind, p0, p1) with per-sample fp32
        parameters [B]: kind 0 = scale2d(p0, p1), 1 = rotate2d(p0), 2 = translate2d(p0, p1); random draws in the
        reference's order (augment.py:196-264) -- the only torch work left per call."""
        ones = torch.ones([batch], device=device)
        p = self.p
        ops = []
        if self.xflip > 0:
            i = torch.floor(torch.rand([batch], device=device) * 2)
            i = torch.where(torch.rand([batch], device=device) < self.xflip * p, i, torch.zeros_like(i))
            ops.append((0, 1 / (1 - 2 * i), 1 / ones))
        if self.rotate90 > 0:
            i = torch.floor(torch.rand([batch], device=device) * 4)
            i = torch.where(torch.rand([batch], device=device) < self.rotate90 * p, i, torch.zeros_like(i))
            ops.append((1, -(-math.pi / 2 * i), None))
        if self.xint > 0:
            t = (torch.rand([batch, 2], device=device) * 2 - 1) * self.xint_max
            t = torch.where(torch.rand([batch, 1], device=device) < self.xint * p, t, torch.zeros_like(t))
            ops.append((2, -torch.round(t[:, 0] * width), -torch.round(t[:, 1] * height)))
        if self.scale > 0:
            s = torch.exp2(torch.randn([batch], device=device) * self.scale_std)
            s = torch.where(torch.rand([batch], device=device) < self.scale * p, s, torch.ones_like(s))
            ops.append((0, 1 / s, 1 / s))
        p_rot = 1 - torch.sqrt((1 - self.rotate * p).clamp(0, 1))
        if self.rotate > 0:
            th = (torch.rand([batch], device=device) * 2 - 1) * math.pi * self.rotate_max
            th = torch.where(torch.rand([batch], device=device) < p_rot, th, torch.zeros_like(th))
            ops.append((1, -(-th), None))
        if self.aniso > 0:
            s = torch.exp2(torch.randn([batch], device=device) * self.aniso_std)
            s = torch.where(torch.rand([batch], device=device) < self.aniso * p, s, torch.ones_like(s))
            ops.append((0, 1 / s, 1 / (1 / s)))
        if self.rotate > 0:
            th = (torch.rand([batch], device=device) * 2 - 1) * math.pi * self.rotate_max
            th = torch.where(torch.rand([batch], device=device) < p_rot, th, torch.zeros_like(th))
            ops.append((1, -(-th), None))
        if self.xfrac > 0:
            t = torch.randn([batch, 2], device=device) * self.xfrac_std
            t = torch.where(torch.rand([batch, 1], device=device) < self.xfrac * p, t, torch.zeros_like(t))
            ops.append((2, -(t[:, 0] * width), -(t[:, 1] * height)))
        return ops

    def sample_raw(self, batch, device):
        """The random draws of `sample_ops` and nothing else -- same torch calls in the same order, so the RNG stream
        is the reference's -- as a list of (form, draw, gate, prob, param) for oi_augment_geom_setup_raw, which applies
        the gating and the parameter arithmetic of augment.py:196-264 itself."""
        raw = []
        rand = lambda *shape: torch.rand(list(shape), device=device)
        randn = lambda *shape: torch.randn(list(shape), device=device)
        if self.xflip > 0:
            raw.append((_lib.AUG_XFLIP, rand(batch), rand(batch), self.xflip, 0.0))
        if self.rotate90 > 0:
            raw.append((_lib.AUG_ROTATE90, rand(batch), rand(batch), self.rotate90, 0.0))
        if self.xint > 0:
            raw.append((_lib.AUG_XINT, rand(batch, 2), rand(batch, 1), self.xint, self.xint_max))
        if self.scale > 0:
            raw.append((_lib.AUG_SCALE, randn(batch), rand(batch), self.scale, self.scale_std))
        if self.rotate > 0:
            raw.append((_lib.AUG_ROTATE, rand(batch), rand(batch), self.rotate, self.rotate_max))
        if self.aniso > 0:
            raw.append((_lib.AUG_ANISO, randn(batch), rand(batch), self.aniso, self.aniso_std))
        if self.rotate > 0:
            raw.append((_lib.AUG_ROTATE, rand(batch), rand(batch), self.rotate, self.rotate_max))
        if self.xfrac > 0:
            raw.append((_lib.AUG_XFRAC, randn(batch, 2), rand(batch, 1), self.xfrac, self.xfrac_std))
        return raw

    def sample_inverse_transform(self, batch, width, height, device):
        """G_inv [B,3,3] composed with torch ops exactly as the reference does (augment.py:196-264); the CPU tests pin
        it to the reference bit for bit.  `forward` composes the same factors in one kernel instead."""
        ops = self.sample_ops(batch, width, height, device)
        if not ops:
            return None
        ones = torch.ones([batch], device=device)
        G = torch.eye(3, device=device)
        for kind, p0, p1 in ops:
            G = G @ (_scale(p0, p1, ones) if kind == 0 else _rotate(p0) if kind == 1 else _translate(p0, p1, ones))
        return G

    def forward(self, images, debug_percentile=None):
        if debug_percentile is not None:
            raise NotImplementedError("debug_percentile is a debugging aid of the reference and is not implemented")
        assert isinstance(images, torch.Tensor) and images.ndim == 4
        B, _, H, W = images.shape
        if not images.is_cuda:
            raise RuntimeError("object_intrinsics_b200.augment has no CPU path: images must be CUDA tensors")
        raw = self.sample_raw(B, images.device)
        if not raw:
            return images
        p = self.p.detach().to(device=images.device, dtype=torch.float32)     # no-op when the module lives with the images
        theta, margins, _ = geometric_setup_raw_cuda(raw, p, B, H, W, len(self._taps), images.device)
        return _GeomForward.apply(images, theta, margins, self._taps)
