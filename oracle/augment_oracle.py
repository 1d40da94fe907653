"""CPU ORACLE for the geometric path of the ADA AugmentPipe -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, with plain torch ops, what `src/third_party/ada/augment.py:AugmentPipe.forward` does for the
pixel-blitting / general geometric options (lines 192-301; paths relative to /root/reference):

  * `sample_inverse_transform`: the per-image inverse homogeneous 2-D transform G_inv (pixel_out -> pixel_in),
    drawing the random numbers in the reference's order (augment.py:196-264) so that a run under the same torch
    seed reproduces the reference;
  * `geometric_path`: margin computation, reflect padding, 2x up-sampling with the sym6 low-pass, bilinear affine
    resampling (`affine_grid` + `grid_sample`, zeros outside), 2x down-sampling with crop (augment.py:270-301).

Pinned by tests/golden/augment_golden.npz: outputs of the UNMODIFIED reference `AugmentPipe(...)(images)` run on
the CPU in the authoring container (oracle/gen_golden_augment.py), compared in tests/test_augment_oracle.py.
The reference's colour / filtering / noise / cutout options (augment.py:303-421) are disabled in its configs
(configs/train.yaml:80-100 enables `scale` and `xint` only) and are not restated.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ops_oracle as OPS

SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148]   # augment.py:24 wavelets['sym6']


def hz_geom(dtype=torch.float32):
    """upfirdn2d.setup_filter(wavelets['sym6']) (augment.py:116): separable 12-tap filter normalised to unit DC."""
    f = torch.tensor(SYM6, dtype=torch.float32)
    return (f / f.sum()).to(dtype)


def _mat(rows, like):
    """3x3 matrices [B,3,3] from rows of scalars / [B] tensors (augment.py:49-58 `matrix`)."""
    elems = [x if isinstance(x, torch.Tensor) else torch.full_like(like, float(x)) for row in rows for x in row]
    return torch.stack(elems, dim=-1).reshape(like.shape + (3, 3))


def translate2d(tx, ty, like):
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], like)


def scale2d(sx, sy, like):
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], like)


def rotate2d(theta):
    return _mat([[torch.cos(theta), torch.sin(-theta), 0], [torch.sin(theta), torch.cos(theta), 0], [0, 0, 1]], theta)


def sample_inverse_transform(batch, width, height, *, p=1.0, xflip=0, rotate90=0, xint=0, xint_max=0.125, scale=0,
                             rotate=0, aniso=0, xfrac=0, scale_std=0.2, rotate_max=1, aniso_std=0.2, xfrac_std=0.125,
                             device="cpu"):
    """augment.py:192-264 with debug_percentile=None.  Returns G_inv [B,3,3] or None (identity: nothing enabled)."""
    G = None
    ones = torch.ones([batch], device=device)

    def mul(G, M):
        return M if G is None else G @ M

    if xflip > 0:
        i = torch.floor(torch.rand([batch], device=device) * 2)
        i = torch.where(torch.rand([batch], device=device) < xflip * p, i, torch.zeros_like(i))
        G = mul(torch.eye(3, device=device), scale2d(1 / (1 - 2 * i), 1 / ones, ones))
    if rotate90 > 0:
        i = torch.floor(torch.rand([batch], device=device) * 4)
        i = torch.where(torch.rand([batch], device=device) < rotate90 * p, i, torch.zeros_like(i))
        G = mul(torch.eye(3, device=device) if G is None else G, rotate2d(-(-math.pi / 2 * i)))
    if xint > 0:
        t = (torch.rand([batch, 2], device=device) * 2 - 1) * xint_max
        t = torch.where(torch.rand([batch, 1], device=device) < xint * p, t, torch.zeros_like(t))
        G = mul(torch.eye(3, device=device) if G is None else G,
                translate2d(-torch.round(t[:, 0] * width), -torch.round(t[:, 1] * height), ones))
    if scale > 0:
        s = torch.exp2(torch.randn([batch], device=device) * scale_std)
        s = torch.where(torch.rand([batch], device=device) < scale * p, s, torch.ones_like(s))
        G = mul(torch.eye(3, device=device) if G is None else G, scale2d(1 / s, 1 / s, ones))
    p_rot = 1 - math.sqrt(min(max(1 - rotate * p, 0.0), 1.0))
    if rotate > 0:
        theta = (torch.rand([batch], device=device) * 2 - 1) * math.pi * rotate_max
        theta = torch.where(torch.rand([batch], device=device) < p_rot, theta, torch.zeros_like(theta))
        G = mul(torch.eye(3, device=device) if G is None else G, rotate2d(-(-theta)))
    if aniso > 0:
        s = torch.exp2(torch.randn([batch], device=device) * aniso_std)
        s = torch.where(torch.rand([batch], device=device) < aniso * p, s, torch.ones_like(s))
        G = mul(torch.eye(3, device=device) if G is None else G, scale2d(1 / s, 1 / (1 / s), ones))
    if rotate > 0:
        theta = (torch.rand([batch], device=device) * 2 - 1) * math.pi * rotate_max
        theta = torch.where(torch.rand([batch], device=device) < p_rot, theta, torch.zeros_like(theta))
        G = mul(torch.eye(3, device=device) if G is None else G, rotate2d(-(-theta)))
    if xfrac > 0:
        t = torch.randn([batch, 2], device=device) * xfrac_std
        t = torch.where(torch.rand([batch, 1], device=device) < xfrac * p, t, torch.zeros_like(t))
        G = mul(torch.eye(3, device=device) if G is None else G, translate2d(-(t[:, 0] * width), -(t[:, 1] * height), ones))
    return G


def margins(G_inv, width, height, hz_pad=3):
    """augment.py:274-283: reflect-padding extents [mx0, my0, mx1, my1] (integers, shared by the whole batch)."""
    cx, cy = (width - 1) / 2, (height - 1) / 2
    cp = torch.tensor([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]], dtype=G_inv.dtype, device=G_inv.device)
    cp = G_inv @ cp.t()
    m = cp[:, :2, :].permute(1, 0, 2).flatten(1)
    m = torch.cat([-m, m]).max(dim=1).values
    m = m + torch.tensor([hz_pad * 2 - cx, hz_pad * 2 - cy] * 2, dtype=G_inv.dtype, device=G_inv.device)
    m = m.max(torch.zeros(4, dtype=G_inv.dtype, device=G_inv.device))
    m = m.min(torch.tensor([width - 1, height - 1] * 2, dtype=G_inv.dtype, device=G_inv.device))
    return [int(v) for v in m.ceil().to(torch.int32)]


def geometric_path(images, G_inv, hz=None):
    """augment.py:270-301.  images [B,C,H,W], G_inv [B,3,3] (pixel_out -> pixel_in, centred pixel coordinates)."""
    B, C, H, W = images.shape
    hz = hz_geom(images.dtype) if hz is None else hz
    hz_pad = hz.shape[0] // 4
    like = torch.ones([B], dtype=G_inv.dtype, device=G_inv.device)
    mx0, my0, mx1, my1 = margins(G_inv, W, H, hz_pad)
    x = F.pad(images, [mx0, mx1, my0, my1], mode="reflect")
    G = translate2d((mx0 - mx1) / 2, (my0 - my1) / 2, like) @ G_inv
    # up-sample (upfirdn2d.upsample2d, up = 2, gain = up^2)
    x = OPS.upfirdn2d_separable(x, hz, up=2, down=1, padding=OPS.upsample2d_padding(hz.shape[0], hz.shape[0], 2),
                                flip_filter=False, gain=4.0)
    G = scale2d(2, 2, like) @ G @ scale2d(1 / 2, 1 / 2, like)
    G = translate2d(-0.5, -0.5, like) @ G @ translate2d(0.5, 0.5, like)
    # resample
    shape = [B, C, (H + hz_pad * 2) * 2, (W + hz_pad * 2) * 2]
    G = scale2d(2 / x.shape[3], 2 / x.shape[2], like) @ G @ scale2d(1 / (2 / shape[3]), 1 / (2 / shape[2]), like)
    grid = F.affine_grid(theta=G[:, :2, :], size=shape, align_corners=False)
    x = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    # down-sample and crop (upfirdn2d.downsample2d, down = 2, padding = -2 hz_pad, flip_filter = True)
    pad = OPS.downsample2d_padding(hz.shape[0], hz.shape[0], 2, (-hz_pad * 2,) * 4)
    return OPS.upfirdn2d_separable(x, hz, up=1, down=2, padding=pad, flip_filter=True, gain=1.0)
