"""The geometric-path oracle (oracle/augment_oracle.py) against outputs of the unmodified reference AugmentPipe
(tests/golden/augment_golden.npz, oracle/gen_golden_augment.py): same seed -> same transform -> same image."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import augment_oracle as AO

CASES = ["train_rgb", "train_mask", "blit", "general", "all_geom", "nonsquare", "single"]


def load(name):
    with np.load(os.path.join(GOLDEN, "augment_golden.npz")) as f:
        meta = json.loads(str(f[f"{name}/meta"]))
        return meta, torch.from_numpy(f[f"{name}/x"]), torch.from_numpy(f[f"{name}/y"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    meta, x, y_ref = load(name)
    B, C, H, W = x.shape
    torch.manual_seed(meta["seed"])
    G = AO.sample_inverse_transform(B, W, H, **meta["kwargs"])
    y = AO.geometric_path(x, G)
    assert y.shape == y_ref.shape
    err = float((y - y_ref).abs().max())
    assert err < 2e-6, (name, err)     # same ops, same order; only the separable-filter pass order differs


def test_identity_transform_is_near_identity():
    # G_inv = I: pad + up-sample + resample at the original sample positions + down-sample ~ the input
    # (sym6 is an orthogonal low-pass: the residual is its stop-band leakage, not zero)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    x = torch.nn.functional.avg_pool2d(torch.nn.functional.interpolate(x, scale_factor=4, mode="bilinear"), 4)
    y = AO.geometric_path(x, torch.eye(3).repeat(2, 1, 1))
    assert float((y - x).abs().max()) < 0.2 * float(x.abs().max())


@pytest.mark.parametrize("name", CASES)
def test_product_sampling_and_setup_match_oracle(name):
    """Host logic of the drop-in (object_intrinsics_b200/augment.py): same random draws -> same G_inv as the oracle,
    and the device-side margins / affine_grid matrices equal the oracle's (computed there with host integers)."""
    from object_intrinsics_b200.augment import AugmentPipe, geometric_setup
    meta, x, _ = load(name)
    B, C, H, W = x.shape
    torch.manual_seed(meta["seed"])
    G_ref = AO.sample_inverse_transform(B, W, H, **meta["kwargs"])
    torch.manual_seed(meta["seed"])
    G = AugmentPipe(**meta["kwargs"]).sample_inverse_transform(B, W, H, torch.device("cpu"))
    assert torch.equal(G, G_ref)
    theta, margins = geometric_setup(G, H, W, 3)
    assert [int(v) for v in margins] == AO.margins(G_ref, W, H, 3)
    # the oracle's final matrix, rebuilt step by step with python integers (augment.py:287-297)
    mx0, my0, mx1, my1 = AO.margins(G_ref, W, H, 3)
    like = torch.ones(B)
    Gm = AO.translate2d((mx0 - mx1) / 2, (my0 - my1) / 2, like) @ G_ref
    Gm = AO.scale2d(2, 2, like) @ Gm @ AO.scale2d(1 / 2, 1 / 2, like)
    Gm = AO.translate2d(-0.5, -0.5, like) @ Gm @ AO.translate2d(0.5, 0.5, like)
    Wu, Hu, Wr, Hr = 2 * (W + mx0 + mx1), 2 * (H + my0 + my1), 2 * (W + 6), 2 * (H + 6)
    Gm = AO.scale2d(2 / Wu, 2 / Hu, like) @ Gm @ AO.scale2d(1 / (2 / Wr), 1 / (2 / Hr), like)
    assert float((theta - Gm[:, :2, :]).abs().max()) < 1e-6


def test_unsupported_options_raise():
    from object_intrinsics_b200.augment import AugmentPipe
    with pytest.raises(NotImplementedError):
        AugmentPipe(brightness=1)
    with pytest.raises(RuntimeError):
        AugmentPipe(xint=1)(torch.zeros(1, 3, 8, 8))      # no CPU path
