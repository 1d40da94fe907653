"""The geometric-path oracle (oracle/augment_oracle.py) against outputs of the unmodified reference AugmentPipe
(tests/golden/augment_golden.npz, oracle/gen_golden_augment.py): same seed -> same transform -> same image."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import augment_oracle as AO

CASES = ["train_rgb", "train_mask", "blit", "general", "all_geom"]


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
