"""TEST INFRASTRUCTURE -- generates tests/golden/augment_golden.npz from the UNMODIFIED reference AugmentPipe
(src/third_party/ada/augment.py) run on the CPU in the authoring container (needs /root/reference):

    python -m oracle.gen_golden_augment

Each case stores the constructor kwargs, the torch seed set immediately before the call, the input images and the
reference output.  tests/test_augment_oracle.py replays the seed through oracle/augment_oracle.py.
"""
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as RH  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, kwargs, batch, channels, size, seed
    ("train_rgb", dict(scale=1, xint=1), 4, 3, 32, 11),          # configs/train.yaml:80-85 (discriminator)
    ("train_mask", dict(scale=1, xint=1), 3, 1, 48, 12),         # configs/train.yaml:96-101 (mask discriminator)
    ("blit", dict(xflip=1, rotate90=1, xint=1), 4, 2, 24, 13),
    ("general", dict(scale=1, rotate=1, aniso=1, xfrac=1), 4, 3, 40, 14),
    ("all_geom", dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1), 5, 3, 32, 15),
    ("nonsquare", dict(xint=1, scale=1, rotate=1, xfrac=1), 3, 2, (24, 40), 16),   # H != W
    ("single", dict(scale=1, xint=1, xint_max=0.4), 1, 4, 16, 17),                # large integer shifts, B = 1
]


def main():
    RH.import_reference()

    class EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    sys.modules.setdefault("dnnlib", types.SimpleNamespace(EasyDict=EasyDict))
    from src.third_party.ada import augment as A
    blob = {}
    for name, kw, b, c, s, seed in CASES:
        pipe = A.AugmentPipe(**kw)
        hh, ww = (s, s) if isinstance(s, int) else s
        x = torch.randn(b, c, hh, ww, generator=torch.Generator().manual_seed(seed))
        torch.manual_seed(seed)
        y = pipe(x)
        blob[f"{name}/x"], blob[f"{name}/y"] = x.numpy(), y.numpy()
        blob[f"{name}/meta"] = np.array(json.dumps({"kwargs": kw, "seed": seed}))
        print(name, tuple(y.shape), float(y.abs().mean()))
    ref_sd = A.AugmentPipe(scale=1, xint=1).state_dict()       # buffers of the reference pipe (augment.py:123,167,179)
    for k, v in ref_sd.items():
        blob[f"state_dict/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "augment_golden.npz"), **blob)


if __name__ == "__main__":
    main()
