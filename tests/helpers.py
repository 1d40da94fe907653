"""Shared test helpers: golden-fixture loading (tests/golden/*.npz written by oracle/gen_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

OUT_KEYS = ['s_val', 'cdf_fine', 'weight_sum', 'weight_max', 'gradients', 'weights', 'gradient_error',
            'inside_sphere', 'mid_z_vals', 'surface_loss', 'sdf', 'pts_norm', 'pts', 'color_fine', 'raw_color']

CASES = ["cfg1_n16_m0", "cfg1_n16_m4", "cfg1_n16_m4_jit", "cfg2_n64_m0", "cfg2_n64_m0_jit", "cfg4_n64_m64",
         "cfgd_n16_m4_D8",
         "cfgs_n16_m8_s2", "cfgs_n16_m12_s4_D8"]   # the last two: up_sample_steps = 2 / 4 (renderer.py:400-413)


def load_params(fname, dtype=torch.float32):
    with np.load(os.path.join(GOLDEN, fname)) as f:
        return {k: torch.from_numpy(f[k]).to(dtype) for k in f.files}


def load_case(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        meta = json.loads(str(f["meta"]))
        inp = {k[3:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("in/")}
        ref32 = {k[6:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("ref32/")}
        ref64 = {k[6:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("ref64/")}
    return meta, inp, ref32, ref64


def linf(a, b):
    return float((a.double() - b.double()).abs().max())
