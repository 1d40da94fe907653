"""Range of the fp16 operands of the tensor-core backward (library built with -DOI_BWD_RANGE_STATS=1):
largest forward-type (scaled by 2^(e_m - e_ref)) and adjoint-type (normalised by the point's adjoint exponent) value
written, against fp16's 65504.  usage (GPU box): OI_LIB_PATH=.../variants/rstats/liboi_b200.so python tools_bwd_range.py"""
import os
import struct
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_inputs as BI  # noqa: E402
from object_intrinsics_b200 import fields  # noqa: E402
from object_intrinsics_b200.renderer import NeuSRenderer  # noqa: E402


def f(bits):
    return struct.unpack("f", struct.pack("I", bits))[0]


def probe(tag, r, ro, rd, near, far, w, z, loss_fn):
    for p in list(r.sdf_network.parameters()) + list(r.color_network.parameters()):
        p.grad = None
    out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w, perturb_overwrite=0)
    loss_fn(out).backward()
    cw = r.last_backward_control_words()
    print(f"{tag:34s} format {r.last_backward_operand_format()}  max|adj| {f(cw[1]):.3e}  low/total mass "
          f"{(cw[4] | cw[5] << 32) / max(1, cw[2] | cw[3] << 32):.2e}  max forward-type operand {f(cw[6]):.4g}  "
          f"max adjoint-type operand {f(cw[7]):.4g}")


P = BI.load_flat_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
losses = {"image + eikonal": lambda o: ((o["color_fine"] + 1 - o["weight_sum"]) ** 2).mean() + 0.1 * o["gradient_error"],
          "sum of every output": lambda o: sum(o[k].sum() for k in ("color_fine", "weight_sum", "weights", "gradients", "sdf",
                                                                  "raw_color", "cdf_fine", "weight_max")) + o["gradient_error"] + o["surface_loss"]}
for name, bs, patch, n, m in [("cfg2 bs=1", 1, 64, 64, 0), ("cfg2 bs=4", 4, 64, 64, 0), ("128x128 16+4", 1, 128, 16, 4)]:
    ro, rd, near, far = [t.cuda() for t in BI.synthetic_rays(bs, patch, seed=1)]
    z = BI.latent(bs, 1).cuda()
    r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=0)
    for ln, lf in losses.items():
        probe(f"{name}, {ln}", r, ro, rd, near, far, sdf.style(z), z, lf)
# random-init networks with a wide FiLM range
torch.manual_seed(0)
sdf2, col2, dev2 = fields.build_networks(D=8, device="cuda")
ro, rd, near, far = [t.cuda() for t in BI.synthetic_rays(1, 64, seed=2)]
z = BI.latent(1, 3).cuda()
r = NeuSRenderer(None, sdf2, dev2, col2, n_samples=64, n_importance=0, n_outside=0, up_sample_steps=1, perturb=0)
for ln, lf in losses.items():
    probe(f"random init, {ln}", r, ro, rd, near, far, sdf2.style(z), z, lf)
