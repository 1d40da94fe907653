"""Operand formats of the tensor-core backward (flags bit 5 = TF32, bit 6 = scaled fp16, default = chosen on the device):
gradients of every parameter vs the exact-fp32 FFMA backward (rel. L-inf per tensor, worst reported) and kernel time.
usage (GPU box): python tools_bwd_modes.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_inputs as BI  # noqa: E402
from object_intrinsics_b200 import fields  # noqa: E402
from object_intrinsics_b200.renderer import NeuSRenderer  # noqa: E402

P = BI.load_flat_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
named = [("sdf." + n, p) for n, p in sdf.named_parameters()] + [("col." + n, p) for n, p in col.named_parameters()] + \
        [("dev." + n, p) for n, p in dev.named_parameters()]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(r, ro, rd, near, far, z, loss_scale, outlier):
    for _, p in named:
        p.grad = None
    out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=sdf.style(z), perturb_overwrite=0)
    img = out["color_fine"] + (1.0 - out["weight_sum"])
    loss = (img ** 2).mean() + 0.1 * out["gradient_error"]
    if outlier:   # one ray with a huge upstream adjoint: heavy-tailed adjoint statistics
        loss = loss + outlier * out["color_fine"][5].sum()
    (loss * loss_scale).backward()
    return {n: p.grad.detach().clone() for n, p in named if p.grad is not None}


for name, bs, patch, n, m in [("cfg2 bs=1", 1, 64, 64, 0), ("cfg2 bs=2", 2, 64, 64, 0), ("32x32 16+4", 1, 32, 16, 4)]:
    ro, rd, near, far = [t.cuda() for t in BI.synthetic_rays(bs, patch, seed=1)]
    z = BI.latent(bs, 1).cuda()
    for loss_scale, outlier in [(1.0, 0.0), (1e-6, 0.0), (1e4, 0.0), (1.0, 1e9)]:
        ref_r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=0,
                             grad_impl="cuda")
        ref_r.bwd_impl = "ffma"
        ref = run(ref_r, ro, rd, near, far, z, loss_scale, outlier)
        row = {"workload": name, "loss_scale": loss_scale, "outlier": outlier}
        for mode, bits in [("auto", 0), ("tf32", 32), ("f16", 64)]:
            r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=0)
            r.flags |= bits
            r.bwd_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            for e in r.bwd_events:
                e.record()
            g = run(r, ro, rd, near, far, z, loss_scale, outlier)
            worst = max((float((g[k] - ref[k]).abs().max() / (ref[k].abs().max() + 1e-30)), k) for k in ref)
            ks = []
            for i in range(5):
                flush.fill_(i)
                run(r, ro, rd, near, far, z, loss_scale, outlier)
                torch.cuda.synchronize()
                ks.append(r.bwd_events[0].elapsed_time(r.bwd_events[1]))
            row[mode] = {"worst_rel_err": worst[0], "tensor": worst[1], "bwd_kernels_ms": sorted(ks)[2]}
        print(json.dumps(row), flush=True)
