"""Timing of the grad-mode render step on the GPU box: forward ms, backward kernels ms (events inside the C-ABI call),
whole step ms.  usage: [OI_LIB_PATH=...] python tools_time_bwd.py [tag]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench_inputs as BI  # noqa: E402
from object_intrinsics_b200 import fields  # noqa: E402
from object_intrinsics_b200.renderer import NeuSRenderer  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
P = BI.load_flat_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
params = list(sdf.parameters()) + list(col.parameters()) + list(dev.parameters())
rows = []
for name, bs, patch, n, m in [("cfg2 bs=1", 1, 64, 64, 0), ("cfg2 bs=4", 4, 64, 64, 0), ("128x128 16+4", 1, 128, 16, 4)]:
    ro, rd, near, far = [t.cuda() for t in BI.synthetic_rays(bs, patch, seed=1)]
    z = BI.latent(bs, 1).cuda()
    r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=0)
    r.bwd_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for e in r.bwd_events:
        e.record()

    def gstep():
        for p in params:
            p.grad = None
        w = sdf.style(z)
        out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w)
        img = out["color_fine"] + (1.0 - out["weight_sum"])
        ((img ** 2).mean() + 0.1 * out["gradient_error"]).backward()
    for _ in range(3):
        gstep()
    torch.cuda.synchronize()
    ts, ks = [], []
    for i in range(7):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gstep()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
        ks.append(r.bwd_events[0].elapsed_time(r.bwd_events[1]))
    ts.sort()
    ks.sort()
    rows.append({"tag": tag, "workload": name, "step_ms": ts[3], "bwd_kernels_ms": ks[3]})
    print(json.dumps(rows[-1]), flush=True)
