"""Grad-mode render step (forward + backward of the generator's render #1) on the GPU box: CUDA backward
(oi_render_backward) vs the torch formulation (torch_graph) on the same inputs.  Writes gpurun_out/bench_backward.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_params
from oracle import neus_oracle as O
from object_intrinsics_b200 import fields
from object_intrinsics_b200.renderer import NeuSRenderer

P = load_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
params = list(sdf.parameters()) + list(col.parameters()) + list(dev.parameters())


def timeit(fn, n=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
for name, bs, patch, n, m in [("cfg2 bs=1 (64x64 x 64)", 1, 64, 64, 0), ("cfg2 bs=4 (headline rays)", 4, 64, 64, 0),
                              ("shipped config 128x128, 16+4", 1, 128, 16, 4)]:
    ro, rd, near, far = [t.cuda() for t in O.synthetic_rays(bs, patch, seed=1)]
    z = torch.randn(bs, 64, device="cuda")
    res = {"workload": name, "rays": ro.shape[0], "samples": n + m}
    for gi in ("cuda", "torch"):
        if gi == "torch" and ro.shape[0] * (n + m) > 400000:
            continue
        r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=1,
                         grad_impl=gi)
        r.bwd_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for e in r.bwd_events:
            e.record()   # torch creates the cudaEvent_t lazily; the C-ABI needs the handle

        def gstep():
            for p in params:
                p.grad = None
            w = sdf.style(z)
            out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w)
            img = out["color_fine"] + (1.0 - out["weight_sum"])
            ((img ** 2).mean() + 0.1 * out["gradient_error"] + (out["weights"][..., None] * out["gradients"]).sum()
             * 1e-3).backward()

        def fwd_only():
            with torch.no_grad():
                w = sdf.style(z)
                r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w)

        ms = timeit(gstep)
        res[f"{gi}_step_ms"] = ms
        if gi == "cuda":
            torch.cuda.synchronize()
            res["mlp_bwd_kernel_ms"] = r.bwd_events[0].elapsed_time(r.bwd_events[1])
            res["nograd_forward_ms"] = timeit(fwd_only)
    if "torch_step_ms" in res:
        res["speedup_vs_torch"] = res["torch_step_ms"] / res["cuda_step_ms"]
    print(json.dumps(res), flush=True)
    rows.append(res)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_backward.json"), "w"), indent=1)
