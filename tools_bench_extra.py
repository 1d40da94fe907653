"""Secondary measurements for DESIGN.md / profiles (run on the GPU box): other BASELINE configs, both cores, and the
torch-eager formulation of the same path on the same GPU (the 'kernel to beat' of SURVEY 8d)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_params
from oracle import neus_oracle as O
from object_intrinsics_b200 import fields, torch_graph
from object_intrinsics_b200.renderer import NeuSRenderer

P = load_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10, warm=3):
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
for name, bs, patch, n, m, impl in [("cfg2 bs=1", 1, 64, 64, 0, "auto"), ("cfg2 bs=4 (headline)", 4, 64, 64, 0, "auto"),
                                    ("cfg2 bs=8", 8, 64, 64, 0, "auto"), ("cfg2 bs=4 FFMA core", 4, 64, 64, 0, "ffma"),
                                    ("shipped config 128x128, 16+4", 1, 128, 16, 4, "auto"),
                                    ("cfg4 128x128, 64+64", 1, 128, 64, 64, "auto"),
                                    ("cfg4 128x128, 64+64 FFMA core", 1, 128, 64, 64, "ffma")]:
    ro, rd, near, far = [t.cuda() for t in O.synthetic_rays(bs, patch, seed=1)]
    z = torch.randn(bs, 64, device="cuda")
    r = NeuSRenderer(None, sdf, dev, col, n_samples=n, n_importance=m, n_outside=0, up_sample_steps=1, perturb=0, impl=impl)
    with torch.no_grad():
        w = sdf.style(z)
        ms = timeit(lambda: r.render(ro, rd, near, far, cos_anneal_ratio=1.0, perturb_overwrite=0, z=z, w=w))
    R = ro.shape[0]
    rows.append((name, R, n + m, ms, R / ms * 1e3))
    print(f"{name:36s} R={R:6d} S={n+m:3d}  {ms:8.3f} ms  {R/ms*1e3/1e6:7.3f} M rays/s", flush=True)

# torch-eager formulation on the same GPU (package's own differentiable path, no_grad), cfg2 bs=1 and bs=4
for bs in (1, 4):
    ro, rd, near, far = [t.cuda() for t in O.synthetic_rays(bs, 64, seed=1)]
    z = torch.randn(bs, 64, device="cuda")
    r = NeuSRenderer(None, sdf, dev, col, n_samples=64, n_importance=0, n_outside=0, up_sample_steps=1, perturb=0)
    with torch.no_grad():
        w = sdf.style(z)
        ms = timeit(lambda: torch_graph.render_differentiable(r, ro, rd, near, far, w, 1.0), n=5, warm=2)
    print(f"torch eager (cuBLAS fp32 + elementwise), cfg2 bs={bs}: {ms:8.3f} ms  {ro.shape[0]/ms*1e3/1e6:7.3f} M rays/s", flush=True)
    rows.append((f"torch eager cfg2 bs={bs}", ro.shape[0], 64, ms, ro.shape[0] / ms * 1e3))
# grad-mode step (forward + backward through oi_render_backward), cfg2 bs=1; see tools_bench_backward.py for the
# comparison with the torch formulation
ro, rd, near, far = [t.cuda() for t in O.synthetic_rays(1, 64, seed=1)]
z = torch.randn(1, 64, device="cuda")
r = NeuSRenderer(None, sdf, dev, col, n_samples=64, n_importance=0, n_outside=0, up_sample_steps=1, perturb=0)
def gstep():
    w = sdf.style(z)
    out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, perturb_overwrite=0, z=z, w=w)
    (out["color_fine"].sum() + out["weight_sum"].sum() + 10 * out["gradient_error"]).backward()
ms = timeit(gstep, n=5, warm=2)
print(f"grad-mode render + backward (oi_render_backward), cfg2 bs=1: {ms:8.3f} ms", flush=True)
rows.append(("grad-mode fwd+bwd cfg2 bs=1", 4096, 64, ms, 4096 / ms * 1e3))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_extra.json"), "w"), indent=1)
