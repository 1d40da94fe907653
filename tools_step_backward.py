"""One grad-mode render step (cfg2 bs=1) for profiling under ncu: python tools_step_backward.py [n_steps]"""
import os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch
import bench_inputs as BI
from object_intrinsics_b200 import fields
from object_intrinsics_b200.renderer import NeuSRenderer
P = BI.load_flat_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
bs = int(os.environ.get("BS", "1"))
ro, rd, near, far = [t.cuda() for t in BI.synthetic_rays(bs, 64, seed=1)]
z = torch.randn(bs, 64, device="cuda")
r = NeuSRenderer(None, sdf, dev, col, n_samples=64, n_importance=0, n_outside=0, up_sample_steps=1, perturb=1)
params = list(sdf.parameters()) + list(col.parameters()) + list(dev.parameters())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for p in params:
        p.grad = None          # as optimizer.zero_grad(set_to_none=True): no accumulation kernels in the launch list
    w = sdf.style(z)
    out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w)
    img = out["color_fine"] + (1.0 - out["weight_sum"])
    ((img ** 2).mean() + 0.1 * out["gradient_error"]).backward()
torch.cuda.synchronize()
