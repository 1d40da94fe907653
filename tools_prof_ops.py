"""Host-side cost of one grad-mode render step (cfg2 bs=1): cProfile of 30 steps + which gradients autograd cloned.
usage (GPU box): python tools_prof_ops.py"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import bench_inputs as BI
from object_intrinsics_b200 import fields
from object_intrinsics_b200.renderer import NeuSRenderer

P = BI.load_flat_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
ro, rd, near, far = [t.cuda() for t in BI.synthetic_rays(1, 64, seed=1)]
z = torch.randn(1, 64, device="cuda")
r = NeuSRenderer(None, sdf, dev, col, n_samples=64, n_importance=0, n_outside=0, up_sample_steps=1, perturb=1)
named = [("sdf." + n, p) for n, p in sdf.named_parameters()] + [("col." + n, p) for n, p in col.named_parameters()] + \
        [("dev." + n, p) for n, p in dev.named_parameters()]
params = [p for _, p in named]


def step():
    for p in params:
        p.grad = None
    w = sdf.style(z)
    out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w)
    img = out["color_fine"] + (1.0 - out["weight_sum"])
    ((img ** 2).mean() + 0.1 * out["gradient_error"]).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
own = [n for n, p in named if p.grad is not None and p.grad.untyped_storage().nbytes() == p.grad.numel() * 4]
shared = [n for n, p in named if p.grad is not None and p.grad.untyped_storage().nbytes() != p.grad.numel() * 4]
print("grads with their own storage (cloned or fresh):", len(own), "| views of a larger buffer (stolen):", len(shared))
print("  views:", shared[:30])
n = 30
t0 = time.perf_counter()
for _ in range(n):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue time {1e3 * t_host / n:.3f} ms/step, wall incl. final sync {1e3 * t_all / n:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(40)
