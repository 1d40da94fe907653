import os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_params
from oracle import neus_oracle as O
from object_intrinsics_b200 import fields
from object_intrinsics_b200.renderer import NeuSRenderer
P = load_params("params_D8.npz")
sdf, col, dev = fields.build_networks(D=8, device="cuda")
fields.load_flat_params(sdf, col, dev, P)
ro, rd, near, far = [t.cuda() for t in O.synthetic_rays(1, 64, seed=1)]
z = torch.randn(1, 64, device="cuda")
r = NeuSRenderer(None, sdf, dev, col, n_samples=64, n_importance=0, n_outside=0, up_sample_steps=1, perturb=1)
params = list(sdf.parameters()) + list(col.parameters()) + list(dev.parameters())
def step(timing=None):
    t0 = time.perf_counter()
    for p in params: p.grad = None
    w = sdf.style(z)
    t1 = time.perf_counter()
    out = r.render(ro, rd, near, far, cos_anneal_ratio=1.0, z=z, w=w)
    t2 = time.perf_counter()
    img = out["color_fine"] + (1.0 - out["weight_sum"])
    loss = (img ** 2).mean() + 0.1 * out["gradient_error"]
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    if timing is not None: timing.append((t1-t0, t2-t1, t3-t2, t4-t3, t5-t4))
for _ in range(5): step()
T = []
for _ in range(10): step(T)
import statistics
names = ["zero+style", "render() cpu", "loss cpu", "backward() cpu", "final sync wait"]
for i, n in enumerate(names): print(f"{n:18s} {statistics.median(t[i] for t in T)*1e3:7.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
