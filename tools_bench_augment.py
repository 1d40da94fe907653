"""AugmentPipe geometric path on the GPU box: the CUDA kernels (oi_augment_geom_*) vs the same chain in torch eager
(the reference's op sequence: F.pad(reflect) with host-side margins, upfirdn2d as conv2d, affine_grid + grid_sample,
conv2d) on the same inputs.  Writes gpurun_out/bench_augment.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import augment_oracle as AO
from object_intrinsics_b200.augment import AugmentPipe, geometric_transform


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
HZ = AO.hz_geom().cuda()
for name, B, C, S in [("discriminator input (configs/train.yaml:86) 4x3x128x128", 4, 3, 128),
                      ("mask discriminator 4x1x128x128", 4, 1, 128), ("8x3x64x64", 8, 3, 64)]:
    pipe = AugmentPipe(scale=1, xint=1).cuda()
    x = torch.rand(B, C, S, S, device="cuda", requires_grad=True)
    torch.manual_seed(0)
    G = pipe.sample_inverse_transform(B, S, S, x.device)
    gy = torch.randn_like(x)

    def ours_fwd():
        with torch.no_grad():
            return geometric_transform(x, G)

    def eager_fwd():
        with torch.no_grad():
            return AO.geometric_path(x, G, HZ)

    def ours_fwd_bwd():
        y = geometric_transform(x, G)
        torch.autograd.grad(y, x, gy)

    def eager_fwd_bwd():
        y = AO.geometric_path(x, G, HZ)
        torch.autograd.grad(y, x, gy)

    def ours_r1():
        y = geometric_transform(x, G)
        (g,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        gy2 = torch.autograd.grad(g.pow(2).sum(), x, allow_unused=True)

    err = float((ours_fwd() - eager_fwd()).abs().max())
    res = {"workload": name, "max_abs_diff_vs_eager": err, "ours_fwd_ms": timeit(ours_fwd), "eager_fwd_ms": timeit(eager_fwd),
           "ours_fwd_bwd_ms": timeit(ours_fwd_bwd), "eager_fwd_bwd_ms": timeit(eager_fwd_bwd)}
    res["fwd_speedup"] = res["eager_fwd_ms"] / res["ours_fwd_ms"]
    res["fwd_bwd_speedup"] = res["eager_fwd_bwd_ms"] / res["ours_fwd_bwd_ms"]
    # sampler + setup + kernels, as the discriminator calls it
    res["pipe_call_ms"] = timeit(lambda: pipe(x.detach()))
    print(json.dumps(res), flush=True)
    rows.append(res)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_augment.json"), "w"), indent=1)
