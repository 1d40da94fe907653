"""Tiny invocations of the round-1 additions for `compute-sanitizer --tool memcheck` (GPU box):
backward (tcgen05 and FP32 cores), AugmentPipe geometric path forward / adjoint / double backward."""
import os, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_backward_gpu as T
from object_intrinsics_b200.augment import AugmentPipe
for impl in ("tcgen05", "ffma"):
    meta, c, w = T._inputs("cfgd_n16_m4_D8", 40, n_inst=2)     # ragged tiles, 2 instances
    r = T._build(meta, n_importance=0, impl=impl)
    out = r.render(c["rays_o"], c["rays_d"], c["near"], c["far"], cos_anneal_ratio=1.0, perturb_overwrite=0, w=w)
    (out["color_fine"].sum() + out["weight_sum"].sum() + out["gradient_error"] + out["weight_max"].sum()).backward()
    torch.cuda.synchronize()
    print("backward", impl, "ok", float(r.sdf_network.pts_linears[1].weight.grad.abs().max()))
pipe = AugmentPipe(scale=1, xint=1, rotate=1).cuda()
x = torch.rand(2, 3, 24, 20, device="cuda", requires_grad=True)
wgt = torch.randn(3, 24, 20, device="cuda", requires_grad=True)
y = pipe(x)
(g,) = torch.autograd.grad((y * wgt).sum(), x, create_graph=True)     # adjoint
g.pow(2).sum().backward()                                             # double backward = forward again
torch.cuda.synchronize()
print("augment ok", float(y.abs().mean()))
