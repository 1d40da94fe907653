"""Bring-up diagnostics for the tcgen05 path (run on the GPU box): selftest GEMM error patterns + render parity."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import torch
from test_tc_gpu import _selftest

g = torch.Generator().manual_seed(0)
a = (torch.rand(128, 128, generator=g) * 2 - 1).cuda()
b = (torch.randn(128, 128, generator=g) * 0.5).cuda()
d = _selftest(a, b)
ref = (a.double() @ b.double().T).float()
err = (d - ref).abs()
print("selftest max err", float(err.max()), "ref max", float(ref.abs().max()))
if err.max() > 1e-3:
    # pattern analysis: identity-ish probes
    eye = torch.eye(128).cuda()
    d1 = _selftest(eye, b)      # expect b^T: d1[m][n] = b[n][m]
    print("A=I : matches b^T:", float((d1 - b.T).abs().max()))
    d2 = _selftest(a, eye)      # expect a
    print("B=I : matches a:", float((d2 - a).abs().max()))
    idx = (d2 - a).abs().argmax()
    print("first rows of d2 vs a", d2[0, :8].tolist(), a[0, :8].tolist())
    # where does each a[0,k] land?
    for k in range(0, 128, 8):
        probe = torch.zeros(128, 128).cuda(); probe[:, k] = 1.0
        dd = _selftest(probe, eye)
        nz = dd[0].nonzero().flatten().tolist()
        print("k", k, "-> n", nz[:8])
