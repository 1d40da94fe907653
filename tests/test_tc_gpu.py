"""GPU tests of the tcgen05 building blocks (UMMA descriptors, TMEM operand layout, SWIZZLE_128B panels)."""
import ctypes as C

import pytest
import torch

from helpers import load_params

pytestmark = pytest.mark.gpu


def _selftest(a, b=None, blob=None, depth=0, panel=0):
    from object_intrinsics_b200 import _lib
    d = torch.empty(128, 128, device="cuda")
    rc = _lib.lib().oi_selftest_tc(a.data_ptr(), _lib.ptr(b), _lib.ptr(blob), depth, panel, d.data_ptr(),
                                   _lib.current_stream_ptr())
    _lib.check(rc, "oi_selftest_tc")
    torch.cuda.synchronize()
    return d


def test_split_fp16_umma_matches_fp64_gemm():
    g = torch.Generator().manual_seed(0)
    a = (torch.rand(128, 128, generator=g) * 2 - 1).cuda()
    b = (torch.randn(128, 128, generator=g) * 0.5).cuda()
    d = _selftest(a, b)
    ref = (a.double() @ b.double().T)
    err = float((d.double() - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 2e-6 * scale, (err, scale)


def test_packed_panels_match_network_weights():
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import PackedWeights, collect_params
    P = load_params("params_D8.npz")
    sdf, col, dev = fields.build_networks(D=8, device="cuda")
    fields.load_flat_params(sdf, col, dev, P)
    pw = PackedWeights()
    blob = pw.get(collect_params(sdf, col, dev, with_style=False))
    g = torch.Generator().manual_seed(1)
    a = (torch.rand(128, 128, generator=g) * 2 - 1).cuda()
    D = 8
    for panel in (0, 3, D - 1, D, 2 * (D - 1)):
        if panel < D - 1:
            w = P[f"sdf_network.pts_linears.{panel + 1}.weight"]            # [n=out][k=in]
        elif panel == D - 1:
            w = P["color_network.views_linears.weight"][:, :128]
        else:
            l = (D - 1) - (panel - D)
            w = P[f"sdf_network.pts_linears.{l}.weight"].T                  # [n=in][k=out]
        ref = a.double().cpu() @ (256.0 * w.double()).T
        d = _selftest(a, blob=blob, depth=D, panel=panel).cpu()
        err = float((d.double() - ref).abs().max())
        assert err <= 2e-6 * float(ref.abs().max()), (panel, err)
