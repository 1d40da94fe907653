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


@pytest.mark.parametrize("n_tiles,n_splits,x_tf,y_tf,mult", [(1, 1, 0, 0, -1), (5, 2, 0, 1, 3), (7, 7, 1, 0, -1)])
def test_wgrad_point_contraction(n_tiles, n_splits, x_tf, y_tf, mult):
    """MN-major bf16x2-split UMMA contraction over points + column sums (csrc/oi_wgrad_tc.cu) vs fp64 einsum."""
    import ctypes as C
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(n_tiles)
    spt = 3
    # values spanning many binades: adjoint-like dynamic range
    slabs = torch.randn(n_tiles, spt, 32, 128, 4, device="cuda", generator=g) * \
        torch.exp(4.0 * torch.randn(n_tiles, spt, 1, 128, 1, device="cuda", generator=g))
    if x_tf:      # sin operands are FiLM pre-activations: |a| up to ~100 rad (MUFU sin, as in the forward core)
        slabs[:, 2] = 30.0 * torch.randn(n_tiles, 32, 128, 4, device="cuda", generator=g)
    if y_tf:
        slabs[:, 0] = 30.0 * torch.randn(n_tiles, 32, 128, 4, device="cuda", generator=g)
    aux = torch.randn(n_tiles, 16, 128, device="cuda", generator=g)
    d = torch.zeros(128, 128, device="cuda")
    col = torch.zeros(128, device="cuda")
    _lib.check(L.oi_selftest_wgrad(slabs.data_ptr(), aux.data_ptr(), n_tiles, spt, 2, 0, x_tf, y_tf, mult, n_splits,
                                   d.data_ptr(), col.data_ptr(), None), "oi_selftest_wgrad")
    torch.cuda.synchronize()
    tf = lambda t, k: torch.sin(t) if k else t
    # [tile, q, m, 4] -> [tile*m, channel]
    X = tf(slabs[:, 2].double(), x_tf).permute(0, 2, 1, 3).reshape(n_tiles * 128, 128)
    Y = tf(slabs[:, 0].double(), y_tf).permute(0, 2, 1, 3).reshape(n_tiles * 128, 128)
    ref = X.T @ Y
    mm = aux[:, mult].double().reshape(-1, 1) if mult >= 0 else 1.0
    ref_col = (X * mm).sum(0)
    bound = (X.abs().T @ Y.abs())
    # MUFU sin on |a| ~ 100 rad carries ~1e-5 absolute error (same as the forward core)
    assert float(((d.double() - ref).abs() / (bound + 1e-30)).max()) < (1e-4 if (x_tf or y_tf) else 3e-5)
    assert float(((col.double() - ref_col).abs() / ((X * mm).abs().sum(0) + 1e-30)).max()) < (1e-4 if x_tf else 1e-5)
