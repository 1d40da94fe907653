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


@pytest.mark.parametrize("n_tiles,n_ctas,tiles_per_inst", [(1, 1, 1), (5, 2, 5), (12, 3, 4), (7, 7, 7)])
def test_wgrad_point_contraction(n_tiles, n_ctas, tiles_per_inst):
    """TMA-fed MN-major TF32 UMMA contraction over points + narrow aux products (csrc/oi_wgrad_tc.cu) vs fp64."""
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(n_tiles)
    spt = 3

    def tf32(t):   # round-to-nearest-even to 10 mantissa bits (what the producer kernel stores)
        i = t.view(torch.int32)
        return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)

    def image(t):   # logical [tile, 128 points, rows] -> [tile, 4 blocks, rows, 32 points] with the 128B swizzle
        nt, _, rows = t.shape
        b = t.reshape(nt, 4, 32, rows).permute(0, 1, 3, 2).contiguous()          # [tile, block, row, 32 pts]
        b = b.reshape(nt, 4, rows, 8, 4)                                          # 8 chunks of 4 points
        r = torch.arange(rows, device=t.device)
        src = torch.arange(8, device=t.device)[None, :] ^ (r[:, None] & 7)        # chunk c' holds logical chunk c' ^ (r&7)
        return torch.gather(b, 3, src[None, None, :, :, None].expand(nt, 4, rows, 8, 4)).reshape(nt, 4, rows, 32)

    # values spanning many binades: adjoint-like dynamic range
    logical = tf32(torch.randn(n_tiles, spt, 128, 128, device="cuda", generator=g) *
                   torch.exp(4.0 * torch.randn(n_tiles, spt, 128, 1, device="cuda", generator=g)))
    slabs = torch.stack([image(logical[:, i]) for i in range(spt)], 1).contiguous()
    A32 = tf32(torch.randn(n_tiles, 128, 4, device="cuda", generator=g))
    aux = image(A32).contiguous()
    n_inst = (n_tiles + tiles_per_inst - 1) // tiles_per_inst
    d = torch.zeros(128, 128, device="cuda")
    col = torch.zeros(n_inst, 128, 4, device="cuda")
    _lib.check(L.oi_selftest_wgrad(slabs.data_ptr(), aux.data_ptr(), n_tiles, spt, 2, 0, tiles_per_inst, n_ctas,
                                   d.data_ptr(), col.data_ptr(), None), "oi_selftest_wgrad")
    torch.cuda.synchronize()
    X, Y, A = logical[:, 2].double(), logical[:, 0].double(), A32.double()
    ref = torch.einsum("tmi,tmj->ij", X, Y)
    bound = torch.einsum("tmi,tmj->ij", X.abs(), Y.abs())
    assert float(((d.double() - ref).abs() / (bound + 1e-30)).max()) < 1e-5   # exact products, fp32 accumulation
    inst = torch.arange(n_tiles, device="cuda") // tiles_per_inst
    for b in range(n_inst):
        sel = inst == b
        ref_c = torch.einsum("tmi,tmc->ic", X[sel], A[sel])
        bnd_c = torch.einsum("tmi,tmc->ic", X[sel].abs(), A[sel].abs())
        assert float(((col[b].double() - ref_c).abs() / (bnd_c + 1e-30)).max()) < 1e-5
