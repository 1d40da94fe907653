"""CPU tests: C-ABI library loads and exports every symbol include/oi_b200.h declares, argument validation fails
loudly without touching the GPU, instance sharding / throughput aggregation over a world_size-2 gloo group."""
import ctypes as C
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "oi_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(oi_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"liboi_b200.so does not export {s}"
    assert set(syms) == set(_lib.EXPORTS), (set(syms) ^ set(_lib.EXPORTS))
    assert L.oi_abi_version() == 1
    assert b"sm_100a" in L.oi_build_info()


def test_argument_validation_without_gpu():
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    n = C.c_size_t(0)
    assert L.oi_packed_weights_bytes(8, C.byref(n)) == 0 and n.value > 2_000_000
    assert L.oi_packed_weights_bytes(0, C.byref(n)) == -1
    assert b"depth" in L.oi_last_error()
    d = _lib.OiRenderDesc()
    assert L.oi_render_forward(C.byref(d), None) == -1                 # n_rays == 0
    d.n_rays, d.rays_per_instance, d.n_samples, d.depth = 10, 3, 16, 8
    assert L.oi_render_forward(C.byref(d), None) == -1 and b"multiple" in L.oi_last_error()
    d.rays_per_instance, d.n_importance, d.up_sample_steps = 5, 4, 3
    assert L.oi_render_workspace_bytes(C.byref(d), C.byref(n)) == -1 and b"multiple of up_sample_steps" in L.oi_last_error()
    d.up_sample_steps, d.n_samples = 2, 1 << 30                          # 32-bit point indices
    assert L.oi_render_workspace_bytes(C.byref(d), C.byref(n)) == -2   # OI_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _lib.check(-2, "x")
    u = _lib.OiUpfirdnDesc()
    assert L.oi_upfirdn2d(C.byref(u), None) == -1
    b = _lib.OiBiasActDesc()
    assert L.oi_bias_act(C.byref(b), None) == -1
    f = _lib.OiFusedBiasActDesc()
    assert L.oi_fused_bias_act(C.byref(f), None) == -1
    p = _lib.OiNetParams()
    p.depth, p.width, p.style_dim = 8, 256, 64
    assert L.oi_pack_weights(C.byref(p), None, 0, None) == -2          # W != 128 unsupported


def test_backward_and_augment_argument_validation_without_gpu():
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    n = C.c_size_t(0)
    b = _lib.OiRenderBwdDesc()
    assert L.oi_render_backward(C.byref(b), None) == -1                       # n_rays == 0
    b.n_rays, b.rays_per_instance, b.n_samples_total, b.n_samples, b.depth = 64, 64, 16, 16, 1
    assert L.oi_render_backward_workspace_bytes(C.byref(b), C.byref(n)) == -2   # depth < 2: OI_ERR_UNSUPPORTED
    b.depth = 8
    assert L.oi_render_backward_workspace_bytes(C.byref(b), C.byref(n)) == -1 and b"NULL" in L.oi_last_error()
    b.n_samples_total = 8                                                      # S < n
    assert L.oi_render_backward(C.byref(b), None) == -1 and b"n_samples" in L.oi_last_error()
    a = _lib.OiAugmentGeomDesc()
    assert L.oi_augment_geom_forward(C.byref(a), None) == -1
    a.batch, a.channels, a.height, a.width, a.filter_taps = 4, 3, 128, 128, 10
    assert L.oi_augment_geom_workspace_bytes(C.byref(a), C.byref(n)) == -1 and b"filter_taps" in L.oi_last_error()
    a.filter_taps = 12
    assert L.oi_augment_geom_workspace_bytes(C.byref(a), C.byref(n)) == 0
    # worst-case up-sampled extent (margins <= size - 1 per side) + the backward's resampled gradient
    assert n.value == 4 * (4 * 3 * (2 * 382) ** 2 + 4 * 3 * (2 * 134) ** 2)
    assert L.oi_augment_geom_backward(C.byref(a), None) == -1 and b"NULL" in L.oi_last_error()
    assert L.oi_augment_geom_setup(None, 4, 128, 128, 12, None, None, None) == -1
    assert L.oi_selftest_wgrad(None, None, 1, 1, 0, 0, 1, 1, None, None, None) == -1


def test_maps_argument_validation_without_gpu():
    """oi_render_maps / oi_render_maps_backward / OiRenderDesc.maps (contract B): descriptors are validated before any
    CUDA call, struct sizes match the header (ctypes mirrors are laid out field by field)."""
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    n = C.c_size_t(0)
    m = _lib.OiRenderMapsDesc()
    assert L.oi_render_maps(C.byref(m), None) == -1                           # n_rays == 0
    m.n_rays, m.rays_per_instance, m.n_samples = 12, 5, 4
    assert L.oi_render_maps(C.byref(m), None) == -1 and b"bad sizes" in L.oi_last_error()
    m.rays_per_instance = 6
    assert L.oi_render_maps(C.byref(m), None) == -1 and b"NULL" in L.oi_last_error()
    bd = _lib.OiRenderMapsBwdDesc()
    assert L.oi_render_maps_backward(C.byref(bd), None) == -1
    bd.fwd.n_rays, bd.fwd.rays_per_instance, bd.fwd.n_samples = 12, 6, 4
    assert L.oi_render_maps_backward(C.byref(bd), None) == -1 and b"NULL" in L.oi_last_error()
    assert C.sizeof(_lib.OiRenderMapsDesc) == 256 and C.sizeof(_lib.OiRenderMapsBwdDesc) == 408
    assert C.sizeof(_lib.OiRenderDesc) == 288
    # OiRenderDesc.maps: `weights` may be NULL only together with a maps descriptor; the maps descriptor is checked
    d = _lib.OiRenderDesc()
    d.n_rays, d.rays_per_instance, d.n_samples, d.depth = 12, 6, 16, 8
    for k in ("rays_o", "rays_d", "near", "far", "style_w"):
        setattr(d, k, 256)
    d.packed_weights = 256
    assert L.oi_render_workspace_bytes(C.byref(d), C.byref(n)) == -1 and b"weights" in L.oi_last_error()
    d.maps = C.addressof(m)
    assert L.oi_render_workspace_bytes(C.byref(d), C.byref(n)) == -1 and b"light_dir" in L.oi_last_error()
    m.light_dir, m.bg_color = 256, 256
    assert L.oi_render_workspace_bytes(C.byref(d), C.byref(n)) == 0 and n.value > 0
    launches = C.c_int32(0)
    assert L.oi_render_launch_count(C.byref(d), C.byref(launches)) == 0
    assert launches.value == 3        # film, core (compositing in its tail: 128 % 16 == 0), maps kernel
    d.flags = 16                      # maps composited in the tile tail as well
    assert L.oi_render_launch_count(C.byref(d), C.byref(launches)) == 0 and launches.value == 2
    d.n_samples, d.n_importance, d.up_sample_steps, d.flags = 16, 4, 1, 0      # S = 20: separate composite kernel
    assert L.oi_render_launch_count(C.byref(d), C.byref(launches)) == 0 and launches.value == 6


def test_grad_mode_renderer_rejects_ray_gradients_and_bad_grad_impl():
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import NeuSRenderer, film_tables
    sdf = fields.ShapeNetwork(D=2)
    col, dev = fields.ColorNetwork(D=2), fields.SingleVarianceNetwork()
    with pytest.raises(ValueError):
        NeuSRenderer(None, sdf, dev, col, 8, 0, 0, 1, 0, grad_impl="jax")
    # the FiLM-table graph that routes dL/dgamma, dL/dbeta to the FiLM linears and to w (pure torch, runs anywhere)
    w = torch.randn(2, 64, requires_grad=True)
    gam, bet = film_tables(sdf, col, w)
    assert gam.shape == (2, 3, 128) and bet.shape == (2, 3, 128)
    (gam.sum() + bet.sum()).backward()
    assert w.grad is not None and sdf.pts_linears[1].gamma.weight.grad is not None
    assert col.views_linears.beta.bias.grad is not None
    ref = 15.0 * (w.detach() @ sdf.pts_linears[0].gamma.weight.T + sdf.pts_linears[0].gamma.bias) + 30.0
    assert float((gam[:, 0].detach() - ref).abs().max()) < 1e-4


def test_renderer_rejects_cpu_and_unsupported_arguments():
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import NeuSRenderer
    sdf = fields.ShapeNetwork(D=2)
    col, dev = fields.ColorNetwork(D=2), fields.SingleVarianceNetwork()
    r = NeuSRenderer(None, sdf, dev, col, n_samples=8, n_importance=0, n_outside=0, up_sample_steps=1, perturb=0)
    ro = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.render(ro, ro, ro[:, :1], ro[:, :1], w=torch.zeros(1, 64))
    with pytest.raises(NotImplementedError):
        r.render(ro, ro, ro[:, :1], ro[:, :1], w=torch.zeros(1, 64), second_order=True)
    with pytest.raises(ValueError):
        NeuSRenderer(None, sdf, dev, col, 8, 0, 0, 1, 0, impl="triton")
    r2 = NeuSRenderer(None, sdf, dev, col, 8, 0, 1, 1, 0)
    with pytest.raises(NotImplementedError):
        r2.render(ro, ro, ro[:, :1], ro[:, :1], w=torch.zeros(1, 64))


def test_shard_instances_is_a_partition():
    from object_intrinsics_b200.parallel import shard_instances
    for total in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in shard_instances(total, world, r)]
            assert got == list(range(total))
            sizes = [len(shard_instances(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_instances(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from object_intrinsics_b200.parallel import aggregate_throughput, env_rank_world, rank_seed, shard_instances
    assert env_rank_world() == (rank, rank, world)
    mine = shard_instances(5, world, rank)                      # 5 instances over 2 ranks -> 3 + 2
    units = len(mine) * 4096.0
    secs = 0.5 if rank == 0 else 0.8
    thr, total, tmax = aggregate_throughput(units, secs)
    q.put((rank, list(mine), rank_seed(1234, rank), thr, total, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2_sharding_and_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    assert res[0][2] == 1234 and res[1][2] == 1235
    for _, _, _, thr, total, tmax in res:                        # every rank sees the same aggregate
        assert total == 5 * 4096.0 and tmax == 0.8 and abs(thr - total / 0.8) < 1e-9


def test_augment_pipe_state_dict_matches_reference_checkpoints():
    """The reference loads discriminator checkpoints with strict=True (src/utils/checkpoint.py:109-130); the pipe is
    the submodule `aug`, so its buffers must be exactly the reference's: p, Hz_geom, Hz_fbank (augment.py:123,167,179).
    Fixture: state_dict of the unmodified reference AugmentPipe (oracle/gen_golden_augment.py)."""
    import numpy as np
    from helpers import GOLDEN
    from object_intrinsics_b200.augment import AugmentPipe
    with np.load(os.path.join(GOLDEN, "augment_golden.npz")) as f:
        ref_sd = {k.split("/", 1)[1]: torch.from_numpy(f[k]) for k in f.files if k.startswith("state_dict/")}
    pipe = AugmentPipe(scale=1, xint=1)
    sd = pipe.state_dict()
    assert set(sd) == {"p", "Hz_geom", "Hz_fbank"} == set(ref_sd)
    for k in sd:
        assert sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype, k
        assert float((sd[k] - ref_sd[k]).abs().max()) <= 1e-7, k
    ref_sd["p"] = torch.tensor(0.37)                   # an ADA-adjusted probability survives the round trip
    pipe.load_state_dict(ref_sd, strict=True)
    assert float(pipe.p) == pytest.approx(0.37)
    AugmentPipe().load_state_dict(pipe.state_dict(), strict=True)


def test_film_graph_edge_matches_the_torch_formulation():
    """renderer._FilmGraph (place-holder forward, hand-written backward of gamma = 15 (Wg w + bg) + 30,
    beta = 0.25 (Wb w + bb)) routes the same gradients to the FiLM linears and to w as autograd through film_tables."""
    import torch
    from object_intrinsics_b200 import fields
    from object_intrinsics_b200.renderer import film_graph, film_tables
    torch.manual_seed(0)
    sdf, col, _ = fields.build_networks(D=4, device="cpu")
    w = torch.randn(3, 64, requires_grad=True)
    gam, bet = film_tables(sdf, col, w)
    gg, gb = torch.randn_like(gam), torch.randn_like(bet)
    params = [p for m in list(sdf.pts_linears) + [col.views_linears]
              for p in (m.gamma.weight, m.gamma.bias, m.beta.weight, m.beta.bias)]
    ref = torch.autograd.grad((gam * gg).sum() + (bet * gb).sum(), [w] + params)
    g2, b2 = film_graph(sdf, col, w)
    assert g2.shape == gam.shape and b2.shape == bet.shape
    got = torch.autograd.grad([g2, b2], [w] + params, [gg, gb])
    for a, b in zip(got, ref):
        assert float((a - b).abs().max()) <= 1e-5 * (float(b.abs().max()) + 1e-30)
    # w without grad: no gradient is formed for it
    g3, _ = film_graph(sdf, col, w.detach())
    assert torch.autograd.grad([g3], params[:1], [gg])[0] is not None


def test_backward_operand_format_rule_without_gpu():
    """bwd_mode (csrc/oi_wgrad.cuh), the rule every kernel of a backward call evaluates on the device, through its host
    twin: fp16 needs a safe guard, finite adjoints and < 2^-12 of the adjoint mass more than 2^18 below the maximum;
    flags bit 5 / 6 force TF32 / fp16 (a forced fp16 still yields to non-finite adjoints)."""
    import ctypes as C
    import struct
    from object_intrinsics_b200 import _lib
    L = _lib.lib()

    def rule(amax, total, low, flags=0, guard=1):
        w = (C.c_uint32 * 12)()
        w[1] = struct.unpack("I", struct.pack("f", amax))[0]
        w[2], w[3] = total & 0xFFFFFFFF, total >> 32
        w[4], w[5] = low & 0xFFFFFFFF, low >> 32
        f16, e_ref = C.c_int32(-1), C.c_int32(0)
        assert L.oi_selftest_bwd_mode(w, flags, guard, C.byref(f16), C.byref(e_ref)) == 0
        return f16.value, e_ref.value

    assert rule(1.0, 1 << 30, 0) == (1, -8)                      # e_ref = exponent of the maximum - 8
    assert rule(3.0e-5, 1 << 30, 0)[1] == -16 - 8                # 3e-5 = 1.97 * 2^-16
    assert rule(1.0, 1 << 30, (1 << 18) - 1)[0] == 1             # low mass just under 2^-12 of the total
    assert rule(1.0, 1 << 30, (1 << 18) + 1)[0] == 0             # ... just over: TF32
    assert rule(1.0, 1 << 40, 1 << 27)[0] == 1 and rule(1.0, 1 << 40, 1 << 29)[0] == 0   # 64-bit sums
    assert rule(1.0, 1 << 30, 0, guard=0)[0] == 0 and rule(1.0, 1 << 30, 0, guard=2)[0] == 0
    assert rule(0.0, 0, 0)[0] == 0                               # all-zero adjoints: nothing to scale
    assert rule(float("inf"), 1 << 30, 0)[0] == 0 and rule(float("inf"), 1 << 30, 0, flags=64)[0] == 0
    assert rule(1.0, 1 << 30, 0, flags=32)[0] == 0
    assert rule(1.0, 1 << 30, 1 << 29, flags=64, guard=2)[0] == 1  # forced fp16 ignores statistics and guard
    assert L.oi_selftest_bwd_mode(None, 0, 1, None, None) == -1
    w = (C.c_uint32 * 12)()
    f = C.c_int32()
    assert L.oi_selftest_bwd_mode(w, 0, 3, C.byref(f), C.byref(f)) == -1


def test_augment_setup_entry_points_validate_without_gpu():
    """oi_augment_geom_setup / _setup_ops / _setup_raw reject NULL pointers, bad sizes, bad factor kinds and forms
    before any launch (status -1 = OI_ERR_INVALID_ARGUMENT, message available through oi_last_error)."""
    import ctypes as C
    from object_intrinsics_b200 import _lib
    L = _lib.lib()
    p = C.c_void_p(0x1000)          # never dereferenced: validation fails first
    assert L.oi_augment_geom_setup(None, 4, 64, 64, 12, p, p, None) == -1
    assert L.oi_augment_geom_setup(p, 0, 64, 64, 12, p, p, None) == -1
    assert L.oi_augment_geom_setup(p, 4, 64, 64, 10, p, p, None) == -1          # taps must be a multiple of 4
    ops = (_lib.OiAugmentOp * 2)()
    ops[0].kind, ops[0].p0, ops[0].p1 = 0, C.cast(p, _lib.f32p), C.cast(p, _lib.f32p)
    ops[1].kind, ops[1].p0 = 5, C.cast(p, _lib.f32p)
    assert L.oi_augment_geom_setup_ops(ops, 2, 4, 64, 64, 12, p, p, p, None) == -1   # kind 5 does not exist
    assert b"kind" in L.oi_last_error()
    assert L.oi_augment_geom_setup_ops(ops, 9, 4, 64, 64, 12, p, p, p, None) == -1   # > OI_AUGMENT_MAX_OPS
    ops[1].kind = 0                                                                  # scale2d needs p1
    assert L.oi_augment_geom_setup_ops(ops, 2, 4, 64, 64, 12, p, p, p, None) == -1
    raw = (_lib.OiAugmentRawOp * 1)()
    raw[0].form, raw[0].draw, raw[0].gate = 7, C.cast(p, _lib.f32p), C.cast(p, _lib.f32p)
    assert L.oi_augment_geom_setup_raw(raw, 1, p, 4, 64, 64, 12, p, p, p, None) == -1   # form 7 does not exist
    assert b"form" in L.oi_last_error()
    raw[0].form, raw[0].gate = _lib.AUG_SCALE, None
    assert L.oi_augment_geom_setup_raw(raw, 1, p, 4, 64, 64, 12, p, p, p, None) == -1   # NULL gate draw
    raw[0].gate = C.cast(p, _lib.f32p)
    assert L.oi_augment_geom_setup_raw(raw, 1, None, 4, 64, 64, 12, p, p, p, None) == -1  # NULL p
    assert L.oi_augment_geom_setup_raw(raw, 0, p, 4, 64, 64, 12, p, p, p, None) == -1
