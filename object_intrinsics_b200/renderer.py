"""Drop-in replacement for the reference's `NeuSRenderer` backed by the fused sm_100a render kernels.

Boundary: the reference builds its renderer from
`configs/train.yaml:69-76` (`renderer.__target__`) through `build_from_config(renderer, nerf=None,
sdf_network=..., deviation_network=..., color_network=...)` at `src/models/generator.py:51-57` and calls
exactly one method, `renderer.render(...)`, at `generator.py:245-252`.  This class keeps the reference's
constructor keywords, attributes and `render` signature / return dict
(`src/third_party/neus/models/renderer.py:77-96, 351-473`), so selecting it is a one-line override:

    python scripts/train.py -d data/example -t fused \
        model.generator.kwargs.renderer.__target__=object_intrinsics_b200.renderer.NeuSRenderer

The renderer owns no parameters; the three networks stay the `nn.Module`s registered on the Generator
(checkpoints / EMA / DDP unchanged).  Packed weight blobs are derived caches keyed on the parameters'
(data_ptr, _version) and rebuilt when any parameter changes.

Every call runs the hand-written CUDA path through the C-ABI library.  Grad-mode calls (render #1 of a training
step, the generator update) go through `_RenderFunction`: the same forward kernels, and `oi_render_backward`
(csrc/oi_render_bwd.cu) for d/dtheta of every parameter the path reads, including the second-order terms through
the analytic normal that the reference obtains from autograd (`fields.py:104-122`, create_graph=True).  Rays are
constants of the training path (poses are sampled, `generator.py:66-78`); a call whose rays require grad raises
unless the renderer was built with `grad_impl="torch"` (`torch_graph.render_differentiable`, the differentiable
torch formulation kept for that case and as the A/B reference of the backward kernel).
There is NO CPU path and no silent fallback: a missing library or an unsupported argument raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _lib

_IMPL = {"auto": _lib.OI_IMPL_AUTO, "ffma": _lib.OI_IMPL_FFMA, "tcgen05": _lib.OI_IMPL_TCGEN05}

OUT_KEYS_PER_POINT = ("cdf_fine", "weights", "inside_sphere", "mid_z_vals", "sdf", "pts_norm")
OUT_KEYS_PER_POINT3 = ("gradients", "pts", "raw_color")
OUT_KEYS_PER_RAY = ("s_val", "weight_sum", "weight_max")


def _film_modules(sdf_network, color_network):
    return list(sdf_network.pts_linears) + [color_network.views_linears]


def collect_params(sdf_network, color_network, deviation_network, with_style=True):
    """Ordered list of (name, tensor) of every parameter the path reads (state_dict names of the reference)."""
    out = []
    for i, f in enumerate(sdf_network.pts_linears):
        for nm, t in (("weight", f.weight), ("bias", f.bias), ("gamma.weight", f.gamma.weight),
                      ("gamma.bias", f.gamma.bias), ("beta.weight", f.beta.weight), ("beta.bias", f.beta.bias)):
            out.append((f"sdf_network.pts_linears.{i}.{nm}", t))
    out.append(("sdf_network.sigma_linear.weight", sdf_network.sigma_linear.weight))
    out.append(("sdf_network.sigma_linear.bias", sdf_network.sigma_linear.bias))
    v = color_network.views_linears
    for nm, t in (("weight", v.weight), ("bias", v.bias), ("gamma.weight", v.gamma.weight),
                  ("gamma.bias", v.gamma.bias), ("beta.weight", v.beta.weight), ("beta.bias", v.beta.bias)):
        out.append((f"color_network.views_linears.{nm}", t))
    out.append(("color_network.rgb_linear.weight", color_network.rgb_linear.weight))
    out.append(("color_network.rgb_linear.bias", color_network.rgb_linear.bias))
    out.append(("deviation_network.variance", deviation_network.variance))
    if with_style and hasattr(sdf_network, "style"):
        for i, s in enumerate(sdf_network.style):
            out.append((f"sdf_network.style.{i}.weight", s.weight))
            out.append((f"sdf_network.style.{i}.bias", s.bias))
    return out


def fill_net_params(named: Dict[str, torch.Tensor]) -> _lib.OiNetParams:
    """OiNetParams from a {state_dict name: CUDA fp32 contiguous tensor} mapping."""
    depth = 0
    while f"sdf_network.pts_linears.{depth}.weight" in named:
        depth += 1
    p = _lib.OiNetParams()
    p.depth, p.width, p.style_dim = depth, named["sdf_network.sigma_linear.weight"].shape[1], \
        named["sdf_network.pts_linears.0.gamma.weight"].shape[1]
    for k, t in named.items():
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"parameter {k} must be a contiguous fp32 CUDA tensor (got {t.dtype}, {t.device})")
    for i in range(depth):
        pre = f"sdf_network.pts_linears.{i}."
        p.pts_weight[i] = named[pre + "weight"].data_ptr()
        p.pts_bias[i] = named[pre + "bias"].data_ptr()
        p.gamma_weight[i] = named[pre + "gamma.weight"].data_ptr()
        p.gamma_bias[i] = named[pre + "gamma.bias"].data_ptr()
        p.beta_weight[i] = named[pre + "beta.weight"].data_ptr()
        p.beta_bias[i] = named[pre + "beta.bias"].data_ptr()
    pre = "color_network.views_linears."
    j = _lib.OI_MAX_DEPTH
    p.gamma_weight[j] = named[pre + "gamma.weight"].data_ptr()
    p.gamma_bias[j] = named[pre + "gamma.bias"].data_ptr()
    p.beta_weight[j] = named[pre + "beta.weight"].data_ptr()
    p.beta_bias[j] = named[pre + "beta.bias"].data_ptr()
    p.views_weight = named[pre + "weight"].data_ptr()
    p.views_bias = named[pre + "bias"].data_ptr()
    p.sigma_weight = named["sdf_network.sigma_linear.weight"].data_ptr()
    p.sigma_bias = named["sdf_network.sigma_linear.bias"].data_ptr()
    p.rgb_weight = named["color_network.rgb_linear.weight"].data_ptr()
    p.rgb_bias = named["color_network.rgb_linear.bias"].data_ptr()
    p.variance = named["deviation_network.variance"].data_ptr()
    for i in range(3):
        if f"sdf_network.style.{i}.weight" in named:
            p.style_weight[i] = named[f"sdf_network.style.{i}.weight"].data_ptr()
            p.style_bias[i] = named[f"sdf_network.style.{i}.bias"].data_ptr()
    return p


class PackedWeights:
    """Device blob in the layout the kernels stream (oi_pack_weights); rebuilt when a parameter changes."""

    def __init__(self):
        self.key = None
        self.blob: Optional[torch.Tensor] = None
        self.depth = 0
        self.repacks = 0
        self.always = os.environ.get("OI_REPACK_ALWAYS", "0") == "1"   # repack on every call (7.5 us kernel)

    def invalidate(self):
        """Forces a repack on the next call.  Needed after writes that bypass the autograd version counter:
        `p.data.copy_()`, `p.data.mul_()`, ... (they do not bump `p._version`, which keys the cache)."""
        self.key = None

    def get(self, named_list, extra=()):
        key = tuple((t.data_ptr(), t._version) for _, t in named_list) + tuple(extra)
        if key != self.key or self.always:
            named = {k: t.detach() for k, t in named_list}
            p = fill_net_params(named)
            L = _lib.lib()
            nbytes = C.c_size_t(0)
            _lib.check(L.oi_packed_weights_bytes(p.depth, C.byref(nbytes)), "oi_packed_weights_bytes")
            dev = named["deviation_network.variance"].device
            if self.blob is None or self.blob.numel() < nbytes.value or self.blob.device != dev:
                self.blob = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.oi_pack_weights(C.byref(p), self.blob.data_ptr(), self.blob.numel(),
                                             _lib.current_stream_ptr(dev)), "oi_pack_weights")
            self.key, self.depth = key, p.depth
            self.repacks += 1
        return self.blob


FN_OUT_KEYS = OUT_KEYS_PER_POINT + OUT_KEYS_PER_POINT3 + OUT_KEYS_PER_RAY + ("color_fine", "gradient_error",
                                                                           "surface_loss", "z_vals")
FN_NON_DIFF = ("inside_sphere", "mid_z_vals", "pts_norm", "pts", "z_vals")


def film_tables(sdf_network, color_network, w):
    """gamma, beta [bs, 9, 128] as differentiable torch ops (volume_renderer.py:27-30,47-48,56-57).  Only the
    autograd graph of this result is used (it routes dL/dgamma, dL/dbeta to the FiLM linears and to w); the
    kernels compute their own tables from the packed blob."""
    mods = _film_modules(sdf_network, color_network)
    Wg = torch.stack([m.gamma.weight for m in mods])      # [L,128,64]
    bg = torch.stack([m.gamma.bias for m in mods])
    Wb = torch.stack([m.beta.weight for m in mods])
    bb = torch.stack([m.beta.bias for m in mods])
    gam = 15.0 * (torch.einsum("bk,lnk->bln", w, Wg) + bg[None]) + 30.0
    bet = 0.25 * (torch.einsum("bk,lnk->bln", w, Wb) + bb[None])
    return gam, bet


class _FilmGraph(torch.autograd.Function):
    """The autograd edge of `film_tables` without its forward arithmetic: the CUDA kernels compute their own FiLM tables
    from the packed blob and `w`, so the forward only hands out two uninitialised [bs, L, 128] place-holders (no
    launch); the backward turns dL/dgamma, dL/dbeta (written by oi_render_backward) into the gradients of the L FiLM
    linears and of `w` with two batched products instead of the ~25 nodes of the einsum / stack graph:
        gamma = 15 (Wg w + bg) + 30,  beta = 0.25 (Wb w + bb)          (volume_renderer.py:27-30,47-48,56-57)."""

    @staticmethod
    def forward(ctx, w, *flat):
        L = len(flat) // 4
        ctx.L = L
        ctx.save_for_backward(w, *flat[:L], *flat[2 * L:3 * L])          # w, gamma weights, beta weights
        gam = torch.empty(w.shape[0], L, flat[0].shape[0], device=w.device, dtype=torch.float32)
        return gam, torch.empty_like(gam)

    @staticmethod
    def backward(ctx, g_gam, g_bet):
        L = ctx.L
        w, Ws = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        if g_gam is None and g_bet is None:
            return (None,) * (1 + 4 * L)
        zeros = lambda: torch.zeros(w.shape[0], L, Ws[0].shape[0], device=w.device, dtype=torch.float32)
        G = torch.cat([15.0 * (g_gam if g_gam is not None else zeros()),
                       0.25 * (g_bet if g_bet is not None else zeros())], 1)           # [bs, 2L, 128]
        wf = w.detach().to(torch.float32)
        dW = torch.einsum("bln,bk->lnk", G, wf)                                          # [2L, 128, 64]
        db = G.sum(0)                                                                    # [2L, 128]
        dw = None
        if ctx.needs_input_grad[0]:
            dw = torch.einsum("bln,lnk->bk", G, torch.stack([t.detach() for t in Ws])).to(w.dtype)
        return (dw, *dW[:L].unbind(0), *db[:L].unbind(0), *dW[L:].unbind(0), *db[L:].unbind(0))


def film_graph(sdf_network, color_network, w):
    """gamma / beta place-holders [bs, L, 128] carrying the autograd edges to the FiLM linears and to `w` (_FilmGraph);
    `film_tables` is the plain-torch statement of the same map (tests compare the two)."""
    mods = _film_modules(sdf_network, color_network)
    flat = [m.gamma.weight for m in mods] + [m.gamma.bias for m in mods] + \
           [m.beta.weight for m in mods] + [m.beta.bias for m in mods]
    return _FilmGraph.apply(w, *flat)


class _RenderFunction(torch.autograd.Function):
    """Forward = the fused CUDA render; backward = oi_render_backward.  Differentiable inputs: the FiLM tables
    (depth+1 slots) and the tensors of `_direct_params`."""

    @staticmethod
    def forward(ctx, renderer, geo, gam, bet, *direct):
        params, rays_o, rays_d, near, far, w, cos_anneal_ratio, t_rand, z_vals = geo
        out = renderer._render_cuda(params, rays_o, rays_d, near, far, w, cos_anneal_ratio, t_rand, z_vals, True)
        ctx.renderer = renderer
        ctx.cos_anneal_ratio = cos_anneal_ratio
        ctx.blob = renderer._packed.blob
        ctx.blob_key = renderer._packed.key
        ctx.depth = renderer._packed.depth
        ctx.n_film = gam.shape[1]
        ctx.direct_shapes = [t.shape for t in direct]
        ctx.set_materialize_grads(False)   # unused outputs arrive as None -> NULL adjoint pointers in the C-ABI
        f32c = lambda t: t.detach().to(torch.float32).contiguous()
        ctx.save_for_backward(f32c(rays_o), f32c(rays_d), out["z_vals"], f32c(w), out["sdf"], out["gradients"],
                              out["raw_color"])
        res = tuple(out[k] for k in FN_OUT_KEYS)
        ctx.mark_non_differentiable(*[out[k] for k in FN_NON_DIFF])
        return res

    @staticmethod
    def backward(ctx, *gouts):
        rays_o, rays_d, z_vals, w, sdf, gradients, raw_color = ctx.saved_tensors
        r = ctx.renderer
        if r._packed.key != ctx.blob_key:
            raise RuntimeError("a parameter of the render path was modified between forward and backward")
        L = _lib.lib()
        dev = rays_o.device
        R, S = z_vals.shape
        n_inst = w.shape[0]
        D = ctx.depth
        gmap = dict(zip(FN_OUT_KEYS, gouts))
        keep = []

        def adj(k):
            g = gmap.get(k)
            if g is None:
                return None
            g = g.detach().to(torch.float32).contiguous()
            keep.append(g)
            return g.data_ptr()

        # one zeroed flat buffer carved into the gradient tensors (shapes of the parameters)
        shapes = [("pts_weight%d" % l, (128, 3 if l == 0 else 128)) for l in range(D)] + \
                 [("pts_bias%d" % l, (128,)) for l in range(D)] + \
                 [("sigma_weight", (1, 128)), ("sigma_bias", (1,)), ("views_weight", (128, 131)),
                  ("views_bias", (128,)), ("rgb_weight", (3, 128)), ("rgb_bias", (3,)), ("variance", ()),
                  ("film_gamma", (n_inst, _lib.OI_MAX_DEPTH + 1, 128)), ("film_beta", (n_inst, _lib.OI_MAX_DEPTH + 1, 128))]
        offs, total = {}, 0
        for k, shp in shapes:
            n = 1
            for v in shp:
                n *= v
            offs[k] = (total, n, shp)
            total += (n + 3) // 4 * 4          # keep every tensor 16-byte aligned (vector reductions)
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        G = {k: flat[o:o + n].view(shp) for k, (o, n, shp) in offs.items()}

        d = _lib.OiRenderBwdDesc()
        d.n_rays, d.rays_per_instance, d.n_samples_total, d.n_samples = R, R // n_inst, S, r.n_samples
        d.depth, d.flags, d.cos_anneal_ratio = D, r.flags, ctx.cos_anneal_ratio
        d.impl = _IMPL[r.bwd_impl]
        d.rays_o, d.rays_d, d.z_vals, d.style_w = rays_o.data_ptr(), rays_d.data_ptr(), z_vals.data_ptr(), w.data_ptr()
        d.packed_weights = ctx.blob.data_ptr()
        d.sdf, d.gradients, d.raw_color = sdf.data_ptr(), gradients.data_ptr(), raw_color.data_ptr()
        for k in _lib.BWD_ADJOINT_KEYS:
            setattr(d, "g_" + k, adj(k))
        for l in range(D):
            d.grads.pts_weight[l] = G["pts_weight%d" % l].data_ptr()
            d.grads.pts_bias[l] = G["pts_bias%d" % l].data_ptr()
        for k in ("sigma_weight", "sigma_bias", "views_weight", "views_bias", "rgb_weight", "rgb_bias", "variance",
                  "film_gamma", "film_beta"):
            setattr(d.grads, k, G[k].data_ptr())
        if r.bwd_events is not None:
            d.evt_core_start, d.evt_core_stop = r.bwd_events[0].cuda_event, r.bwd_events[1].cuda_event
        nbytes = C.c_size_t(0)
        with torch.cuda.device(dev):
            _lib.check(L.oi_render_backward_workspace_bytes(C.byref(d), C.byref(nbytes)),
                       "oi_render_backward_workspace_bytes")
            if r._bwd_workspace is None or r._bwd_workspace.numel() < nbytes.value or r._bwd_workspace.device != dev:
                # zeroed once: the workspace holds the state of the fp16 overflow guard (OiRenderBwdDesc.flags)
                r._bwd_workspace = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
            d.workspace, d.workspace_bytes = r._bwd_workspace.data_ptr(), r._bwd_workspace.numel()
            _lib.check(L.oi_render_backward(C.byref(d), _lib.current_stream_ptr(dev)), "oi_render_backward")
        r._last_bwd_desc = d
        # slots 0..D-1 and OI_MAX_DEPTH (slices, not a list index: that would copy an index tensor host->device
        # and stall the host until the kernels above have finished)
        last = _lib.OI_MAX_DEPTH
        g_gam = torch.cat([G["film_gamma"][:, :D], G["film_gamma"][:, last:last + 1]], 1)
        g_bet = torch.cat([G["film_beta"][:, :D], G["film_beta"][:, last:last + 1]], 1)
        direct = [G["pts_weight%d" % l] for l in range(D)] + [G["pts_bias%d" % l] for l in range(D)] + \
                 [G[k] for k in ("sigma_weight", "sigma_bias", "views_weight", "views_bias", "rgb_weight", "rgb_bias",
                                 "variance")]
        direct = [g.reshape(shp) for g, shp in zip(direct, ctx.direct_shapes)]
        return (None, None, g_gam, g_bet, *direct)


def _direct_params(sdf_network, color_network, deviation_network):
    """Parameters whose gradient oi_render_backward writes directly (order fixed by _RenderFunction.backward)."""
    return [f.weight for f in sdf_network.pts_linears] + [f.bias for f in sdf_network.pts_linears] + \
           [sdf_network.sigma_linear.weight, sdf_network.sigma_linear.bias, color_network.views_linears.weight,
            color_network.views_linears.bias, color_network.rgb_linear.weight, color_network.rgb_linear.bias,
            deviation_network.variance]


class NeuSRenderer:
    """Same constructor keywords as the reference class (renderer.py:77-96) plus `impl` and `grad_impl`."""

    def __init__(self, nerf, sdf_network, deviation_network, color_network, n_samples, n_importance, n_outside,
                 up_sample_steps, perturb, impl: str = "auto", grad_impl: str = "cuda"):
        self.nerf = nerf
        self.sdf_network = sdf_network
        self.deviation_network = deviation_network
        self.color_network = color_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.up_sample_steps = up_sample_steps
        self.perturb = perturb
        self.impl = impl
        if impl not in _IMPL:
            raise ValueError(f"impl must be one of {sorted(_IMPL)}")
        if grad_impl not in ("cuda", "torch"):
            raise ValueError("grad_impl must be 'cuda' or 'torch'")
        self.grad_impl = os.environ.get("OI_GRAD_IMPL", grad_impl)
        self.bwd_impl = os.environ.get("OI_BWD_IMPL", impl)   # core of oi_render_backward: auto/tcgen05 or ffma
        self._packed = PackedWeights()
        self._workspace: Optional[torch.Tensor] = None
        self._bwd_workspace: Optional[torch.Tensor] = None
        self.bwd_events = None    # optional (torch.cuda.Event, torch.cuda.Event) around the MLP backward kernel
        self._lin = {}
        self.flags = int(os.environ.get("OI_RENDER_FLAGS", "1"))   # OiRenderDesc.flags: bit 0 = L2 discard of dead scratch
        self.last_launches = 0
        self.core_events = None   # optional (torch.cuda.Event, torch.cuda.Event) recorded around the core kernel
        # the packed blob is keyed on (data_ptr, _version); load_state_dict goes through copy_ (bumps _version) but
        # hooks are cheap insurance against loaders that write through `.data`
        for mod in (sdf_network, color_network, deviation_network):
            if hasattr(mod, "register_load_state_dict_post_hook"):
                mod.register_load_state_dict_post_hook(lambda *_a, _p=self._packed: _p.invalidate())

    def invalidate(self):
        """Drop the packed-weight cache.  Call after modifying a parameter through `.data` (EMA copies, manual
        clipping, custom initialisers): such writes do not change `param._version`, which keys the cache."""
        self._packed.invalidate()

    # -------------------------------------------------------------------------------------------
    def _linspaces(self, device):
        key = (str(device), self.n_samples, self.n_importance, self.up_sample_steps)
        if key not in self._lin:
            n, m = self.n_samples, self.n_importance // max(self.up_sample_steps, 1)   # new samples per step (:404)
            lin_c = torch.linspace(0.0, 1.0, n, device=device, dtype=torch.float32)            # renderer.py:359
            lin_f = (torch.linspace(0.5 / m, 1.0 - 0.5 / m, m, device=device, dtype=torch.float32)
                     if m > 0 else None)                                                      # renderer.py:53
            self._lin[key] = (lin_c, lin_f)
        return self._lin[key]

    def _mode_key(self):
        # a train()/eval() toggle also forces a repack (EMA swaps and test-time copies usually sit on that boundary)
        return tuple(bool(getattr(m, "training", False)) for m in (self.sdf_network, self.color_network,
                                                                   self.deviation_network))

    def packed_weights(self):
        return self._packed.get(collect_params(self.sdf_network, self.color_network, self.deviation_network,
                                               with_style=False), self._mode_key())

    # -------------------------------------------------------------------------------------------
    def last_backward_operand_format(self) -> str:
        """'fp16' or 'tf32': the format of the per-point operands of the weight-gradient contraction that the rule of
        the tensor-core backward yields for the adjoint statistics of the LAST call and the CURRENT state of the
        workspace's overflow guard -- i.e. what that call ran on, except that the first call on a workspace (the TF32
        range probe) already answers for the calls after it, and a call that tripped the guard answers 'tf32'
        although it still ran on fp16 (OiRenderBwdDesc.flags; `self.flags |= 32` forces TF32, `|= 64` fp16).
        Diagnostic: synchronises the stream."""
        d = getattr(self, "_last_bwd_desc", None)
        if d is None:
            raise RuntimeError("no backward has run on this renderer yet")
        fmt = C.c_int32(-1)
        dev = self._bwd_workspace.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().oi_render_backward_operand_format(C.byref(d), C.byref(fmt),
                                                                    _lib.current_stream_ptr(dev)),
                       "oi_render_backward_operand_format")
        return "fp16" if fmt.value == 1 else "tf32"

    def last_backward_control_words(self):
        """The 12 control words of the last backward (oi_render_backward_control_words): diagnostic, synchronises."""
        d = getattr(self, "_last_bwd_desc", None)
        if d is None:
            raise RuntimeError("no backward has run on this renderer yet")
        words = (C.c_uint32 * 12)()
        dev = self._bwd_workspace.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().oi_render_backward_control_words(C.byref(d), words, _lib.current_stream_ptr(dev)),
                       "oi_render_backward_control_words")
        return list(words)

    def render(self, rays_o, rays_d, near, far, perturb_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0,
               siren_network=None, z=None, w=None, second_order=None, compute_color=True,
               compute_sample_dist=False, blend_background=False, *, t_rand=None, z_vals=None,
               return_z_vals=False):
        """renderer.py:351-473.  Keyword-only extras (not in the reference): `t_rand` injects the per-ray jitter
        draw of renderer.py:372, `z_vals` [R,S] renders given section starts, `return_z_vals` adds 'z_vals'."""
        if siren_network is not None or second_order or compute_sample_dist or blend_background:
            raise NotImplementedError("siren_network / second_order / compute_sample_dist / blend_background are "
                                      "not used by the reference's configs and are not implemented")
        if self.n_outside > 0 or self.nerf is not None:
            raise NotImplementedError("n_outside > 0 (NeRF++ background) is not implemented (configs/train.yaml:73)")
        if w is None:
            if z is None:
                raise ValueError("either z or w must be given")
            w = self.sdf_network.style(z)                                                      # fields.py:57-58
        if not rays_o.is_cuda:
            raise RuntimeError("object_intrinsics_b200 has no CPU path: rays must be CUDA tensors")

        perturb = self.perturb
        if perturb_overwrite >= 0:
            perturb = perturb_overwrite
        R = rays_o.shape[0]
        if R % w.shape[0] != 0:                                                               # fields.py:55
            raise ValueError(f"number of rays ({R}) must be a multiple of the number of instances ({w.shape[0]})")
        if t_rand is None and perturb > 0:
            t_rand = torch.rand([R, 1], device=rays_o.device) - 0.5                           # renderer.py:372

        params = collect_params(self.sdf_network, self.color_network, self.deviation_network, with_style=False)
        rays_grad = any(t is not None and t.requires_grad for t in (rays_o, rays_d, near, far, z_vals))
        needs_grad = torch.is_grad_enabled() and (any(t.requires_grad for _, t in params) or w.requires_grad or
                                                  rays_grad)
        if needs_grad and self.grad_impl == "torch":
            from . import torch_graph
            ret = torch_graph.render_differentiable(self, rays_o, rays_d, near, far, w, float(cos_anneal_ratio),
                                                    t_rand, z_vals)
        elif needs_grad:
            if rays_grad:
                raise NotImplementedError("gradients w.r.t. rays / near / far / z_vals are not produced by the CUDA "
                                          "backward (the training path samples poses); build the renderer with "
                                          "grad_impl='torch' for that")
            gam, bet = film_graph(self.sdf_network, self.color_network, w)
            geo = (params, rays_o, rays_d, near, far, w, float(cos_anneal_ratio), t_rand, z_vals)
            res = _RenderFunction.apply(self, geo, gam, bet, *_direct_params(self.sdf_network, self.color_network,
                                                                             self.deviation_network))
            ret = dict(zip(FN_OUT_KEYS, res))
            if not return_z_vals:
                ret.pop("z_vals")
        else:
            ret = self._render_cuda(params, rays_o, rays_d, near, far, w, float(cos_anneal_ratio), t_rand, z_vals,
                                    return_z_vals)
        if background_rgb is not None:                                                        # renderer.py:306-307
            ret["color_fine"] = ret["color_fine"] + background_rgb * (1.0 - ret["weight_sum"])
        if not compute_color:
            ret.pop("color_fine", None)
            ret.pop("raw_color", None)
        return ret

    # -------------------------------------------------------------------------------------------
    def render_with_maps(self, rays_o, rays_d, near, far, *, w, light_params, light_dir, bg_color, resolution,
                         return_raw=False, cos_anneal_ratio=0.0, perturb_overwrite=-1, t_rand=None):
        """Contract B, no-grad: `render` (renderer.py:351-473) and `Generator.render_maps`
        (generator.py:80-174) in ONE call of the C-ABI (`OiRenderDesc.maps`).  Returns (per-ray render outputs +
        the two global scalars, maps [bs,C,P,P]); no per-point tensor is allocated or written to HBM when the tcgen05
        core composites in its tile tail (128 % S == 0).  light_params [10] = ambient[3], diffuse[3], specular[3],
        shininess; light_dir, bg_color [bs,3].  Differentiable callers use `render` + `generator_ops.render_maps`."""
        if not rays_o.is_cuda:
            raise RuntimeError("object_intrinsics_b200 has no CPU path: rays must be CUDA tensors")
        perturb = self.perturb if perturb_overwrite < 0 else perturb_overwrite
        R = rays_o.shape[0]
        if t_rand is None and perturb > 0:
            t_rand = torch.rand([R, 1], device=rays_o.device) - 0.5                           # renderer.py:372
        params = collect_params(self.sdf_network, self.color_network, self.deviation_network, with_style=False)
        maps = dict(light_params=light_params, light_dir=light_dir, bg_color=bg_color, resolution=int(resolution),
                    return_raw=bool(return_raw))
        with torch.no_grad():
            return self._render_cuda(params, rays_o, rays_d, near, far, w, float(cos_anneal_ratio), t_rand, None, False,
                                     maps=maps)

    def _render_cuda(self, params, rays_o, rays_d, near, far, w, cos_anneal_ratio, t_rand, z_vals, return_z_vals,
                     maps=None):
        L = _lib.lib()
        dev = rays_o.device
        f32 = dict(device=dev, dtype=torch.float32)

        def prep(t):
            return None if t is None else t.detach().to(**f32).contiguous()

        rays_o, rays_d, near, far, w, t_rand, z_vals = map(prep, (rays_o, rays_d, near, far, w, t_rand, z_vals))
        R = rays_o.shape[0]
        n_inst = w.shape[0]
        if R % n_inst != 0:
            raise ValueError(f"number of rays ({R}) must be a multiple of the number of instances ({n_inst})")
        n, m = self.n_samples, self.n_importance
        S = n + m
        if z_vals is not None and tuple(z_vals.shape) != (R, S):
            raise ValueError(f"z_vals must have shape {(R, S)}")
        blob = self._packed.get(params, self._mode_key())
        lin_c, lin_f = self._linspaces(dev)

        out = {}
        if maps is None:
            for k in OUT_KEYS_PER_POINT:
                out[k] = torch.empty((R, S), **f32)
            for k in OUT_KEYS_PER_POINT3:
                out[k] = torch.empty((R, S, 3), **f32)
        for k in OUT_KEYS_PER_RAY:
            out[k] = torch.empty((R, 1), **f32)
        out["color_fine"] = torch.empty((R, 3), **f32)
        out["gradient_error"] = torch.empty((), **f32)
        out["surface_loss"] = torch.empty((), **f32)
        if return_z_vals:
            out["z_vals"] = torch.empty((R, S), **f32)

        d = _lib.OiRenderDesc()
        d.n_rays, d.rays_per_instance, d.n_samples, d.n_importance = R, R // n_inst, n, m
        d.up_sample_steps, d.depth, d.impl, d.flags = self.up_sample_steps, self._packed.depth, _IMPL[self.impl], \
            self.flags
        d.cos_anneal_ratio = cos_anneal_ratio
        d.rays_o, d.rays_d, d.near, d.far = rays_o.data_ptr(), rays_d.data_ptr(), near.data_ptr(), far.data_ptr()
        d.t_rand = _lib.ptr(t_rand)
        d.lin_coarse, d.lin_fine = lin_c.data_ptr(), _lib.ptr(lin_f)
        d.z_vals_in = _lib.ptr(z_vals)
        d.style_w = w.data_ptr()
        d.packed_weights = blob.data_ptr()
        for k in OUT_KEYS_PER_POINT + OUT_KEYS_PER_POINT3 + OUT_KEYS_PER_RAY + ("color_fine", "gradient_error",
                                                                                "surface_loss"):
            setattr(d, k, _lib.ptr(out.get(k)))
        d.z_vals_out = _lib.ptr(out.get("z_vals"))
        map_out = None
        if maps is not None:
            from .generator_ops import _MAPS_BASE, _MAPS_RAW, _MAP_CHANNELS
            P = maps["resolution"]
            names = _MAPS_BASE + (_MAPS_RAW if maps["return_raw"] else ())
            map_out = {k: torch.empty((n_inst, _MAP_CHANNELS.get(k, 3), P, P), **f32) for k in names}
            keep = [prep(maps[k]) for k in ("light_params", "light_dir", "bg_color")]
            md = _lib.OiRenderMapsDesc()
            md.n_rays, md.rays_per_instance, md.n_samples = R, P * P, S
            md.light_params, md.light_dir, md.bg_color = [t.data_ptr() for t in keep]
            for k, t in map_out.items():
                setattr(md, k, t.data_ptr())
            if maps["return_raw"]:
                zmin = torch.empty((R,), **f32)
                md.z_min_per_ray = zmin.data_ptr()
            d.maps = C.addressof(md)
        if self.core_events is not None:
            d.evt_core_start, d.evt_core_stop = self.core_events[0].cuda_event, self.core_events[1].cuda_event

        nbytes = C.c_size_t(0)
        with torch.cuda.device(dev):
            _lib.check(L.oi_render_workspace_bytes(C.byref(d), C.byref(nbytes)), "oi_render_workspace_bytes")
            if self._workspace is None or self._workspace.numel() < nbytes.value or self._workspace.device != dev:
                self._workspace = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            d.workspace, d.workspace_bytes = self._workspace.data_ptr(), self._workspace.numel()
            _lib.check(L.oi_render_forward(C.byref(d), _lib.current_stream_ptr(dev)), "oi_render_forward")
            nl = C.c_int32(0)
            L.oi_render_launch_count(C.byref(d), C.byref(nl))
        self.last_launches = nl.value
        if map_out is not None:
            if maps["return_raw"]:
                map_out["z_min"] = zmin.reshape(n_inst, -1).min(-1).values
            return out, map_out
        return out

    # -------------------------------------------------------------------------------------------
    def extract_geometry(self, *args, **kwargs):
        raise NotImplementedError("mesh extraction (renderer.py:475-492, PyMCubes) is outside the hot path")


FusedNeuSRenderer = NeuSRenderer
