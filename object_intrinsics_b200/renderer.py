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

No-grad calls (renders #2/#3 of every training step, and inference) run the hand-written CUDA path through
the C-ABI library.  Grad-mode calls (render #1, the generator step) need d/dtheta through the analytic normal
(second order); they run `torch_graph.render_differentiable`, this package's own differentiable torch
formulation on the GPU (the hand-written backward is the next step).
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

    def get(self, named_list):
        key = tuple((t.data_ptr(), t._version) for _, t in named_list)
        if key != self.key:
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


class NeuSRenderer:
    """Same constructor keywords as the reference class (renderer.py:77-96) plus `impl`."""

    def __init__(self, nerf, sdf_network, deviation_network, color_network, n_samples, n_importance, n_outside,
                 up_sample_steps, perturb, impl: str = "auto"):
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
        self._packed = PackedWeights()
        self._workspace: Optional[torch.Tensor] = None
        self._lin = {}
        self.flags = int(os.environ.get("OI_RENDER_FLAGS", "1"))   # OiRenderDesc.flags: bit 0 = L2 discard of dead scratch
        self.last_launches = 0
        self.core_events = None   # optional (torch.cuda.Event, torch.cuda.Event) recorded around the core kernel

    # -------------------------------------------------------------------------------------------
    def _linspaces(self, device):
        key = (str(device), self.n_samples, self.n_importance)
        if key not in self._lin:
            n, m = self.n_samples, self.n_importance
            lin_c = torch.linspace(0.0, 1.0, n, device=device, dtype=torch.float32)            # renderer.py:359
            lin_f = (torch.linspace(0.5 / m, 1.0 - 0.5 / m, m, device=device, dtype=torch.float32)
                     if m > 0 else None)                                                      # renderer.py:53
            self._lin[key] = (lin_c, lin_f)
        return self._lin[key]

    def packed_weights(self):
        return self._packed.get(collect_params(self.sdf_network, self.color_network, self.deviation_network,
                                               with_style=False))

    # -------------------------------------------------------------------------------------------
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
        needs_grad = torch.is_grad_enabled() and (
            any(t.requires_grad for _, t in params) or w.requires_grad or rays_o.requires_grad or rays_d.requires_grad)
        if needs_grad:
            from . import torch_graph
            ret = torch_graph.render_differentiable(self, rays_o, rays_d, near, far, w, float(cos_anneal_ratio),
                                                    t_rand, z_vals)
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
    def _render_cuda(self, params, rays_o, rays_d, near, far, w, cos_anneal_ratio, t_rand, z_vals, return_z_vals):
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
        blob = self._packed.get(params)
        lin_c, lin_f = self._linspaces(dev)

        out = {}
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
            setattr(d, k, out[k].data_ptr())
        d.z_vals_out = _lib.ptr(out.get("z_vals"))
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
        return out

    # -------------------------------------------------------------------------------------------
    def extract_geometry(self, *args, **kwargs):
        raise NotImplementedError("mesh extraction (renderer.py:475-492, PyMCubes) is outside the hot path")


FusedNeuSRenderer = NeuSRenderer
