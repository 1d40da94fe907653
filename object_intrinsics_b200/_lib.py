"""ctypes binding of liboi_b200.so (C-ABI declared in include/oi_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

OI_MAX_DEPTH = 8
OI_WIDTH = 128
OI_STYLE_DIM = 64
OI_IMPL_AUTO, OI_IMPL_FFMA, OI_IMPL_TCGEN05 = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
# OI_LIB_PATH selects an experiment build (object_intrinsics_b200/build.py --variant ...) for A/B timing; the product
# path is the in-tree default
LIB_PATH = os.environ.get("OI_LIB_PATH") or os.path.join(_HERE, "lib", "liboi_b200.so")

f32p = C.c_void_p  # device pointers travel as integers


class OiNetParams(C.Structure):
    _fields_ = [
        ("depth", C.c_int32), ("width", C.c_int32), ("style_dim", C.c_int32), ("reserved", C.c_int32),
        ("pts_weight", f32p * OI_MAX_DEPTH), ("pts_bias", f32p * OI_MAX_DEPTH),
        ("gamma_weight", f32p * (OI_MAX_DEPTH + 1)), ("gamma_bias", f32p * (OI_MAX_DEPTH + 1)),
        ("beta_weight", f32p * (OI_MAX_DEPTH + 1)), ("beta_bias", f32p * (OI_MAX_DEPTH + 1)),
        ("sigma_weight", f32p), ("sigma_bias", f32p), ("views_weight", f32p), ("views_bias", f32p),
        ("rgb_weight", f32p), ("rgb_bias", f32p), ("variance", f32p),
        ("style_weight", f32p * 3), ("style_bias", f32p * 3),
    ]


class OiRenderDesc(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int32), ("rays_per_instance", C.c_int32), ("n_samples", C.c_int32),
        ("n_importance", C.c_int32), ("up_sample_steps", C.c_int32), ("depth", C.c_int32),
        ("impl", C.c_int32), ("flags", C.c_int32), ("cos_anneal_ratio", C.c_float), ("reserved_f", C.c_float),
        ("rays_o", f32p), ("rays_d", f32p), ("near", f32p), ("far", f32p), ("t_rand", f32p),
        ("lin_coarse", f32p), ("lin_fine", f32p), ("z_vals_in", f32p), ("style_w", f32p),
        ("packed_weights", C.c_void_p),
        ("s_val", f32p), ("cdf_fine", f32p), ("weight_sum", f32p), ("weight_max", f32p), ("gradients", f32p),
        ("weights", f32p), ("gradient_error", f32p), ("inside_sphere", f32p), ("mid_z_vals", f32p),
        ("surface_loss", f32p), ("sdf", f32p), ("pts_norm", f32p), ("pts", f32p), ("color_fine", f32p),
        ("raw_color", f32p), ("z_vals_out", f32p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("evt_core_start", C.c_void_p), ("evt_core_stop", C.c_void_p),
        ("maps", C.c_void_p),      # const OiRenderMapsDesc* (host pointer) or NULL
    ]


class OiNetGrads(C.Structure):
    _fields_ = [
        ("pts_weight", f32p * OI_MAX_DEPTH), ("pts_bias", f32p * OI_MAX_DEPTH),
        ("sigma_weight", f32p), ("sigma_bias", f32p), ("views_weight", f32p), ("views_bias", f32p),
        ("rgb_weight", f32p), ("rgb_bias", f32p), ("variance", f32p), ("film_gamma", f32p), ("film_beta", f32p),
    ]


BWD_ADJOINT_KEYS = ("weights", "weight_sum", "weight_max", "color_fine", "raw_color", "gradients", "sdf", "cdf_fine",
                    "s_val", "gradient_error", "surface_loss")


class OiRenderBwdDesc(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int32), ("rays_per_instance", C.c_int32), ("n_samples_total", C.c_int32),
        ("n_samples", C.c_int32), ("depth", C.c_int32), ("impl", C.c_int32), ("flags", C.c_int32),
        ("reserved", C.c_int32), ("cos_anneal_ratio", C.c_float), ("reserved_f", C.c_float),
        ("rays_o", f32p), ("rays_d", f32p), ("z_vals", f32p), ("style_w", f32p), ("packed_weights", C.c_void_p),
        ("sdf", f32p), ("gradients", f32p), ("raw_color", f32p),
    ] + [("g_" + k, f32p) for k in BWD_ADJOINT_KEYS] + [
        ("grads", OiNetGrads),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("evt_core_start", C.c_void_p), ("evt_core_stop", C.c_void_p),
    ]


class OiUpfirdnDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("f", C.c_void_p), ("y", C.c_void_p), ("dtype", C.c_int32),
        ("batch", C.c_int32), ("channels", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("x_stride_n", C.c_int64), ("x_stride_c", C.c_int64), ("x_stride_h", C.c_int64), ("x_stride_w", C.c_int64),
        ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("y_stride_n", C.c_int64), ("y_stride_c", C.c_int64), ("y_stride_h", C.c_int64), ("y_stride_w", C.c_int64),
        ("filter_h", C.c_int32), ("filter_w", C.c_int32), ("f_stride_h", C.c_int64), ("f_stride_w", C.c_int64),
        ("up_x", C.c_int32), ("up_y", C.c_int32), ("down_x", C.c_int32), ("down_y", C.c_int32),
        ("pad_x0", C.c_int32), ("pad_x1", C.c_int32), ("pad_y0", C.c_int32), ("pad_y1", C.c_int32),
        ("flip", C.c_int32), ("gain", C.c_float),
    ]


class OiBiasActDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("b", C.c_void_p), ("xref", C.c_void_p), ("yref", C.c_void_p), ("dy", C.c_void_p),
        ("y", C.c_void_p), ("dtype", C.c_int32), ("grad", C.c_int32), ("act", C.c_int32), ("reserved", C.c_int32),
        ("alpha", C.c_float), ("gain", C.c_float), ("clamp", C.c_float),
        ("size_x", C.c_int32), ("size_b", C.c_int32), ("step_b", C.c_int32),
    ]


class OiFusedBiasActDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("bias", C.c_void_p), ("ref", C.c_void_p), ("y", C.c_void_p), ("dtype", C.c_int32),
        ("act", C.c_int32), ("grad", C.c_int32), ("size_x", C.c_int32), ("size_b", C.c_int32), ("step_b", C.c_int32),
        ("alpha", C.c_float), ("scale", C.c_float),
    ]


class OiGenRaysDesc(C.Structure):
    _fields_ = [
        ("n_instances", C.c_int32), ("resolution", C.c_int32), ("scene_resolution", C.c_int32), ("reserved", C.c_int32),
        ("cam_dist", C.c_float), ("reserved_f", C.c_float),
        ("b2w", f32p), ("c2b", f32p), ("w2c", f32p), ("intrinsics_inv", f32p),
        ("rays_o", f32p), ("rays_d", f32p), ("x_offset", f32p), ("y_offset", f32p), ("near", f32p), ("far", f32p),
    ]


class OiAugmentGeomDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("filter_taps", C.c_int32), ("reserved", C.c_int32),
        ("filter", C.POINTER(C.c_float)), ("theta", f32p), ("margins", C.c_void_p), ("x", f32p), ("y", f32p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class OiAugmentOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p0", f32p), ("p1", f32p)]


class OiAugmentRawOp(C.Structure):
    _fields_ = [("form", C.c_int32), ("reserved", C.c_int32), ("draw", f32p), ("gate", f32p), ("prob", C.c_float),
                ("param", C.c_float)]


AUG_XFLIP, AUG_ROTATE90, AUG_XINT, AUG_SCALE, AUG_ROTATE, AUG_ANISO, AUG_XFRAC = range(7)


class OiRenderMapsDesc(C.Structure):
    _fields_ = [
        ("n_rays", C.c_int32), ("rays_per_instance", C.c_int32), ("n_samples", C.c_int32), ("reserved", C.c_int32),
        ("shininess", C.c_float), ("reserved_f", C.c_float),
        ("ambient_color", C.c_float * 3), ("diffuse_color", C.c_float * 3), ("specular_color", C.c_float * 3),
        ("pad", C.c_float),
        ("weights", f32p), ("gradients", f32p), ("raw_color", f32p), ("pts", f32p), ("mid_z_vals", f32p),
        ("weight_sum", f32p), ("color_fine", f32p), ("rays_o", f32p), ("light_dir", f32p), ("bg_color", f32p),
        ("image", f32p), ("image_no_bg", f32p), ("mask", f32p), ("shading_map", f32p), ("color_map", f32p),
        ("weight_sum_map", f32p), ("amb_shading_map", f32p), ("diff_shading_map", f32p), ("normal_map", f32p),
        ("no_specular_map", f32p), ("specular_map", f32p), ("z_map", f32p), ("z_min_per_ray", f32p),
        ("light_params", f32p),
    ]


MAP_NAMES = ("image", "image_no_bg", "mask", "shading_map", "color_map", "weight_sum_map", "amb_shading_map",
             "diff_shading_map", "normal_map", "no_specular_map", "specular_map", "z_map")


class OiRenderMapsBwdDesc(C.Structure):
    _fields_ = ([("fwd", OiRenderMapsDesc)] + [("g_" + n, f32p) for n in MAP_NAMES] +
                [("d_weights", f32p), ("d_gradients", f32p), ("d_raw_color", f32p), ("d_weight_sum", f32p),
                 ("d_color_fine", f32p), ("d_light_params", f32p), ("d_light_dir", f32p)])


EXPORTS = ["oi_packed_weights_bytes", "oi_pack_weights", "oi_style_mlp", "oi_render_workspace_bytes",
           "oi_render_forward", "oi_render_launch_count", "oi_upfirdn2d", "oi_bias_act", "oi_fused_bias_act",
           "oi_last_error", "oi_abi_version", "oi_build_info", "oi_selftest_tc", "oi_gen_rays", "oi_render_maps",
           "oi_render_backward_workspace_bytes", "oi_render_backward", "oi_selftest_wgrad",
           "oi_augment_geom_workspace_bytes", "oi_augment_geom_forward", "oi_augment_geom_backward",
           "oi_augment_geom_setup", "oi_render_maps_backward", "oi_augment_geom_setup_ops",
           "oi_augment_geom_setup_raw", "oi_render_backward_operand_format",
           "oi_render_backward_control_words", "oi_selftest_bwd_mode"]

_lib = None


def lib():
    """Loads (once) and returns the shared library.  Raises if it is not built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m object_intrinsics_b200.build` "
            "(or __graft_entry__.build()).  object_intrinsics_b200 has no CPU/eager fallback.")
    L = C.CDLL(LIB_PATH)
    L.oi_last_error.restype = C.c_char_p
    L.oi_build_info.restype = C.c_char_p
    L.oi_abi_version.restype = C.c_int
    L.oi_packed_weights_bytes.argtypes = [C.c_int32, C.POINTER(C.c_size_t)]
    L.oi_pack_weights.argtypes = [C.POINTER(OiNetParams), C.c_void_p, C.c_size_t, C.c_void_p]
    L.oi_style_mlp.argtypes = [C.POINTER(OiNetParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.oi_render_workspace_bytes.argtypes = [C.POINTER(OiRenderDesc), C.POINTER(C.c_size_t)]
    L.oi_render_forward.argtypes = [C.POINTER(OiRenderDesc), C.c_void_p]
    L.oi_render_launch_count.argtypes = [C.POINTER(OiRenderDesc), C.POINTER(C.c_int32)]
    L.oi_render_backward_workspace_bytes.argtypes = [C.POINTER(OiRenderBwdDesc), C.POINTER(C.c_size_t)]
    L.oi_render_backward.argtypes = [C.POINTER(OiRenderBwdDesc), C.c_void_p]
    L.oi_render_backward_operand_format.argtypes = [C.POINTER(OiRenderBwdDesc), C.POINTER(C.c_int32), C.c_void_p]
    L.oi_render_backward_control_words.argtypes = [C.POINTER(OiRenderBwdDesc), C.POINTER(C.c_uint32), C.c_void_p]
    L.oi_selftest_bwd_mode.argtypes = [C.POINTER(C.c_uint32), C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32)]
    L.oi_selftest_wgrad.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p]
    L.oi_augment_geom_workspace_bytes.argtypes = [C.POINTER(OiAugmentGeomDesc), C.POINTER(C.c_size_t)]
    L.oi_augment_geom_setup.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
    L.oi_augment_geom_setup_ops.argtypes = [C.POINTER(OiAugmentOp), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oi_augment_geom_setup_raw.argtypes = [C.POINTER(OiAugmentRawOp), C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oi_augment_geom_forward.argtypes = [C.POINTER(OiAugmentGeomDesc), C.c_void_p]
    L.oi_augment_geom_backward.argtypes = [C.POINTER(OiAugmentGeomDesc), C.c_void_p]
    L.oi_upfirdn2d.argtypes = [C.POINTER(OiUpfirdnDesc), C.c_void_p]
    L.oi_bias_act.argtypes = [C.POINTER(OiBiasActDesc), C.c_void_p]
    L.oi_fused_bias_act.argtypes = [C.POINTER(OiFusedBiasActDesc), C.c_void_p]
    L.oi_gen_rays.argtypes = [C.POINTER(OiGenRaysDesc), C.c_void_p]
    L.oi_render_maps.argtypes = [C.POINTER(OiRenderMapsDesc), C.c_void_p]
    L.oi_render_maps_backward.argtypes = [C.POINTER(OiRenderMapsBwdDesc), C.c_void_p]
    L.oi_selftest_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    if L.oi_abi_version() != 1:
        raise RuntimeError(f"liboi_b200.so ABI version {L.oi_abi_version()} != 1")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().oi_last_error().decode("utf-8", "replace")
        names = {-1: "OI_ERR_INVALID_ARGUMENT", -2: "OI_ERR_UNSUPPORTED", -3: "OI_ERR_CUDA", -4: "OI_ERR_WORKSPACE"}
        if rc == -2:
            raise NotImplementedError(f"{what}: {names.get(rc, rc)}: {msg}")
        raise RuntimeError(f"{what}: {names.get(rc, rc)}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}
