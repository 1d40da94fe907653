/*
 * oi_b200.h -- C-ABI of the B200-native (sm_100a) hot path of zzyunzhi/object-intrinsics.
 *
 * The reference has NO native interface for the renderer: its hot path is Python/torch
 * (src/third_party/neus/models/renderer.py:351-473 `NeuSRenderer.render`, :199-349 `render_core`,
 * :137-181 `up_sample`, :44-74 `sample_pdf`, :183-197 `cat_z_vals`; src/models/fields.py:49-122
 * ShapeNetwork / ColorNetwork / gradient; src/third_party/stylesdf/volume_renderer.py:27-61
 * LinearLayer / FiLMSiren).  The entry points below are what a binding for that path binds
 * (ctypes stub in INTEGRATION.md; the Python drop-in class is object_intrinsics_b200.renderer.NeuSRenderer).
 * The three StyleGAN2 ops DO have native interfaces in the reference (pybind11); the entry points here
 * replace them one for one and cite them.
 *
 * Conventions: plain pointers + sizes, all pointers are DEVICE pointers unless stated otherwise,
 * all tensors fp32 contiguous row-major unless stated otherwise.  Every call returns OI_OK (0) or a
 * negative OiStatus; the message is available through oi_last_error() (thread-local).  Nothing throws,
 * exits, allocates device memory or synchronises: kernels are launched on the caller's stream
 * (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream).  Re-entrant; no global
 * mutable state besides the thread-local error string and per-device cached function attributes.
 */
#ifndef OI_B200_H_
#define OI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OI_ABI_VERSION 1

typedef enum OiStatus {
  OI_OK = 0,
  OI_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, bad size, misaligned buffer */
  OI_ERR_UNSUPPORTED = -2,      /* configuration outside what the kernels implement */
  OI_ERR_CUDA = -3,             /* a CUDA runtime call failed; message holds cudaGetErrorString */
  OI_ERR_WORKSPACE = -4         /* workspace too small */
} OiStatus;

#define OI_MAX_DEPTH 8 /* FiLM-SIREN SDF layers (configs/train.yaml:41 D: 8) */
#define OI_WIDTH 128   /* hidden width            (configs/train.yaml:42 W: 128) */
#define OI_STYLE_DIM 64

/* Which MLP core evaluates the FiLM-SIREN contraction. */
typedef enum OiRenderImpl {
  OI_IMPL_AUTO = 0,
  OI_IMPL_FFMA = 1,   /* FP32 FFMA register-tiled contraction (exact fp32 products)          */
  OI_IMPL_TCGEN05 = 2 /* tcgen05.mma kind::f16 with a 2-term fp16 split (3 MMAs per product) */
} OiRenderImpl;

/* ------------------------------------------------------------------------------------------------
 * Raw network parameters, PyTorch layouts ([out, in] row-major), as registered by the reference:
 *   pts_*      : ShapeNetwork.pts_linears[l]  = FiLMSiren  (stylesdf/volume_renderer.py:33-48)
 *   sigma_*    : ShapeNetwork.sigma_linear    = LinearLayer(W,1)   (volume_renderer.py:83)
 *   views_*    : ColorNetwork.views_linears   = FiLMSiren(W+3, W)  (volume_renderer.py:80-81)
 *   rgb_*      : ColorNetwork.rgb_linear      = LinearLayer(W,3)   (volume_renderer.py:82)
 *   variance   : SingleVarianceNetwork.variance (neus/models/fields.py:262-268)
 *   style_*    : ShapeNetwork.style[i]        = MappingLinear(64,64) (src/models/fields.py:14-19)
 * FiLM index 0..depth-1 = pts_linears, index OI_MAX_DEPTH = views_linears.
 * ---------------------------------------------------------------------------------------------- */
typedef struct OiNetParams {
  int32_t depth;                            /* D, 1..OI_MAX_DEPTH */
  int32_t width;                            /* must be OI_WIDTH */
  int32_t style_dim;                        /* must be OI_STYLE_DIM */
  int32_t reserved;
  const float* pts_weight[OI_MAX_DEPTH];    /* [W,3] for l=0, [W,W] otherwise */
  const float* pts_bias[OI_MAX_DEPTH];      /* [W] */
  const float* gamma_weight[OI_MAX_DEPTH + 1]; /* [W,style] */
  const float* gamma_bias[OI_MAX_DEPTH + 1];   /* [W] */
  const float* beta_weight[OI_MAX_DEPTH + 1];  /* [W,style] */
  const float* beta_bias[OI_MAX_DEPTH + 1];    /* [W] */
  const float* sigma_weight;                /* [1,W] */
  const float* sigma_bias;                  /* [1] */
  const float* views_weight;                /* [W, W+3] */
  const float* views_bias;                  /* [W] */
  const float* rgb_weight;                  /* [3,W] */
  const float* rgb_bias;                    /* [3] */
  const float* variance;                    /* [] */
  const float* style_weight[3];             /* [style,style] (may be NULL if oi_style_mlp is not used) */
  const float* style_bias[3];               /* [style] */
} OiNetParams;

/* Size in bytes of the packed weight blob for a network of the given depth. */
int oi_packed_weights_bytes(int32_t depth, size_t* bytes);

/* Re-lays the parameters into the streaming order / operand formats the render kernels consume
 * (transposed fp32 panels for the FFMA core, swizzled fp16 hi/lo UMMA panels for the tcgen05 core,
 * folded constants).  Call again whenever a parameter changes (after an optimiser step).
 * `blob` must be 128-byte aligned and at least oi_packed_weights_bytes(depth) long. */
int oi_pack_weights(const OiNetParams* params, void* blob, size_t blob_bytes, void* stream);

/* ShapeNetwork.style: w = f(f(f(z))), f(x) = leaky_relu(x W^T + b, 0.2) * 1
 * (src/models/fields.py:14-19; stylesdf/model.py:49-56).  z, w: [n_instances, 64]. */
int oi_style_mlp(const OiNetParams* params, const float* z, float* w, int32_t n_instances, void* stream);

/* ------------------------------------------------------------------------------------------------
 * NeuSRenderer.render (renderer.py:351-473) for the live configuration of the reference
 * (nerf=None, n_outside=0, siren_network=None, second_order=None, compute_color=True,
 * compute_sample_dist=False, blend_background=False, background_rgb=None).
 * R = n_rays, n = n_samples, m = n_importance, S = n + m.
 * ---------------------------------------------------------------------------------------------- */
struct OiRenderMapsDesc;
typedef struct OiRenderDesc {
  /* sizes */
  int32_t n_rays;            /* R; rays of instance b occupy [b*rays_per_instance, (b+1)*rays_per_instance) */
  int32_t rays_per_instance; /* R / n_instances (src/models/fields.py:55) */
  int32_t n_samples;         /* n >= 2 */
  int32_t n_importance;      /* m >= 0 */
  int32_t up_sample_steps;   /* >= 1, must divide n_importance; step i adds n_importance / steps samples with
                              * inv_s = 64 * 2^i (renderer.py:400-413); `lin_fine` then has n_importance / steps entries */
  int32_t depth;             /* D of the packed network */
  int32_t impl;              /* OiRenderImpl */
  int32_t flags;             /* bit 0 (OI_FLAG_DISCARD_SCRATCH): drop dead reverse-sweep scratch lines from L2 (discard.global.L2);
                              * bit 3 (value 8): keep the per-ray compositing in its own kernel even where the tcgen05 core
                              * could do it (rays aligned with its 128-point tiles: 128 % (n_samples + n_importance) == 0);
                              * bit 4 (value 16): with `maps`, composite the maps in the tile tail of the render kernel too
                              * (nothing per-point is written at all; slower than the default, see below) */
  float cos_anneal_ratio;    /* renderer.py:273-274 */
  float reserved_f;

  /* inputs */
  const float* rays_o;       /* [R,3] */
  const float* rays_d;       /* [R,3] */
  const float* near;         /* [R,1] */
  const float* far;          /* [R,1] */
  const float* t_rand;       /* [R,1] = U[0,1)-0.5 per ray (renderer.py:371-373) or NULL for no jitter */
  const float* lin_coarse;   /* [n]  = linspace(0,1,n) in fp32 (renderer.py:359); NULL = computed in-kernel */
  const float* lin_fine;     /* [m]  = linspace(.5/m, 1-.5/m, m) (renderer.py:53); NULL = computed in-kernel */
  const float* z_vals_in;    /* optional [R,S]: skip sampling, render these z (matched-z parity tests) or NULL */
  const float* style_w;      /* [n_instances, 64] latent w (generator.py:237-238) */
  const void* packed_weights;/* blob written by oi_pack_weights */

  /* outputs (renderer.py:448-468); any pointer may be NULL to skip that tensor except `weights`
   * (it carries alpha between the two phases; optional too when `maps` is given) */
  float* s_val;          /* [R,1] */
  float* cdf_fine;       /* [R,S] */
  float* weight_sum;     /* [R,1] */
  float* weight_max;     /* [R,1] */
  float* gradients;      /* [R,S,3] */
  float* weights;        /* [R,S] */
  float* gradient_error; /* [] */
  float* inside_sphere;  /* [R,S] */
  float* mid_z_vals;     /* [R,S] */
  float* surface_loss;   /* [] */
  float* sdf;            /* [R,S] */
  float* pts_norm;       /* [R,S] */
  float* pts;            /* [R,S,3] */
  float* color_fine;     /* [R,3] */
  float* raw_color;      /* [R,S,3] */
  float* z_vals_out;     /* optional [R,S]: the (sorted) section start z-values actually rendered */

  void* workspace;       /* >= oi_render_workspace_bytes(desc), 256-byte aligned */
  size_t workspace_bytes;

  /* optional cudaEvent_t handles recorded on `stream` immediately before / after the launch of the fused
   * render_core kernel (the dominant kernel), so that callers can time it in isolation; NULL = off */
  void* evt_core_start;
  void* evt_core_stop;

  /* Contract B: when non-NULL the call also produces the shading / compositing maps of
   * Generator.render_maps (generator.py:80-174) for the rays it renders.  Of *maps the light (light_dir, bg_color,
   * light_params or the by-value colours), rays_per_instance, the output-map pointers and z_min_per_ray are used; its
   * input pointers are ignored (the inputs are this render's own results).  Every per-point output above, `weights`
   * included, may then be NULL: the per-point tensors the caller does not ask for live in the workspace only (L2 for
   * typical sizes) and oi_render_maps' kernel runs after the compositing, inside this call.  With flags bit 4, the
   * tcgen05 core and 128 % (n + m) == 0 the maps are composited in the tail of the render kernel instead and nothing
   * per-point is written (measured slower: 2.04 vs 1.78 ms at 16 384 rays x 64 -- kept as the contract-B variant). */
  const struct OiRenderMapsDesc* maps;
} OiRenderDesc;

int oi_render_workspace_bytes(const OiRenderDesc* desc, size_t* bytes);
int oi_render_forward(const OiRenderDesc* desc, void* stream);

/* Number of kernel launches one oi_render_forward(desc) performs (for bench bookkeeping). */
int oi_render_launch_count(const OiRenderDesc* desc, int32_t* launches);

/* ------------------------------------------------------------------------------------------------
 * Backward of NeuSRenderer.render w.r.t. every parameter the path reads.  The reference gets it from
 * torch.autograd (render #1 of each training step, src/trainers/gan_pose_trainer.py:110,141), including the
 * second-order terms through the SDF normal that `fields.py:104-122` builds with create_graph=True.  Here it
 * is one reverse sweep per 128-point tile that recomputes the forward (nothing but the render outputs is
 * kept between forward and backward).  Sampling is not differentiated (renderer.py:390 runs it under
 * no_grad): pass the z-values the forward rendered (OiRenderDesc.z_vals_out).  Rays are treated as constants.
 *
 * Gradients are ACCUMULATED (+=) into the OiNetGrads tensors, which have the shapes of the corresponding
 * OiNetParams tensors; the caller zeroes them when it wants plain gradients.  The FiLM linears
 * (gamma/beta weight and bias) and the latent w receive their gradient through the FiLM tables:
 * d_film_gamma / d_film_beta are dL/dgamma, dL/dbeta with gamma = 15 (G w + g) + 30, beta = 0.25 (B w + c)
 * (volume_renderer.py:27-30); the remaining chain is two tiny matrix products left to the caller.
 * Summation over points uses floating-point atomics: results are reproducible to rounding, not bit-wise.
 * ---------------------------------------------------------------------------------------------- */
typedef struct OiNetGrads {
  float* pts_weight[OI_MAX_DEPTH];  /* [W,3] for l=0, [W,W] otherwise */
  float* pts_bias[OI_MAX_DEPTH];    /* [W] */
  float* sigma_weight;              /* [1,W] */
  float* sigma_bias;                /* [1] */
  float* views_weight;              /* [W, W+3] */
  float* views_bias;                /* [W] */
  float* rgb_weight;                /* [3,W] */
  float* rgb_bias;                  /* [3] */
  float* variance;                  /* [] */
  float* film_gamma;                /* [n_instances, OI_MAX_DEPTH+1, W] (slot OI_MAX_DEPTH = views_linears) */
  float* film_beta;                 /* [n_instances, OI_MAX_DEPTH+1, W] */
} OiNetGrads;

typedef struct OiRenderBwdDesc {
  int32_t n_rays;            /* R */
  int32_t rays_per_instance;
  int32_t n_samples_total;   /* S = n_samples + n_importance */
  int32_t n_samples;         /* n (sample_dist = 2/n, renderer.py:356) */
  int32_t depth;             /* D >= 2 */
  int32_t impl;              /* OiRenderImpl: FFMA = one FP32 kernel; TCGEN05 (= AUTO) = two tensor-core kernels */
  int32_t flags;             /* tensor-core backward, format of the per-point operands of the weight-gradient contraction:
                              * default = chosen per call on the device from the adjoints' dynamic range (scaled fp16 when
                              * the points more than 2^18 below the largest adjoint carry < 2^-12 of the adjoint mass, else
                              * TF32); bit 5 (value 32): always TF32; bit 6 (value 64): always scaled fp16.
                              * Overflow guard (one word of the workspace, which the caller zeroes ONCE after
                              * allocating it): the first call on a workspace keeps TF32 and samples what the fp16
                              * operands would be; fp16 is used from the next call on if none reached an eighth of fp16's
                              * largest number, and is switched off for the life of the workspace as soon as an operand
                              * does -- 8x below saturation */
  int32_t reserved;
  float cos_anneal_ratio;
  float reserved_f;

  /* inputs of the forward call */
  const float* rays_o;       /* [R,3] */
  const float* rays_d;       /* [R,3] */
  const float* z_vals;       /* [R,S] section starts rendered by the forward (z_vals_out) */
  const float* style_w;      /* [n_instances,64] */
  const void* packed_weights;

  /* outputs of the forward call that the tail needs */
  const float* sdf;          /* [R,S] */
  const float* gradients;    /* [R,S,3] */
  const float* raw_color;    /* [R,S,3] */

  /* incoming adjoints dL/d(output); each may be NULL (= zero) */
  const float* g_weights;        /* [R,S] */
  const float* g_weight_sum;     /* [R,1] */
  const float* g_weight_max;     /* [R,1] */
  const float* g_color_fine;     /* [R,3] */
  const float* g_raw_color;      /* [R,S,3] */
  const float* g_gradients;      /* [R,S,3] */
  const float* g_sdf;            /* [R,S] */
  const float* g_cdf_fine;       /* [R,S] */
  const float* g_s_val;          /* [R,1] */
  const float* g_gradient_error; /* [] */
  const float* g_surface_loss;   /* [] */

  OiNetGrads grads;          /* every pointer must be non-NULL */

  void* workspace;           /* >= oi_render_backward_workspace_bytes(desc), 256-byte aligned */
  size_t workspace_bytes;
  void* evt_core_start;      /* optional cudaEvent_t pair around the MLP backward kernel */
  void* evt_core_stop;
} OiRenderBwdDesc;

int oi_render_backward_workspace_bytes(const OiRenderBwdDesc* desc, size_t* bytes);
int oi_render_backward(const OiRenderBwdDesc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * StyleGAN2 ops.
 * ---------------------------------------------------------------------------------------------- */

/* Replaces `_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)`
 * (src/third_party/ada/torch_utils/ops/upfirdn2d.cpp:16-94; kernels upfirdn2d.cu:29-341) and, with
 * NHWC-style strides, `upfirdn2d_op.upfirdn2d(input[N,H,W,1], kernel, ...)`
 * (src/third_party/stylesdf/op/upfirdn2d.cpp:12-22; upfirdn2d_kernel.cu:49-369).
 * dtype: 0 = fp32, 1 = fp16, 2 = fp64 (x and y; the filter is always fp32; accumulation fp32, fp64 for fp64).
 * Strides are in elements.  Output size must be
 *   out_w = (in_w*up_x + pad_x0 + pad_x1 - filter_w + down_x) / down_x   (upfirdn2d.cpp:33-34). */
typedef struct OiUpfirdnDesc {
  const void* x;
  const float* f;      /* [filter_h, filter_w] with strides f_stride_{h,w} */
  void* y;
  int32_t dtype;
  int32_t batch, channels, in_h, in_w;
  int64_t x_stride_n, x_stride_c, x_stride_h, x_stride_w;
  int32_t out_h, out_w;
  int64_t y_stride_n, y_stride_c, y_stride_h, y_stride_w;
  int32_t filter_h, filter_w;
  int64_t f_stride_h, f_stride_w;
  int32_t up_x, up_y, down_x, down_y;
  int32_t pad_x0, pad_x1, pad_y0, pad_y1;
  int32_t flip;        /* 0: convolution (filter flipped), 1: correlation */
  float gain;
} OiUpfirdnDesc;
int oi_upfirdn2d(const OiUpfirdnDesc* desc, void* stream);

/* Replaces `_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)`
 * (src/third_party/ada/torch_utils/ops/bias_act.cpp:32-90; kernel bias_act.cu:23-147).
 * act: 1 linear, 2 relu, 3 lrelu, 4 tanh, 5 sigmoid, 6 elu, 7 selu, 8 softplus, 9 swish.
 * grad: 0 forward, 1 first-order, 2 second-order.  b/xref/yref/dy may be NULL.
 * Elementwise over `size_x` dense elements; bias index = (i / step_b) % size_b. */
typedef struct OiBiasActDesc {
  const void* x;
  const void* b;
  const void* xref;
  const void* yref;
  const void* dy;
  void* y;
  int32_t dtype;       /* 0 = fp32, 1 = fp16, 2 = fp64 */
  int32_t grad;
  int32_t act;
  int32_t reserved;
  float alpha, gain, clamp; /* clamp < 0 disables */
  int32_t size_x, size_b, step_b;
} OiBiasActDesc;
int oi_bias_act(const OiBiasActDesc* desc, void* stream);

/* Replaces `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
 * (src/third_party/stylesdf/op/fused_bias_act.cpp:11-20; kernel fused_bias_act_kernel.cu:19-98).
 * act*10+grad: 10/11 linear, 12 -> 0, 30 lrelu fwd, 31 lrelu grad (sign taken from `ref`), 32 -> 0.
 * bias/ref may be NULL.  bias index = (i / step_b) % size_b. */
typedef struct OiFusedBiasActDesc {
  const void* x;
  const void* bias;
  const void* ref;
  void* y;
  int32_t dtype;       /* 0 = fp32, 1 = fp16, 2 = fp64 */
  int32_t act, grad;
  int32_t size_x, size_b, step_b;
  float alpha, scale;
} OiFusedBiasActDesc;
int oi_fused_bias_act(const OiFusedBiasActDesc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Callers either side of the render path.
 * ---------------------------------------------------------------------------------------------- */

/* Generator.gen_rays_at + build_rays (+ near_far_from_sphere when near/far are given)
 * (src/models/generator.py:255-279, 317-333, 336-342).  4x4 matrices are row-major. */
typedef struct OiGenRaysDesc {
  int32_t n_instances, resolution, scene_resolution, reserved;
  float cam_dist, reserved_f;
  const float* b2w;            /* [bs,4,4] box -> world */
  const float* c2b;            /* [bs,4,4] camera -> box (= w2b @ c2w, generator.py:73) */
  const float* w2c;            /* [4,4] camera.w2c */
  const float* intrinsics_inv; /* [4,4] camera.intrinsics_inv */
  float* rays_o;               /* [bs,P,P,3] (materialised; the reference returns a stride-0 expand) */
  float* rays_d;               /* [bs,P,P,3] */
  float* x_offset;             /* [bs] or NULL */
  float* y_offset;             /* [bs] or NULL */
  float* near;                 /* [bs*P*P,1] or NULL */
  float* far;                  /* [bs*P*P,1] or NULL */
} OiGenRaysDesc;
int oi_gen_rays(const OiGenRaysDesc* desc, void* stream);

/* Generator.render_maps with the directional Phong light (src/models/generator.py:80-174;
 * src/models/lighting.py:61-76,94-119,126-225).  Inputs are the renderer's outputs; every output map is
 * [bs, C, P, P] (C = 3 or 1) and may be NULL. */
typedef struct OiRenderMapsDesc {
  int32_t n_rays, rays_per_instance, n_samples, reserved;
  float shininess, reserved_f;
  float ambient_color[3], diffuse_color[3], specular_color[3], pad;
  const float* weights;     /* [R,S] */
  const float* gradients;   /* [R,S,3] */
  const float* raw_color;   /* [R,S,3] */
  const float* pts;         /* [R,S,3] */
  const float* mid_z_vals;  /* [R,S] or NULL (z_map / z_min) */
  const float* weight_sum;  /* [R,1] */
  const float* color_fine;  /* [R,3] */
  const float* rays_o;      /* [R,3] camera position per ray (box frame) */
  const float* light_dir;   /* [bs,3] light direction in each box frame (lighting.py:115-119) */
  const float* bg_color;    /* [bs,3] */
  float *image, *image_no_bg, *mask, *shading_map, *color_map, *weight_sum_map;
  float *amb_shading_map, *diff_shading_map, *normal_map, *no_specular_map, *specular_map, *z_map;
  float* z_min_per_ray;     /* [R] min_s mid_z (reduced per instance by the caller) or NULL */
  const float* light_params; /* device [10] = ambient[3], diffuse[3], specular[3], shininess, or NULL; when given it
                              * replaces the four by-value fields above (the light's nn.Parameters are then never read
                              * back to the host: lighting.py:13-52 keeps them on the device) */
} OiRenderMapsDesc;
int oi_render_maps(const OiRenderMapsDesc* desc, void* stream);

/* Reverse mode of oi_render_maps: what autograd does for generator.py:80-174 + lighting.py:126-225 in the generator
 * step (gan_pose_trainer.py:141).  `fwd` repeats the forward call's inputs (its output pointers are ignored);
 * g_* are the adjoints of the maps ([bs,C,P,P], NULL = no gradient flows through that map).  Written (not
 * accumulated): d_weights [R,S], d_gradients [R,S,3], d_raw_color [R,S,3], d_weight_sum [R,1], d_color_fine [R,3];
 * d_light_params [10] (same order as light_params) and d_light_dir [bs,3] are zero-filled by the call and reduced
 * over all sample points.  pts, rays_o, mid_z_vals and bg_color are constants of the training path. */
typedef struct OiRenderMapsBwdDesc {
  OiRenderMapsDesc fwd;
  const float *g_image, *g_image_no_bg, *g_mask, *g_shading_map, *g_color_map, *g_weight_sum_map;
  const float *g_amb_shading_map, *g_diff_shading_map, *g_normal_map, *g_no_specular_map, *g_specular_map, *g_z_map;
  float *d_weights, *d_gradients, *d_raw_color, *d_weight_sum, *d_color_fine;
  float* d_light_params;    /* [10] or NULL */
  float* d_light_dir;       /* [bs,3] or NULL */
} OiRenderMapsBwdDesc;
int oi_render_maps_backward(const OiRenderMapsBwdDesc* desc, void* stream);

/* Geometric path of the ADA AugmentPipe (src/third_party/ada/augment.py:270-301): reflect-pad by `margins`,
 * 2x up-sample with the separable low-pass `filter` (gain 4), bilinear affine resample (affine_grid + grid_sample,
 * align_corners = False, zeros outside) onto a [(H + 2 hz_pad) * 2, (W + 2 hz_pad) * 2] grid, 2x down-sample
 * (correlation) with crop; hz_pad = filter_taps / 4.  `theta` are the final affine_grid matrices of augment.py:297
 * and `margins` the integers of augment.py:283, both DEVICE tensors (the reference reads the margins back to the
 * host; here nothing synchronises).  The map x -> y is linear: oi_augment_geom_backward applies its adjoint
 * (x = dL/dy in, y = dL/dx out), and the double backward of the R1 penalty is oi_augment_geom_forward again. */
typedef struct OiAugmentGeomDesc {
  int32_t batch, channels, height, width;
  int32_t filter_taps;    /* even, <= 16 (12 = sym6) */
  int32_t reserved;
  const float* filter;    /* HOST pointer: [filter_taps] taps (normalised to unit DC by the caller) */
  const float* theta;     /* [batch,2,3] */
  const int32_t* margins; /* [4] = mx0, my0, mx1, my1, each in [0, size-1] */
  const float* x;         /* [batch,channels,height,width] */
  float* y;               /* [batch,channels,height,width] */
  void* workspace;        /* >= oi_augment_geom_workspace_bytes(desc) */
  size_t workspace_bytes;
} OiAugmentGeomDesc;
/* Margins [4] (augment.py:274-283) and affine_grid matrices theta [batch,2,3] (augment.py:287-297) from the inverse
 * transforms g_inv [batch,3,3] (pixel_out -> pixel_in, centred pixel coordinates); everything on the device. */
int oi_augment_geom_setup(const float* g_inv, int32_t batch, int32_t height, int32_t width, int32_t filter_taps,
                          float* theta, int32_t* margins, void* stream);
/* The same, with the inverse transform given as the sequence of elementary 3x3 factors the reference composes with
 * ~15 torch launches each (augment.py:196-264: G_inv = G_inv @ scale2d_inv(..) @ rotate2d_inv(..) @ translate2d_inv(..)
 * ...): G_inv[b] = prod_i M_i[b] in the order given.  kind 0: scale2d(p0[b], p1[b]); 1: rotate2d(p0[b]);
 * 2: translate2d(p0[b], p1[b]); the per-sample parameters are the caller's (already gated) random draws.
 * g_inv [batch,3,3] receives the composed transform (required). */
typedef struct OiAugmentOp {
  int32_t kind, reserved;
  const float* p0;   /* [batch] */
  const float* p1;   /* [batch] or NULL (rotate) */
} OiAugmentOp;
#define OI_AUGMENT_MAX_OPS 8
int oi_augment_geom_setup_ops(const OiAugmentOp* ops, int32_t n_ops, int32_t batch, int32_t height, int32_t width,
                              int32_t filter_taps, float* g_inv, float* theta, int32_t* margins, void* stream);
/* The same from the RAW random draws: the kernel also applies the reference's gating and parameter arithmetic
 * (augment.py:196-264) -- `torch.where(gate < prob * p, value(draw), identity)` per factor -- so that the caller's
 * only work per call is the draws themselves (one torch.rand / randn each, in the reference's order: the RNG stream
 * is the reference's).  form: the reference's factor
 *   OI_AUG_XFLIP    i = floor(draw*2), gated                          scale2d(1/(1-2i), 1)          augment.py:196-201
 *   OI_AUG_ROTATE90 i = floor(draw*4), gated                          rotate2d(pi/2 * i)            :203-208
 *   OI_AUG_XINT     t = (draw[b,0:2]*2-1)*param, gated                translate2d(-round(t*size))   :210-215
 *   OI_AUG_SCALE    s = exp2(draw*param), gated (else 1)              scale2d(1/s, 1/s)             :220-224
 *   OI_AUG_ROTATE   th = (draw*2-1)*pi*param, gate < 1-sqrt(1-prob*p) rotate2d(th)                  :226-231,241-245
 *   OI_AUG_ANISO    s = exp2(draw*param), gated (else 1)              scale2d(1/s, s)               :234-238
 *   OI_AUG_XFRAC    t = draw[b,0:2]*param, gated                      translate2d(-t*size)          :248-252
 * draw: [batch] ([batch,2] for XINT / XFRAC), gate: [batch], p: the module's device scalar `p`. */
enum { OI_AUG_XFLIP = 0, OI_AUG_ROTATE90 = 1, OI_AUG_XINT = 2, OI_AUG_SCALE = 3, OI_AUG_ROTATE = 4, OI_AUG_ANISO = 5,
       OI_AUG_XFRAC = 6 };
typedef struct OiAugmentRawOp {
  int32_t form, reserved;
  const float* draw;
  const float* gate;
  float prob;    /* constructor probability multiplier of the factor (xflip, rotate90, xint, scale, rotate, aniso, xfrac) */
  float param;   /* xint_max | scale_std | rotate_max | aniso_std | xfrac_std (unused: XFLIP, ROTATE90) */
} OiAugmentRawOp;
int oi_augment_geom_setup_raw(const OiAugmentRawOp* ops, int32_t n_ops, const float* p, int32_t batch, int32_t height,
                              int32_t width, int32_t filter_taps, float* g_inv, float* theta, int32_t* margins,
                              void* stream);
int oi_augment_geom_workspace_bytes(const OiAugmentGeomDesc* desc, size_t* bytes);
int oi_augment_geom_forward(const OiAugmentGeomDesc* desc, void* stream);
int oi_augment_geom_backward(const OiAugmentGeomDesc* desc, void* stream);

/* Which operand format the LAST oi_render_backward on desc->workspace used for the weight-gradient contraction
 * (see OiRenderBwdDesc.flags): *format = 0 TF32, 1 scaled fp16.  Diagnostic: synchronises `stream`. */
int oi_render_backward_operand_format(const OiRenderBwdDesc* desc, int32_t* format, void* stream);
/* The 12 control words the last oi_render_backward on desc->workspace left behind (diagnostic, synchronises `stream`):
 * [0] points inside the relaxed sphere (gradient_error), [1] float bits of the largest |adjoint component|,
 * [2,3] u64 adjoint mass of all points / [4,5] of the points more than 2^18 below the maximum (2^-20 fixed point,
 * relative to the maximum's binade; zero when the format is forced), [6,7] float bits of the largest forward-type /
 * adjoint-type fp16 operand written (builds with -DOI_BWD_RANGE_STATS=1 only, else 0), [8] state of the workspace's fp16
 * overflow guard: 0 unknown (next call probes), 1 safe, 2 unsafe (see OiRenderBwdDesc.flags), [9..11] 0. */
int oi_render_backward_control_words(const OiRenderBwdDesc* desc, uint32_t* words, void* stream);

/* The device-side rule that picks the operand format of a backward call, evaluated on the host (no GPU needed) for the
 * given control words (layout above; only [1]..[5] are read), desc flags and guard state (0 unknown, 1 safe, 2
 * unsafe): *fp16 = 1 when the scaled fp16 operands would be used, *e_ref = exponent the forward-type operands are
 * referred to. */
int oi_selftest_bwd_mode(const uint32_t* control_words, int32_t flags, int32_t guard_state, int32_t* fp16,
                         int32_t* e_ref);

/* Self-test of the tcgen05 building blocks: d[128,128] = a[128,128] * B^T through the split-fp16 UMMA path.
 * B = b[128,128] ([n][k] row-major) when packed_weights is NULL, else panel `panel` of the packed blob
 * (order: forward l=1..D-1, colour features, reverse l=D-1..1; each is 2^8 * W in [n][k] orientation). */
int oi_selftest_tc(const float* a, const float* b, const void* packed_weights, int32_t depth, int32_t panel,
                   float* d, void* stream);

/* Self-test of the point-contraction kernel of the tensor-core backward (csrc/oi_wgrad_tc.cu):
 *   d[i][j] += sum_m X[m][i] * Y[m][j]                      (TF32 tensor-core products, fp32 accumulation)
 *   col[inst][i][c] += sum_{m in instance} X[m][i] * aux[m][c],  c < 4
 * over n_tiles tiles of 128 points.  slabs: [n_tiles][slabs_per_tile][4 blocks][128 channels][32 points] fp32 with
 * the 16-byte chunks of every 128-byte row XOR-permuted by (channel & 7) (values should be TF32-representable);
 * aux: [n_tiles][4 blocks][4 columns][32 points] fp32, same permutation.  d [128,128] and
 * col [n_instances,128,4] are accumulated into (zero them first).  The work is split over n_ctas CTAs. */
int oi_selftest_wgrad(const float* slabs, const float* aux, int32_t n_tiles, int32_t slabs_per_tile, int32_t x_slab,
                      int32_t y_slab, int32_t tiles_per_instance, int32_t n_ctas, float* d, float* col, void* stream);

/* ------------------------------------------------------------------------------------------------ */
const char* oi_last_error(void);
int oi_abi_version(void);
/* Compile-time facts about the library: "sm_100a;ffma;tcgen05" */
const char* oi_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* OI_B200_H_ */
