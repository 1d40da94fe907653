// Internal declarations shared by the translation units of liboi_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "oi_b200.h"

namespace oi {

// ------------------------------------------------------------------------------------------------
// error plumbing (oi_api.cu)
// ------------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
#define OI_CHECK_ARG(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return ::oi::set_error(OI_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)
#define OI_CHECK_CUDA(expr)                                                                       \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return ::oi::set_error(OI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                 \
  } while (0)

// ------------------------------------------------------------------------------------------------
// packed weight blob layout (all offsets in floats; written by pack_weights_kernel, oi_render_aux.cu)
//
//   [stream]  the FFMA core consumes the weights as a linear stream of 8 KB "chunks" = 16 k-rows x 128
//             n-columns fp32, in exactly the order the kernel needs them:
//               chunk 0                : layer 0,  rows k=0..2 = W_0^T (k = xyz), rows 3..15 zero
//               8 chunks per l=1..D-1  : W_l^T  (row k = input channel, column n = output channel)
//               8 chunks               : views_linears.weight[:, :128]^T (feature part of the colour layer)
//               8 chunks per l=D-1..1  : W_l as stored by torch (row k = output channel, column n = input
//                                        channel) -- the operand of the reverse (input-gradient) sweep
//             (the forward kernels stop here: n_chunks_fine)
//               8 chunks               : views_linears.weight[:, :128] as stored (row k = colour channel,
//                                        column n = feature channel) -- backward kernel only
//   [const]   bias[9][128] (slot 8 = views_linears.bias), w_sigma[128], wc_grad[3][128]
//             (= views_linears.weight[:, 128+j]), w_rgb[3][128], w0t[3][128] (= W_0^T), scalars[8] =
//             {b_sigma, b_rgb[0..2], inv_s, 1/inv_s, 0, 0}
//   [film]    gamma_w[9][128][64], gamma_b[9][128], beta_w[9][128][64], beta_b[9][128]
//   [tc]      fp16 hi/lo UMMA operand panels for the tcgen05 core (see oi_render_tc.cu)
//   [tcb]     bf16 hi/lo UMMA operand panels (unscaled) for the adjoint sweeps of the tcgen05 backward
//             (oi_render_bwd_tc.cu): colour-as-stored, forward orientation l=1..D-1, reverse orientation l=D-1..1
// ------------------------------------------------------------------------------------------------
constexpr int kW = OI_WIDTH;
constexpr int kStyle = OI_STYLE_DIM;
constexpr int kFilm = OI_MAX_DEPTH + 1;  // 9 FiLM layers: 8 SDF + colour
constexpr int kKC = 16;                  // k-rows per streamed chunk
constexpr int kChunkFloats = kKC * kW;   // 2048
constexpr int kChunkBytes = kChunkFloats * 4;

struct BlobLayout {
  int depth;
  int n_chunks_fine;    // 9 + 16 (D-1)
  int n_chunks_coarse;  // 1 + 8 (D-1)
  int n_chunks_stream;  // n_chunks_fine + 8 (chunks only the backward kernel streams)
  size_t stream_off, const_off, film_off, tc_off, tcb_off, total_floats;
  // const section sub-offsets (relative to const_off)
  static constexpr int kBias = 0;                      // [9][128]
  static constexpr int kWsig = kFilm * kW;             // [128]
  static constexpr int kWcg = kWsig + kW;              // [3][128]
  static constexpr int kWrgb = kWcg + 3 * kW;          // [3][128]
  static constexpr int kW0t = kWrgb + 3 * kW;          // [3][128]
  static constexpr int kScalars = kW0t + 3 * kW;       // [8]
  static constexpr int kConstFloats = kScalars + 8;
  // film section sub-offsets (relative to film_off)
  static constexpr int kGammaW = 0;
  static constexpr int kGammaB = kFilm * kW * kStyle;
  static constexpr int kBetaW = kGammaB + kFilm * kW;
  static constexpr int kBetaB = kBetaW + kFilm * kW * kStyle;
  static constexpr int kFilmFloats = kBetaB + kFilm * kW;
};

__host__ __device__ inline size_t tc_section_floats(int depth);

__host__ __device__ inline BlobLayout blob_layout(int depth) {
  BlobLayout L;
  L.depth = depth;
  L.n_chunks_fine = 9 + 16 * (depth - 1);
  L.n_chunks_coarse = 1 + 8 * (depth - 1);
  L.stream_off = 0;
  L.n_chunks_stream = L.n_chunks_fine + 8;
  L.const_off = (size_t)L.n_chunks_stream * kChunkFloats;
  L.film_off = L.const_off + ((BlobLayout::kConstFloats + 31) / 32) * 32;
  L.tc_off = L.film_off + ((BlobLayout::kFilmFloats + 31) / 32) * 32;
  L.tcb_off = L.tc_off + tc_section_floats(depth);
  L.total_floats = L.tcb_off + tc_section_floats(depth);
  return L;
}

// tcgen05 section: per MMA layer one 64 KB panel = {hi, lo} x [128 n][128 k] fp16 in the canonical
// K-major SWIZZLE_128B shared-memory image (2 k-blocks of 64 per operand).  Layer order:
// fwd l=1..D-1, colour-feature, reverse l=D-1..1  -> (2(D-1)+1) panels of 16384 floats.
__host__ __device__ inline size_t tc_section_floats(int depth) { return (size_t)(2 * (depth - 1) + 1) * 16384; }

// ------------------------------------------------------------------------------------------------
// kernel argument blocks
// ------------------------------------------------------------------------------------------------
// Shading maps composited in the tail of the tcgen05 core (contract B): the parts of OiRenderMapsDesc the kernel needs.
struct MapsKArgs {
  int enabled, rays_per_image;    // rays_per_image = P*P (one image per instance)
  const float* light_dir;         // [n_inst,3]
  const float* bg_color;          // [n_inst,3]
  const float* light_params;      // device [10] or NULL -> lp[]
  float lp[10];                   // ambient[3], diffuse[3], specular[3], shininess
  float *image, *image_no_bg, *mask, *shading_map, *color_map, *weight_sum_map;
  float *amb_shading_map, *diff_shading_map, *normal_map, *no_specular_map, *specular_map, *z_map;
  float* z_min_per_ray;
};

struct RenderKArgs {
  int R, rays_per_inst, n_inst;
  int S;         // samples per ray seen by this launch (n for the coarse pass, n+m for the fine pass)
  int n_coarse;  // n (for sample_dist and the in-kernel linspace)
  int D;
  int pts_per_inst, tiles_per_inst, n_tiles;
  int coarse;    // 1: SDF only at z (renderer.py:389-399), 0: full render_core at the section midpoints
  int flags;     // OiRenderDesc.flags (bit 0: drop dead scratch lines from L2 with discard.global.L2)
  float cos_anneal, sample_dist;
  const float *rays_o, *rays_d, *near, *far, *t_rand, *lin;
  const float* z_vals;  // [R,S] section starts or NULL (computed from near/far/lin/t_rand)
  const float* blob;
  const float* film;    // [n_inst][9][2][128] gamma, beta
  const float* film_tc; // [n_inst][9][128] float2 (gamma', delta) for the tcgen05 core
  float* scratch;       // per-CTA slabs
  size_t scratch_stride;  // floats per CTA
  float *cdf_fine, *gradients, *alpha, *inside_sphere, *mid_z, *sdf, *pts_norm, *pts, *raw_color;
  float* z_out;
  float* sdf_coarse;    // [R,n] (coarse pass output)
  // Per-ray compositing inside the tcgen05 core (fine pass, rays aligned with the 128-point tiles: 128 % S == 0):
  // the kernel then writes the final weights and the per-ray outputs itself and reduces the two global scalars in
  // its last CTA; otherwise composite_kernel does it in a second launch.
  int fuse_composite;
  float *weight_sum, *weight_max, *color_fine, *s_val, *gradient_error, *surface_loss;
  float* partials;      // [R,3] per-ray partial sums of the global scalars
  unsigned int* ticket;
  MapsKArgs maps;
};

// launchers (each returns an OiStatus)
int launch_pack_weights(const OiNetParams* p, float* blob, cudaStream_t st);
int launch_style_mlp(const OiNetParams* p, const float* z, float* w, int n_inst, cudaStream_t st);
int launch_film(const float* blob, int depth, const float* style_w, float* film, int n_inst, unsigned int* ticket,
                cudaStream_t st);
int launch_upfirdn2d(const OiUpfirdnDesc& d, cudaStream_t s);
int launch_bias_act(const OiBiasActDesc& d, cudaStream_t s);
int launch_fused_bias_act(const OiFusedBiasActDesc& d, cudaStream_t s);
int launch_gen_rays(const OiGenRaysDesc& d, cudaStream_t st);
int launch_render_maps_bwd(const OiRenderMapsBwdDesc& d, cudaStream_t st);
int launch_augment_geom(const OiAugmentGeomDesc& d, bool backward, cudaStream_t st);
int launch_augment_setup(const float* G_inv, int B, int H, int W, int hz_pad, float* theta, int* margins,
                         cudaStream_t st);
int launch_augment_setup_ops(const OiAugmentOp* ops, int n_ops, int B, int H, int W, int hz_pad, float* g_inv,
                             float* g_tmp, float* theta, int* margins, cudaStream_t st);
int launch_augment_setup_raw(const OiAugmentRawOp* ops, int n_ops, const float* p, int B, int H, int W, int hz_pad,
                             float* g_inv, float* g_tmp, float* theta, int* margins, cudaStream_t st);
size_t augment_u_floats(const OiAugmentGeomDesc& d);
size_t augment_r_floats(const OiAugmentGeomDesc& d);
int launch_render_maps(const OiRenderMapsDesc& d, cudaStream_t st);
int launch_render_ffma(const RenderKArgs& a, cudaStream_t st);
int launch_render_tc(const RenderKArgs& a, cudaStream_t st);
int launch_tc_selftest(const float* A, const float* B, const void* panel, float* D, cudaStream_t st);
int launch_bwd_tail(const OiRenderBwdDesc& d, const RenderKArgs& geo, float* adj, float* invs_partial,
                    unsigned int* relax_count, const unsigned int* guard, float* d_film, bool head_biases,
                    cudaStream_t st);
int launch_render_bwd_ffma(const OiRenderBwdDesc& d, const RenderKArgs& geo, const float* adj,
                           const float* invs_partial, float* d_film, float* scratch, int n_ctas, cudaStream_t st);
int launch_render_bwd_tc(const OiRenderBwdDesc& d, const RenderKArgs& geo, const float* adj, const float* invs_partial,
                         float* d_film, float* scratch, float* slabs, float* aux, float* dw_inst,
                         const unsigned int* ctl, unsigned int* sticky, int chunk_tiles, int n_ctas, cudaStream_t st);
size_t render_bwd_tc_dw_floats(int n_inst, int depth);
int render_bwd_ctas(int n_tiles);
size_t render_bwd_scratch_floats();
int render_bwd_tc_ctas(int n_tiles);
size_t render_bwd_tc_scratch_floats();
size_t render_bwd_tc_slab_floats_per_tile();
size_t render_ffma_scratch_floats(int depth, int* n_ctas, int n_tiles);
size_t render_tc_scratch_floats(int depth, int* n_ctas, int n_tiles);
int launch_upsample(int R, int n, int m, const float* rays_o, const float* rays_d, const float* near,
                    const float* far, const float* t_rand, const float* lin, const float* lin_fine,
                    const float* z_in, float inv_s, const float* sdf_coarse, float* z_fine, cudaStream_t st);
int launch_composite(int R, int S, const float* blob, int depth, float* weights /*in: alpha*/,
                     const float* raw_color, const float* gradients, const float* pts_norm, const float* sdf,
                     float* weight_sum, float* weight_max, float* color_fine, float* s_val,
                     float* gradient_error, float* surface_loss, float* partials, unsigned int* ticket,
                     cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Same, but lets the hardware suspend the thread for up to `hint_ns` per probe (long waits: far fewer
// issue slots burnt on polling than the plain try_wait loop).
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, uint32_t parity, uint32_t hint_ns = 20000u) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(hint_ns)
      : "memory");
}
// Polling with back-off: a failed probe parks the warp for `sleep_ns` (nanosleep) before the next one.  For the
// single-lane service warps (TMA producer, MMA issuers): their probes otherwise take MIO-queue and issue slots from
// the epilogue warps of the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, uint32_t parity, uint32_t sleep_ns) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "nanosleep.u32 %2;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(sleep_ns)
      : "memory");
}
// Drops a 128-byte line from L2 without writing it back (the data is dead: read-once scratch).
__device__ __forceinline__ void l2_discard_128(const void* p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// TMA 1-D bulk copy shared -> global (bulk async-group completion), and the group bookkeeping of the issuing thread.
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's most recent bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// at most N of this thread's most recent bulk groups are incomplete (writes not yet performed)
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes -> visible to the async proxy (TMA) that reads them next
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// sin and cos of a FiLM-SIREN pre-activation (|x| up to a few hundred rad).  Two-constant Cody-Waite
// reduction by 2*pi (exact to ~1e-7 rad for |x| < 1e3) followed by the MUFU approximations, whose
// absolute error on [-pi, pi] (2^-21.4) is below the fp32 rounding noise of the argument itself.
__device__ __forceinline__ void sincos_film(float x, float* s, float* c) {
  const float kInv2Pi = 0.15915494309189535f;
  const float k2PiHi = 6.2831854820251465f;      // fp32(2*pi)
  const float k2PiLo = -1.7484555314695172e-07f; // 2*pi - fp32(2*pi)
  float t = fmaf(x, kInv2Pi, 12582912.0f);
  float k = t - 12582912.0f;  // rint(x / 2pi)
  float r = fmaf(k, -k2PiHi, x);
  r = fmaf(k, -k2PiLo, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
#endif

}  // namespace oi
