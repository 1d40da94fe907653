// extern "C" entry points of liboi_b200.so (declared in include/oi_b200.h): validation, workspace carving,
// launch sequencing.  No allocation, no synchronisation, no exceptions.
#include <stdarg.h>
#include <string.h>

#include "oi_internal.cuh"
#include "oi_wgrad.cuh"

namespace oi {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  size_t film, scratch, partials, ticket, sdf_coarse, z_fine, z_fine2, tmp_weights, tmp_pts, tmp_mid_z, tmp_weight_sum,
      tmp_color_fine, tmp_raw_color, tmp_gradients, tmp_pts_norm, tmp_sdf;
  size_t total;
  int n_ctas;
  size_t scratch_stride;
};

int resolve_impl(const OiRenderDesc* d);
// The tcgen05 core composites per ray itself when rays are aligned runs of its 128-point tiles.
bool fused_composite(const OiRenderDesc* d, int impl) {
  const int S = d->n_samples + d->n_importance;
  return impl == OI_IMPL_TCGEN05 && S <= 128 && 128 % S == 0 && (d->flags & 8) == 0;
}

// Contract B inside the render kernel (flags bit 4): the maps are composited in the tile tail and nothing per-point is
// written.  Off by default: the tail runs on 4 of a slot's 8 epilogue warps and on the slot's critical path, and the
// extra code pushes the kernel past the instruction cache (+0.29 ms at 16 384 x 64 vs +0.035 ms for the maps kernel).
bool maps_in_kernel(const OiRenderDesc* d, int impl) {
  return d->maps != nullptr && (d->flags & 16) != 0 && fused_composite(d, impl);
}

int resolve_impl(const OiRenderDesc* d) {
  int impl = d->impl;
  if (impl == OI_IMPL_AUTO) impl = (d->depth >= 2) ? OI_IMPL_TCGEN05 : OI_IMPL_FFMA;
  return impl;
}

int plan_workspace(const OiRenderDesc* d, Workspace* w) {
  const int R = d->n_rays, S = d->n_samples + d->n_importance;
  const int n_inst = R / d->rays_per_instance;
  const long long pts_per_inst = (long long)d->rays_per_instance * S;
  const int tiles_per_inst = (int)((pts_per_inst + 127) / 128);
  const int n_tiles = tiles_per_inst * n_inst;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  w->film = take((size_t)n_inst * kFilm * 4 * kW * 4);  // (gamma, beta) table + (gamma', delta) table
  int n_ctas = 0;
  size_t stride = (resolve_impl(d) == OI_IMPL_TCGEN05) ? render_tc_scratch_floats(d->depth, &n_ctas, n_tiles)
                                                      : render_ffma_scratch_floats(d->depth, &n_ctas, n_tiles);
  w->n_ctas = n_ctas;
  w->scratch_stride = stride;
  w->scratch = take((size_t)n_ctas * stride * 4);
  w->partials = take((size_t)R * 3 * 4);
  w->ticket = take(256);
  const bool hier = d->n_importance > 0 && d->z_vals_in == nullptr;
  // multi-step up-sampling re-evaluates the SDF at the merged z of every step: [R,S] buffers, two z buffers
  w->sdf_coarse = hier ? take((size_t)R * (d->up_sample_steps > 1 ? S : d->n_samples) * 4) : 0;
  w->z_fine = hier ? take((size_t)R * S * 4) : 0;
  w->z_fine2 = (hier && d->up_sample_steps > 1) ? take((size_t)R * S * 4) : 0;
  // maps without in-kernel compositing: the per-point tensors the caller did not ask for go through the workspace
  const bool maps_unfused = d->maps != nullptr && !maps_in_kernel(d, resolve_impl(d));
  w->tmp_weights = (d->maps && !d->weights && maps_unfused) ? take((size_t)R * S * 4) : 0;
  w->tmp_pts = (maps_unfused && !d->pts) ? take((size_t)R * S * 3 * 4) : 0;
  w->tmp_mid_z = (maps_unfused && !d->mid_z_vals) ? take((size_t)R * S * 4) : 0;
  w->tmp_weight_sum = (maps_unfused && !d->weight_sum) ? take((size_t)R * 4) : 0;
  w->tmp_color_fine = (maps_unfused && !d->color_fine) ? take((size_t)R * 3 * 4) : 0;
  w->tmp_raw_color = d->raw_color ? 0 : take((size_t)R * S * 3 * 4);
  w->tmp_gradients = d->gradients ? 0 : take((size_t)R * S * 3 * 4);
  w->tmp_pts_norm = d->pts_norm ? 0 : take((size_t)R * S * 4);
  w->tmp_sdf = d->sdf ? 0 : take((size_t)R * S * 4);
  w->total = off;
  return OI_OK;
}

int validate_render(const OiRenderDesc* d) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->n_rays > 0, "n_rays must be positive (got %d)", d->n_rays);
  OI_CHECK_ARG(d->rays_per_instance > 0 && d->n_rays % d->rays_per_instance == 0,
               "n_rays (%d) must be a multiple of rays_per_instance (%d)", d->n_rays, d->rays_per_instance);
  OI_CHECK_ARG(d->n_samples >= 2, "n_samples must be >= 2 (got %d)", d->n_samples);
  OI_CHECK_ARG(d->n_importance >= 0, "n_importance must be >= 0");
  OI_CHECK_ARG(d->depth >= 1 && d->depth <= OI_MAX_DEPTH, "depth must be in [1, %d] (got %d)", OI_MAX_DEPTH, d->depth);
  if (d->n_importance > 0) {
    OI_CHECK_ARG(d->up_sample_steps >= 1, "up_sample_steps must be >= 1 (got %d)", d->up_sample_steps);
    // the reference adds n_importance // up_sample_steps samples per step (renderer.py:404) and then assumes
    // n_samples + n_importance samples per ray (:415): only exact divisions are consistent
    OI_CHECK_ARG(d->n_importance % d->up_sample_steps == 0, "n_importance (%d) must be a multiple of up_sample_steps (%d)",
                 d->n_importance, d->up_sample_steps);
  }
  if ((long long)d->n_rays * (d->n_samples + d->n_importance) >= (1ll << 30))
    return set_error(OI_ERR_UNSUPPORTED, "n_rays * samples too large for 32-bit point indices");
  OI_CHECK_ARG(d->impl >= OI_IMPL_AUTO && d->impl <= OI_IMPL_TCGEN05, "bad impl %d", d->impl);
  OI_CHECK_ARG(d->rays_o && d->rays_d && d->near && d->far, "rays_o/rays_d/near/far must be non-NULL");
  OI_CHECK_ARG(d->style_w && d->packed_weights, "style_w and packed_weights must be non-NULL");
  OI_CHECK_ARG(((uintptr_t)d->packed_weights & 127) == 0, "packed_weights must be 128-byte aligned");
  OI_CHECK_ARG(d->weights != nullptr || d->maps != nullptr, "the `weights` output is mandatory (unless `maps` is given)");
  if (d->maps) {
    const OiRenderMapsDesc* mp = d->maps;
    OI_CHECK_ARG(mp->rays_per_instance > 0 && d->n_rays % mp->rays_per_instance == 0,
                 "maps: n_rays (%d) must be a multiple of maps->rays_per_instance (%d)", d->n_rays, mp->rays_per_instance);
    OI_CHECK_ARG(mp->light_dir && mp->bg_color, "maps: light_dir and bg_color must be non-NULL");
  }
  return OI_OK;
}

}  // namespace
}  // namespace oi

using namespace oi;

extern "C" {

const char* oi_last_error(void) { return g_err; }
int oi_abi_version(void) { return OI_ABI_VERSION; }
const char* oi_build_info(void) { return "sm_100a;ffma;tcgen05;backward"; }

int oi_packed_weights_bytes(int32_t depth, size_t* bytes) {
  OI_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  OI_CHECK_ARG(depth >= 1 && depth <= OI_MAX_DEPTH, "depth must be in [1, %d]", OI_MAX_DEPTH);
  *bytes = blob_layout(depth).total_floats * sizeof(float);
  return OI_OK;
}

static int check_params(const OiNetParams* p, bool need_style) {
  OI_CHECK_ARG(p != nullptr, "params is NULL");
  OI_CHECK_ARG(p->depth >= 1 && p->depth <= OI_MAX_DEPTH, "depth must be in [1, %d] (got %d)", OI_MAX_DEPTH, p->depth);
  if (p->width != OI_WIDTH || p->style_dim != OI_STYLE_DIM)
    return set_error(OI_ERR_UNSUPPORTED, "only W=%d, style_dim=%d are implemented (got %d, %d)", OI_WIDTH,
                     OI_STYLE_DIM, p->width, p->style_dim);
  if (need_style) {
    for (int i = 0; i < 3; ++i) OI_CHECK_ARG(p->style_weight[i] && p->style_bias[i], "style layer %d is NULL", i);
    return OI_OK;
  }
  for (int l = 0; l < p->depth; ++l)
    OI_CHECK_ARG(p->pts_weight[l] && p->pts_bias[l] && p->gamma_weight[l] && p->gamma_bias[l] && p->beta_weight[l] &&
                     p->beta_bias[l],
                 "pts_linears[%d] has a NULL tensor", l);
  OI_CHECK_ARG(p->gamma_weight[OI_MAX_DEPTH] && p->gamma_bias[OI_MAX_DEPTH] && p->beta_weight[OI_MAX_DEPTH] &&
                   p->beta_bias[OI_MAX_DEPTH],
               "views_linears FiLM tensors (index %d) are NULL", OI_MAX_DEPTH);
  OI_CHECK_ARG(p->sigma_weight && p->sigma_bias && p->views_weight && p->views_bias && p->rgb_weight && p->rgb_bias &&
                   p->variance,
               "a head tensor is NULL");
  return OI_OK;
}

int oi_pack_weights(const OiNetParams* params, void* blob, size_t blob_bytes, void* stream) {
  int rc = check_params(params, false);
  if (rc) return rc;
  OI_CHECK_ARG(blob != nullptr && ((uintptr_t)blob & 127) == 0, "blob must be non-NULL and 128-byte aligned");
  size_t need = blob_layout(params->depth).total_floats * sizeof(float);
  if (blob_bytes < need) return set_error(OI_ERR_WORKSPACE, "blob too small: %zu < %zu", blob_bytes, need);
  return launch_pack_weights(params, static_cast<float*>(blob), static_cast<cudaStream_t>(stream));
}

int oi_style_mlp(const OiNetParams* params, const float* z, float* w, int32_t n_instances, void* stream) {
  int rc = check_params(params, true);
  if (rc) return rc;
  OI_CHECK_ARG(z && w && n_instances > 0, "z/w NULL or n_instances <= 0");
  return launch_style_mlp(params, z, w, n_instances, static_cast<cudaStream_t>(stream));
}

int oi_render_workspace_bytes(const OiRenderDesc* desc, size_t* bytes) {
  int rc = validate_render(desc);
  if (rc) return rc;
  OI_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  Workspace w;
  plan_workspace(desc, &w);
  *bytes = w.total;
  return OI_OK;
}

int oi_render_launch_count(const OiRenderDesc* desc, int32_t* launches) {
  int rc = validate_render(desc);
  if (rc) return rc;
  OI_CHECK_ARG(launches != nullptr, "launches is NULL");
  const bool hier = desc->n_importance > 0 && desc->z_vals_in == nullptr;
  // film, steps x [coarse, upsample], fine, [composite unless the core does it]
  const bool fused = fused_composite(desc, resolve_impl(desc));
  *launches = 3 + (hier ? 2 * desc->up_sample_steps : 0) - (fused ? 1 : 0) +
              ((desc->maps && !maps_in_kernel(desc, resolve_impl(desc))) ? 1 : 0);
  return OI_OK;
}

int oi_render_forward(const OiRenderDesc* d, void* stream) {
  int rc = validate_render(d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace w;
  plan_workspace(d, &w);
  OI_CHECK_ARG(d->workspace != nullptr && ((uintptr_t)d->workspace & 255) == 0,
               "workspace must be non-NULL and 256-byte aligned");
  if (d->workspace_bytes < w.total)
    return set_error(OI_ERR_WORKSPACE, "workspace too small: %zu < %zu", d->workspace_bytes, w.total);
  char* ws = static_cast<char*>(d->workspace);
  const float* blob = static_cast<const float*>(d->packed_weights);
  const int impl = resolve_impl(d);

  const int R = d->n_rays, n = d->n_samples, m = d->n_importance, S = n + m;
  const int n_inst = R / d->rays_per_instance;
  float* film = reinterpret_cast<float*>(ws + w.film);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + w.ticket);
  rc = launch_film(blob, d->depth, d->style_w, film, n_inst, ticket, st);
  if (rc) return rc;

  RenderKArgs a;
  memset(&a, 0, sizeof(a));
  a.R = R;
  a.rays_per_inst = d->rays_per_instance;
  a.n_inst = n_inst;
  a.n_coarse = n;
  a.D = d->depth;
  a.cos_anneal = d->cos_anneal_ratio;
  a.flags = d->flags;
  a.sample_dist = 2.0f / (float)n;  // renderer.py:356
  a.rays_o = d->rays_o;
  a.rays_d = d->rays_d;
  a.near = d->near;
  a.far = d->far;
  a.t_rand = d->t_rand;
  a.lin = d->lin_coarse;
  a.blob = blob;
  a.film = film;
  a.film_tc = film + (size_t)n_inst * kFilm * 2 * kW;
  a.scratch = reinterpret_cast<float*>(ws + w.scratch);
  a.scratch_stride = w.scratch_stride;

  const float* z_vals = d->z_vals_in;
  const bool hier = m > 0 && z_vals == nullptr;
  if (hier) {
    // up_sample_steps x { SDF-only pass at the current z, up_sample + cat_z_vals } (renderer.py:389-413).  Step 0
    // evaluates the n stratified z; every later step re-evaluates the SDF at the merged (sorted) z of the previous
    // one -- the reference evaluates only the new points and gathers (:186-195); per point the result is the same.
    const int steps = d->up_sample_steps, m_step = m / steps;
    float* zbuf[2] = {reinterpret_cast<float*>(ws + w.z_fine), reinterpret_cast<float*>(ws + w.z_fine2)};
    if (steps % 2 == 0) {   // the last step must land in z_fine
      float* t = zbuf[0];
      zbuf[0] = zbuf[1];
      zbuf[1] = t;
    }
    const float* z_cur = nullptr;
    for (int i = 0; i < steps; ++i) {
      const int n_cur = n + i * m_step;
      RenderKArgs c = a;
      c.coarse = 1;
      c.S = n_cur;
      c.z_vals = z_cur;
      c.pts_per_inst = d->rays_per_instance * n_cur;
      c.tiles_per_inst = (c.pts_per_inst + 127) / 128;
      c.n_tiles = c.tiles_per_inst * n_inst;
      c.sdf_coarse = reinterpret_cast<float*>(ws + w.sdf_coarse);
      rc = (impl == OI_IMPL_TCGEN05) ? launch_render_tc(c, st) : launch_render_ffma(c, st);
      if (rc) return rc;
      float* z_next = zbuf[i & 1];
      rc = launch_upsample(R, n_cur, m_step, d->rays_o, d->rays_d, d->near, d->far, d->t_rand, d->lin_coarse,
                           d->lin_fine, z_cur, 64.0f * (float)(1 << i), c.sdf_coarse, z_next, st);
      if (rc) return rc;
      z_cur = z_next;
    }
    z_vals = z_cur;
  }

  a.coarse = 0;
  a.S = S;
  a.pts_per_inst = d->rays_per_instance * S;
  a.tiles_per_inst = (a.pts_per_inst + 127) / 128;
  a.n_tiles = a.tiles_per_inst * n_inst;
  a.z_vals = z_vals;
  if (z_vals == nullptr && m > 0) return set_error(OI_ERR_INVALID_ARGUMENT, "internal: missing fine z values");
  a.cdf_fine = d->cdf_fine;
  a.gradients = d->gradients ? d->gradients : reinterpret_cast<float*>(ws + w.tmp_gradients);
  a.alpha = d->weights;
  a.inside_sphere = d->inside_sphere;
  a.mid_z = d->mid_z_vals;
  a.sdf = d->sdf ? d->sdf : reinterpret_cast<float*>(ws + w.tmp_sdf);
  a.pts_norm = d->pts_norm ? d->pts_norm : reinterpret_cast<float*>(ws + w.tmp_pts_norm);
  a.pts = d->pts;
  a.raw_color = d->raw_color ? d->raw_color : reinterpret_cast<float*>(ws + w.tmp_raw_color);
  a.z_out = d->z_vals_out;
  a.fuse_composite = fused_composite(d, impl) ? 1 : 0;
  float* weights = d->weights;
  float* weight_sum = d->weight_sum;
  float* color_fine = d->color_fine;
  const bool maps_fused = maps_in_kernel(d, impl);
  if (d->maps && !maps_fused) {   // maps kernel after the render: every input of render_maps_kernel must exist somewhere
    if (!weights) weights = reinterpret_cast<float*>(ws + w.tmp_weights);
    if (!weight_sum) weight_sum = reinterpret_cast<float*>(ws + w.tmp_weight_sum);
    if (!color_fine) color_fine = reinterpret_cast<float*>(ws + w.tmp_color_fine);
    if (!a.pts) a.pts = reinterpret_cast<float*>(ws + w.tmp_pts);
    if (!a.mid_z) a.mid_z = reinterpret_cast<float*>(ws + w.tmp_mid_z);
    a.alpha = weights;
  }
  if (maps_fused) {
    const OiRenderMapsDesc* mp = d->maps;
    MapsKArgs& k = a.maps;
    k.enabled = 1;
    k.rays_per_image = mp->rays_per_instance;
    k.light_dir = mp->light_dir;
    k.bg_color = mp->bg_color;
    k.light_params = mp->light_params;
    for (int c = 0; c < 3; ++c) {
      k.lp[c] = mp->ambient_color[c];
      k.lp[3 + c] = mp->diffuse_color[c];
      k.lp[6 + c] = mp->specular_color[c];
    }
    k.lp[9] = mp->shininess;
    k.image = mp->image;
    k.image_no_bg = mp->image_no_bg;
    k.mask = mp->mask;
    k.shading_map = mp->shading_map;
    k.color_map = mp->color_map;
    k.weight_sum_map = mp->weight_sum_map;
    k.amb_shading_map = mp->amb_shading_map;
    k.diff_shading_map = mp->diff_shading_map;
    k.normal_map = mp->normal_map;
    k.no_specular_map = mp->no_specular_map;
    k.specular_map = mp->specular_map;
    k.z_map = mp->z_map;
    k.z_min_per_ray = mp->z_min_per_ray;
  }
  if (a.fuse_composite) {
    a.weight_sum = weight_sum;
    a.weight_max = d->weight_max;
    a.color_fine = color_fine;
    a.s_val = d->s_val;
    a.gradient_error = d->gradient_error;
    a.surface_loss = d->surface_loss;
    a.partials = reinterpret_cast<float*>(ws + w.partials);
    a.ticket = ticket;
  }
  if (d->evt_core_start) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d->evt_core_start), st));
  rc = (impl == OI_IMPL_TCGEN05) ? launch_render_tc(a, st) : launch_render_ffma(a, st);
  if (rc) return rc;
  if (d->evt_core_stop) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d->evt_core_stop), st));
  if (a.fuse_composite && (!d->maps || maps_fused)) return OI_OK;
  if (!a.fuse_composite) {
    rc = launch_composite(R, S, blob, d->depth, weights, a.raw_color, a.gradients, a.pts_norm, a.sdf, weight_sum,
                          d->weight_max, color_fine, d->s_val, d->gradient_error, d->surface_loss,
                          reinterpret_cast<float*>(ws + w.partials), ticket, st);
    if (rc || !d->maps) return rc;
  }

  OiRenderMapsDesc md = *d->maps;   // the render's own results are the inputs of the maps kernel
  md.n_rays = R;
  md.n_samples = S;
  md.weights = weights;
  md.gradients = a.gradients;
  md.raw_color = a.raw_color;
  md.pts = a.pts;
  md.mid_z_vals = a.mid_z;
  md.weight_sum = weight_sum;
  md.color_fine = color_fine;
  md.rays_o = d->rays_o;
  return launch_render_maps(md, st);
}

namespace {
struct BwdWorkspace {
  size_t film, d_film, adj, invs_partial, relax_count, ticket, scratch, slabs, aux, dw_inst, total;
  int n_ctas, n_inst, tiles_per_inst, n_tiles, chunk_tiles;
  bool tc;
};
constexpr int kBwdStickyWord = 8;   // word of the `ticket` block that holds the fp16 overflow guard (oi_wgrad.cuh); never reset
constexpr int kBwdChunkTiles = 2048;  // tiles per (bwd_tc_kernel, wgrad_tc_kernel) round: 5.8 GB of slabs

int validate_bwd(const OiRenderBwdDesc* d) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->n_rays > 0, "n_rays must be positive (got %d)", d->n_rays);
  OI_CHECK_ARG(d->rays_per_instance > 0 && d->n_rays % d->rays_per_instance == 0,
               "n_rays (%d) must be a multiple of rays_per_instance (%d)", d->n_rays, d->rays_per_instance);
  OI_CHECK_ARG(d->n_samples >= 2 && d->n_samples_total >= d->n_samples, "bad n_samples / n_samples_total");
  if (d->depth < 2 || d->depth > OI_MAX_DEPTH)
    return set_error(OI_ERR_UNSUPPORTED, "backward needs 2 <= depth <= %d (got %d)", OI_MAX_DEPTH, d->depth);
  if ((long long)d->n_rays * d->n_samples_total >= (1ll << 30))
    return set_error(OI_ERR_UNSUPPORTED, "n_rays * samples too large for 32-bit point indices");
  OI_CHECK_ARG(d->impl >= OI_IMPL_AUTO && d->impl <= OI_IMPL_TCGEN05, "bad impl %d", d->impl);
  OI_CHECK_ARG(d->rays_o && d->rays_d && d->z_vals && d->style_w && d->packed_weights, "NULL input pointer");
  OI_CHECK_ARG(((uintptr_t)d->packed_weights & 127) == 0, "packed_weights must be 128-byte aligned");
  OI_CHECK_ARG(d->sdf && d->gradients && d->raw_color, "sdf / gradients / raw_color of the forward are required");
  const OiNetGrads& g = d->grads;
  for (int l = 0; l < d->depth; ++l)
    OI_CHECK_ARG(g.pts_weight[l] && g.pts_bias[l], "grads.pts_weight/bias[%d] is NULL", l);
  for (int l = 1; l < d->depth; ++l)
    OI_CHECK_ARG(((uintptr_t)g.pts_weight[l] & 15) == 0, "grads.pts_weight[%d] must be 16-byte aligned", l);
  OI_CHECK_ARG(g.sigma_weight && g.sigma_bias && g.views_weight && g.views_bias && g.rgb_weight && g.rgb_bias &&
                   g.variance && g.film_gamma && g.film_beta,
               "a grads pointer is NULL");
  return OI_OK;
}

void plan_bwd(const OiRenderBwdDesc* d, BwdWorkspace* w) {
  const int R = d->n_rays, S = d->n_samples_total;
  w->n_inst = R / d->rays_per_instance;
  const long long pts_per_inst = (long long)d->rays_per_instance * S;
  w->tiles_per_inst = (int)((pts_per_inst + 127) / 128);
  w->n_tiles = w->tiles_per_inst * w->n_inst;
  w->tc = d->impl != OI_IMPL_FFMA;
  w->chunk_tiles = w->n_tiles < kBwdChunkTiles ? w->n_tiles : kBwdChunkTiles;
  w->n_ctas = w->tc ? render_bwd_tc_ctas(w->chunk_tiles) : render_bwd_ctas(w->n_tiles);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  w->film = take((size_t)w->n_inst * kFilm * 4 * kW * 4);
  w->d_film = take((size_t)w->n_inst * kFilm * 2 * kW * 4);
  w->adj = take((size_t)R * S * 8 * 4);
  w->invs_partial = take((size_t)R * 4);
  w->relax_count = take(256);
  w->ticket = take(256);
  if (w->tc) {
    w->scratch = take((size_t)w->n_ctas * render_bwd_tc_scratch_floats() * 4);
    w->slabs = take((size_t)w->chunk_tiles * kSlabsPerTile * kSlabFloats * 4);
    w->aux = take((size_t)w->chunk_tiles * 512 * 4);
    w->dw_inst = take(render_bwd_tc_dw_floats(w->n_inst, d->depth) * 4);
  } else {
    w->scratch = take((size_t)w->n_ctas * render_bwd_scratch_floats() * 4);
    w->slabs = w->aux = w->dw_inst = 0;
  }
  w->total = off;
}
}  // namespace

int oi_render_backward_workspace_bytes(const OiRenderBwdDesc* desc, size_t* bytes) {
  int rc = validate_bwd(desc);
  if (rc) return rc;
  OI_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  BwdWorkspace w;
  plan_bwd(desc, &w);
  *bytes = w.total;
  return OI_OK;
}

int oi_render_backward(const OiRenderBwdDesc* d, void* stream) {
  int rc = validate_bwd(d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BwdWorkspace w;
  plan_bwd(d, &w);
  OI_CHECK_ARG(d->workspace != nullptr && ((uintptr_t)d->workspace & 255) == 0,
               "workspace must be non-NULL and 256-byte aligned");
  if (d->workspace_bytes < w.total)
    return set_error(OI_ERR_WORKSPACE, "workspace too small: %zu < %zu", d->workspace_bytes, w.total);
  char* ws = static_cast<char*>(d->workspace);
  const float* blob = static_cast<const float*>(d->packed_weights);
  float* film = reinterpret_cast<float*>(ws + w.film);
  rc = launch_film(blob, d->depth, d->style_w, film, w.n_inst, reinterpret_cast<unsigned int*>(ws + w.ticket), st);
  if (rc) return rc;

  RenderKArgs a;
  memset(&a, 0, sizeof(a));
  a.R = d->n_rays;
  a.rays_per_inst = d->rays_per_instance;
  a.n_inst = w.n_inst;
  a.n_coarse = d->n_samples;
  a.D = d->depth;
  a.cos_anneal = d->cos_anneal_ratio;
  a.flags = d->flags;
  a.sample_dist = 2.0f / (float)d->n_samples;  // renderer.py:356
  a.rays_o = d->rays_o;
  a.rays_d = d->rays_d;
  a.z_vals = d->z_vals;
  a.blob = blob;
  a.film = film;
  a.coarse = 0;
  a.S = d->n_samples_total;
  a.pts_per_inst = d->rays_per_instance * a.S;
  a.tiles_per_inst = w.tiles_per_inst;
  a.n_tiles = w.n_tiles;
  a.film_tc = film + (size_t)w.n_inst * kFilm * 2 * kW;
  float* adj = reinterpret_cast<float*>(ws + w.adj);
  float* invs_partial = reinterpret_cast<float*>(ws + w.invs_partial);
  float* d_film = reinterpret_cast<float*>(ws + w.d_film);
  rc = launch_bwd_tail(*d, a, adj, invs_partial, reinterpret_cast<unsigned int*>(ws + w.relax_count),
                       reinterpret_cast<const unsigned int*>(ws + w.ticket) + kBwdStickyWord, d_film, w.tc, st);
  if (rc) return rc;
  if (!w.tc)
    return launch_render_bwd_ffma(*d, a, adj, invs_partial, d_film, reinterpret_cast<float*>(ws + w.scratch), w.n_ctas,
                                  st);
  return launch_render_bwd_tc(*d, a, adj, invs_partial, d_film,
                              reinterpret_cast<float*>(ws + w.scratch), reinterpret_cast<float*>(ws + w.slabs),
                              reinterpret_cast<float*>(ws + w.aux), reinterpret_cast<float*>(ws + w.dw_inst),
                              reinterpret_cast<const unsigned int*>(ws + w.relax_count),
                              reinterpret_cast<unsigned int*>(ws + w.ticket) + kBwdStickyWord, w.chunk_tiles, w.n_ctas, st);
}

int oi_render_backward_control_words(const OiRenderBwdDesc* d, uint32_t* words, void* stream) {
  int rc = validate_bwd(d);
  if (rc) return rc;
  OI_CHECK_ARG(words != nullptr && d->workspace != nullptr, "NULL pointer");
  BwdWorkspace w;
  plan_bwd(d, &w);
  if (d->workspace_bytes < w.total)
    return set_error(OI_ERR_WORKSPACE, "workspace too small: %zu < %zu", d->workspace_bytes, w.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OI_CHECK_CUDA(cudaMemcpyAsync(words, static_cast<const char*>(d->workspace) + w.relax_count, 8 * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, st));
  uint32_t sticky = 0;
  OI_CHECK_CUDA(cudaMemcpyAsync(&sticky, static_cast<const char*>(d->workspace) + w.ticket + kBwdStickyWord * 4,
                                sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  OI_CHECK_CUDA(cudaStreamSynchronize(st));
  words[8] = sticky == kF16Safe ? 1u : (sticky == kF16Unsafe ? 2u : 0u);
  words[9] = words[10] = words[11] = 0u;
  return OI_OK;
}

int oi_render_backward_operand_format(const OiRenderBwdDesc* d, int32_t* format, void* stream) {
  OI_CHECK_ARG(format != nullptr, "NULL pointer");
  unsigned int ctl[12];
  int rc = oi_render_backward_control_words(d, ctl, stream);
  if (rc) return rc;
  // NOTE: the guard word read here is the one the NEXT call will see (a probing call has just marked the workspace
  // safe and answers 1 although it ran on TF32 itself; a call that has just tripped the guard answers 0)
  *format = (d->impl != OI_IMPL_FFMA &&
             bwd_mode(ctl, d->flags, ctl[8] == 1u ? kF16Safe : (ctl[8] == 2u ? kF16Unsafe : 0u)).f16) ? 1 : 0;
  return OI_OK;
}

int oi_selftest_bwd_mode(const uint32_t* control_words, int32_t flags, int32_t guard_state, int32_t* fp16,
                         int32_t* e_ref) {
  OI_CHECK_ARG(control_words && fp16 && e_ref, "NULL pointer");
  OI_CHECK_ARG(guard_state >= 0 && guard_state <= 2, "guard_state must be 0 (unknown), 1 (safe) or 2 (unsafe)");
  const BwdMode m = bwd_mode(control_words, flags, guard_state == 1 ? kF16Safe : (guard_state == 2 ? kF16Unsafe : 0u));
  *fp16 = m.f16 ? 1 : 0;
  *e_ref = m.e_ref;
  return OI_OK;
}

int oi_selftest_tc(const float* a, const float* b, const void* packed_weights, int32_t depth, int32_t panel,
                   float* d, void* stream) {
  OI_CHECK_ARG(a && d, "a and d must be non-NULL");
  const void* img = nullptr;
  if (packed_weights) {
    OI_CHECK_ARG(depth >= 2 && depth <= OI_MAX_DEPTH && panel >= 0 && panel < 2 * (depth - 1) + 1, "bad depth/panel");
    img = static_cast<const char*>(packed_weights) + blob_layout(depth).tc_off * sizeof(float) + (size_t)panel * 65536;
  } else {
    OI_CHECK_ARG(b != nullptr, "b must be non-NULL when no packed panel is given");
  }
  return launch_tc_selftest(a, b, img, d, static_cast<cudaStream_t>(stream));
}

int oi_selftest_wgrad(const float* slabs, const float* aux, int32_t n_tiles, int32_t slabs_per_tile, int32_t x_slab,
                      int32_t y_slab, int32_t tiles_per_instance, int32_t n_ctas, float* d, float* col, void* stream) {
  OI_CHECK_ARG(slabs && aux && d && col, "NULL pointer");
  OI_CHECK_ARG(n_tiles > 0 && slabs_per_tile > 0 && n_ctas > 0 && tiles_per_instance > 0, "bad sizes");
  OI_CHECK_ARG(x_slab >= 0 && x_slab < slabs_per_tile && y_slab >= 0 && y_slab < slabs_per_tile, "bad slab index");
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.n_tiles = n_tiles;
  a.tiles_per_inst = tiles_per_instance;
  a.n_groups = 1;
  a.n_ctas = n_ctas;
  a.slabs_per_tile = slabs_per_tile;
  a.slabs = slabs;
  a.aux = aux;
  WgGroup& g = a.groups[0];
  g.n_pairs = 1;
  g.pairs[0] = WgPair{x_slab, y_slab};
  g.use_aux = 1;
  for (int c = 0; c < 4; ++c) {
    g.aux_out[c] = col + c;
    g.aux_inst_stride[c] = 128 * 4;
    g.aux_ch_stride[c] = 4;
  }
  g.out = d;
  g.out_ld = 128;
  g.weight = 1;
  return launch_wgrad_tc(a, static_cast<cudaStream_t>(stream));
}

static int check_augment(const OiAugmentGeomDesc* d, bool need_ws) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->batch > 0 && d->channels > 0 && d->height >= 2 && d->width >= 2, "bad image sizes");
  OI_CHECK_ARG(d->filter_taps >= 4 && d->filter_taps <= 16 && d->filter_taps % 4 == 0,
               "filter_taps must be a multiple of 4 in [4, 16] (got %d)", d->filter_taps);
  if (!need_ws) return OI_OK;
  OI_CHECK_ARG(d->filter && d->theta && d->margins && d->x && d->y, "NULL pointer");
  const size_t need = (augment_u_floats(*d) + augment_r_floats(*d)) * sizeof(float);
  OI_CHECK_ARG(d->workspace != nullptr, "workspace is NULL");
  if (d->workspace_bytes < need) return set_error(OI_ERR_WORKSPACE, "workspace too small: %zu < %zu", d->workspace_bytes, need);
  return OI_OK;
}

int oi_augment_geom_setup(const float* g_inv, int32_t batch, int32_t height, int32_t width, int32_t filter_taps,
                          float* theta, int32_t* margins, void* stream) {
  OI_CHECK_ARG(g_inv && theta && margins, "NULL pointer");
  OI_CHECK_ARG(batch > 0 && height >= 2 && width >= 2 && filter_taps >= 4 && filter_taps % 4 == 0, "bad sizes");
  return launch_augment_setup(g_inv, batch, height, width, filter_taps / 4, theta, margins,
                              static_cast<cudaStream_t>(stream));
}

int oi_augment_geom_setup_ops(const OiAugmentOp* ops, int32_t n_ops, int32_t batch, int32_t height, int32_t width,
                              int32_t filter_taps, float* g_inv, float* theta, int32_t* margins, void* stream) {
  OI_CHECK_ARG(ops && g_inv && theta && margins, "NULL pointer");
  OI_CHECK_ARG(n_ops >= 1 && n_ops <= OI_AUGMENT_MAX_OPS, "n_ops must be in [1, %d] (got %d)", OI_AUGMENT_MAX_OPS, n_ops);
  OI_CHECK_ARG(batch > 0 && height >= 2 && width >= 2 && filter_taps >= 4 && filter_taps % 4 == 0, "bad sizes");
  for (int i = 0; i < n_ops; ++i) {
    OI_CHECK_ARG(ops[i].kind >= 0 && ops[i].kind <= 2, "op %d: bad kind %d", i, ops[i].kind);
    OI_CHECK_ARG(ops[i].p0 != nullptr && (ops[i].kind == 1 || ops[i].p1 != nullptr), "op %d: NULL parameter", i);
  }
  return launch_augment_setup_ops(ops, n_ops, batch, height, width, filter_taps / 4, g_inv, g_inv, theta, margins,
                                  static_cast<cudaStream_t>(stream));
}

int oi_augment_geom_setup_raw(const OiAugmentRawOp* ops, int32_t n_ops, const float* p, int32_t batch, int32_t height,
                              int32_t width, int32_t filter_taps, float* g_inv, float* theta, int32_t* margins,
                              void* stream) {
  OI_CHECK_ARG(ops && p && g_inv && theta && margins, "NULL pointer");
  OI_CHECK_ARG(n_ops >= 1 && n_ops <= OI_AUGMENT_MAX_OPS, "n_ops must be in [1, %d] (got %d)", OI_AUGMENT_MAX_OPS, n_ops);
  OI_CHECK_ARG(batch > 0 && height >= 2 && width >= 2 && filter_taps >= 4 && filter_taps % 4 == 0, "bad sizes");
  for (int i = 0; i < n_ops; ++i) {
    OI_CHECK_ARG(ops[i].form >= OI_AUG_XFLIP && ops[i].form <= OI_AUG_XFRAC, "op %d: bad form %d", i, ops[i].form);
    OI_CHECK_ARG(ops[i].draw != nullptr && ops[i].gate != nullptr, "op %d: NULL draw", i);
  }
  return launch_augment_setup_raw(ops, n_ops, p, batch, height, width, filter_taps / 4, g_inv, g_inv, theta, margins,
                                  static_cast<cudaStream_t>(stream));
}

int oi_augment_geom_workspace_bytes(const OiAugmentGeomDesc* d, size_t* bytes) {
  int rc = check_augment(d, false);
  if (rc) return rc;
  OI_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  *bytes = (augment_u_floats(*d) + augment_r_floats(*d)) * sizeof(float);
  return OI_OK;
}

int oi_augment_geom_forward(const OiAugmentGeomDesc* d, void* stream) {
  int rc = check_augment(d, true);
  if (rc) return rc;
  return launch_augment_geom(*d, false, static_cast<cudaStream_t>(stream));
}

int oi_augment_geom_backward(const OiAugmentGeomDesc* d, void* stream) {
  int rc = check_augment(d, true);
  if (rc) return rc;
  return launch_augment_geom(*d, true, static_cast<cudaStream_t>(stream));
}

int oi_gen_rays(const OiGenRaysDesc* d, void* stream) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->n_instances > 0 && d->resolution >= 2 && d->scene_resolution > 0, "bad sizes");
  OI_CHECK_ARG(d->b2w && d->c2b && d->w2c && d->intrinsics_inv && d->rays_o && d->rays_d, "NULL pointer");
  OI_CHECK_ARG((d->near == nullptr) == (d->far == nullptr), "near and far must be given together");
  return launch_gen_rays(*d, static_cast<cudaStream_t>(stream));
}

int oi_render_maps(const OiRenderMapsDesc* d, void* stream) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->n_rays > 0 && d->rays_per_instance > 0 && d->n_rays % d->rays_per_instance == 0 && d->n_samples > 0,
               "bad sizes");
  OI_CHECK_ARG(d->weights && d->gradients && d->raw_color && d->pts && d->weight_sum && d->color_fine && d->rays_o &&
                   d->light_dir && d->bg_color,
               "NULL input pointer");
  return launch_render_maps(*d, static_cast<cudaStream_t>(stream));
}

int oi_render_maps_backward(const OiRenderMapsBwdDesc* bd, void* stream) {
  OI_CHECK_ARG(bd != nullptr, "desc is NULL");
  const OiRenderMapsDesc* d = &bd->fwd;
  OI_CHECK_ARG(d->n_rays > 0 && d->rays_per_instance > 0 && d->n_rays % d->rays_per_instance == 0 && d->n_samples > 0,
               "bad sizes");
  OI_CHECK_ARG(d->weights && d->gradients && d->raw_color && d->pts && d->weight_sum && d->rays_o && d->light_dir &&
                   d->bg_color,
               "NULL input pointer");
  OI_CHECK_ARG(!bd->g_z_map || d->mid_z_vals, "g_z_map needs mid_z_vals");
  return launch_render_maps_bwd(*bd, static_cast<cudaStream_t>(stream));
}

int oi_upfirdn2d(const OiUpfirdnDesc* d, void* stream) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->x && d->f && d->y, "x, f, y must be non-NULL");
  OI_CHECK_ARG(d->batch > 0 && d->channels > 0 && d->in_h > 0 && d->in_w > 0, "empty input");
  OI_CHECK_ARG(d->filter_h >= 1 && d->filter_w >= 1, "f must be at least 1x1");
  OI_CHECK_ARG(d->up_x >= 1 && d->up_y >= 1, "upsampling factor must be at least 1");
  OI_CHECK_ARG(d->down_x >= 1 && d->down_y >= 1, "downsampling factor must be at least 1");
  const int ow = (d->in_w * d->up_x + d->pad_x0 + d->pad_x1 - d->filter_w + d->down_x) / d->down_x;
  const int oh = (d->in_h * d->up_y + d->pad_y0 + d->pad_y1 - d->filter_h + d->down_y) / d->down_y;
  OI_CHECK_ARG(ow >= 1 && oh >= 1, "output must be at least 1x1");
  OI_CHECK_ARG(ow == d->out_w && oh == d->out_h, "out size mismatch: expected %dx%d, got %dx%d", oh, ow, d->out_h,
               d->out_w);
  return launch_upfirdn2d(*d, static_cast<cudaStream_t>(stream));
}

int oi_bias_act(const OiBiasActDesc* d, void* stream) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->size_x >= 0, "size_x must be non-negative");
  OI_CHECK_ARG(d->grad >= 0 && d->grad <= 2, "grad must be 0, 1 or 2");
  OI_CHECK_ARG(d->act >= 1 && d->act <= 9, "act must be in [1, 9] (got %d)", d->act);
  if (d->size_x == 0) return OI_OK;   // empty tensors are legal (their data_ptr() is NULL), as in the reference op
  OI_CHECK_ARG(d->x && d->y, "x and y must be non-NULL");
  OI_CHECK_ARG(d->b == nullptr || (d->size_b > 0 && d->step_b > 0), "bias given but size_b/step_b invalid");
  return launch_bias_act(*d, static_cast<cudaStream_t>(stream));
}

int oi_fused_bias_act(const OiFusedBiasActDesc* d, void* stream) {
  OI_CHECK_ARG(d != nullptr, "desc is NULL");
  OI_CHECK_ARG(d->size_x >= 0, "size_x must be non-negative");
  OI_CHECK_ARG(d->act >= 1 && d->grad >= 0 && d->grad <= 2, "act must be >= 1 and grad in 0..2 (got %d, %d)", d->act,
               d->grad);
  if (d->size_x == 0) return OI_OK;   // empty tensors are legal (their data_ptr() is NULL)
  OI_CHECK_ARG(d->x && d->y, "x and y must be non-NULL");
  OI_CHECK_ARG(d->bias == nullptr || (d->size_b > 0 && d->step_b > 0), "bias given but size_b/step_b invalid");
  return launch_fused_bias_act(*d, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
