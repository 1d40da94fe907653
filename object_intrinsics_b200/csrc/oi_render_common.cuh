// Per-point prologue / tail shared by the FFMA and tcgen05 render cores.
#pragma once
#include "oi_internal.cuh"

namespace oi {

struct PointCtx {
  int ray, si;
  bool valid;
  float px, py, pz;   // sample position fed to the SDF network
  float dx, dy, dz;   // ray direction
  float dist;         // section length (renderer.py:219-222)
  float mid;          // section midpoint z
  float z0;           // section start z
};

// Sample position of point `p_local` of tile `tin` of instance `inst` (renderer.py:359-373, 219-228; coarse pass
// :391).  Also writes the position-only outputs (pts, mid_z_vals, pts_norm, inside_sphere, z_vals).
__device__ __forceinline__ PointCtx point_prologue(const RenderKArgs& a, int inst, int tin, int p_local,
                                                   bool write_outputs = true) {
  PointCtx c;
  const int p = tin * 128 + p_local;
  c.valid = p < a.pts_per_inst;
  const int pc = c.valid ? p : a.pts_per_inst - 1;
  const int rl = pc / a.S;
  c.si = pc - rl * a.S;
  c.ray = inst * a.rays_per_inst + rl;
  const int ray = c.ray, si = c.si;
  const float ox = a.rays_o[ray * 3 + 0], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  c.dx = a.rays_d[ray * 3 + 0];
  c.dy = a.rays_d[ray * 3 + 1];
  c.dz = a.rays_d[ray * 3 + 2];
  float z0, z1 = 0.f;
  if (a.z_vals) {
    z0 = a.z_vals[(size_t)ray * a.S + si];
    if (si + 1 < a.S) z1 = a.z_vals[(size_t)ray * a.S + si + 1];
  } else {
    const float nr = a.near[ray], fr = a.far[ray];
    const float span = fr - nr;
    const float jit = a.t_rand ? a.t_rand[ray] * 2.0f / (float)a.n_coarse : 0.f;
    const float l0 = a.lin ? a.lin[si] : (float)si / (float)(a.n_coarse - 1);
    z0 = nr + span * l0;
    if (a.t_rand) z0 = z0 + jit;
    if (si + 1 < a.S) {
      const float l1 = a.lin ? a.lin[si + 1] : (float)(si + 1) / (float)(a.n_coarse - 1);
      z1 = nr + span * l1;
      if (a.t_rand) z1 = z1 + jit;
    }
  }
  c.z0 = z0;
  float zp;
  if (a.coarse) {
    c.dist = 0.f;
    c.mid = z0;
    zp = z0;
  } else {
    c.dist = (si + 1 < a.S) ? (z1 - z0) : a.sample_dist;
    c.mid = z0 + c.dist * 0.5f;
    zp = c.mid;
  }
  c.px = ox + c.dx * zp;
  c.py = oy + c.dy * zp;
  c.pz = oz + c.dz * zp;
  if (c.valid && !a.coarse && write_outputs) {
    const size_t gp = (size_t)ray * a.S + si;
    const float nrm = sqrtf(c.px * c.px + c.py * c.py + c.pz * c.pz);
    if (a.pts) {
      a.pts[gp * 3 + 0] = c.px;
      a.pts[gp * 3 + 1] = c.py;
      a.pts[gp * 3 + 2] = c.pz;
    }
    if (a.mid_z) a.mid_z[gp] = c.mid;
    if (a.pts_norm) a.pts_norm[gp] = nrm;
    if (a.inside_sphere) a.inside_sphere[gp] = nrm < 1.0f ? 1.f : 0.f;
    if (a.z_out) a.z_out[gp] = z0;
  }
  return c;
}

// NeuS section opacity and the per-point outputs (renderer.py:261-286).  rgb_pre = W_rgb h_c (bias not yet added).
// `tail_out` (optional): {alpha, r, g, b} handed back for in-kernel compositing; the opacity is then NOT stored.
__device__ __forceinline__ void point_tail(const RenderKArgs& a, const PointCtx& c, const float* __restrict__ cst,
                                           float sdf, float gx, float gy, float gz, const float (&rgb_pre)[3],
                                           float* tail_out = nullptr) {
  if (!c.valid) return;
  const size_t gp = (size_t)c.ray * a.S + c.si;
  const float inv_s = cst[BlobLayout::kScalars + 4];
  const float r = sigmoidf_acc(rgb_pre[0] + cst[BlobLayout::kScalars + 1]);
  const float g = sigmoidf_acc(rgb_pre[1] + cst[BlobLayout::kScalars + 2]);
  const float b = sigmoidf_acc(rgb_pre[2] + cst[BlobLayout::kScalars + 3]);
  const float true_cos = c.dx * gx + c.dy * gy + c.dz * gz;
  const float iter_cos =
      -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - a.cos_anneal) + fmaxf(-true_cos, 0.f) * a.cos_anneal);
  const float half_step = iter_cos * c.dist * 0.5f;
  const float prev_cdf = sigmoidf_acc((sdf - half_step) * inv_s);
  const float next_cdf = sigmoidf_acc((sdf + half_step) * inv_s);
  float alpha = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
  alpha = fminf(fmaxf(alpha, 0.f), 1.f);
  if (a.sdf) a.sdf[gp] = sdf;
  if (a.cdf_fine) a.cdf_fine[gp] = prev_cdf;
  if (tail_out) {
    tail_out[0] = alpha;
    tail_out[1] = r;
    tail_out[2] = g;
    tail_out[3] = b;
  } else {
    a.alpha[gp] = alpha;
  }
  if (a.gradients) {
    a.gradients[gp * 3 + 0] = gx;
    a.gradients[gp * 3 + 1] = gy;
    a.gradients[gp * 3 + 2] = gz;
  }
  if (a.raw_color) {
    a.raw_color[gp * 3 + 0] = r;
    a.raw_color[gp * 3 + 1] = g;
    a.raw_color[gp * 3 + 2] = b;
  }
}

}  // namespace oi
