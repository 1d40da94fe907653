// The two callers either side of the render path:
//   gen_rays_kernel     Generator.gen_rays_at + build_rays + near_far_from_sphere (generator.py:255-279,317-342)
//   render_maps_kernel  Generator.render_maps with the Phong light (generator.py:80-174; lighting.py:126-225)
// Both are HBM/latency-bound streaming kernels: coalesced reads of the renderer's per-point outputs, one pass.
#include "oi_internal.cuh"

namespace oi {

namespace {

__device__ __forceinline__ float linspace01(int i, int n) {  // torch.linspace(0, 1, n)[i] in fp32 (two-sided form)
  const float step = 1.0f / (float)(n - 1);
  return (i < n / 2) ? step * (float)i : 1.0f - step * (float)(n - 1 - i);
}

__global__ void gen_rays_kernel(const OiGenRaysDesc d) {
  const int P = d.resolution, PP = P * P;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.n_instances * PP) return;
  const int b = r / PP, pix = r - b * PP, hh = pix / P, ww = pix - hh * P;
  const float* b2w = d.b2w + b * 16;
  const float* c2b = d.c2b + b * 16;
  // b2c translation = (w2c @ b2w)[:3, 3]
  float t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    t[i] = d.w2c[i * 4 + 0] * b2w[3] + d.w2c[i * 4 + 1] * b2w[7] + d.w2c[i * 4 + 2] * b2w[11] + d.w2c[i * 4 + 3] * b2w[15];
  const float res = (float)P, sres = (float)d.scene_resolution;
  const float cx = d.cam_dist / t[2] * t[0] * res / 2.0f + 0.5f * sres;
  const float cy = d.cam_dist / t[2] * t[1] * res / 2.0f + 0.5f * sres;
  const float xo = cx - res / 2.0f, yo = cy - res / 2.0f;
  if (pix == 0) {
    if (d.x_offset) d.x_offset[b] = xo;
    if (d.y_offset) d.y_offset[b] = yo;
  }
  const float px = linspace01(ww, P) * res + xo;
  const float py = linspace01(hh, P) * res + yo;
  const float* K = d.intrinsics_inv;
  float v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) v[i] = K[i * 4 + 0] * px + K[i * 4 + 1] * py + K[i * 4 + 2];
  const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] *= inv;
  v[1] *= inv;
  v[2] *= inv;
  float dd[3], o[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dd[i] = c2b[i * 4 + 0] * v[0] + c2b[i * 4 + 1] * v[1] + c2b[i * 4 + 2] * v[2];
    o[i] = c2b[i * 4 + 3];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d.rays_d[(size_t)r * 3 + i] = dd[i];
    d.rays_o[(size_t)r * 3 + i] = o[i];
  }
  if (d.near && d.far) {  // near_far_from_sphere (generator.py:336-342)
    const float aa = dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2];
    const float bb = 2.0f * (o[0] * dd[0] + o[1] * dd[1] + o[2] * dd[2]);
    const float mid = 0.5f * (-bb) / aa;
    d.near[r] = mid - 1.0f;
    d.far[r] = mid + 1.0f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct LightParams {
  float amb[3], dif[3], spe[3], shin;
};
__device__ __forceinline__ LightParams load_light(const OiRenderMapsDesc& d) {
  LightParams p;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    p.amb[c] = d.light_params ? d.light_params[c] : d.ambient_color[c];
    p.dif[c] = d.light_params ? d.light_params[3 + c] : d.diffuse_color[c];
    p.spe[c] = d.light_params ? d.light_params[6 + c] : d.specular_color[c];
  }
  p.shin = d.light_params ? d.light_params[9] : d.shininess;
  return p;
}

// One warp per ray; lanes stride over the S samples.
__global__ void render_maps_kernel(const OiRenderMapsDesc d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int ray = blockIdx.x * wpb + (threadIdx.x >> 5);
  if (ray >= d.n_rays) return;
  const int PP = d.rays_per_instance, S = d.n_samples;
  const int b = ray / PP, pix = ray - b * PP;
  float L[3] = {d.light_dir[b * 3], d.light_dir[b * 3 + 1], d.light_dir[b * 3 + 2]};
  {
    const float n = fmaxf(sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]), 1e-6f);  // F.normalize eps
    L[0] /= n;
    L[1] /= n;
    L[2] /= n;
  }
  const LightParams lp = load_light(d);
  const float ox = d.rays_o[(size_t)ray * 3], oy = d.rays_o[(size_t)ray * 3 + 1], oz = d.rays_o[(size_t)ray * 3 + 2];
  float acc_sh[3] = {0, 0, 0}, acc_ns[3] = {0, 0, 0}, acc_sp[3] = {0, 0, 0}, acc_df[3] = {0, 0, 0}, acc_n[3] = {0, 0, 0};
  float acc_w = 0.f, acc_z = 0.f, zmin = 3.0e38f;
  for (int s = lane; s < S; s += 32) {
    const size_t gp = (size_t)ray * S + s;
    const float w = d.weights[gp];
    const float nx = d.gradients[gp * 3], ny = d.gradients[gp * 3 + 1], nz = d.gradients[gp * 3 + 2];
    const float nn = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-6f);
    const float ux = nx / nn, uy = ny / nn, uz = nz / nn;
    const float cosv = ux * L[0] + uy * L[1] + uz * L[2];
    const float ang = fmaxf(cosv, 0.f);
    // specular (lighting.py:205-225)
    const float px = d.pts[gp * 3], py = d.pts[gp * 3 + 1], pz = d.pts[gp * 3 + 2];
    float vx = ox - px, vy = oy - py, vz = oz - pz;
    const float vn = fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-6f);
    vx /= vn;
    vy /= vn;
    vz /= vn;
    const float rx = -L[0] + 2.0f * (cosv * ux), ry = -L[1] + 2.0f * (cosv * uy), rz = -L[2] + 2.0f * (cosv * uz);
    const float al = fmaxf(vx * rx + vy * ry + vz * rz, 0.f) * (cosv > 0.f ? 1.f : 0.f);
    const float sp = powf(al, lp.shin);
    acc_w += w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float diffc = lp.dif[c] * ang;
      const float shade = lp.amb[c] + diffc;
      acc_df[c] = fmaf(diffc, w, acc_df[c]);
      acc_sh[c] = fmaf(shade, w, acc_sh[c]);
      acc_ns[c] = fmaf(shade * d.raw_color[gp * 3 + c], w, acc_ns[c]);
      acc_sp[c] = fmaf(lp.spe[c] * sp, w, acc_sp[c]);
    }
    acc_n[0] = fmaf(nx, w, acc_n[0]);
    acc_n[1] = fmaf(ny, w, acc_n[1]);
    acc_n[2] = fmaf(nz, w, acc_n[2]);
    if (d.mid_z_vals) {
      const float z = d.mid_z_vals[gp];
      acc_z = fmaf(z, w, acc_z);
      zmin = fminf(zmin, z);
    }
  }
  acc_w = warp_sum(acc_w);
  acc_z = warp_sum(acc_z);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    acc_sh[c] = warp_sum(acc_sh[c]);
    acc_ns[c] = warp_sum(acc_ns[c]);
    acc_sp[c] = warp_sum(acc_sp[c]);
    acc_df[c] = warp_sum(acc_df[c]);
    acc_n[c] = warp_sum(acc_n[c]);
  }
  if (lane != 0) return;
  const float ws = d.weight_sum[ray];
  auto put3 = [&](float* dst, const float (&v)[3]) {
    if (!dst) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[((size_t)b * 3 + c) * PP + pix] = v[c];
  };
  auto put1 = [&](float* dst, float v) {
    if (dst) dst[(size_t)b * PP + pix] = v;
  };
  float rgb[3], img[3], amb[3], col[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    rgb[c] = acc_ns[c] + acc_sp[c];
    img[c] = rgb[c] + d.bg_color[b * 3 + c] * (1.0f - ws);
    amb[c] = lp.amb[c] * acc_w;
    col[c] = d.color_fine[(size_t)ray * 3 + c];
  }
  put3(d.image, img);
  put3(d.image_no_bg, rgb);
  put3(d.shading_map, acc_sh);
  put3(d.color_map, col);
  put1(d.weight_sum_map, ws);
  put1(d.mask, fminf(fmaxf(ws, 1e-3f), 1.0f - 1e-3f));
  put3(d.amb_shading_map, amb);
  put3(d.diff_shading_map, acc_df);
  put3(d.normal_map, acc_n);
  put3(d.no_specular_map, acc_ns);
  put3(d.specular_map, acc_sp);
  put1(d.z_map, acc_z);
  if (d.z_min_per_ray) d.z_min_per_ray[ray] = zmin;
}


// Reverse mode of render_maps_kernel, same decomposition (one warp per ray, lanes stride over the samples).
// Per sample:  A_c = amb_c + dif_c ang,  ang = max(u.L, 0),  u = n / |n|,  r = -L + 2 (u.L) u,
//              al = max(v.r, 0) [u.L > 0],  sp = al^shin;   maps = sum_s w_s * {A_c, A_c col_c, spe_c sp, dif_c ang,
//              amb_c, n_c, mid_z}.  The light's parameters and direction collect block / warp partial sums.
constexpr int kMapsBwdWarps = 8;
__global__ void __launch_bounds__(kMapsBwdWarps * 32) render_maps_bwd_kernel(const OiRenderMapsBwdDesc bd) {
  const OiRenderMapsDesc& d = bd.fwd;
  __shared__ float red[kMapsBwdWarps][10];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ray = blockIdx.x * kMapsBwdWarps + wib;
  const int PP = d.rays_per_instance, S = d.n_samples;
  float gl[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // d amb[3], dif[3], spe[3], shin (this lane's partial sums)
  if (ray < d.n_rays) {
    const int b = ray / PP, pix = ray - b * PP;
    const LightParams lp = load_light(d);
    float Lr[3] = {d.light_dir[b * 3], d.light_dir[b * 3 + 1], d.light_dir[b * 3 + 2]};
    const float Ln = fmaxf(sqrtf(Lr[0] * Lr[0] + Lr[1] * Lr[1] + Lr[2] * Lr[2]), 1e-6f);
    const float L[3] = {Lr[0] / Ln, Lr[1] / Ln, Lr[2] / Ln};
    auto get3 = [&](const float* src, float (&v)[3]) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = src ? src[((size_t)b * 3 + c) * PP + pix] : 0.f;
    };
    auto get1 = [&](const float* src) { return src ? src[(size_t)b * PP + pix] : 0.f; };
    float g_img[3], g_nobg[3], g_sh[3], g_col[3], g_amb[3], g_df[3], g_n[3], g_ns[3], g_sp[3];
    get3(bd.g_image, g_img);
    get3(bd.g_image_no_bg, g_nobg);
    get3(bd.g_shading_map, g_sh);
    get3(bd.g_color_map, g_col);
    get3(bd.g_amb_shading_map, g_amb);
    get3(bd.g_diff_shading_map, g_df);
    get3(bd.g_normal_map, g_n);
    get3(bd.g_no_specular_map, g_ns);
    get3(bd.g_specular_map, g_sp);
    const float g_z = get1(bd.g_z_map);
    float ns_bar[3], sp_bar[3];   // adjoints of the composited no-specular / specular colours
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float rgb_bar = g_img[c] + g_nobg[c];
      ns_bar[c] = rgb_bar + g_ns[c];
      sp_bar[c] = rgb_bar + g_sp[c];
    }
    const float ox = d.rays_o[(size_t)ray * 3], oy = d.rays_o[(size_t)ray * 3 + 1], oz = d.rays_o[(size_t)ray * 3 + 2];
    float dL[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
    for (int s = lane; s < S; s += 32) {
      const size_t gp = (size_t)ray * S + s;
      const float w = d.weights[gp];
      const float n[3] = {d.gradients[gp * 3], d.gradients[gp * 3 + 1], d.gradients[gp * 3 + 2]};
      const float col[3] = {d.raw_color[gp * 3], d.raw_color[gp * 3 + 1], d.raw_color[gp * 3 + 2]};
      const float nlen = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      const float nn = fmaxf(nlen, 1e-6f);
      const float u[3] = {n[0] / nn, n[1] / nn, n[2] / nn};
      const float cosv = u[0] * L[0] + u[1] * L[1] + u[2] * L[2];
      const float ang = fmaxf(cosv, 0.f);
      float v[3] = {ox - d.pts[gp * 3], oy - d.pts[gp * 3 + 1], oz - d.pts[gp * 3 + 2]};
      const float vn = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-6f);
      v[0] /= vn;
      v[1] /= vn;
      v[2] /= vn;
      const float r[3] = {-L[0] + 2.0f * (cosv * u[0]), -L[1] + 2.0f * (cosv * u[1]), -L[2] + 2.0f * (cosv * u[2])};
      const float vr = v[0] * r[0] + v[1] * r[1] + v[2] * r[2];
      const float lit = cosv > 0.f ? 1.f : 0.f;
      const float al = fmaxf(vr, 0.f) * lit;
      const float sp = powf(al, lp.shin);
      wsum += w;
      // ---- adjoint of the weight, of the albedo, and of (ang, sp); light colour sums
      float w_bar = g_z * (d.mid_z_vals ? d.mid_z_vals[gp] : 0.f);
      float ang_bar = 0.f, sp_bar_s = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float A = lp.amb[c] + lp.dif[c] * ang;
        const float a_bar = g_sh[c] + ns_bar[c] * col[c];          // adjoint of w * A_c
        w_bar += a_bar * A + sp_bar[c] * lp.spe[c] * sp + g_df[c] * lp.dif[c] * ang + g_amb[c] * lp.amb[c] + g_n[c] * n[c];
        if (bd.d_raw_color) bd.d_raw_color[gp * 3 + c] = w * ns_bar[c] * A;
        ang_bar += w * lp.dif[c] * (a_bar + g_df[c]);
        sp_bar_s += w * sp_bar[c] * lp.spe[c];
        gl[c] += w * a_bar + g_amb[c] * w;                          // d ambient_c
        gl[3 + c] += w * ang * (a_bar + g_df[c]);                   // d diffuse_c
        gl[6 + c] += w * sp_bar[c] * sp;                            // d specular_c
      }
      if (bd.d_weights) bd.d_weights[gp] = w_bar;
      // ---- sp = al^shin
      float al_bar = 0.f;
      if (al > 0.f) {
        al_bar = sp_bar_s * lp.shin * powf(al, lp.shin - 1.0f);
        gl[9] += sp_bar_s * sp * logf(al);
      }
      const float vr_bar = (vr > 0.f ? al_bar : 0.f) * lit;
      // r = -L + 2 cosv u  ->  cosv, u, L
      const float rb[3] = {vr_bar * v[0], vr_bar * v[1], vr_bar * v[2]};
      float cos_bar = 2.0f * (rb[0] * u[0] + rb[1] * u[1] + rb[2] * u[2]) + (cosv > 0.f ? ang_bar : 0.f);
      float ub[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ub[c] = 2.0f * cosv * rb[c] + cos_bar * L[c];
        dL[c] += cos_bar * u[c] - rb[c];
      }
      // u = n / max(|n|, eps)
      float nb[3];
      if (nlen > 1e-6f) {
        const float uu = ub[0] * u[0] + ub[1] * u[1] + ub[2] * u[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) nb[c] = (ub[c] - uu * u[c]) / nn;
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) nb[c] = ub[c] / nn;
      }
      if (bd.d_gradients) {
#pragma unroll
        for (int c = 0; c < 3; ++c) bd.d_gradients[gp * 3 + c] = nb[c] + w * g_n[c];
      }
    }
    // ---- per-ray outputs and the light direction of this instance
#pragma unroll
    for (int c = 0; c < 3; ++c) dL[c] = warp_sum(dL[c]);
    if (lane == 0) {
      const float ws = d.weight_sum[ray];
      if (bd.d_weight_sum) {
        float g = get1(bd.g_weight_sum_map) + ((ws > 1e-3f && ws < 1.0f - 1e-3f) ? get1(bd.g_mask) : 0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) g -= g_img[c] * d.bg_color[b * 3 + c];
        bd.d_weight_sum[ray] = g;
      }
      if (bd.d_color_fine) {
#pragma unroll
        for (int c = 0; c < 3; ++c) bd.d_color_fine[(size_t)ray * 3 + c] = g_col[c];
      }
      if (bd.d_light_dir) {   // L = light_dir / max(|light_dir|, eps)
        const float ll = dL[0] * L[0] + dL[1] * L[1] + dL[2] * L[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(&bd.d_light_dir[b * 3 + c], (dL[c] - ll * L[c]) / Ln);
      }
    }
  }
  if (!bd.d_light_params) return;
#pragma unroll
  for (int i = 0; i < 10; ++i) gl[i] = warp_sum(gl[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 10; ++i) red[wib][i] = gl[i];
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    float v = 0.f;
    for (int w = 0; w < kMapsBwdWarps; ++w) v += red[w][threadIdx.x];
    atomicAdd(&bd.d_light_params[threadIdx.x], v);
  }
}

}  // namespace

int launch_gen_rays(const OiGenRaysDesc& d, cudaStream_t st) {
  const int n = d.n_instances * d.resolution * d.resolution;
  gen_rays_kernel<<<(n + 127) / 128, 128, 0, st>>>(d);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_render_maps(const OiRenderMapsDesc& d, cudaStream_t st) {
  const int wpb = 8;
  render_maps_kernel<<<(d.n_rays + wpb - 1) / wpb, wpb * 32, 0, st>>>(d);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_render_maps_bwd(const OiRenderMapsBwdDesc& d, cudaStream_t st) {
  const int bs = d.fwd.n_rays / d.fwd.rays_per_instance;
  if (d.d_light_params) OI_CHECK_CUDA(cudaMemsetAsync(d.d_light_params, 0, 10 * sizeof(float), st));
  if (d.d_light_dir) OI_CHECK_CUDA(cudaMemsetAsync(d.d_light_dir, 0, (size_t)bs * 3 * sizeof(float), st));
  render_maps_bwd_kernel<<<(d.fwd.n_rays + kMapsBwdWarps - 1) / kMapsBwdWarps, kMapsBwdWarps * 32, 0, st>>>(d);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
