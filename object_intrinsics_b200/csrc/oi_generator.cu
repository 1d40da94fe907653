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
    const float sp = powf(al, d.shininess);
    acc_w += w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float diffc = d.diffuse_color[c] * ang;
      const float shade = d.ambient_color[c] + diffc;
      acc_df[c] = fmaf(diffc, w, acc_df[c]);
      acc_sh[c] = fmaf(shade, w, acc_sh[c]);
      acc_ns[c] = fmaf(shade * d.raw_color[gp * 3 + c], w, acc_ns[c]);
      acc_sp[c] = fmaf(d.specular_color[c] * sp, w, acc_sp[c]);
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
    amb[c] = d.ambient_color[c] * acc_w;
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

}  // namespace oi
