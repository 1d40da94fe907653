// Small kernels around the fused render core: weight packing, style MLP, FiLM gamma/beta,
// hierarchical up-sampling (renderer.py:137-197 + 44-74) and per-ray compositing (renderer.py:300-338,448-455).
#include <cuda_fp16.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "oi_internal.cuh"

namespace oi {

namespace {

struct PackArgs {
  OiNetParams p;
};

// canonical K-major SWIZZLE_128B image of a [128 rows][64 k] fp16 block: row r occupies 128 bytes at r*128;
// its eight 16-byte chunks are XOR-permuted by (r & 7).  Returns the element index inside the 8192-element block.
__device__ __forceinline__ int sw128_index(int r, int k) {
  const int chunk = (k >> 3) ^ (r & 7);
  return r * 64 + chunk * 8 + (k & 7);
}

__global__ void pack_weights_kernel(const PackArgs a, float* __restrict__ blob) {
  const OiNetParams& p = a.p;
  const int D = p.depth;
  const BlobLayout L = blob_layout(D);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;

  // ---- stream section
  const size_t stream_floats = (size_t)L.n_chunks_stream * kChunkFloats;
  for (size_t i = tid; i < stream_floats; i += nthreads) {
    const int chunk = (int)(i / kChunkFloats);
    const int r = (int)(i % kChunkFloats) / kW;  // row in chunk
    const int n = (int)(i % kW);
    float v = 0.f;
    if (chunk == 0) {
      if (r < 3) v = p.pts_weight[0][n * 3 + r];
    } else {
      int c = chunk - 1;
      const int k = (c % 8) * kKC + r;
      const int seg = c / 8;  // 0..D-2 forward, D-1 colour, D..2D-2 reverse, 2D-1 colour as stored (backward)
      if (seg == 2 * D - 1) {
        v = p.views_weight[k * (kW + 3) + n];
      } else if (seg < D - 1) {
        v = p.pts_weight[seg + 1][n * kW + k];  // W_l^T
      } else if (seg == D - 1) {
        v = p.views_weight[n * (kW + 3) + k];
      } else {
        const int l = (D - 1) - (seg - D);  // D-1 .. 1
        v = p.pts_weight[l][k * kW + n];    // W_l as stored: row = output channel
      }
    }
    blob[L.stream_off + i] = v;
  }
  // ---- const section
  float* cst = blob + L.const_off;
  for (size_t i = tid; i < (size_t)BlobLayout::kConstFloats; i += nthreads) {
    const int ii = (int)i;
    float v = 0.f;
    if (ii < BlobLayout::kWsig) {
      const int l = ii / kW, n = ii % kW;
      if (l < D) v = p.pts_bias[l][n];
      else if (l == OI_MAX_DEPTH) v = p.views_bias[n];
    } else if (ii < BlobLayout::kWcg) {
      v = p.sigma_weight[ii - BlobLayout::kWsig];
    } else if (ii < BlobLayout::kWrgb) {
      const int j = (ii - BlobLayout::kWcg) / kW, n = (ii - BlobLayout::kWcg) % kW;
      v = p.views_weight[n * (kW + 3) + kW + j];
    } else if (ii < BlobLayout::kW0t) {
      v = p.rgb_weight[ii - BlobLayout::kWrgb];
    } else if (ii < BlobLayout::kScalars) {
      const int j = (ii - BlobLayout::kW0t) / kW, n = (ii - BlobLayout::kW0t) % kW;
      v = p.pts_weight[0][n * 3 + j];
    } else {
      const int s = ii - BlobLayout::kScalars;
      // inv_s = clip(exp(10 * variance), 1e-6, 1e6): neus/models/fields.py:267-268, renderer.py:266
      const float inv_s = fminf(fmaxf(expf(p.variance[0] * 10.0f), 1e-6f), 1e6f);
      if (s == 0) v = p.sigma_bias[0];
      else if (s >= 1 && s <= 3) v = p.rgb_bias[s - 1];
      else if (s == 4) v = inv_s;
      else if (s == 5) v = 1.0f / inv_s;
    }
    cst[i] = v;
  }
  // ---- film section (copied so that a render call needs only the blob)
  float* fl = blob + L.film_off;
  for (size_t i = tid; i < (size_t)BlobLayout::kFilmFloats; i += nthreads) {
    const int ii = (int)i;
    float v = 0.f;
    int rel, l;
    if (ii < BlobLayout::kGammaB) {
      rel = ii;
      l = rel / (kW * kStyle);
      const float* src = (l < D) ? p.gamma_weight[l] : (l == OI_MAX_DEPTH ? p.gamma_weight[OI_MAX_DEPTH] : nullptr);
      if (src) v = src[rel % (kW * kStyle)];
    } else if (ii < BlobLayout::kBetaW) {
      rel = ii - BlobLayout::kGammaB;
      l = rel / kW;
      const float* src = (l < D) ? p.gamma_bias[l] : (l == OI_MAX_DEPTH ? p.gamma_bias[OI_MAX_DEPTH] : nullptr);
      if (src) v = src[rel % kW];
    } else if (ii < BlobLayout::kBetaB) {
      rel = ii - BlobLayout::kBetaW;
      l = rel / (kW * kStyle);
      const float* src = (l < D) ? p.beta_weight[l] : (l == OI_MAX_DEPTH ? p.beta_weight[OI_MAX_DEPTH] : nullptr);
      if (src) v = src[rel % (kW * kStyle)];
    } else {
      rel = ii - BlobLayout::kBetaB;
      l = rel / kW;
      const float* src = (l < D) ? p.beta_bias[l] : (l == OI_MAX_DEPTH ? p.beta_bias[OI_MAX_DEPTH] : nullptr);
      if (src) v = src[rel % kW];
    }
    fl[i] = v;
  }
  // ---- tcgen05 section: fp16 {hi, lo} split of each 128x128 panel, B operand = [n][k] K-major SW128.
  // Weights are pre-scaled by 2^kTcWShift so that hi AND lo stay in fp16's normal range.
  __half* tc = reinterpret_cast<__half*>(blob + L.tc_off);
  const int n_panels = 2 * (D - 1) + 1;
  const size_t tc_elems = (size_t)n_panels * kW * kW;
  for (size_t i = tid; i < tc_elems; i += nthreads) {
    const int panel = (int)(i / (kW * kW));
    const int n = (int)(i % (kW * kW)) / kW;  // MMA N index (output column of D = A * B^T)
    const int k = (int)(i % kW);              // MMA K index
    float v;
    if (panel < D - 1) {
      v = p.pts_weight[panel + 1][n * kW + k];  // forward: N = out channel, K = in channel
    } else if (panel == D - 1) {
      v = p.views_weight[n * (kW + 3) + k];
    } else {
      const int l = (D - 1) - (panel - D);
      v = p.pts_weight[l][k * kW + n];          // reverse: N = in channel, K = out channel
    }
    v *= 256.0f;  // kTcWShift = 8 (undone in the epilogue)
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    // panel image: [hi: kblock0 (8192) | kblock1 (8192)] [lo: kblock0 | kblock1]  (halves)
    __half* base = tc + (size_t)panel * (4 * 8192);
    const int kb = k >> 6, kk = k & 63;
    base[kb * 8192 + sw128_index(n, kk)] = hi;
    base[2 * 8192 + kb * 8192 + sw128_index(n, kk)] = lo;
  }
  // ---- bf16 section for the adjoint sweeps of the tcgen05 backward (unscaled: bf16 has fp32's exponent range).
  // Panel order: colour as stored (N = feature channel, K = colour channel), forward orientation l = 1..D-1
  // (t_bar_l = W_l g_bar_l), reverse orientation l = D-1..1 (h_bar_l = W_l^T u_bar_l).
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(blob + L.tcb_off);
  for (size_t i = tid; i < tc_elems; i += nthreads) {
    const int panel = (int)(i / (kW * kW));
    const int n = (int)(i % (kW * kW)) / kW;
    const int k = (int)(i % kW);
    float v;
    if (panel == 0) {
      v = p.views_weight[k * (kW + 3) + n];
    } else if (panel < D) {
      v = p.pts_weight[panel][n * kW + k];
    } else {
      const int l = (D - 1) - (panel - D);
      v = p.pts_weight[l][k * kW + n];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* base = tb + (size_t)panel * (4 * 8192);
    const int kb = k >> 6, kk = k & 63;
    base[kb * 8192 + sw128_index(n, kk)] = hi;
    base[2 * 8192 + kb * 8192 + sw128_index(n, kk)] = lo;
  }
}

// ShapeNetwork.style (fields.py:14-19): three 64x64 linears, bias + leaky_relu(0.2) * 1 fused.
__global__ void style_mlp_kernel(const OiNetParams p, const float* __restrict__ z, float* __restrict__ w) {
  __shared__ float buf[2][kStyle];
  const int b = blockIdx.x, t = threadIdx.x;
  buf[0][t] = z[b * kStyle + t];
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < 3; ++l) {
    const float* W = p.style_weight[l] + t * kStyle;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < kStyle; ++k) s = fmaf(W[k], buf[cur][k], s);
    s += p.style_bias[l][t];
    s = (s > 0.f) ? s : s * 0.2f;
    buf[cur ^ 1][t] = s;
    __syncthreads();
    cur ^= 1;
  }
  w[b * kStyle + t] = buf[cur][t];
}

// gamma = 15 (G w + g) + 30, beta = 0.25 (B w + c)   (volume_renderer.py:27-30,47-48,56-57)
__global__ void film_kernel(const float* __restrict__ blob, int depth, const float* __restrict__ style_w,
                            float* __restrict__ film, unsigned int* __restrict__ ticket) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *ticket = 0u;  // arm the compositing reduction
  const BlobLayout L = blob_layout(depth);
  const float* fl = blob + L.film_off;
  const int inst = blockIdx.y, l = blockIdx.x, n = threadIdx.x;
  __shared__ float ws[kStyle];
  if (n < kStyle) ws[n] = style_w[inst * kStyle + n];
  __syncthreads();
  const float* G = fl + BlobLayout::kGammaW + ((size_t)l * kW + n) * kStyle;
  const float* B = fl + BlobLayout::kBetaW + ((size_t)l * kW + n) * kStyle;
  float sg = 0.f, sb = 0.f;
#pragma unroll 8
  for (int k = 0; k < kStyle; ++k) {
    sg = fmaf(G[k], ws[k], sg);
    sb = fmaf(B[k], ws[k], sb);
  }
  sg += fl[BlobLayout::kGammaB + l * kW + n];
  sb += fl[BlobLayout::kBetaB + l * kW + n];
  float* out = film + ((size_t)inst * kFilm + l) * 2 * kW;
  const float gamma = 15.0f * sg + 30.0f, beta = 0.25f * sb;
  out[n] = gamma;
  out[kW + n] = beta;
  // second table for the tcgen05 core: (gamma', delta) with arg = gamma' * acc + delta, where acc is the
  // accumulator of the 2^8-scaled weight panels (layer 0 runs on the FMA pipe, unscaled)
  float2* tc_tab = reinterpret_cast<float2*>(film + (size_t)gridDim.y * kFilm * 2 * kW);
  const float bias = blob[L.const_off + BlobLayout::kBias + l * kW + n];
  // pair layout: [gamma'_{2p}, gamma'_{2p+1}, delta_{2p}, delta_{2p+1}] so that one 16-byte load feeds an FFMA2
  float* row = reinterpret_cast<float*>(tc_tab) + ((size_t)inst * kFilm + l) * kW * 2;
  row[(n >> 1) * 4 + (n & 1)] = (l == 0) ? gamma : gamma * (1.0f / 256.0f);
  row[(n >> 1) * 4 + 2 + (n & 1)] = fmaf(gamma, bias, beta);
}

// ---------------------------------------------------------------------------------------------------
// up_sample + sample_pdf + cat_z_vals for up_sample_steps == 1 (renderer.py:137-181, 44-74, 183-197).
// One thread per ray; all scans are sequential in the reference's order (cumprod, cumsum) so that the
// ill-conditioned inverse-CDF step sees the same roundings as the fp32 reference.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxCoarse = 256;
constexpr int kMaxFine = 256;

__global__ void upsample_kernel(int R, int n, int m, const float* __restrict__ rays_o,
                                const float* __restrict__ rays_d, const float* __restrict__ near,
                                const float* __restrict__ far, const float* __restrict__ t_rand,
                                const float* __restrict__ lin, const float* __restrict__ lin_fine,
                                const float* __restrict__ z_in, float inv_s,
                                const float* __restrict__ sdf_coarse, float* __restrict__ z_fine) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= R) return;
  float z[kMaxCoarse];
  float cdf[kMaxCoarse];  // cdf[0] = 0, cdf[i] = cumsum(pdf)[i-1], i < n
  const float ox = rays_o[ray * 3], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const float nr = near[ray], span = far[ray] - near[ray];
  const float jit = t_rand ? t_rand[ray] * 2.0f / (float)n : 0.f;
  if (z_in) {   // up-sampling step i > 0: the merged z of the previous step (renderer.py:400-413)
    for (int i = 0; i < n; ++i) z[i] = z_in[(size_t)ray * n + i];
  } else {
    for (int i = 0; i < n; ++i) {
      const float l = lin ? lin[i] : (float)i / (float)(n - 1);
      float zi = nr + span * l;
      if (t_rand) zi = zi + jit;
      z[i] = zi;
    }
  }
  // inv_s = 64 * 2^i in up-sampling step i (renderer.py:406)
  const float* sdf = sdf_coarse + (size_t)ray * n;
  // pass 1: weights (stored temporarily in cdf[1..n-1]) and their sum
  float T = 1.0f, prev_cos = 0.f, wsum = 0.f;
  float px = ox + dx * z[0], py = oy + dy * z[0], pz = oz + dz * z[0];
  float rad_prev = sqrtf(px * px + py * py + pz * pz);
  for (int i = 0; i < n - 1; ++i) {
    px = ox + dx * z[i + 1];
    py = oy + dy * z[i + 1];
    pz = oz + dz * z[i + 1];
    const float rad_next = sqrtf(px * px + py * py + pz * pz);
    const float inside = (rad_prev < 1.0f || rad_next < 1.0f) ? 1.f : 0.f;
    const float s0 = sdf[i], s1 = sdf[i + 1];
    const float mid_sdf = (s0 + s1) * 0.5f;
    const float dzv = z[i + 1] - z[i];
    const float cosv = (s1 - s0) / (dzv + 1e-5f);
    float cv = fminf(prev_cos, cosv);
    prev_cos = cosv;
    cv = fminf(fmaxf(cv, -1e3f), 0.0f) * inside;
    const float prev_esti = mid_sdf - cv * dzv * 0.5f;
    const float next_esti = mid_sdf + cv * dzv * 0.5f;
    const float prev_cdf = sigmoidf_acc(prev_esti * inv_s);
    const float next_cdf = sigmoidf_acc(next_esti * inv_s);
    const float alpha = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
    const float w = alpha * T + 1e-5f;       // weights + 1e-5 (renderer.py:47)
    T = T * (1.0f - alpha + 1e-7f);          // exclusive cumprod (renderer.py:177-178)
    cdf[i + 1] = w;
    wsum += w;
    rad_prev = rad_next;
  }
  // pass 2: cdf = cumsum(w / sum)
  cdf[0] = 0.f;
  float run = 0.f;
  for (int i = 1; i < n; ++i) {
    run += cdf[i] / wsum;
    cdf[i] = run;
  }
  // pass 3: inverse-CDF samples (monotone in u) merged with the coarse z on the fly (sorted output)
  float* zo = z_fine + (size_t)ray * (n + m);
  int ci = 0, oi = 0, idx = 0;  // idx: searchsorted(right=True) cursor, monotone because u is increasing
  for (int j = 0; j < m; ++j) {
    const float u = lin_fine ? lin_fine[j] : (0.5f + (float)j) / (float)m;
    while (idx < n && cdf[idx] <= u) ++idx;  // number of cdf entries <= u
    const int below = max(idx - 1, 0), above = min(idx, n - 1);
    float denom = cdf[above] - cdf[below];
    if (denom < 1e-5f) denom = 1.0f;
    const float t = (u - cdf[below]) / denom;
    const float zs = z[below] + t * (z[above] - z[below]);
    while (ci < n && z[ci] <= zs) zo[oi++] = z[ci++];
    zo[oi++] = zs;
  }
  while (ci < n) zo[oi++] = z[ci++];
}

// ---------------------------------------------------------------------------------------------------
// Per-ray compositing.  One warp per ray.  In: alpha (in the `weights` buffer), raw_color, gradients,
// pts_norm, sdf.  Out: weights = alpha * exclusive_cumprod(1 - alpha + 1e-7) (renderer.py:300),
// weight_sum/max, color_fine (:304), s_val (:449) and the two global scalars gradient_error (:309-311)
// and surface_loss (:338), reduced deterministically (per-ray partials, last block sums in fixed order).
// ---------------------------------------------------------------------------------------------------
__global__ void composite_kernel(int R, int S, const float* __restrict__ blob, int depth, float* __restrict__ weights,
                                 const float* __restrict__ raw_color, const float* __restrict__ gradients,
                                 const float* __restrict__ pts_norm, const float* __restrict__ sdf,
                                 float* __restrict__ weight_sum, float* __restrict__ weight_max,
                                 float* __restrict__ color_fine, float* __restrict__ s_val,
                                 float* __restrict__ gradient_error, float* __restrict__ surface_loss,
                                 float* __restrict__ partials, unsigned int* __restrict__ ticket) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int ray = blockIdx.x * wpb + wib;
  const BlobLayout L = blob_layout(depth);
  const float inv_s_recip = blob[L.const_off + BlobLayout::kScalars + 5];
  if (ray < R) {
    float carry = 1.0f;  // running exclusive product entering this 32-sample block
    float wsum = 0.f, wmax = -1e30f, cr = 0.f, cg = 0.f, cb = 0.f;
    float ge = 0.f, gcnt = 0.f, sl = 0.f;
    for (int base = 0; base < S; base += 32) {
      const int i = base + lane;
      const bool ok = i < S;
      const size_t gp = (size_t)ray * S + (ok ? i : S - 1);
      const float alpha = ok ? weights[gp] : 0.f;
      const float f = ok ? (1.0f - alpha + 1e-7f) : 1.0f;
      // inclusive product scan over the warp
      float incl = f;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl *= o;
      }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float w = alpha * (carry * excl);
      carry = carry * __shfl_sync(0xffffffffu, incl, 31);
      if (ok) {
        weights[gp] = w;
        wsum += w;
        wmax = fmaxf(wmax, w);
        if (raw_color) {
          cr = fmaf(raw_color[gp * 3 + 0], w, cr);
          cg = fmaf(raw_color[gp * 3 + 1], w, cg);
          cb = fmaf(raw_color[gp * 3 + 2], w, cb);
        }
        if (gradients && pts_norm) {
          const float gx = gradients[gp * 3], gy = gradients[gp * 3 + 1], gz = gradients[gp * 3 + 2];
          const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
          const float relax = pts_norm[gp] < 1.2f ? 1.f : 0.f;
          ge = fmaf(relax, e * e, ge);
          gcnt += relax;
        }
        if (sdf) sl += expf(-100.0f * fabsf(sdf[gp]));
      }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
      wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
      cr += __shfl_xor_sync(0xffffffffu, cr, d);
      cg += __shfl_xor_sync(0xffffffffu, cg, d);
      cb += __shfl_xor_sync(0xffffffffu, cb, d);
      ge += __shfl_xor_sync(0xffffffffu, ge, d);
      gcnt += __shfl_xor_sync(0xffffffffu, gcnt, d);
      sl += __shfl_xor_sync(0xffffffffu, sl, d);
    }
    if (lane == 0) {
      if (weight_sum) weight_sum[ray] = wsum;
      if (weight_max) weight_max[ray] = wmax;
      if (color_fine) {
        color_fine[ray * 3 + 0] = cr;
        color_fine[ray * 3 + 1] = cg;
        color_fine[ray * 3 + 2] = cb;
      }
      if (s_val) s_val[ray] = inv_s_recip;
      partials[(size_t)ray * 3 + 0] = ge;
      partials[(size_t)ray * 3 + 1] = gcnt;
      partials[(size_t)ray * 3 + 2] = sl;
    }
  }
  // last block reduces the per-ray partials in a fixed order
  __shared__ bool is_last;
  __shared__ double red[3][32];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    s0 += (double)__ldcg(partials + (size_t)r * 3 + 0);
    s1 += (double)__ldcg(partials + (size_t)r * 3 + 1);
    s2 += (double)__ldcg(partials + (size_t)r * 3 + 2);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, d);
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane == 0) {
    red[0][wib] = s0;
    red[1][wib] = s1;
    red[2][wib] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < wpb; ++i) {
      t0 += red[0][i];
      t1 += red[1][i];
      t2 += red[2][i];
    }
    if (gradient_error) *gradient_error = (float)(t0 / (t1 + 1e-5));
    if (surface_loss) *surface_loss = (float)(t2 / ((double)R * (double)S));
    *ticket = 0;  // self-reset so that the workspace can be reused without a memset
  }
}

}  // namespace

int launch_pack_weights(const OiNetParams* p, float* blob, cudaStream_t st) {
  PackArgs a;
  a.p = *p;
  pack_weights_kernel<<<296, 256, 0, st>>>(a, blob);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_style_mlp(const OiNetParams* p, const float* z, float* w, int n_inst, cudaStream_t st) {
  style_mlp_kernel<<<n_inst, kStyle, 0, st>>>(*p, z, w);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_film(const float* blob, int depth, const float* style_w, float* film, int n_inst, unsigned int* ticket,
                cudaStream_t st) {
  film_kernel<<<dim3(kFilm, n_inst), kW, 0, st>>>(blob, depth, style_w, film, ticket);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_upsample(int R, int n, int m, const float* rays_o, const float* rays_d, const float* near,
                    const float* far, const float* t_rand, const float* lin, const float* lin_fine,
                    const float* z_in, float inv_s, const float* sdf_coarse, float* z_fine, cudaStream_t st) {
  if (n > kMaxCoarse || m > kMaxFine)
    return set_error(OI_ERR_UNSUPPORTED, "at most %d samples entering an up-sampling step and %d new ones per step",
                     kMaxCoarse, kMaxFine);
  upsample_kernel<<<(R + 63) / 64, 64, 0, st>>>(R, n, m, rays_o, rays_d, near, far, t_rand, lin, lin_fine, z_in,
                                                inv_s, sdf_coarse, z_fine);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_composite(int R, int S, const float* blob, int depth, float* weights, const float* raw_color,
                     const float* gradients, const float* pts_norm, const float* sdf, float* weight_sum,
                     float* weight_max, float* color_fine, float* s_val, float* gradient_error,
                     float* surface_loss, float* partials, unsigned int* ticket, cudaStream_t st) {
  const int wpb = 8;
  composite_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, st>>>(R, S, blob, depth, weights, raw_color, gradients,
                                                             pts_norm, sdf, weight_sum, weight_max, color_fine,
                                                             s_val, gradient_error, surface_loss, partials, ticket);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
