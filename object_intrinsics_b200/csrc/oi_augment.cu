// Geometric path of the ADA AugmentPipe (src/third_party/ada/augment.py:270-301 of the reference) without the
// reference's intermediate tensors and without its device->host sync:
//     reflect-pad (margins)  ->  2x up-sample (sym6 FIR, gain 4)  ->  bilinear affine resample (zeros outside)
//     ->  2x down-sample (sym6 FIR, correlation) + crop
// The reference materialises the padded image, the up-sampled image, the sampling grid and the resampled image
// (augment.py:286-301) and reads the padding margins back to the host (`margin.ceil().to(torch.int32)`, :283).
// Here the margins stay on the device (a [4] int32 tensor read by the kernels), and the chain is two kernels:
//   augment_up_kernel       : U = upsample(reflect_pad(x)) evaluated polyphase straight from x (36 taps / pixel)
//   augment_resample_kernel : y = downsample(resample(U)) per output pixel: 12x12 FIR taps, each a bilinear lookup
// The map x -> y is linear, so the backward is its adjoint and the double backward (R1 regularisation,
// src/loss/gan.py:5-14) is the forward again:
//   augment_down_adj_kernel     : dR = downsample^T dy            (gather, 6x6 taps)
//   augment_resample_adj_kernel : dU += resample^T dR             (4 reductions per resampled pixel)
//   augment_up_adj_kernel       : dx = (upsample o reflect_pad)^T dU   (gather over the <= 3x3 reflected pre-images)
// Filters are separable with an even number of taps T (T = 12, hz_pad = T / 4 = 3 in the reference).
#include "oi_internal.cuh"

namespace oi {

namespace {

constexpr int kMaxTaps = 16;

struct AugArgs {
  int B, C, H, W, T;
  int HzPad;               // T / 4
  const float* theta;      // [B,2,3] affine_grid matrices (normalised output coords -> normalised U coords)
  const int* margins;      // [4] mx0, my0, mx1, my1
  float f[kMaxTaps];       // filter taps
  const float* x;          // forward: images; backward: dy
  float* y;                // forward: output; backward: dx
  float* U;                // workspace [B,C,Hu_max,Wu_max] with row stride Wu (actual), plane stride Hu*Wu (actual)
  float* R;                // backward workspace [B,C,Hr,Wr]
};

__device__ __forceinline__ int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// U[u] = 2 * sum_k f[T-1-k] z[u + k - p0], z[2i] = P[i]  (per dimension; p0 = (T + 1) / 2)
__global__ void augment_up_kernel(const AugArgs a) {
  const int mx0 = a.margins[0], my0 = a.margins[1], mx1 = a.margins[2], my1 = a.margins[3];
  const int Wp = a.W + mx0 + mx1, Hp = a.H + my0 + my1;
  const int Wu = 2 * Wp, Hu = 2 * Hp;
  const int p0 = (a.T + 1) / 2;   // (fw + up - 1) // 2
  const int nc = blockIdx.z;
  const float* src = a.x + (size_t)nc * a.H * a.W;
  float* dst = a.U + (size_t)nc * Hu * Wu;
  for (int uy = blockIdx.y * blockDim.y + threadIdx.y; uy < Hu; uy += gridDim.y * blockDim.y) {
    for (int ux = blockIdx.x * blockDim.x + threadIdx.x; ux < Wu; ux += gridDim.x * blockDim.x) {
      // taps with (u + k - p0) even: k = k0, k0 + 2, ...
      const int kx0 = (ux + p0) & 1, ky0 = (uy + p0) & 1;
      float acc = 0.f;
      for (int ky = ky0; ky < a.T; ky += 2) {
        const int zy = uy + ky - p0;
        if (zy < 0 || zy >= Hu) continue;
        const int iy = reflect((zy >> 1) - my0, a.H);
        const float fy = a.f[a.T - 1 - ky];
        float row = 0.f;
        for (int kx = kx0; kx < a.T; kx += 2) {
          const int zx = ux + kx - p0;
          if (zx < 0 || zx >= Wu) continue;
          const int ix = reflect((zx >> 1) - mx0, a.W);
          row = fmaf(a.f[a.T - 1 - kx], src[iy * a.W + ix], row);
        }
        acc = fmaf(fy, row, acc);
      }
      dst[(size_t)uy * Wu + ux] = 4.0f * acc;
    }
  }
}

struct SamplePos {
  int x0, y0;
  float wx, wy;
};
// source position in U of resampled pixel (rx, ry): affine_grid + grid_sample, align_corners = False
__device__ __forceinline__ SamplePos sample_pos(const float* th, int rx, int ry, int Wr, int Hr, int Wu, int Hu) {
  const float xn = (2.0f * rx + 1.0f) / Wr - 1.0f;
  const float yn = (2.0f * ry + 1.0f) / Hr - 1.0f;
  const float xs = th[0] * xn + th[1] * yn + th[2];
  const float ys = th[3] * xn + th[4] * yn + th[5];
  const float sx = ((xs + 1.0f) * Wu - 1.0f) * 0.5f;
  const float sy = ((ys + 1.0f) * Hu - 1.0f) * 0.5f;
  SamplePos p;
  const float fx = floorf(sx), fy = floorf(sy);
  p.x0 = (int)fx;
  p.y0 = (int)fy;
  p.wx = sx - fx;
  p.wy = sy - fy;
  return p;
}

// y[o] = sum_k f[k] R[2 o + k + c0], c0 = -((T - 1) / 2 - 2 hz_pad)  (downsample2d padding, correlation)
__global__ void augment_resample_kernel(const AugArgs a) {
  const int mx0 = a.margins[0], my0 = a.margins[1], mx1 = a.margins[2], my1 = a.margins[3];
  const int Wu = 2 * (a.W + mx0 + mx1), Hu = 2 * (a.H + my0 + my1);
  const int Wr = 2 * (a.W + 2 * a.HzPad), Hr = 2 * (a.H + 2 * a.HzPad);
  const int c0 = 2 * a.HzPad - (a.T - 1) / 2;   // R index of tap 0 for output 0 (= 1 for T = 12)
  const int n = blockIdx.z;
  const float* th = a.theta + n * 6;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= a.W || oy >= a.H) return;
  float acc[8];
  for (int c0_ = 0; c0_ < a.C; c0_ += 8) {
    const int cn = min(8, a.C - c0_);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int ky = 0; ky < a.T; ++ky) {
      const int ry = 2 * oy + ky + c0;
      for (int kx = 0; kx < a.T; ++kx) {
        const int rx = 2 * ox + kx + c0;
        const SamplePos p = sample_pos(th, rx, ry, Wr, Hr, Wu, Hu);
        const float wgt = a.f[ky] * a.f[kx];
        const bool x0in = p.x0 >= 0 && p.x0 < Wu, x1in = p.x0 + 1 >= 0 && p.x0 + 1 < Wu;
        const bool y0in = p.y0 >= 0 && p.y0 < Hu, y1in = p.y0 + 1 >= 0 && p.y0 + 1 < Hu;
        const float w00 = (1.f - p.wx) * (1.f - p.wy), w01 = p.wx * (1.f - p.wy);
        const float w10 = (1.f - p.wx) * p.wy, w11 = p.wx * p.wy;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < cn) {
            const float* u = a.U + ((size_t)(n * a.C + c0_ + c) * Hu) * Wu;
            float v = 0.f;
            if (y0in && x0in) v = fmaf(w00, u[(size_t)p.y0 * Wu + p.x0], v);
            if (y0in && x1in) v = fmaf(w01, u[(size_t)p.y0 * Wu + p.x0 + 1], v);
            if (y1in && x0in) v = fmaf(w10, u[(size_t)(p.y0 + 1) * Wu + p.x0], v);
            if (y1in && x1in) v = fmaf(w11, u[(size_t)(p.y0 + 1) * Wu + p.x0 + 1], v);
            acc[c] = fmaf(wgt, v, acc[c]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < cn) a.y[((size_t)(n * a.C + c0_ + c) * a.H + oy) * a.W + ox] = acc[c];
  }
}

// dR[r] = sum over (o, k) with 2 o + k + c0 = r of f[k] dy[o]
__global__ void augment_down_adj_kernel(const AugArgs a) {
  const int Wr = 2 * (a.W + 2 * a.HzPad), Hr = 2 * (a.H + 2 * a.HzPad);
  const int c0 = 2 * a.HzPad - (a.T - 1) / 2;
  const int nc = blockIdx.z;
  const float* dy = a.x + (size_t)nc * a.H * a.W;
  const int rx = blockIdx.x * blockDim.x + threadIdx.x, ry = blockIdx.y * blockDim.y + threadIdx.y;
  if (rx >= Wr || ry >= Hr) return;
  float acc = 0.f;
  for (int ky = (ry - c0) & 1; ky < a.T; ky += 2) {
    const int oy = (ry - c0 - ky) / 2;
    if (ry - c0 - ky < 0 || oy >= a.H) continue;
    float row = 0.f;
    for (int kx = (rx - c0) & 1; kx < a.T; kx += 2) {
      const int ox = (rx - c0 - kx) / 2;
      if (rx - c0 - kx < 0 || ox >= a.W) continue;
      row = fmaf(a.f[kx], dy[oy * a.W + ox], row);
    }
    acc = fmaf(a.f[ky], row, acc);
  }
  a.R[((size_t)nc * Hr + ry) * Wr + rx] = acc;
}

// dU += resample^T dR (U must be zeroed)
__global__ void augment_resample_adj_kernel(const AugArgs a) {
  const int mx0 = a.margins[0], my0 = a.margins[1], mx1 = a.margins[2], my1 = a.margins[3];
  const int Wu = 2 * (a.W + mx0 + mx1), Hu = 2 * (a.H + my0 + my1);
  const int Wr = 2 * (a.W + 2 * a.HzPad), Hr = 2 * (a.H + 2 * a.HzPad);
  const int n = blockIdx.z;
  const float* th = a.theta + n * 6;
  const int rx = blockIdx.x * blockDim.x + threadIdx.x, ry = blockIdx.y * blockDim.y + threadIdx.y;
  if (rx >= Wr || ry >= Hr) return;
  const SamplePos p = sample_pos(th, rx, ry, Wr, Hr, Wu, Hu);
  const bool x0in = p.x0 >= 0 && p.x0 < Wu, x1in = p.x0 + 1 >= 0 && p.x0 + 1 < Wu;
  const bool y0in = p.y0 >= 0 && p.y0 < Hu, y1in = p.y0 + 1 >= 0 && p.y0 + 1 < Hu;
  const float w00 = (1.f - p.wx) * (1.f - p.wy), w01 = p.wx * (1.f - p.wy);
  const float w10 = (1.f - p.wx) * p.wy, w11 = p.wx * p.wy;
  for (int c = 0; c < a.C; ++c) {
    const float g = a.R[((size_t)(n * a.C + c) * Hr + ry) * Wr + rx];
    float* u = a.U + ((size_t)(n * a.C + c) * Hu) * Wu;
    if (y0in && x0in) atomicAdd(u + (size_t)p.y0 * Wu + p.x0, w00 * g);
    if (y0in && x1in) atomicAdd(u + (size_t)p.y0 * Wu + p.x0 + 1, w01 * g);
    if (y1in && x0in) atomicAdd(u + (size_t)(p.y0 + 1) * Wu + p.x0, w10 * g);
    if (y1in && x1in) atomicAdd(u + (size_t)(p.y0 + 1) * Wu + p.x0 + 1, w11 * g);
  }
}

// dx[i] = sum over padded positions p with reflect(p - m0) = i of dP[p];
// dP[p] = 4 sum_k f[T-1-ky] f[T-1-kx] dU[2 py + p0 - ky][2 px + p0 - kx]
__global__ void augment_up_adj_kernel(const AugArgs a) {
  const int mx0 = a.margins[0], my0 = a.margins[1], mx1 = a.margins[2], my1 = a.margins[3];
  const int Wp = a.W + mx0 + mx1, Hp = a.H + my0 + my1;
  const int Wu = 2 * Wp, Hu = 2 * Hp;
  const int p0 = (a.T + 1) / 2;
  const int nc = blockIdx.z;
  const float* dU = a.U + (size_t)nc * Hu * Wu;
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ix >= a.W || iy >= a.H) return;
  // pre-images of i under the reflect padding: i + m0; m0 - i (i >= 1, left margin); 2(n-1) - i + m0 (i <= n-2, right)
  int pys[3], pxs[3], ny = 0, nx = 0;
  pys[ny++] = iy + my0;
  if (iy >= 1 && iy <= my0) pys[ny++] = my0 - iy;
  if (iy <= a.H - 2 && (a.H - 1 - iy) <= my1) pys[ny++] = 2 * (a.H - 1) - iy + my0;
  pxs[nx++] = ix + mx0;
  if (ix >= 1 && ix <= mx0) pxs[nx++] = mx0 - ix;
  if (ix <= a.W - 2 && (a.W - 1 - ix) <= mx1) pxs[nx++] = 2 * (a.W - 1) - ix + mx0;
  float acc = 0.f;
  for (int a_ = 0; a_ < ny; ++a_) {
    const int py = pys[a_];
    for (int b_ = 0; b_ < nx; ++b_) {
      const int px = pxs[b_];
      for (int ky = 0; ky < a.T; ++ky) {
        const int uy = 2 * py + p0 - ky;
        if (uy < 0 || uy >= Hu) continue;
        float row = 0.f;
        for (int kx = 0; kx < a.T; ++kx) {
          const int ux = 2 * px + p0 - kx;
          if (ux < 0 || ux >= Wu) continue;
          row = fmaf(a.f[a.T - 1 - kx], dU[(size_t)uy * Wu + ux], row);
        }
        acc = fmaf(a.f[a.T - 1 - ky], row, acc);
      }
    }
  }
  a.y[((size_t)nc * a.H + iy) * a.W + ix] = 4.0f * acc;
}

// Margins (augment.py:274-283) and affine_grid matrices (augment.py:287-297) from the inverse transforms: one
// block; replaces ~40 tiny torch launches and the reference's host read-back of the margins.
struct M3 {
  double m[9];
};
__device__ __forceinline__ M3 mul3(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
__device__ __forceinline__ M3 scale3(double sx, double sy) { return M3{{sx, 0, 0, 0, sy, 0, 0, 0, 1}}; }
__device__ __forceinline__ M3 trans3(double tx, double ty) { return M3{{1, 0, tx, 0, 1, ty, 0, 0, 1}}; }

struct AugOps {
  int n;
  OiAugmentOp op[OI_AUGMENT_MAX_OPS];
};

// G_inv[b] = prod_i M_i[b] in fp32, each product accumulated in the order of a 3x3 matmul (augment.py:196-264)
__device__ __forceinline__ void compose_ops(const AugOps& o, int b, float (&g)[9]) {
  g[0] = 1.f; g[1] = 0.f; g[2] = 0.f; g[3] = 0.f; g[4] = 1.f; g[5] = 0.f; g[6] = 0.f; g[7] = 0.f; g[8] = 1.f;
  for (int i = 0; i < o.n; ++i) {
    float m[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    const float p0 = o.op[i].p0[b];
    if (o.op[i].kind == 0) {
      m[0] = p0;
      m[4] = o.op[i].p1[b];
    } else if (o.op[i].kind == 1) {
      const float c = cosf(p0), s = sinf(p0);
      m[0] = c; m[1] = -s; m[3] = s; m[4] = c;
    } else {
      m[2] = p0;
      m[5] = o.op[i].p1[b];
    }
    float r[9];
    for (int rr = 0; rr < 3; ++rr)
      for (int cc = 0; cc < 3; ++cc)
        r[rr * 3 + cc] = fmaf(g[rr * 3 + 2], m[6 + cc], fmaf(g[rr * 3 + 1], m[3 + cc], g[rr * 3] * m[cc]));
    for (int k = 0; k < 9; ++k) g[k] = r[k];
  }
}

struct AugRawOps {
  int n;
  const float* p;
  OiAugmentRawOp op[OI_AUGMENT_MAX_OPS];
};

// g = g @ m (fp32, accumulation order of a 3x3 matmul)
__device__ __forceinline__ void mul3f(float (&g)[9], const float (&m)[9]) {
  float r[9];
  for (int rr = 0; rr < 3; ++rr)
    for (int cc = 0; cc < 3; ++cc)
      r[rr * 3 + cc] = fmaf(g[rr * 3 + 2], m[6 + cc], fmaf(g[rr * 3 + 1], m[3 + cc], g[rr * 3] * m[cc]));
  for (int k = 0; k < 9; ++k) g[k] = r[k];
}

// The reference's per-factor arithmetic on the raw draws (augment.py:196-264), every step rounded as torch rounds it
// (fp32 products of a tensor with a python scalar, IEEE division / square root, round-half-even).
__device__ __forceinline__ void compose_raw(const AugRawOps& o, int b, int H, int W, float (&g)[9]) {
  g[0] = 1.f; g[1] = 0.f; g[2] = 0.f; g[3] = 0.f; g[4] = 1.f; g[5] = 0.f; g[6] = 0.f; g[7] = 0.f; g[8] = 1.f;
  const float p = *o.p;
  const float kPi = 3.14159265358979323846f;
  for (int i = 0; i < o.n; ++i) {
    const OiAugmentRawOp& r = o.op[i];
    float m[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    float thr = __fmul_rn(r.prob, p);
    if (r.form == OI_AUG_ROTATE)
      thr = __fsub_rn(1.f, __fsqrt_rn(fminf(fmaxf(__fsub_rn(1.f, thr), 0.f), 1.f)));
    const bool on = r.gate[b] < thr;
    switch (r.form) {
      case OI_AUG_XFLIP: {
        const float k = on ? floorf(__fmul_rn(r.draw[b], 2.f)) : 0.f;
        m[0] = __fdiv_rn(1.f, __fsub_rn(1.f, __fmul_rn(2.f, k)));
        break;
      }
      case OI_AUG_ROTATE90:
      case OI_AUG_ROTATE: {
        float th;
        if (r.form == OI_AUG_ROTATE90) {
          const float k = on ? floorf(__fmul_rn(r.draw[b], 4.f)) : 0.f;
          th = -__fmul_rn(-0.5f * kPi, k);
        } else {
          th = on ? __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(r.draw[b], 2.f), 1.f), kPi), r.param) : 0.f;
        }
        const float c = cosf(th), sn = sinf(th);
        m[0] = c; m[1] = -sn; m[3] = sn; m[4] = c;
        break;
      }
      case OI_AUG_XINT: {
        const float t0 = on ? __fmul_rn(__fsub_rn(__fmul_rn(r.draw[2 * b], 2.f), 1.f), r.param) : 0.f;
        const float t1 = on ? __fmul_rn(__fsub_rn(__fmul_rn(r.draw[2 * b + 1], 2.f), 1.f), r.param) : 0.f;
        m[2] = -rintf(__fmul_rn(t0, (float)W));
        m[5] = -rintf(__fmul_rn(t1, (float)H));
        break;
      }
      case OI_AUG_SCALE:
      case OI_AUG_ANISO: {
        const float sc = on ? (float)exp2((double)__fmul_rn(r.draw[b], r.param)) : 1.f;
        const float inv = __fdiv_rn(1.f, sc);
        m[0] = inv;
        m[4] = (r.form == OI_AUG_SCALE) ? inv : __fdiv_rn(1.f, inv);
        break;
      }
      default: {   // OI_AUG_XFRAC
        const float t0 = on ? __fmul_rn(r.draw[2 * b], r.param) : 0.f;
        const float t1 = on ? __fmul_rn(r.draw[2 * b + 1], r.param) : 0.f;
        m[2] = -__fmul_rn(t0, (float)W);
        m[5] = -__fmul_rn(t1, (float)H);
        break;
      }
    }
    mul3f(g, m);
  }
}

// G_in != NULL: the transform is given; else it is composed from `ops` / `raw` (and written to G_out when given).
__global__ void augment_setup_kernel(const float* __restrict__ G_in, const AugOps ops, const AugRawOps raw,
                                     float* __restrict__ G_out,
                                     float* __restrict__ G_tmp, int B, int H, int W, int hz_pad,
                                     float* __restrict__ theta, int* __restrict__ margins) {
  __shared__ float red[4][32];
  __shared__ int marg[4];
  const int tid = threadIdx.x;   // 32 threads
  const float* G_inv = G_in;
  if (G_in == nullptr) {
    for (int b = tid; b < B; b += 32) {
      float g[9];
      if (raw.n > 0) compose_raw(raw, b, H, W, g);
      else compose_ops(ops, b, g);
      for (int k = 0; k < 9; ++k) {
        G_tmp[b * 9 + k] = g[k];
        if (G_out) G_out[b * 9 + k] = g[k];
      }
    }
    __syncthreads();
    G_inv = G_tmp;
  }
  const float cx = (W - 1) * 0.5f, cy = (H - 1) * 0.5f;
  float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;   // max(-x), max(-y), max(x), max(y)
  for (int b = tid; b < B; b += 32) {
    const float* g = G_inv + b * 9;
    const float xs[4] = {-cx, cx, cx, -cx}, ys[4] = {-cy, -cy, cy, cy};
    for (int k = 0; k < 4; ++k) {
      const float px = g[0] * xs[k] + g[1] * ys[k] + g[2];
      const float py = g[3] * xs[k] + g[4] * ys[k] + g[5];
      m0 = fmaxf(m0, -px);
      m1 = fmaxf(m1, -py);
      m2 = fmaxf(m2, px);
      m3 = fmaxf(m3, py);
    }
  }
  red[0][tid] = m0;
  red[1][tid] = m1;
  red[2][tid] = m2;
  red[3][tid] = m3;
  __syncthreads();
  if (tid < 4) {
    float v = -1e30f;
    for (int i = 0; i < 32; ++i) v = fmaxf(v, red[tid][i]);
    v += (tid & 1) ? (hz_pad * 2 - cy) : (hz_pad * 2 - cx);
    v = fminf(fmaxf(v, 0.f), (tid & 1) ? (float)(H - 1) : (float)(W - 1));
    marg[tid] = (int)ceilf(v);
    margins[tid] = marg[tid];
  }
  __syncthreads();
  const int mx0 = marg[0], my0 = marg[1], mx1 = marg[2], my1 = marg[3];
  const double Wu = 2.0 * (W + mx0 + mx1), Hu = 2.0 * (H + my0 + my1);
  const double Wr = 2.0 * (W + 2 * hz_pad), Hr = 2.0 * (H + 2 * hz_pad);
  for (int b = tid; b < B; b += 32) {
    M3 G;
    for (int i = 0; i < 9; ++i) G.m[i] = G_inv[b * 9 + i];
    G = mul3(trans3((mx0 - mx1) * 0.5, (my0 - my1) * 0.5), G);
    G = mul3(mul3(scale3(2, 2), G), scale3(0.5, 0.5));
    G = mul3(mul3(trans3(-0.5, -0.5), G), trans3(0.5, 0.5));
    G = mul3(mul3(scale3(2.0 / Wu, 2.0 / Hu), G), scale3(Wr * 0.5, Hr * 0.5));
    for (int i = 0; i < 6; ++i) theta[b * 6 + i] = (float)G.m[i];
  }
}

int fill_args(const OiAugmentGeomDesc& d, AugArgs* a) {
  a->B = d.batch;
  a->C = d.channels;
  a->H = d.height;
  a->W = d.width;
  a->T = d.filter_taps;
  a->HzPad = d.filter_taps / 4;
  a->theta = d.theta;
  a->margins = d.margins;
  for (int i = 0; i < kMaxTaps; ++i) a->f[i] = i < d.filter_taps ? d.filter[i] : 0.f;
  a->x = d.x;
  a->y = d.y;
  return OI_OK;
}

}  // namespace

size_t augment_u_floats(const OiAugmentGeomDesc& d) {
  // margins are clamped to [0, n-1] per side (augment.py:282): padded size <= 3n - 2
  return (size_t)d.batch * d.channels * (2 * (size_t)(3 * d.height - 2)) * (2 * (size_t)(3 * d.width - 2));
}
size_t augment_r_floats(const OiAugmentGeomDesc& d) {
  const int hp = d.filter_taps / 4;
  return (size_t)d.batch * d.channels * (2 * (size_t)(d.height + 2 * hp)) * (2 * (size_t)(d.width + 2 * hp));
}

int launch_augment_setup(const float* G_inv, int B, int H, int W, int hz_pad, float* theta, int* margins,
                         cudaStream_t st) {
  AugOps none;
  none.n = 0;
  AugRawOps no_raw;
  no_raw.n = 0;
  augment_setup_kernel<<<1, 32, 0, st>>>(G_inv, none, no_raw, nullptr, nullptr, B, H, W, hz_pad, theta, margins);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

// `theta` ([B,2,3]) is followed in the caller's buffer by [B,3,3] floats of scratch for the composed transform
int launch_augment_setup_ops(const OiAugmentOp* ops, int n_ops, int B, int H, int W, int hz_pad, float* g_inv,
                             float* g_tmp, float* theta, int* margins, cudaStream_t st) {
  AugOps o;
  o.n = n_ops;
  for (int i = 0; i < n_ops; ++i) o.op[i] = ops[i];
  AugRawOps no_raw;
  no_raw.n = 0;
  augment_setup_kernel<<<1, 32, 0, st>>>(nullptr, o, no_raw, g_inv, g_tmp, B, H, W, hz_pad, theta, margins);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_augment_setup_raw(const OiAugmentRawOp* ops, int n_ops, const float* p, int B, int H, int W, int hz_pad,
                             float* g_inv, float* g_tmp, float* theta, int* margins, cudaStream_t st) {
  AugOps none;
  none.n = 0;
  AugRawOps r;
  r.n = n_ops;
  r.p = p;
  for (int i = 0; i < n_ops; ++i) r.op[i] = ops[i];
  augment_setup_kernel<<<1, 32, 0, st>>>(nullptr, none, r, g_inv, g_tmp, B, H, W, hz_pad, theta, margins);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_augment_geom(const OiAugmentGeomDesc& d, bool backward, cudaStream_t st) {
  AugArgs a;
  fill_args(d, &a);
  a.U = static_cast<float*>(d.workspace);
  a.R = a.U + augment_u_floats(d);
  const dim3 blk(32, 8);
  const int Hu_max = 2 * (3 * d.height - 2), Wu_max = 2 * (3 * d.width - 2);
  const int Hr = 2 * (d.height + 2 * a.HzPad), Wr = 2 * (d.width + 2 * a.HzPad);
  if (!backward) {
    // grid-stride over the (data-dependent) up-sampled extent: sized for a typical margin, loops cover the rest
    const dim3 g1((2 * 2 * d.width + 31) / 32, (2 * 2 * d.height + 7) / 8, d.batch * d.channels);
    augment_up_kernel<<<g1, blk, 0, st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
    const dim3 g2((d.width + 31) / 32, (d.height + 7) / 8, d.batch);
    augment_resample_kernel<<<g2, blk, 0, st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
  } else {
    OI_CHECK_CUDA(cudaMemsetAsync(a.U, 0, augment_u_floats(d) * sizeof(float), st));
    const dim3 g1((Wr + 31) / 32, (Hr + 7) / 8, d.batch * d.channels);
    augment_down_adj_kernel<<<g1, blk, 0, st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
    const dim3 g2((Wr + 31) / 32, (Hr + 7) / 8, d.batch);
    augment_resample_adj_kernel<<<g2, blk, 0, st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
    const dim3 g3((d.width + 31) / 32, (d.height + 7) / 8, d.batch * d.channels);
    augment_up_adj_kernel<<<g3, blk, 0, st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
  }
  (void)Hu_max;
  (void)Wu_max;
  return OI_OK;
}

}  // namespace oi
