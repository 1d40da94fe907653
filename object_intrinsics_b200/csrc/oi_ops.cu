// StyleGAN2 ops for sm_100a: upfirdn2d (ADA + stylesdf flavours), bias_act (ADA), fused_bias_act (stylesdf).
// These are HBM-/launch-bound element-wise and short-FIR kernels: coalesced, one pass, no staging copies.
#include <cuda_fp16.h>

#include "oi_internal.cuh"

namespace oi {

namespace {

template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };

template <class T> __device__ __forceinline__ typename Acc<T>::type ld(const T* p) { return (typename Acc<T>::type)(*p); }
template <> __device__ __forceinline__ float ld<__half>(const __half* p) { return __half2float(*p); }
template <class T> __device__ __forceinline__ void st(T* p, typename Acc<T>::type v) { *p = (T)v; }
template <> __device__ __forceinline__ void st<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// ---------------------------------------------------------------------------------------------------
// upfirdn2d.  Definition (ada/torch_utils/ops/upfirdn2d.py:120-165): zero-insert upsample by `up`, pad
// (negative = crop), correlate with the (flipped unless `flip`) FIR, keep every `down`-th sample, times gain.
// Per output sample only the taps that land on a real (non-inserted) input sample are visited (polyphase).
// One thread per output sample, x fastest -> coalesced stores; the FIR sits in shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 1024;

template <class T>
__global__ void upfirdn2d_kernel(const OiUpfirdnDesc d) {
  typedef typename Acc<T>::type acc_t;
  __shared__ float fs[kMaxTaps];
  const int ntaps = d.filter_w * d.filter_h;
  const bool f_in_smem = ntaps <= kMaxTaps;
  if (f_in_smem) {
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < ntaps; i += blockDim.x * blockDim.y) {
      int ty = i / d.filter_w, tx = i - ty * d.filter_w;
      // tap (ty, tx) multiplies padded sample (oy*down + ty, ox*down + tx): convolution flips the filter
      const int fy = d.flip ? ty : d.filter_h - 1 - ty;
      const int fx = d.flip ? tx : d.filter_w - 1 - tx;
      fs[i] = d.f[fy * d.f_stride_h + fx * d.f_stride_w];
    }
    __syncthreads();
  }
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= d.out_w || oy >= d.out_h) return;

  // first tap index t >= 0 with (o*down + t - pad0) a non-negative multiple of up
  const int bx = ox * d.down_x - d.pad_x0, by = oy * d.down_y - d.pad_y0;
  int tx0 = ((-bx) % d.up_x + d.up_x) % d.up_x;
  if (bx + tx0 < 0) tx0 += ((-(bx + tx0) + d.up_x - 1) / d.up_x) * d.up_x;
  int ty0 = ((-by) % d.up_y + d.up_y) % d.up_y;
  if (by + ty0 < 0) ty0 += ((-(by + ty0) + d.up_y - 1) / d.up_y) * d.up_y;

  const T* x = static_cast<const T*>(d.x);
  T* y = static_cast<T*>(d.y);
  for (int nc = blockIdx.z; nc < d.batch * d.channels; nc += gridDim.z) {
    const int n = nc / d.channels, c = nc - n * d.channels;
    const T* xp = x + n * d.x_stride_n + c * d.x_stride_c;
    acc_t v = 0;
    for (int ty = ty0; ty < d.filter_h; ty += d.up_y) {
      const int iy = (by + ty) / d.up_y;
      if (iy >= d.in_h) break;
      for (int tx = tx0; tx < d.filter_w; tx += d.up_x) {
        const int ix = (bx + tx) / d.up_x;
        if (ix >= d.in_w) break;
        float fv;
        if (f_in_smem) {
          fv = fs[ty * d.filter_w + tx];
        } else {
          const int fy = d.flip ? ty : d.filter_h - 1 - ty;
          const int fx = d.flip ? tx : d.filter_w - 1 - tx;
          fv = d.f[fy * d.f_stride_h + fx * d.f_stride_w];
        }
        v += ld<T>(xp + iy * d.x_stride_h + ix * d.x_stride_w) * (acc_t)fv;
      }
    }
    v *= (acc_t)d.gain;
    st<T>(y + n * d.y_stride_n + c * d.y_stride_c + oy * d.y_stride_h + ox * d.y_stride_w, v);
  }
}


// ---------------------------------------------------------------------------------------------------
// Fast path of upfirdn2d for what the reference actually issues (the AugmentPipe applies the
// 12-tap sym6 filter as four separable passes, up or down by 2 along ONE axis per pass): fp32, dense NCHW,
// a 1-D filter (fw x 1 or 1 x fh, <= 32 taps), up/down in {1, 2} along the filter axis and 1 across it.
// One thread per output sample, x fastest; rows = blockIdx.y (no per-sample division: UP / DOWN are compile-time,
// 32-bit indexing, the polyphase tap loop is unrolled); a warp's taps overlap in L1.  Everything else takes
// the generic kernel above.
// ---------------------------------------------------------------------------------------------------
constexpr int kFastTaps = 32;

template <int UP, int DOWN, int AXIS>   // AXIS 0: filter along x, 1: along y
__global__ void __launch_bounds__(256) upfirdn1d_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ f, int f_stride, int taps, int flip,
                                                        int in_h, int in_w, int out_h, int out_w, int pad_a0,
                                                        int pad_c0, int n_rows, float gain) {
  __shared__ float fs[kFastTaps];
  if (threadIdx.x < kFastTaps) {
    const int t = threadIdx.x;
    fs[t] = t < taps ? f[(flip ? t : taps - 1 - t) * f_stride] : 0.f;
  }
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= out_w) return;
  for (int row = blockIdx.y; row < n_rows; row += gridDim.y) {   // row = nc * out_h + oy
    const int nc = row / out_h, oy = row - nc * out_h;
    const float* xp = x + (size_t)nc * in_h * in_w;
    float v = 0.f;
    if (AXIS == 0) {
      const int iy = oy - pad_c0;                      // across the filter: pure crop / zero pad
      if (iy >= 0 && iy < in_h) {
        const float* xr = xp + iy * in_w;
        const int b = ox * DOWN - pad_a0;              // padded-upsampled coordinate of tap 0
        int t0 = (UP == 1) ? 0 : ((-b) & (UP - 1));   // first tap on a real (non-inserted) sample
        if (b + t0 < 0) t0 += ((-(b + t0) + UP - 1) / UP) * UP;
#pragma unroll 4
        for (int t = t0; t < taps; t += UP) {
          const int ix = (b + t) / UP;
          if (ix >= in_w) break;
          v = fmaf(__ldg(xr + ix), fs[t], v);
        }
      }
    } else {
      const int ix = ox - pad_c0;
      if (ix >= 0 && ix < in_w) {
        const int b = oy * DOWN - pad_a0;
        int t0 = (UP == 1) ? 0 : ((-b) & (UP - 1));
        if (b + t0 < 0) t0 += ((-(b + t0) + UP - 1) / UP) * UP;
#pragma unroll 4
        for (int t = t0; t < taps; t += UP) {
          const int iy = (b + t) / UP;
          if (iy >= in_h) break;
          v = fmaf(__ldg(xp + iy * in_w + ix), fs[t], v);
        }
      }
    }
    y[(size_t)row * out_w + ox] = v * gain;
  }
}

bool upfirdn_fast_ok(const OiUpfirdnDesc& d, int* axis) {
  if (d.dtype != 0) return false;
  const bool fx = d.filter_h == 1 && d.up_y == 1 && d.down_y == 1;
  const bool fy = d.filter_w == 1 && d.up_x == 1 && d.down_x == 1;
  if (!fx && !fy) return false;
  *axis = fx ? 0 : 1;
  const int taps = fx ? d.filter_w : d.filter_h, up = fx ? d.up_x : d.up_y, down = fx ? d.down_x : d.down_y;
  if (taps > kFastTaps || up > 2 || down > 2) return false;
  const long long hw_in = (long long)d.in_h * d.in_w, hw_out = (long long)d.out_h * d.out_w;
  if (d.x_stride_w != 1 || d.x_stride_h != d.in_w || d.x_stride_c != hw_in || d.x_stride_n != hw_in * d.channels)
    return false;
  if (d.y_stride_w != 1 || d.y_stride_h != d.out_w || d.y_stride_c != hw_out || d.y_stride_n != hw_out * d.channels)
    return false;
  return (long long)d.batch * d.channels * hw_in < (1ll << 31) && (long long)d.batch * d.channels * hw_out < (1ll << 31);
}

int launch_upfirdn_fast(const OiUpfirdnDesc& d, int axis, cudaStream_t s) {
  const int taps = axis == 0 ? d.filter_w : d.filter_h;
  const int up = axis == 0 ? d.up_x : d.up_y, down = axis == 0 ? d.down_x : d.down_y;
  const int pad_a0 = axis == 0 ? d.pad_x0 : d.pad_y0, pad_c0 = axis == 0 ? d.pad_y0 : d.pad_x0;
  const int f_stride = (int)(axis == 0 ? d.f_stride_w : d.f_stride_h);
  const int n_rows = d.batch * d.channels * d.out_h;
  const int threads = d.out_w >= 192 ? 256 : (d.out_w >= 96 ? 128 : 64);
  dim3 grid((d.out_w + threads - 1) / threads, n_rows < 65535 ? n_rows : 65535);
  const float* x = static_cast<const float*>(d.x);
  float* y = static_cast<float*>(d.y);
#define OI_UF(U_, D_, A_)                                                                                          \
  upfirdn1d_kernel<U_, D_, A_><<<grid, threads, 0, s>>>(x, y, d.f, f_stride, taps, d.flip, d.in_h, d.in_w, d.out_h, \
                                                        d.out_w, pad_a0, pad_c0, n_rows, d.gain)
  const int key = (up - 1) * 4 + (down - 1) * 2 + axis;
  switch (key) {
    case 0: OI_UF(1, 1, 0); break;
    case 1: OI_UF(1, 1, 1); break;
    case 2: OI_UF(1, 2, 0); break;
    case 3: OI_UF(1, 2, 1); break;
    case 4: OI_UF(2, 1, 0); break;
    case 5: OI_UF(2, 1, 1); break;
    case 6: OI_UF(2, 2, 0); break;
    default: OI_UF(2, 2, 1); break;
  }
#undef OI_UF
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

// ---------------------------------------------------------------------------------------------------
// bias_act (ada/torch_utils/ops/bias_act.py:23-33 table; bias_act.cu:23-147 semantics):
//   grad 0: y = clamp(act(x + b) * gain)
//   grad 1: y = x * act'(.) * gain, expressed through yref (or xref for swish), zeroed where the forward clamped
//   grad 2: y = x * act''(.) * gain * dy
// ---------------------------------------------------------------------------------------------------
template <class A, int ACT>
__device__ __forceinline__ A act_eval(int G, A x, A xref, A yy, A alpha, A* yref_io, A gain) {
  const A one = 1, two = 2, zero = 0;
  const A kExpRange = 80, kHalfExpRange = 40;
  const A kSeluScale = (A)1.0507009873554804934193349852946;
  const A kSeluAlpha = (A)1.6732632423543772848170429916717;
  A y = zero;
  if (ACT == 1) {  // linear
    if (G <= 1) y = x;
  } else if (ACT == 2) {  // relu
    if (G == 0) y = x > zero ? x : zero;
    if (G == 1) y = yy > zero ? x : zero;
  } else if (ACT == 3) {  // lrelu
    if (G == 0) y = x > zero ? x : x * alpha;
    if (G == 1) y = yy > zero ? x : x * alpha;
  } else if (ACT == 4) {  // tanh
    if (G == 0) {
      if (x < -kExpRange) y = -one;
      else if (x > kExpRange) y = one;
      else { const A e = exp(x), r = one / e; y = (e - r) / (e + r); }
    }
    if (G == 1) y = x * (one - yy * yy);
    if (G == 2) y = x * (one - yy * yy) * (-two * yy);
  } else if (ACT == 5) {  // sigmoid
    if (G == 0) y = x < -kExpRange ? zero : one / (exp(-x) + one);
    if (G == 1) y = x * yy * (one - yy);
    if (G == 2) y = x * yy * (one - yy) * (one - two * yy);
  } else if (ACT == 6) {  // elu
    if (G == 0) y = x >= zero ? x : exp(x) - one;
    if (G == 1) y = yy >= zero ? x : x * (yy + one);
    if (G == 2) y = yy >= zero ? zero : x * (yy + one);
  } else if (ACT == 7) {  // selu
    if (G == 0) y = x >= zero ? kSeluScale * x : (kSeluScale * kSeluAlpha) * (exp(x) - one);
    if (G == 1) y = yy >= zero ? x * kSeluScale : x * (yy + kSeluScale * kSeluAlpha);
    if (G == 2) y = yy >= zero ? zero : x * (yy + kSeluScale * kSeluAlpha);
  } else if (ACT == 8) {  // softplus
    if (G == 0) y = x > kExpRange ? x : log(exp(x) + one);
    if (G == 1) y = x * (one - exp(-yy));
    if (G == 2) { const A e = exp(-yy); y = x * e * (one - e); }
  } else if (ACT == 9) {  // swish
    if (G == 0) {
      y = x < -kExpRange ? zero : x / (exp(-x) + one);
    } else {
      const A e = exp(xref), dd = e + one;
      if (G == 1) y = xref > kHalfExpRange ? x : x * e * (xref + dd) / (dd * dd);
      else y = xref > kHalfExpRange ? zero : x * e * (xref * (two - dd) + two * dd) / (dd * dd * dd);
      *yref_io = xref < -kExpRange ? zero : xref / (exp(-xref) + one) * gain;
    }
  }
  return y;
}

template <class T, int ACT>
__global__ void bias_act_kernel(const OiBiasActDesc d) {
  typedef typename Acc<T>::type A;
  const T* x = static_cast<const T*>(d.x);
  const T* b = static_cast<const T*>(d.b);
  const T* xr = static_cast<const T*>(d.xref);
  const T* yr = static_cast<const T*>(d.yref);
  const T* dy = static_cast<const T*>(d.dy);
  T* y = static_cast<T*>(d.y);
  const int G = d.grad;
  const A alpha = (A)d.alpha, gain = (A)d.gain, clampv = (A)d.clamp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.size_x; i += gridDim.x * blockDim.x) {
    A xv = ld<T>(x + i);
    const A bv = b ? ld<T>(b + (i / d.step_b) % d.size_b) : (A)0;
    A xref = xr ? ld<T>(xr + i) : (A)0;
    A yref = yr ? ld<T>(yr + i) : (A)0;
    const A dyv = dy ? ld<T>(dy + i) : (A)1;
    const A yy = (gain != (A)0) ? yref / gain : (A)0;
    if (G == 0) xv += bv; else xref += bv;
    A v = act_eval<A, ACT>(G, xv, xref, yy, alpha, &yref, gain);
    v *= gain * dyv;
    if (clampv >= (A)0) {
      if (G == 0) v = (v > -clampv && v < clampv) ? v : (v >= (A)0 ? clampv : -clampv);
      else v = (yref > -clampv && yref < clampv) ? v : (A)0;
    }
    st<T>(y + i, v);
  }
}


// Fast path of bias_act: fp32, 16-byte aligned dense tensors whose bias run length (step_b) is a multiple of 4.
// One row of `step_b` elements per blockIdx.y (bias index = row % size_b: no per-element division), float4
// accesses, four independent 16-byte loads in flight per thread.  HBM-bound: 8 bytes of traffic per element.
template <int ACT>
__global__ void __launch_bounds__(256) bias_act_rows_kernel(const OiBiasActDesc d, int n_rows, int row4) {
  const float4* x = static_cast<const float4*>(d.x);
  const float* b = static_cast<const float*>(d.b);
  const float4* xr = static_cast<const float4*>(d.xref);
  const float4* yr = static_cast<const float4*>(d.yref);
  const float4* dy = static_cast<const float4*>(d.dy);
  float4* y = static_cast<float4*>(d.y);
  const int G = d.grad;
  const float alpha = d.alpha, gain = d.gain, clampv = d.clamp;
  const float inv_gain = gain != 0.f ? 1.0f / gain : 0.f;
  for (int row = blockIdx.y; row < n_rows; row += gridDim.y) {
    const float bv = b ? __ldg(b + row % d.size_b) : 0.f;
    const size_t base = (size_t)row * row4;
    for (int i0 = blockIdx.x * blockDim.x * 4 + threadIdx.x; i0 < row4; i0 += gridDim.x * blockDim.x * 4) {
      float4 xv[4], xf[4], yf[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < row4) {
          xv[u] = __ldcs(x + base + i);
          if (xr) xf[u] = __ldcs(xr + base + i);
          if (yr) yf[u] = __ldcs(yr + base + i);
          if (dy) dv[u] = __ldcs(dy + base + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i >= row4) continue;
        float xe[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, out[4];
        const float xre[4] = {xf[u].x, xf[u].y, xf[u].z, xf[u].w}, yre[4] = {yf[u].x, yf[u].y, yf[u].z, yf[u].w};
        const float dye[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float xin = xe[e], xref = xr ? xre[e] : 0.f, yref = yr ? yre[e] : 0.f;
          const float dyv = dy ? dye[e] : 1.f;
          const float yy = yref * inv_gain;
          if (G == 0) xin += bv; else xref += bv;
          float v = act_eval<float, ACT>(G, xin, xref, yy, alpha, &yref, gain);
          v *= gain * dyv;
          if (clampv >= 0.f) {
            if (G == 0) v = (v > -clampv && v < clampv) ? v : (v >= 0.f ? clampv : -clampv);
            else v = (yref > -clampv && yref < clampv) ? v : 0.f;
          }
          out[e] = v;
        }
        __stcs(y + base + i, make_float4(out[0], out[1], out[2], out[3]));
      }
    }
  }
}

bool bias_act_fast_ok(const OiBiasActDesc& d, int* n_rows, int* row4) {
  if (d.dtype != 0 || d.size_x < 4096) return false;
  int step = d.b ? d.step_b : 4096;
  if (d.b == nullptr) {   // no bias: any row length works; pick one that divides size_x
    while (step > 4 && d.size_x % step != 0) step >>= 1;
  }
  if (step < 256 || step % 4 != 0 || d.size_x % step != 0) return false;
  const uintptr_t al = (uintptr_t)d.x | (uintptr_t)d.y | (uintptr_t)d.xref | (uintptr_t)d.yref | (uintptr_t)d.dy;
  if (al & 15) return false;
  *n_rows = d.size_x / step;
  *row4 = step / 4;
  return true;
}

template <class T>
int launch_bias_act_t(const OiBiasActDesc& d, cudaStream_t s) {
  const int threads = 256;
  int blocks = (d.size_x + threads * 4 - 1) / (threads * 4);
  if (blocks < 1) blocks = 1;
  switch (d.act) {
#define OI_CASE(A_) case A_: bias_act_kernel<T, A_><<<blocks, threads, 0, s>>>(d); break;
    OI_CASE(1) OI_CASE(2) OI_CASE(3) OI_CASE(4) OI_CASE(5) OI_CASE(6) OI_CASE(7) OI_CASE(8) OI_CASE(9)
#undef OI_CASE
    default: return set_error(OI_ERR_INVALID_ARGUMENT, "bias_act: no kernel for act=%d", d.act);
  }
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

// ---------------------------------------------------------------------------------------------------
// stylesdf fused_bias_act (stylesdf/op/fused_bias_act_kernel.cu:19-52): y = f(x + b) * scale with
// act*10+grad in {10,11: identity; 30: lrelu; 31: lrelu slope selected by sign of ref; 12,32: zero}.
// ---------------------------------------------------------------------------------------------------
template <class T>
__global__ void fused_bias_act_kernel(const OiFusedBiasActDesc d) {
  typedef typename Acc<T>::type A;
  const T* x = static_cast<const T*>(d.x);
  const T* b = static_cast<const T*>(d.bias);
  const T* r = static_cast<const T*>(d.ref);
  T* y = static_cast<T*>(d.y);
  const int mode = d.act * 10 + d.grad;
  const A alpha = (A)d.alpha, scale = (A)d.scale;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.size_x; i += gridDim.x * blockDim.x) {
    A v = ld<T>(x + i);
    if (b) v += ld<T>(b + (i / d.step_b) % d.size_b);
    const A ref = r ? ld<T>(r + i) : (A)0;
    A out;
    switch (mode) {
      case 12: case 32: out = (A)0; break;
      case 30: out = v > (A)0 ? v : v * alpha; break;
      case 31: out = ref > (A)0 ? v : v * alpha; break;
      default: out = v; break;
    }
    st<T>(y + i, out * scale);
  }
}

}  // namespace

int launch_upfirdn2d(const OiUpfirdnDesc& d, cudaStream_t s) {
  int axis = 0;
  if (upfirdn_fast_ok(d, &axis)) return launch_upfirdn_fast(d, axis, s);
  dim3 block(32, 8, 1);
  int nc = d.batch * d.channels;
  dim3 grid((d.out_w + 31) / 32, (d.out_h + 7) / 8, nc < 65535 ? nc : 65535);
  switch (d.dtype) {
    case 0: upfirdn2d_kernel<float><<<grid, block, 0, s>>>(d); break;
    case 1: upfirdn2d_kernel<__half><<<grid, block, 0, s>>>(d); break;
    case 2: upfirdn2d_kernel<double><<<grid, block, 0, s>>>(d); break;
    default: return set_error(OI_ERR_INVALID_ARGUMENT, "upfirdn2d: bad dtype %d", d.dtype);
  }
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_bias_act(const OiBiasActDesc& d, cudaStream_t s) {
  int n_rows = 0, row4 = 0;
  if (bias_act_fast_ok(d, &n_rows, &row4)) {
    const int threads = 256;
    dim3 grid((row4 + threads * 4 - 1) / (threads * 4), n_rows < 65535 ? n_rows : 65535);
    switch (d.act) {
#define OI_CASE(A_) case A_: bias_act_rows_kernel<A_><<<grid, threads, 0, s>>>(d, n_rows, row4); break;
      OI_CASE(1) OI_CASE(2) OI_CASE(3) OI_CASE(4) OI_CASE(5) OI_CASE(6) OI_CASE(7) OI_CASE(8) OI_CASE(9)
#undef OI_CASE
      default: return set_error(OI_ERR_INVALID_ARGUMENT, "bias_act: no kernel for act=%d", d.act);
    }
    OI_CHECK_CUDA(cudaGetLastError());
    return OI_OK;
  }
  switch (d.dtype) {
    case 0: return launch_bias_act_t<float>(d, s);
    case 1: return launch_bias_act_t<__half>(d, s);
    case 2: return launch_bias_act_t<double>(d, s);
    default: return set_error(OI_ERR_INVALID_ARGUMENT, "bias_act: bad dtype %d", d.dtype);
  }
}

int launch_fused_bias_act(const OiFusedBiasActDesc& d, cudaStream_t s) {
  const int threads = 256;
  int blocks = (d.size_x + threads * 4 - 1) / (threads * 4);
  if (blocks < 1) blocks = 1;
  switch (d.dtype) {
    case 0: fused_bias_act_kernel<float><<<blocks, threads, 0, s>>>(d); break;
    case 1: fused_bias_act_kernel<__half><<<blocks, threads, 0, s>>>(d); break;
    case 2: fused_bias_act_kernel<double><<<blocks, threads, 0, s>>>(d); break;
    default: return set_error(OI_ERR_INVALID_ARGUMENT, "fused_bias_act: bad dtype %d", d.dtype);
  }
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
