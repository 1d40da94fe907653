// FP32-FFMA core of the fused SDF render kernel (OI_IMPL_FFMA).
//
// One CTA owns a tile of 128 sample points of one object instance and keeps the whole [128 ch x 128 pt]
// activation tile in shared memory through
//   forward  : D FiLM-SIREN layers  h <- sin(gamma*(W h + b) + beta)        (volume_renderer.py:50-61)
//   sdf      : w_sigma . h + b_sigma                                        (fields.py:64)
//   reverse  : grad_x sdf by an explicit reverse sweep  g <- W^T (g * gamma*cos(arg))
//              (what autograd.grad computes at fields.py:104-122, without the reference's second forward)
//   colour   : sigmoid(W_rgb sin(gamma_c*(W_c [h; grad] + b_c) + beta_c) + b_rgb)   (fields.py:89-101)
//   alpha    : NeuS section opacity                                         (renderer.py:266-286)
// The 128x128 weight panels are streamed from the packed blob (L2-resident, ~1 MB) through a 4-stage
// shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier); each thread accumulates an 8x8
// register tile.  gamma*cos(arg) of every layer (needed by the reverse sweep) goes to a per-CTA scratch
// slab that persistent CTAs reuse, so it lives in L2.
//
// The same kernel runs the coarse pass of hierarchical sampling (SDF only, points at z instead of
// section midpoints, renderer.py:389-399) with args.coarse = 1.
#include "oi_internal.cuh"
#include "oi_render_common.cuh"

namespace oi {

namespace {

constexpr int kTP = 128;  // points per tile
constexpr int kStages = 4;
constexpr int kThreads = 256;

struct __align__(128) FfmaSmem {
  float act[kW * kTP];                   // [k][m], 16-byte chunks XOR-swizzled by (k>>2)&7   (64 KB)
  float wring[kStages][kChunkFloats];    // streamed weight chunks                            (32 KB)
  float red[2][4][kTP];                  // two-half partial sums of the narrow (<=3 output) contractions
  float sdfv[kTP];
  float grad[3][kTP];
  unsigned long long full[kStages];
};

struct Pipe {
  int cc;         // chunks consumed so far by this CTA
  int pc;         // chunks issued so far
  int total;      // chunks this CTA will consume over its lifetime
  int per_tile;   // chunks per tile (stream length)
  const float* stream;
};

__device__ __forceinline__ void pipe_issue(FfmaSmem& sm, Pipe& p) {
  int stage = p.pc % kStages;
  const float* src = p.stream + (size_t)(p.pc % p.per_tile) * kChunkFloats;
  mbar_expect_tx(&sm.full[stage], kChunkBytes);
  tma_bulk_g2s(sm.wring[stage], src, kChunkBytes, &sm.full[stage]);
  p.pc++;
}

// acc[im][jn] += sum_k act[k][m(im)] * wchunk[k][n(jn)] over `nchunks` chunks of `krows` rows each.
__device__ __forceinline__ void gemm_chunks(FfmaSmem& sm, Pipe& p, float (&acc)[8][8], int nchunks, int krows,
                                            int tx, int ty, int tid) {
  const float4* act4 = reinterpret_cast<const float4*>(sm.act);
  for (int c = 0; c < nchunks; ++c) {
    const int stage = p.cc % kStages;
    mbar_wait(&sm.full[stage], (p.cc / kStages) & 1);
    const float4* w4 = reinterpret_cast<const float4*>(sm.wring[stage]);
    const int kbase = c * kKC;
    for (int kk = 0; kk < krows; kk += 4) {
      const int sw = ((kbase + kk) >> 2) & 7;
      const int ca = ty ^ sw;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a0 = act4[(kbase + kk + q) * 32 + ca];
        const float4 a1 = act4[(kbase + kk + q) * 32 + 16 + ca];
        const float4 b0 = w4[(kk + q) * 32 + tx];
        const float4 b1 = w4[(kk + q) * 32 + 16 + tx];
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();  // every warp is done with this stage (and, on the last chunk, with `act`)
    if (tid == 0 && p.pc < p.total) pipe_issue(sm, p);
    p.cc++;
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ int col_of(int tx, int jn) { return (jn < 4) ? tx * 4 + jn : 64 + tx * 4 + (jn - 4); }

// writes one activation row segment (this thread's 8 points of channel n) into the swizzled tile
__device__ __forceinline__ void store_act(FfmaSmem& sm, int n, int ty, const float (&v)[8]) {
  float4* act4 = reinterpret_cast<float4*>(sm.act);
  const int ca = ty ^ ((n >> 2) & 7);
  act4[n * 32 + ca] = make_float4(v[0], v[1], v[2], v[3]);
  act4[n * 32 + 16 + ca] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float load_act(const FfmaSmem& sm, int n, int m) {
  return sm.act[n * kTP + ((((m >> 2) ^ ((n >> 2) & 7)) << 2) | (m & 3))];
}

// out[j] (j < J) = sum_n act[n][m] * V[j][n]; result valid for tid < 128 (point m = tid) after return.
template <int J>
__device__ __forceinline__ void narrow_contract(FfmaSmem& sm, const float* __restrict__ V, float (&out)[J], int tid) {
  const int m = tid & (kTP - 1), half = tid >> 7;
  float s[J];
#pragma unroll
  for (int j = 0; j < J; ++j) s[j] = 0.f;
  const int n0 = half * 64;
#pragma unroll 8
  for (int n = n0; n < n0 + 64; ++n) {
    const float a = load_act(sm, n, m);
#pragma unroll
    for (int j = 0; j < J; ++j) s[j] = fmaf(a, __ldg(V + j * kW + n), s[j]);
  }
#pragma unroll
  for (int j = 0; j < J; ++j) sm.red[half][j][m] = s[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J; ++j) out[j] = sm.red[0][j][m] + sm.red[1][j][m];
}

__global__ void __launch_bounds__(kThreads, 2) render_ffma_kernel(const RenderKArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FfmaSmem& sm = *reinterpret_cast<FfmaSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = (warp & 1) * 8 + (lane & 7);
  const int ty = (warp >> 1) * 4 + (lane >> 3);
  const BlobLayout L = blob_layout(a.D);
  const float* cst = a.blob + L.const_off;
  const int D = a.D;

  const int my_tiles = (a.n_tiles > (int)blockIdx.x) ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  Pipe pipe;
  pipe.cc = 0;
  pipe.pc = 0;
  pipe.per_tile = a.coarse ? L.n_chunks_coarse : L.n_chunks_fine;
  pipe.total = pipe.per_tile * my_tiles;
  pipe.stream = a.blob + L.stream_off;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kStages && pipe.pc < pipe.total; ++s) pipe_issue(sm, pipe);

  float* scr = a.scratch + (size_t)blockIdx.x * a.scratch_stride;  // [(D+1)][128 n][128 m]

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int inst = tile / a.tiles_per_inst;
    const int tin = tile - inst * a.tiles_per_inst;
    const float* film = a.film + (size_t)inst * kFilm * 2 * kW;

    // ---------------- prologue: sample positions of the 128 points of this tile ----------------
    PointCtx pc;
    pc.valid = false;
    if (tid < kTP) {
      pc = point_prologue(a, inst, tin, tid);
      sm.act[0 * kTP + tid] = pc.px;  // rows 0..3: (k>>2)&7 == 0 -> no swizzle
      sm.act[1 * kTP + tid] = pc.py;
      sm.act[2 * kTP + tid] = pc.pz;
      sm.act[3 * kTP + tid] = 0.f;
    }
    __syncthreads();

    float acc[8][8];
    // ---------------- forward sweep ----------------
    for (int l = 0; l < D; ++l) {
      zero_acc(acc);
      gemm_chunks(sm, pipe, acc, l == 0 ? 1 : 8, l == 0 ? 4 : kKC, tx, ty, tid);
      const float* gam = film + (l * 2 + 0) * kW;
      const float* bet = film + (l * 2 + 1) * kW;
      const float* bia = cst + BlobLayout::kBias + l * kW;
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const int n = col_of(tx, jn);
        const float g = __ldg(gam + n), be = __ldg(bet + n), bi = __ldg(bia + n);
        float hv[8], cv[8];
#pragma unroll
        for (int im = 0; im < 8; ++im) {
          const float arg = fmaf(g, acc[im][jn] + bi, be);
          float s, c;
          sincos_film(arg, &s, &c);
          hv[im] = s;
          cv[im] = g * c;
        }
        store_act(sm, n, ty, hv);
        if (!a.coarse) {
          float4* dst = reinterpret_cast<float4*>(scr + ((size_t)l * kW + n) * kTP);
          dst[ty] = make_float4(cv[0], cv[1], cv[2], cv[3]);
          dst[16 + ty] = make_float4(cv[4], cv[5], cv[6], cv[7]);
        }
      }
      __syncthreads();
    }

    // ---------------- sdf = w_sigma . h + b_sigma ----------------
    {
      float o[1];
      narrow_contract<1>(sm, cst + BlobLayout::kWsig, o, tid);
      if (tid < kTP) sm.sdfv[tid] = o[0] + cst[BlobLayout::kScalars + 0];
    }
    if (a.coarse) {
      if (tid < kTP && pc.valid) a.sdf_coarse[(size_t)pc.ray * a.S + pc.si] = sm.sdfv[tid];
      __syncthreads();
      continue;
    }

    // ---------------- colour layer, feature part: u_c = W_c[:, :128] h  (parked in scratch slot D) ----------------
    zero_acc(acc);
    gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
      const int n = col_of(tx, jn);
      float4* dst = reinterpret_cast<float4*>(scr + ((size_t)D * kW + n) * kTP);
      dst[ty] = make_float4(acc[0][jn], acc[1][jn], acc[2][jn], acc[3][jn]);
      dst[16 + ty] = make_float4(acc[4][jn], acc[5][jn], acc[6][jn], acc[7][jn]);
    }

    // ---------------- reverse sweep: t_{D-1} = w_sigma * c_{D-1};  g_l = W_l^T t_l;  t_{l-1} = g_l * c_{l-1} ----------------
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
      const int n = col_of(tx, jn);
      const float ws = __ldg(cst + BlobLayout::kWsig + n);
      const float4* src = reinterpret_cast<const float4*>(scr + ((size_t)(D - 1) * kW + n) * kTP);
      const float4 c0 = src[ty], c1 = src[16 + ty];
      const float tv[8] = {ws * c0.x, ws * c0.y, ws * c0.z, ws * c0.w, ws * c1.x, ws * c1.y, ws * c1.z, ws * c1.w};
      store_act(sm, n, ty, tv);
    }
    __syncthreads();
    for (int l = D - 1; l >= 1; --l) {
      zero_acc(acc);
      gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const int n = col_of(tx, jn);
        const float4* src = reinterpret_cast<const float4*>(scr + ((size_t)(l - 1) * kW + n) * kTP);
        const float4 c0 = src[ty], c1 = src[16 + ty];
        const float tv[8] = {acc[0][jn] * c0.x, acc[1][jn] * c0.y, acc[2][jn] * c0.z, acc[3][jn] * c0.w,
                             acc[4][jn] * c1.x, acc[5][jn] * c1.y, acc[6][jn] * c1.z, acc[7][jn] * c1.w};
        store_act(sm, n, ty, tv);
      }
      __syncthreads();
    }
    // grad_x sdf = W_0^T t_0
    {
      float o[3];
      narrow_contract<3>(sm, cst + BlobLayout::kW0t, o, tid);
      if (tid < kTP) {
        sm.grad[0][tid] = o[0];
        sm.grad[1][tid] = o[1];
        sm.grad[2][tid] = o[2];
      }
    }
    __syncthreads();

    // ---------------- colour layer epilogue: h_c = sin(gamma_c (u_c + W_c[:,128:] grad + b_c) + beta_c) ----------------
    {
      const float* gam = film + (OI_MAX_DEPTH * 2 + 0) * kW;
      const float* bet = film + (OI_MAX_DEPTH * 2 + 1) * kW;
      const float* bia = cst + BlobLayout::kBias + OI_MAX_DEPTH * kW;
      float gx[8], gy[8], gz[8];
#pragma unroll
      for (int im = 0; im < 8; ++im) {
        const int m = (im < 4) ? ty * 4 + im : 64 + ty * 4 + (im - 4);
        gx[im] = sm.grad[0][m];
        gy[im] = sm.grad[1][m];
        gz[im] = sm.grad[2][m];
      }
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const int n = col_of(tx, jn);
        const float g = __ldg(gam + n), be = __ldg(bet + n), bi = __ldg(bia + n);
        const float w0 = __ldg(cst + BlobLayout::kWcg + 0 * kW + n);
        const float w1 = __ldg(cst + BlobLayout::kWcg + 1 * kW + n);
        const float w2 = __ldg(cst + BlobLayout::kWcg + 2 * kW + n);
        const float4* src = reinterpret_cast<const float4*>(scr + ((size_t)D * kW + n) * kTP);
        const float4 u0 = src[ty], u1 = src[16 + ty];
        const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        float hv[8];
#pragma unroll
        for (int im = 0; im < 8; ++im) {
          float pre = fmaf(w0, gx[im], u[im]);
          pre = fmaf(w1, gy[im], pre);
          pre = fmaf(w2, gz[im], pre);
          const float arg = fmaf(g, pre + bi, be);
          float s, c;
          sincos_film(arg, &s, &c);
          hv[im] = s;
        }
        store_act(sm, n, ty, hv);
      }
    }
    __syncthreads();
    float rgb[3];
    narrow_contract<3>(sm, cst + BlobLayout::kWrgb, rgb, tid);

    // ---------------- per-point tail: colour, NeuS alpha (renderer.py:266-286) ----------------
    if (tid < kTP) point_tail(a, pc, cst, sm.sdfv[tid], sm.grad[0][tid], sm.grad[1][tid], sm.grad[2][tid], rgb);
    __syncthreads();
  }
}

}  // namespace

size_t render_ffma_scratch_floats(int depth, int* n_ctas, int n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ctas = 2 * sms;
  if (ctas > n_tiles) ctas = n_tiles;
  if (ctas < 1) ctas = 1;
  if (n_ctas) *n_ctas = ctas;
  return (size_t)(depth + 1) * kW * kTP;
}

int launch_render_ffma(const RenderKArgs& a, cudaStream_t st) {
  int n_ctas = 0;
  render_ffma_scratch_floats(a.D, &n_ctas, a.n_tiles);
  static_assert(sizeof(FfmaSmem) <= 113 * 1024, "two CTAs per SM");
  OI_CHECK_CUDA(cudaFuncSetAttribute(render_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(FfmaSmem)));
  render_ffma_kernel<<<n_ctas, kThreads, sizeof(FfmaSmem), st>>>(a);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
