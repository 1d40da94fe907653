// Backward of the fused SDF render path (oi_render_backward): dL/dtheta for every parameter NeuSRenderer.render
// reads, including the second-order terms through the analytic normal (the reference builds them with
// autograd.grad(create_graph=True), src/models/fields.py:104-122, and differentiates again in
// src/trainers/gan_pose_trainer.py:141).  The algorithm is restated on the CPU in oracle/backward_oracle.py and
// checked there against torch.autograd; this file is that sweep as kernels.
//
//   relax_count_kernel : number of points with |p| < 1.2 (denominator of gradient_error, renderer.py:295-297)
//   tail_bwd_kernel    : one thread per ray; adjoint of the compositing (weights = alpha * excl-cumprod,
//                        weight_sum/max, colour) and of the NeuS alpha (renderer.py:266-305) -> per-point
//                        adjoints {sdf_bar, normal_bar[3], rgb-pre-activation_bar[3]}
//   mlp_bwd_kernel     : per 128-point tile (same tiling / register blocking / TMA weight ring as the FP32 FFMA
//                        forward core): recomputes the forward sweep (keeps u_l) and the reverse sweep
//                        (keeps g_l), then runs the backward of the reverse sweep (ascending l, operands W_l),
//                        the backward of the forward sweep (descending l, operands W_l^T), and finally the
//                        weight-gradient contractions over the 128 points of the tile
//                            dW_l += u_bar_l (x) h_l + t_l (x) g_bar_l
//                        with both operands streamed back from per-CTA scratch by TMA in [point][channel]
//                        chunks, accumulated in registers and flushed with vector reductions (red.global.v4).
//   finalize_bwd_kernel: bias gradients from the FiLM-table gradients (db = sum_inst gamma * dbeta), variance.
//
// Per point: u_l = W_l h_l + b_l, a_l = gamma_l u_l + beta_l, h_{l+1} = sin a_l, c_l = gamma_l cos a_l,
// t_{D-1} = w_s c_{D-1}, g_l = W_l^T t_l, t_{l-1} = g_l c_{l-1}, normal = g_0.
#include "oi_internal.cuh"
#include "oi_render_common.cuh"
#include "oi_wgrad.cuh"

namespace oi {

namespace {

constexpr int kTP = 128;
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kSlot = kW * kTP;  // floats per scratch slot (64 KB)

// scratch slots per CTA.  [ch][pt] slots are written and read by the same thread; [pt][ch] slots are TMA sources.
constexpr int kSlotU = 0;     // U[l], l = 0..7          [ch][pt]  pre-activation u_l
constexpr int kSlotG = 8;     // G[l], l = 1..7 (8+l)    [ch][pt]  g_l, later overwritten by c_bar_{l-1}
constexpr int kSlotUC = 16;   //                         [ch][pt]  W_cf h_D
constexpr int kSlotHB = 17;   //                         [ch][pt]  h_bar_D
constexpr int kSlotHT = 18;   // HT[l], l = 1..8 (17+l)  [pt][ch]  h_l
constexpr int kSlotTT = 26;   // TT[l], l = 1..7 (25+l)  [pt][ch]  t_l
constexpr int kSlotGBT = 33;  // GBT[l], l = 1..7 (32+l) [pt][ch]  g_bar_l
constexpr int kSlotUBT = 40;  // UBT[l], l = 1..7 (39+l) [pt][ch]  u_bar_l
constexpr int kSlotUCT = 47;  //                         [pt][ch]  u_bar_c
constexpr int kNumSlots = 48;

struct BwdKArgs {
  RenderKArgs r;     // geometry (z_vals set, all outputs NULL), blob, film, D, tiling
  const float* adj;  // [N][8] per-point adjoints written by tail_bwd_kernel
  float* scratch;
  size_t scratch_stride;
  OiNetGrads g;
  float* d_film;     // [n_inst][9][2][128]  (dgamma, dbeta)
};

struct __align__(128) BwdSmem {
  float act[kW * kTP];                 // [k][m], 16-byte chunks XOR-swizzled by (k>>2)&7   (64 KB)
  float wring[kStages][kChunkFloats];  // streamed chunks (weights or scratch operands)     (32 KB)
  float red[2][4][kTP];
  float pt[10][kTP];                   // x,y,z | sdf_bar | normal_bar[3] | zrgb_bar[3]
  float nrm[3][kTP];
  unsigned long long full[kStages];
};

struct Pipe {
  int cc, pc, limit, total, per_tile, lw, nf, D;
  const float* stream;
  const float* scr;
};

// Source of chunk `pos` (position inside one tile's sequence).
__device__ __forceinline__ const float* chunk_src(const Pipe& p, int pos) {
  const int D = p.D;
  if (pos < p.lw) {
    int id;
    const int c0 = p.nf + 8;             // end of: forward stream + colour-as-stored
    const int c1 = c0 + 8 * (D - 1);     // end of: backward of the reverse sweep (W_l^T chunks, l = 1..D-1)
    if (pos < c0) id = pos;
    else if (pos < c1) id = 1 + (pos - c0);
    else id = 1 + 8 * (D - 1) + 8 + (pos - c1);  // backward of the forward sweep: W_l chunks, l = D-1..1
    return p.stream + (size_t)id * kChunkFloats;
  }
  const int q = pos - p.lw;
  const int step = q >> 1, which = q & 1;
  int slot, c;
  if (step < 16 * (D - 1)) {
    const int l = 1 + step / 16, s = step % 16;
    c = s & 7;
    if (s < 8) slot = which ? (kSlotHT + l - 1) : (kSlotUBT + l - 1);
    else slot = which ? (kSlotGBT + l - 1) : (kSlotTT + l - 1);
  } else {
    c = step - 16 * (D - 1);
    slot = which ? (kSlotHT + D - 1) : kSlotUCT;
  }
  return p.scr + (size_t)slot * kSlot + (size_t)c * kChunkFloats;
}

__device__ __forceinline__ void pipe_fill(BwdSmem& sm, Pipe& p) {
  while (p.pc < p.limit && p.pc - p.cc < kStages) {
    const int stage = p.pc % kStages;
    const float* src = chunk_src(p, p.pc % p.per_tile);
    mbar_expect_tx(&sm.full[stage], kChunkBytes);
    tma_bulk_g2s(sm.wring[stage], src, kChunkBytes, &sm.full[stage]);
    p.pc++;
  }
}

// acc[im][jn] += sum_k act[k][m(im)] * chunk[k][n(jn)]
__device__ __forceinline__ void gemm_chunks(BwdSmem& sm, Pipe& p, float (&acc)[8][8], int nchunks, int krows, int tx,
                                            int ty, int tid) {
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
    __syncthreads();
    p.cc++;
    if (tid == 0) pipe_fill(sm, p);
  }
}

// Weight-gradient contraction: acc[im][jn] += sum_{pt} X[pt][i(im)] * Y[pt][j(jn)], X and Y chunks of 16 points
// arriving as consecutive ring stages.
__device__ __forceinline__ void wgrad_chunks(BwdSmem& sm, Pipe& p, float (&acc)[8][8], int nsteps, int tx, int ty,
                                             int tid) {
  for (int s = 0; s < nsteps; ++s) {
    const int sx = p.cc % kStages, sy = (p.cc + 1) % kStages;
    mbar_wait(&sm.full[sx], (p.cc / kStages) & 1);
    mbar_wait(&sm.full[sy], ((p.cc + 1) / kStages) & 1);
    const float4* x4 = reinterpret_cast<const float4*>(sm.wring[sx]);
    const float4* y4 = reinterpret_cast<const float4*>(sm.wring[sy]);
#pragma unroll 4
    for (int k = 0; k < kKC; ++k) {
      const float4 a0 = x4[k * 32 + ty];
      const float4 a1 = x4[k * 32 + 16 + ty];
      const float4 b0 = y4[k * 32 + tx];
      const float4 b1 = y4[k * 32 + 16 + tx];
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    p.cc += 2;
    if (tid == 0) pipe_fill(sm, p);
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// index maps of the 8x8 register tile: jn -> channel, im -> point
__device__ __forceinline__ int col_of(int tx, int jn) { return (jn < 4) ? tx * 4 + jn : 64 + tx * 4 + (jn - 4); }
__device__ __forceinline__ int row_of(int ty, int im) { return (im < 4) ? ty * 4 + im : 64 + ty * 4 + (im - 4); }

__device__ __forceinline__ void store_act(BwdSmem& sm, int n, int ty, const float (&v)[8]) {
  float4* act4 = reinterpret_cast<float4*>(sm.act);
  const int ca = ty ^ ((n >> 2) & 7);
  act4[n * 32 + ca] = make_float4(v[0], v[1], v[2], v[3]);
  act4[n * 32 + 16 + ca] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float load_act(const BwdSmem& sm, int n, int m) {
  return sm.act[n * kTP + ((((m >> 2) ^ ((n >> 2) & 7)) << 2) | (m & 3))];
}
// [ch][pt] scratch slot, this thread's 8 points of channel n
__device__ __forceinline__ void store_cp(float* slot, int n, int ty, const float (&v)[8]) {
  float4* dst = reinterpret_cast<float4*>(slot + (size_t)n * kTP);
  dst[ty] = make_float4(v[0], v[1], v[2], v[3]);
  dst[16 + ty] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load_cp(const float* slot, int n, int ty, float (&v)[8]) {
  const float4* src = reinterpret_cast<const float4*>(slot + (size_t)n * kTP);
  const float4 a = src[ty], b = src[16 + ty];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// [pt][ch] scratch slot (TMA operand): o[j4][im] = value of channel half*64 + tx*4 + j4 at point row_of(ty, im)
__device__ __forceinline__ void store_pc(float* slot, int half, int tx, int ty, const float (&o)[4][8]) {
#pragma unroll
  for (int im = 0; im < 8; ++im) {
    float4* dst = reinterpret_cast<float4*>(slot + (size_t)row_of(ty, im) * kW + half * 64 + tx * 4);
    *dst = make_float4(o[0][im], o[1][im], o[2][im], o[3][im]);
  }
}

// sum over the points held by the 4 lanes of this warp that share a channel, then one atomic per channel
__device__ __forceinline__ void chan_add(float v, float* dst, int lane) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  if ((lane >> 3) == 0) atomicAdd(dst, v);
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int J>
__device__ __forceinline__ void narrow_contract(BwdSmem& sm, const float* __restrict__ V, float (&out)[J], int tid) {
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__global__ void __launch_bounds__(kThreads, 2) mlp_bwd_kernel(const BwdKArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = (warp & 1) * 8 + (lane & 7);
  const int ty = (warp >> 1) * 4 + (lane >> 3);
  const int D = a.r.D;
  const BlobLayout L = blob_layout(D);
  const float* cst = a.r.blob + L.const_off;
  float* scr = a.scratch + (size_t)blockIdx.x * a.scratch_stride;

  const int my_tiles =
      (a.r.n_tiles > (int)blockIdx.x) ? (a.r.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  Pipe pipe;
  pipe.cc = 0;
  pipe.pc = 0;
  pipe.D = D;
  pipe.nf = L.n_chunks_fine;
  pipe.lw = L.n_chunks_fine + 8 + 16 * (D - 1);
  pipe.per_tile = pipe.lw + 2 * (16 * (D - 1) + 8);
  pipe.total = pipe.per_tile * my_tiles;
  pipe.limit = my_tiles > 0 ? pipe.lw : 0;
  pipe.stream = a.r.blob + L.stream_off;
  pipe.scr = scr;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) pipe_fill(sm, pipe);

  int tile_local = 0;
  for (int tile = blockIdx.x; tile < a.r.n_tiles; tile += gridDim.x, ++tile_local) {
    const int inst = tile / a.r.tiles_per_inst;
    const int tin = tile - inst * a.r.tiles_per_inst;
    const float* film = a.r.film + (size_t)inst * kFilm * 2 * kW;
    float* dfilm = a.d_film + (size_t)inst * kFilm * 2 * kW;

    // ---------------- prologue: positions and incoming per-point adjoints ----------------
    if (tid < kTP) {
      const PointCtx pc = point_prologue(a.r, inst, tin, tid, false);
      sm.act[0 * kTP + tid] = pc.px;
      sm.act[1 * kTP + tid] = pc.py;
      sm.act[2 * kTP + tid] = pc.pz;
      sm.act[3 * kTP + tid] = 0.f;
      sm.pt[0][tid] = pc.px;
      sm.pt[1][tid] = pc.py;
      sm.pt[2][tid] = pc.pz;
      float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
      if (pc.valid) {
        const float4* ap = reinterpret_cast<const float4*>(a.adj + ((size_t)pc.ray * a.r.S + pc.si) * 8);
        q0 = ap[0];
        q1 = ap[1];
      }
      sm.pt[3][tid] = q0.x;
      sm.pt[4][tid] = q0.y;
      sm.pt[5][tid] = q0.z;
      sm.pt[6][tid] = q0.w;
      sm.pt[7][tid] = q1.x;
      sm.pt[8][tid] = q1.y;
      sm.pt[9][tid] = q1.z;
      // d b_sigma = sum sdf_bar ; d b_rgb = sum zrgb_bar
      const float s0 = warp_sum(q0.x), s1 = warp_sum(q1.x), s2 = warp_sum(q1.y), s3 = warp_sum(q1.z);
      if (lane == 0) {
        atomicAdd(a.g.sigma_bias, s0);
        atomicAdd(a.g.rgb_bias + 0, s1);
        atomicAdd(a.g.rgb_bias + 1, s2);
        atomicAdd(a.g.rgb_bias + 2, s3);
      }
    }
    __syncthreads();

    float acc[8][8];
    // ---------------- recompute: forward sweep, keeps u_l (slot U[l]) and h_{l+1} (slot HT[l+1]) ----------------
    for (int l = 0; l < D; ++l) {
      zero_acc(acc);
      gemm_chunks(sm, pipe, acc, l == 0 ? 1 : 8, l == 0 ? 4 : kKC, tx, ty, tid);
      const float* gam = film + (l * 2 + 0) * kW;
      const float* bet = film + (l * 2 + 1) * kW;
      const float* bia = cst + BlobLayout::kBias + l * kW;
      float sb[8];
      if (l == D - 1) {
#pragma unroll
        for (int im = 0; im < 8; ++im) sb[im] = sm.pt[3][row_of(ty, im)];
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int jn = half * 4 + j4;
          const int n = col_of(tx, jn);
          const float g = __ldg(gam + n), be = __ldg(bet + n), bi = __ldg(bia + n);
          float uv[8];
          float dws = 0.f;
#pragma unroll
          for (int im = 0; im < 8; ++im) {
            uv[im] = acc[im][jn] + bi;
            float s, c;
            sincos_film(fmaf(g, uv[im], be), &s, &c);
            o[j4][im] = s;
            if (l == D - 1) dws = fmaf(sb[im], s, dws);
          }
          store_cp(scr + (size_t)(kSlotU + l) * kSlot, n, ty, uv);
          store_act(sm, n, ty, o[j4]);
          if (l == D - 1) chan_add(dws, a.g.sigma_weight + n, lane);  // d w_s += sdf_bar * h_D
        }
        store_pc(scr + (size_t)(kSlotHT + l) * kSlot, half, tx, ty, o);  // HT[l+1]
      }
      __syncthreads();
    }

    // ---------------- colour layer, feature part: W_cf h_D -> slot UC ----------------
    zero_acc(acc);
    gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
      float v[8];
#pragma unroll
      for (int im = 0; im < 8; ++im) v[im] = acc[im][jn];
      store_cp(scr + (size_t)kSlotUC * kSlot, col_of(tx, jn), ty, v);
    }

    // ---------------- recompute: reverse sweep, keeps g_l (slot G[l]) and t_l (slot TT[l]) ----------------
    {
      const int l = D - 1;
      const float* gam = film + (l * 2 + 0) * kW;
      const float* bet = film + (l * 2 + 1) * kW;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int n = col_of(tx, half * 4 + j4);
          const float g = __ldg(gam + n), be = __ldg(bet + n), ws = __ldg(cst + BlobLayout::kWsig + n);
          float uv[8];
          load_cp(scr + (size_t)(kSlotU + l) * kSlot, n, ty, uv);
#pragma unroll
          for (int im = 0; im < 8; ++im) {
            float s, c;
            sincos_film(fmaf(g, uv[im], be), &s, &c);
            o[j4][im] = ws * (g * c);
          }
          store_act(sm, n, ty, o[j4]);
        }
        store_pc(scr + (size_t)(kSlotTT + l - 1) * kSlot, half, tx, ty, o);  // TT[D-1]
      }
    }
    __syncthreads();
    for (int l = D - 1; l >= 1; --l) {
      zero_acc(acc);
      gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);  // g_l = W_l^T t_l
      const float* gam = film + ((l - 1) * 2 + 0) * kW;
      const float* bet = film + ((l - 1) * 2 + 1) * kW;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int jn = half * 4 + j4;
          const int n = col_of(tx, jn);
          const float g = __ldg(gam + n), be = __ldg(bet + n);
          float uv[8], gv[8];
          load_cp(scr + (size_t)(kSlotU + l - 1) * kSlot, n, ty, uv);
#pragma unroll
          for (int im = 0; im < 8; ++im) {
            float s, c;
            sincos_film(fmaf(g, uv[im], be), &s, &c);
            gv[im] = acc[im][jn];
            o[j4][im] = gv[im] * (g * c);  // t_{l-1}
          }
          store_cp(scr + (size_t)(kSlotG + l) * kSlot, n, ty, gv);
          store_act(sm, n, ty, o[j4]);
        }
        if (l - 1 >= 1) store_pc(scr + (size_t)(kSlotTT + l - 2) * kSlot, half, tx, ty, o);  // TT[l-1]
      }
      __syncthreads();
    }
    // normal = W_0^T t_0
    {
      float o[3];
      narrow_contract<3>(sm, cst + BlobLayout::kW0t, o, tid);
      if (tid < kTP) {
        sm.nrm[0][tid] = o[0];
        sm.nrm[1][tid] = o[1];
        sm.nrm[2][tid] = o[2];
      }
    }
    __syncthreads();

    // ---------------- colour layer: recompute + backward of the rgb head and the colour FiLM layer ----------------
    {
      const float* gam = film + (OI_MAX_DEPTH * 2 + 0) * kW;
      const float* bet = film + (OI_MAX_DEPTH * 2 + 1) * kW;
      const float* bia = cst + BlobLayout::kBias + OI_MAX_DEPTH * kW;
      float* dgam = dfilm + (OI_MAX_DEPTH * 2 + 0) * kW;
      float* dbet = dfilm + (OI_MAX_DEPTH * 2 + 1) * kW;
      float nx[8], ny[8], nz[8], z0[8], z1[8], z2[8];
#pragma unroll
      for (int im = 0; im < 8; ++im) {
        const int m = row_of(ty, im);
        nx[im] = sm.nrm[0][m];
        ny[im] = sm.nrm[1][m];
        nz[im] = sm.nrm[2][m];
        z0[im] = sm.pt[7][m];
        z1[im] = sm.pt[8][m];
        z2[im] = sm.pt[9][m];
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int n = col_of(tx, half * 4 + j4);
          const float g = __ldg(gam + n), be = __ldg(bet + n), bi = __ldg(bia + n);
          const float w0 = __ldg(cst + BlobLayout::kWcg + 0 * kW + n);
          const float w1 = __ldg(cst + BlobLayout::kWcg + 1 * kW + n);
          const float w2 = __ldg(cst + BlobLayout::kWcg + 2 * kW + n);
          const float r0 = __ldg(cst + BlobLayout::kWrgb + 0 * kW + n);
          const float r1 = __ldg(cst + BlobLayout::kWrgb + 1 * kW + n);
          const float r2 = __ldg(cst + BlobLayout::kWrgb + 2 * kW + n);
          float uv[8];
          load_cp(scr + (size_t)kSlotUC * kSlot, n, ty, uv);
          float s_db = 0.f, s_dg = 0.f, s_r0 = 0.f, s_r1 = 0.f, s_r2 = 0.f, s_g0 = 0.f, s_g1 = 0.f, s_g2 = 0.f;
#pragma unroll
          for (int im = 0; im < 8; ++im) {
            float pre = fmaf(w0, nx[im], uv[im]);
            pre = fmaf(w1, ny[im], pre);
            pre = fmaf(w2, nz[im], pre);
            const float uu = pre + bi;
            float s, c;
            sincos_film(fmaf(g, uu, be), &s, &c);
            const float hb = fmaf(r0, z0[im], fmaf(r1, z1[im], r2 * z2[im]));  // h_c_bar = W_rgb^T z_bar
            const float ab = hb * c;
            const float ub = ab * g;
            s_db += ab;
            s_dg = fmaf(ab, uu, s_dg);
            s_r0 = fmaf(z0[im], s, s_r0);
            s_r1 = fmaf(z1[im], s, s_r1);
            s_r2 = fmaf(z2[im], s, s_r2);
            s_g0 = fmaf(ub, nx[im], s_g0);
            s_g1 = fmaf(ub, ny[im], s_g1);
            s_g2 = fmaf(ub, nz[im], s_g2);
            o[j4][im] = ub;
          }
          store_act(sm, n, ty, o[j4]);
          chan_add(s_db, dbet + n, lane);
          chan_add(s_dg, dgam + n, lane);
          chan_add(s_r0, a.g.rgb_weight + 0 * kW + n, lane);
          chan_add(s_r1, a.g.rgb_weight + 1 * kW + n, lane);
          chan_add(s_r2, a.g.rgb_weight + 2 * kW + n, lane);
          chan_add(s_g0, a.g.views_weight + (size_t)n * (kW + 3) + kW + 0, lane);
          chan_add(s_g1, a.g.views_weight + (size_t)n * (kW + 3) + kW + 1, lane);
          chan_add(s_g2, a.g.views_weight + (size_t)n * (kW + 3) + kW + 2, lane);
        }
        store_pc(scr + (size_t)kSlotUCT * kSlot, half, tx, ty, o);
      }
    }
    __syncthreads();
    // normal_bar += W_cg^T u_bar_c
    {
      float o[3];
      narrow_contract<3>(sm, cst + BlobLayout::kWcg, o, tid);
      if (tid < kTP) {
        sm.pt[4][tid] += o[0];
        sm.pt[5][tid] += o[1];
        sm.pt[6][tid] += o[2];
      }
    }
    // ---------------- h_bar_D = W_cf^T u_bar_c + sdf_bar w_s -> slot HB ----------------
    zero_acc(acc);
    gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);
    {
      float sb[8];
#pragma unroll
      for (int im = 0; im < 8; ++im) sb[im] = sm.pt[3][row_of(ty, im)];
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        const int n = col_of(tx, jn);
        const float ws = __ldg(cst + BlobLayout::kWsig + n);
        float v[8];
#pragma unroll
        for (int im = 0; im < 8; ++im) v[im] = fmaf(sb[im], ws, acc[im][jn]);
        store_cp(scr + (size_t)kSlotHB * kSlot, n, ty, v);
      }
    }

    // ---------------- backward of the reverse sweep, l = 0: t_bar_0 = W_0 normal_bar (K = 3) ----------------
    {
      const float* gam = film + (0 * 2 + 0) * kW;
      const float* bet = film + (0 * 2 + 1) * kW;
      float b0[8], b1[8], b2[8];
#pragma unroll
      for (int im = 0; im < 8; ++im) {
        const int m = row_of(ty, im);
        b0[im] = sm.pt[4][m];
        b1[im] = sm.pt[5][m];
        b2[im] = sm.pt[6][m];
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int n = col_of(tx, half * 4 + j4);
          const float g = __ldg(gam + n), be = __ldg(bet + n);
          const float w0 = __ldg(cst + BlobLayout::kW0t + 0 * kW + n);
          const float w1 = __ldg(cst + BlobLayout::kW0t + 1 * kW + n);
          const float w2 = __ldg(cst + BlobLayout::kW0t + 2 * kW + n);
          float uv[8], gv[8], cb[8];
          load_cp(scr + (size_t)(kSlotU + 0) * kSlot, n, ty, uv);
          load_cp(scr + (size_t)(kSlotG + 1) * kSlot, n, ty, gv);  // g_1
          float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
          for (int im = 0; im < 8; ++im) {
            const float tb = fmaf(w0, b0[im], fmaf(w1, b1[im], w2 * b2[im]));
            float s, c;
            sincos_film(fmaf(g, uv[im], be), &s, &c);
            const float c0 = g * c;
            const float t0 = gv[im] * c0;
            d0 = fmaf(t0, b0[im], d0);
            d1 = fmaf(t0, b1[im], d1);
            d2 = fmaf(t0, b2[im], d2);
            cb[im] = tb * gv[im];   // c_bar_0
            o[j4][im] = tb * c0;    // g_bar_1
          }
          store_cp(scr + (size_t)(kSlotG + 1) * kSlot, n, ty, cb);
          store_act(sm, n, ty, o[j4]);
          chan_add(d0, a.g.pts_weight[0] + n * 3 + 0, lane);  // dW_0 += t_0 (x) normal_bar
          chan_add(d1, a.g.pts_weight[0] + n * 3 + 1, lane);
          chan_add(d2, a.g.pts_weight[0] + n * 3 + 2, lane);
        }
        store_pc(scr + (size_t)(kSlotGBT + 0) * kSlot, half, tx, ty, o);  // GBT[1]
      }
    }
    __syncthreads();

    // ---------------- backward of the reverse sweep, l = 1..D-1: t_bar_l = W_l g_bar_l ----------------
    for (int l = 1; l < D; ++l) {
      zero_acc(acc);
      gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);
      const float* gam = film + (l * 2 + 0) * kW;
      const float* bet = film + (l * 2 + 1) * kW;
      float* dgam = dfilm + (l * 2 + 0) * kW;
      float* dbet = dfilm + (l * 2 + 1) * kW;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int jn = half * 4 + j4;
          const int n = col_of(tx, jn);
          const float g = __ldg(gam + n), be = __ldg(bet + n);
          float uv[8];
          load_cp(scr + (size_t)(kSlotU + l) * kSlot, n, ty, uv);
          if (l < D - 1) {
            float gv[8], cb[8];
            load_cp(scr + (size_t)(kSlotG + l + 1) * kSlot, n, ty, gv);  // g_{l+1}
#pragma unroll
            for (int im = 0; im < 8; ++im) {
              float s, c;
              sincos_film(fmaf(g, uv[im], be), &s, &c);
              const float tb = acc[im][jn];
              cb[im] = tb * gv[im];         // c_bar_l
              o[j4][im] = tb * (g * c);     // g_bar_{l+1}
            }
            store_cp(scr + (size_t)(kSlotG + l + 1) * kSlot, n, ty, cb);
          } else {
            // top of the reverse sweep: t_{D-1} = w_s c_{D-1}; then straight into the backward of the forward
            // sweep for layer D-1 (h_bar_D is parked in slot HB)
            const float ws = __ldg(cst + BlobLayout::kWsig + n);
            float hb[8];
            load_cp(scr + (size_t)kSlotHB * kSlot, n, ty, hb);
            float s_ws = 0.f, s_db = 0.f, s_dg = 0.f;
#pragma unroll
            for (int im = 0; im < 8; ++im) {
              float s, c;
              sincos_film(fmaf(g, uv[im], be), &s, &c);
              const float tb = acc[im][jn];
              s_ws = fmaf(tb, g * c, s_ws);
              const float cbar = tb * ws;
              const float ab = hb[im] * c - cbar * g * s;
              s_db += ab;
              s_dg += fmaf(ab, uv[im], cbar * c);
              o[j4][im] = ab * g;  // u_bar_{D-1}
            }
            chan_add(s_ws, a.g.sigma_weight + n, lane);
            chan_add(s_db, dbet + n, lane);
            chan_add(s_dg, dgam + n, lane);
          }
          store_act(sm, n, ty, o[j4]);
        }
        if (l < D - 1) store_pc(scr + (size_t)(kSlotGBT + l) * kSlot, half, tx, ty, o);      // GBT[l+1]
        else store_pc(scr + (size_t)(kSlotUBT + l - 1) * kSlot, half, tx, ty, o);            // UBT[D-1]
      }
      __syncthreads();
    }

    // ---------------- backward of the forward sweep: h_bar_l = W_l^T u_bar_l, then layer l-1 ----------------
    for (int l = D - 1; l >= 1; --l) {
      zero_acc(acc);
      gemm_chunks(sm, pipe, acc, 8, kKC, tx, ty, tid);
      const int k = l - 1;
      const float* gam = film + (k * 2 + 0) * kW;
      const float* bet = film + (k * 2 + 1) * kW;
      float* dgam = dfilm + (k * 2 + 0) * kW;
      float* dbet = dfilm + (k * 2 + 1) * kW;
      float x0[8], x1[8], x2[8];
      if (k == 0) {
#pragma unroll
        for (int im = 0; im < 8; ++im) {
          const int m = row_of(ty, im);
          x0[im] = sm.pt[0][m];
          x1[im] = sm.pt[1][m];
          x2[im] = sm.pt[2][m];
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float o[4][8];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int jn = half * 4 + j4;
          const int n = col_of(tx, jn);
          const float g = __ldg(gam + n), be = __ldg(bet + n);
          float uv[8], cb[8];
          load_cp(scr + (size_t)(kSlotU + k) * kSlot, n, ty, uv);
          load_cp(scr + (size_t)(kSlotG + k + 1) * kSlot, n, ty, cb);  // c_bar_k
          float s_db = 0.f, s_dg = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
          for (int im = 0; im < 8; ++im) {
            float s, c;
            sincos_film(fmaf(g, uv[im], be), &s, &c);
            const float ab = acc[im][jn] * c - cb[im] * g * s;
            s_db += ab;
            s_dg += fmaf(ab, uv[im], cb[im] * c);
            const float ub = ab * g;
            o[j4][im] = ub;
            if (k == 0) {
              d0 = fmaf(ub, x0[im], d0);
              d1 = fmaf(ub, x1[im], d1);
              d2 = fmaf(ub, x2[im], d2);
            }
          }
          chan_add(s_db, dbet + n, lane);
          chan_add(s_dg, dgam + n, lane);
          if (k == 0) {
            chan_add(d0, a.g.pts_weight[0] + n * 3 + 0, lane);  // dW_0 += u_bar_0 (x) x
            chan_add(d1, a.g.pts_weight[0] + n * 3 + 1, lane);
            chan_add(d2, a.g.pts_weight[0] + n * 3 + 2, lane);
          } else {
            store_act(sm, n, ty, o[j4]);
          }
        }
        if (k >= 1) store_pc(scr + (size_t)(kSlotUBT + k - 1) * kSlot, half, tx, ty, o);  // UBT[k]
      }
      __syncthreads();
    }

    // ---------------- weight gradients: contraction over the 128 points of the tile ----------------
    asm volatile("fence.proxy.async.global;" ::: "memory");  // scratch written above is read by TMA below
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
      const int next = (tile_local + 1) * pipe.per_tile + pipe.lw;
      pipe.limit = next < pipe.total ? next : pipe.total;
      pipe_fill(sm, pipe);
    }
    for (int l = 1; l < D; ++l) {
      zero_acc(acc);
      wgrad_chunks(sm, pipe, acc, 16, tx, ty, tid);  // u_bar_l (x) h_l  +  t_l (x) g_bar_l
      float* dst = a.g.pts_weight[l];
#pragma unroll
      for (int im = 0; im < 8; ++im) {
        float* row = dst + (size_t)row_of(ty, im) * kW;
        red_add_v4(row + tx * 4, acc[im][0], acc[im][1], acc[im][2], acc[im][3]);
        red_add_v4(row + 64 + tx * 4, acc[im][4], acc[im][5], acc[im][6], acc[im][7]);
      }
    }
    zero_acc(acc);
    wgrad_chunks(sm, pipe, acc, 8, tx, ty, tid);  // u_bar_c (x) h_D
#pragma unroll
    for (int im = 0; im < 8; ++im) {
      float* row = a.g.views_weight + (size_t)row_of(ty, im) * (kW + 3);
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) atomicAdd(row + col_of(tx, jn), acc[im][jn]);
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// per-ray tail
// -----------------------------------------------------------------------------------------------------
struct TailArgs {
  int R, S;
  float cos_anneal, sample_dist;
  const float *rays_o, *rays_d, *z_vals, *blob;
  int depth;
  const float *sdf, *gradients, *raw_color;
  const float *g_weights, *g_weight_sum, *g_weight_max, *g_color_fine, *g_raw_color, *g_gradients, *g_sdf, *g_cdf_fine,
      *g_s_val, *g_gradient_error, *g_surface_loss;
  float* adj;               // [N][8]
  float* invs_partial;      // [R]
  unsigned int* relax_count;
  float* d_sigma_bias;      // += sum sdf_bar, or NULL (the FFMA MLP kernel does it itself)
  float* d_rgb_bias;        // [3] += sum zrgb_bar, or NULL
};

__device__ __forceinline__ float section_dist(const TailArgs& a, const float* zr, int i) {
  return (i + 1 < a.S) ? (zr[i + 1] - zr[i]) : a.sample_dist;
}

__global__ void relax_count_kernel(const TailArgs a) {
  const size_t N = (size_t)a.R * a.S;
  unsigned int cnt = 0;
  for (size_t gp = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gp < N; gp += (size_t)gridDim.x * blockDim.x) {
    const int ray = (int)(gp / a.S), i = (int)(gp - (size_t)ray * a.S);
    const float* zr = a.z_vals + (size_t)ray * a.S;
    const float mid = zr[i] + section_dist(a, zr, i) * 0.5f;
    const float px = a.rays_o[ray * 3 + 0] + a.rays_d[ray * 3 + 0] * mid;
    const float py = a.rays_o[ray * 3 + 1] + a.rays_d[ray * 3 + 1] * mid;
    const float pz = a.rays_o[ray * 3 + 2] + a.rays_d[ray * 3 + 2] * mid;
    cnt += (sqrtf(px * px + py * py + pz * pz) < 1.2f) ? 1u : 0u;
  }
  for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(a.relax_count, cnt);
}

struct AlphaTerms {
  float tc, hs, p, q, raw, alpha, dist;
};
__device__ __forceinline__ AlphaTerms alpha_terms(const TailArgs& a, const float* zr, int i, float sdf, float gx,
                                                  float gy, float gz, float dx, float dy, float dz, float inv_s) {
  AlphaTerms t;
  t.dist = section_dist(a, zr, i);
  t.tc = dx * gx + dy * gy + dz * gz;
  const float ic = -(fmaxf(-t.tc * 0.5f + 0.5f, 0.f) * (1.0f - a.cos_anneal) + fmaxf(-t.tc, 0.f) * a.cos_anneal);
  t.hs = ic * t.dist * 0.5f;
  t.p = sigmoidf_acc((sdf - t.hs) * inv_s);
  t.q = sigmoidf_acc((sdf + t.hs) * inv_s);
  t.raw = (t.p - t.q + 1e-5f) / (t.p + 1e-5f);
  t.alpha = fminf(fmaxf(t.raw, 0.f), 1.f);
  return t;
}

// One warp per ray, lanes = consecutive samples (coalesced): pass 1 front to back (alpha, exclusive transmittance
// by a warp product scan with carry -- the same scan as composite_kernel --, argmax of the weights), pass 2 back to
// front (suffix sums by a reverse warp scan with carry).
__global__ void tail_bwd_kernel(const TailArgs a) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int ray_raw = blockIdx.x * wpb + wib;
  const bool live = ray_raw < a.R;
  const int ray = live ? ray_raw : a.R - 1;   // dead warps shadow the last ray (no stores)
  const int S = a.S;
  const BlobLayout L = blob_layout(a.depth);
  const float inv_s = a.blob[L.const_off + BlobLayout::kScalars + 4];
  const float* zr = a.z_vals + (size_t)ray * S;
  const float ox = a.rays_o[ray * 3 + 0], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  const float dx = a.rays_d[ray * 3 + 0], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
  const size_t base = (size_t)ray * S;
  const float gws = a.g_weight_sum ? a.g_weight_sum[ray] : 0.f;
  const float gwm = a.g_weight_max ? a.g_weight_max[ray] : 0.f;
  float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f;
  if (a.g_color_fine) {
    gc0 = a.g_color_fine[ray * 3 + 0];
    gc1 = a.g_color_fine[ray * 3 + 1];
    gc2 = a.g_color_fine[ray * 3 + 2];
  }
  const float ge_coef = a.g_gradient_error ? a.g_gradient_error[0] / ((float)(*a.relax_count) + 1e-5f) : 0.f;
  const float sl_coef = a.g_surface_loss ? a.g_surface_loss[0] / ((float)a.R * (float)S) : 0.f;
  const int n_blk = (S + 31) / 32;

  // ---- pass 1: alpha_i and T_i parked in adj[.][0] / adj[.][7]; argmax of w_i = alpha_i T_i (first maximum)
  float carry = 1.0f, wmax = -1e30f;
  int imax = 0;
  for (int blk = 0; blk < n_blk; ++blk) {
    const int i = blk * 32 + lane;
    const bool ok = i < S;
    const size_t gp = base + (ok ? i : S - 1);
    const AlphaTerms t = alpha_terms(a, zr, ok ? i : S - 1, a.sdf[gp], a.gradients[gp * 3], a.gradients[gp * 3 + 1],
                                     a.gradients[gp * 3 + 2], dx, dy, dz, inv_s);
    const float alpha = ok ? t.alpha : 0.f;
    float incl = ok ? (1.0f - alpha + 1e-7f) : 1.0f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl *= o;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float T = carry * excl;
    carry = carry * __shfl_sync(0xffffffffu, incl, 31);
    if (ok && live) {
      a.adj[gp * 8 + 0] = alpha;
      a.adj[gp * 8 + 7] = T;
    }
    const float w = ok ? alpha * T : -1e30f;
    if (w > wmax) {
      wmax = w;
      imax = i;
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const float ow = __shfl_xor_sync(0xffffffffu, wmax, d);
    const int oi = __shfl_xor_sync(0xffffffffu, imax, d);
    if (ow > wmax || (ow == wmax && oi < imax)) {
      wmax = ow;
      imax = oi;
    }
  }
  __syncwarp();

  // ---- pass 2: back to front
  float suffix_carry = 0.f, invs_bar = 0.f;
  float sum_sb = 0.f, sum_z0 = 0.f, sum_z1 = 0.f, sum_z2 = 0.f;
  float amax = 0.f;   // max |adjoint component| over this warp's points
  for (int blk = n_blk - 1; blk >= 0; --blk) {
    const int i = blk * 32 + lane;
    const bool ok = i < S;
    const size_t gp = base + (ok ? i : S - 1);
    const float sdf = a.sdf[gp];
    const float gx = a.gradients[gp * 3], gy = a.gradients[gp * 3 + 1], gz = a.gradients[gp * 3 + 2];
    const float r = a.raw_color[gp * 3], g = a.raw_color[gp * 3 + 1], b = a.raw_color[gp * 3 + 2];
    const AlphaTerms t = alpha_terms(a, zr, ok ? i : S - 1, sdf, gx, gy, gz, dx, dy, dz, inv_s);
    const float alpha = live ? a.adj[gp * 8 + 0] : 0.f, Ti = live ? a.adj[gp * 8 + 7] : 0.f;
    const float w = alpha * Ti;
    float wbar = gws + gc0 * r + gc1 * g + gc2 * b;
    if (a.g_weights) wbar += a.g_weights[gp];
    if (i == imax) wbar += gwm;
    // suffix_i = sum_{k > i} wbar_k w_k: reverse exclusive scan inside the block + carry from the later blocks
    const float ww = ok ? wbar * w : 0.f;
    float incl = ww;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_down_sync(0xffffffffu, incl, d);
      if (lane + d < 32) incl += o;
    }
    const float suffix = suffix_carry + (incl - ww);
    suffix_carry += __shfl_sync(0xffffffffu, incl, 0);
    if (!ok) continue;
    const float f = 1.0f - alpha + 1e-7f;
    const float alpha_bar = wbar * Ti - suffix / f;
    const float raw_bar = (t.raw >= 0.f && t.raw <= 1.f) ? alpha_bar : 0.f;
    const float pe = t.p + 1e-5f;
    float p_bar = raw_bar * t.q / (pe * pe);
    if (a.g_cdf_fine) p_bar += a.g_cdf_fine[gp];
    const float q_bar = -raw_bar / pe;
    const float A_bar = p_bar * t.p * (1.0f - t.p), B_bar = q_bar * t.q * (1.0f - t.q);
    float sdf_bar = (A_bar + B_bar) * inv_s;
    if (a.g_sdf) sdf_bar += a.g_sdf[gp];
    const float hs_bar = (B_bar - A_bar) * inv_s;
    invs_bar += A_bar * (sdf - t.hs) + B_bar * (sdf + t.hs);
    const float ic_bar = hs_bar * t.dist * 0.5f;
    const float tc_bar = ic_bar * (0.5f * (1.0f - a.cos_anneal) * ((-t.tc * 0.5f + 0.5f > 0.f) ? 1.f : 0.f) +
                                   a.cos_anneal * ((-t.tc > 0.f) ? 1.f : 0.f));
    float nb0 = tc_bar * dx, nb1 = tc_bar * dy, nb2 = tc_bar * dz;
    if (a.g_gradients) {
      nb0 += a.g_gradients[gp * 3 + 0];
      nb1 += a.g_gradients[gp * 3 + 1];
      nb2 += a.g_gradients[gp * 3 + 2];
    }
    if (a.g_gradient_error) {
      const float mid = zr[i] + t.dist * 0.5f;
      const float px = ox + dx * mid, py = oy + dy * mid, pz = oz + dz * mid;
      if (sqrtf(px * px + py * py + pz * pz) < 1.2f) {
        const float nn = sqrtf(gx * gx + gy * gy + gz * gz);
        const float k = ge_coef * 2.0f * (nn - 1.0f) / nn;
        nb0 = fmaf(k, gx, nb0);
        nb1 = fmaf(k, gy, nb1);
        nb2 = fmaf(k, gz, nb2);
      }
    }
    if (a.g_surface_loss) {
      const float sg = (sdf > 0.f) ? 1.f : ((sdf < 0.f) ? -1.f : 0.f);
      sdf_bar += sl_coef * (-100.0f * sg) * expf(-100.0f * fabsf(sdf));
    }
    float rb0 = w * gc0, rb1 = w * gc1, rb2 = w * gc2;
    if (a.g_raw_color) {
      rb0 += a.g_raw_color[gp * 3 + 0];
      rb1 += a.g_raw_color[gp * 3 + 1];
      rb2 += a.g_raw_color[gp * 3 + 2];
    }
    const float z0 = rb0 * r * (1.0f - r), z1 = rb1 * g * (1.0f - g), z2 = rb2 * b * (1.0f - b);
    if (live) {
      float4* out = reinterpret_cast<float4*>(a.adj + gp * 8);
      out[0] = make_float4(sdf_bar, nb0, nb1, nb2);
      out[1] = make_float4(z0, z1, z2, 0.f);
      amax = fmaxf(amax, fmaxf(fmaxf(fmaxf(fabsf(sdf_bar), fabsf(nb0)), fmaxf(fabsf(nb1), fabsf(nb2))),
                               fmaxf(fmaxf(fabsf(z0), fabsf(z1)), fabsf(z2))));
      sum_sb += sdf_bar;
      sum_z0 += z0;
      sum_z1 += z1;
      sum_z2 += z2;
    }
  }
  __syncwarp();
  invs_bar = warp_sum(invs_bar);
  for (int d = 16; d >= 1; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
  if (live && lane == 0) {
    if (a.g_s_val) invs_bar -= a.g_s_val[ray] / (inv_s * inv_s);
    a.invs_partial[ray] = invs_bar;
    // global max |adj| (non-negative floats order like their bit patterns; inf / NaN saturate the exponent field and
    // bwd_mode() then keeps the TF32 operands)
    if (amax > 0.f) atomicMax(a.relax_count + 1, __float_as_uint(amax));
  }
  if (a.d_sigma_bias) {
    sum_sb = warp_sum(sum_sb);
    sum_z0 = warp_sum(sum_z0);
    sum_z1 = warp_sum(sum_z1);
    sum_z2 = warp_sum(sum_z2);
    if (live && lane == 0) {
      atomicAdd(a.d_sigma_bias, sum_sb);
      atomicAdd(a.d_rgb_bias + 0, sum_z0);
      atomicAdd(a.d_rgb_bias + 1, sum_z1);
      atomicAdd(a.d_rgb_bias + 2, sum_z2);
    }
  }
}

// db_l = sum_inst gamma * dbeta (u_bar = a_bar * gamma), d variance = inv_s_bar * 10 * inv_s (inside the clip).
__global__ void finalize_bwd_kernel(int D, int n_inst, int R, const float* __restrict__ film,
                                    const float* __restrict__ d_film, const float* __restrict__ invs_partial,
                                    const float* __restrict__ blob, OiNetGrads g) {
  const int n = threadIdx.x;  // 128 threads
  const int l = blockIdx.x;   // 0..D-1 SDF layers, D = colour layer, D+1 = variance
  if (l <= D) {
    const int slot = (l < D) ? l : OI_MAX_DEPTH;
    float s = 0.f;
    for (int i = 0; i < n_inst; ++i) {
      const size_t o = ((size_t)i * kFilm + slot) * 2 * kW;
      s = fmaf(film[o + n], d_film[o + kW + n], s);
    }
    float* dst = (l < D) ? g.pts_bias[l] : g.views_bias;
    dst[n] += s;
    // split the (dgamma, dbeta) table into the two caller-visible tensors
    for (int i = 0; i < n_inst; ++i) {
      const size_t o = ((size_t)i * kFilm + slot) * 2 * kW;
      g.film_gamma[((size_t)i * kFilm + slot) * kW + n] += d_film[o + n];
      g.film_beta[((size_t)i * kFilm + slot) * kW + n] += d_film[o + kW + n];
    }
  } else {
    __shared__ double red[4];
    double s = 0.0;
    for (int r = n; r < R; r += blockDim.x) s += (double)invs_partial[r];
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((n & 31) == 0) red[n >> 5] = s;
    __syncthreads();
    if (n == 0) {
      const BlobLayout L = blob_layout(D);
      const float inv_s = blob[L.const_off + BlobLayout::kScalars + 4];
      const double tot = red[0] + red[1] + red[2] + red[3];
      const bool inside = inv_s > 1e-6f && inv_s < 1e6f;  // clip(exp(10 v), 1e-6, 1e6), fields.py:267-268
      if (inside) g.variance[0] += (float)(tot * 10.0 * (double)inv_s);
    }
  }
}

}  // namespace

int render_bwd_ctas(int n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ctas = 2 * sms;
  if (ctas > n_tiles) ctas = n_tiles;
  return ctas < 1 ? 1 : ctas;
}
size_t render_bwd_scratch_floats() { return (size_t)kNumSlots * kSlot; }

// relax count + per-ray tail: fills adj [N][8] and invs_partial [R]; zeroes d_film.
// Adjoint statistics for bwd_mode() (oi_wgrad.cuh): total "mass" sum_m A_m / 2^e_max in 2^-20 fixed point (A_m = max
// |adjoint component| of point m) and the mass of the points more than 2^kF16LowShift below the global maximum --
// integer sums, so the decision is the same in every run.
__global__ void adj_stats_kernel(const float* __restrict__ adj, size_t n_points, unsigned int* ctl,
                                 const unsigned int* __restrict__ guard) {
  // snapshot of the workspace's fp16 overflow guard: every kernel of this call decides on the SAME state, also when
  // the sweep trips the guard half-way through the call (the trip then takes effect from the next call on)
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl[kCtlGuardSnapshot] = *guard;
  const unsigned int mb = ctl[1];
  const int ex = (int)((mb >> 23) & 0xFFu);
  if (ex == 0 || ex == 255) return;
  const int e_max = ex - 127;
  const float unit = pow2i(20 - e_max);
  unsigned long long tot = 0, low = 0;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += (size_t)gridDim.x * blockDim.x) {
    const float4* q = reinterpret_cast<const float4*>(adj + p * 8);
    const float4 q0 = q[0], q1 = q[1];
    const float am = fmaxf(fmaxf(fmaxf(fabsf(q0.x), fabsf(q0.y)), fmaxf(fabsf(q0.z), fabsf(q0.w))),
                           fmaxf(fmaxf(fabsf(q1.x), fabsf(q1.y)), fabsf(q1.z)));
    const unsigned long long v = (unsigned long long)(am * unit);
    tot += v;
    if (adj_exponent(q0, q1) < e_max - kF16LowShift) low += v;
  }
  for (int d = 16; d >= 1; d >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, d);
    low += __shfl_xor_sync(0xffffffffu, low, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(ctl + 2), tot);
    if (low) atomicAdd(reinterpret_cast<unsigned long long*>(ctl + 4), low);
  }
}

int launch_bwd_tail(const OiRenderBwdDesc& d, const RenderKArgs& geo, float* adj, float* invs_partial,
                    unsigned int* relax_count, const unsigned int* guard, float* d_film, bool head_biases,
                    cudaStream_t st) {
  TailArgs t;
  t.R = d.n_rays;
  t.S = d.n_samples_total;
  t.cos_anneal = d.cos_anneal_ratio;
  t.sample_dist = geo.sample_dist;
  t.rays_o = d.rays_o;
  t.rays_d = d.rays_d;
  t.z_vals = d.z_vals;
  t.blob = geo.blob;
  t.depth = d.depth;
  t.sdf = d.sdf;
  t.gradients = d.gradients;
  t.raw_color = d.raw_color;
  t.g_weights = d.g_weights;
  t.g_weight_sum = d.g_weight_sum;
  t.g_weight_max = d.g_weight_max;
  t.g_color_fine = d.g_color_fine;
  t.g_raw_color = d.g_raw_color;
  t.g_gradients = d.g_gradients;
  t.g_sdf = d.g_sdf;
  t.g_cdf_fine = d.g_cdf_fine;
  t.g_s_val = d.g_s_val;
  t.g_gradient_error = d.g_gradient_error;
  t.g_surface_loss = d.g_surface_loss;
  t.adj = adj;
  t.invs_partial = invs_partial;
  t.relax_count = relax_count;
  t.d_sigma_bias = head_biases ? d.grads.sigma_bias : nullptr;
  t.d_rgb_bias = head_biases ? d.grads.rgb_bias : nullptr;
  OI_CHECK_CUDA(cudaMemsetAsync(relax_count, 0, 256, st));
  OI_CHECK_CUDA(cudaMemsetAsync(d_film, 0, (size_t)geo.n_inst * kFilm * 2 * kW * sizeof(float), st));
  if (d.g_gradient_error) {
    relax_count_kernel<<<296, 256, 0, st>>>(t);
    OI_CHECK_CUDA(cudaGetLastError());
  }
  tail_bwd_kernel<<<(t.R + 7) / 8, 256, 0, st>>>(t);   // 8 warps per block, one ray per warp
  OI_CHECK_CUDA(cudaGetLastError());
  if (head_biases && !(d.flags & (OI_BWD_FLAG_FORCE_TF32 | OI_BWD_FLAG_FORCE_F16))) {   // tensor-core backward, automatic
    adj_stats_kernel<<<296, 256, 0, st>>>(adj, (size_t)t.R * t.S, relax_count, guard);
    OI_CHECK_CUDA(cudaGetLastError());
  }
  return OI_OK;
}

// FP32-FFMA MLP backward (one kernel) + finalize.
int launch_render_bwd_ffma(const OiRenderBwdDesc& d, const RenderKArgs& geo, const float* adj,
                           const float* invs_partial, float* d_film, float* scratch, int n_ctas, cudaStream_t st) {
  BwdKArgs a;
  a.r = geo;
  a.adj = adj;
  a.scratch = scratch;
  a.scratch_stride = render_bwd_scratch_floats();
  a.g = d.grads;
  a.d_film = d_film;
  OI_CHECK_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(BwdSmem)));
  static_assert(sizeof(BwdSmem) <= 113 * 1024, "two CTAs per SM");
  if (d.evt_core_start) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d.evt_core_start), st));
  mlp_bwd_kernel<<<n_ctas, kThreads, sizeof(BwdSmem), st>>>(a);
  OI_CHECK_CUDA(cudaGetLastError());
  if (d.evt_core_stop) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d.evt_core_stop), st));

  finalize_bwd_kernel<<<d.depth + 2, kW, 0, st>>>(d.depth, geo.n_inst, d.n_rays, geo.film, d_film, invs_partial,
                                                  geo.blob, d.grads);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
