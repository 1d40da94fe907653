// tcgen05 first kernel of the render backward (OI_IMPL_TCGEN05): every per-point sweep of the reverse-mode
// algorithm of oi_render_bwd.cu / oracle/backward_oracle.py on the tensor cores, in the structure of the forward
// core (oi_render_tc.cu): one persistent CTA per SM, two 128-point tile slots, row m of a tile <-> TMEM lane m <->
// two epilogue threads (64 channels each), activations / adjoints written straight into TMEM as the A operand of
// the next tcgen05.mma, 64 KB weight panels streamed by TMA through a 3-stage ring.
//
// Per tile, 4D-2 MMA layers:
//   recompute  forward l=1..D-1, colour features, reverse l=D-1..1      (fp16 2-term split, panels of the forward)
//   adjoint    colour^T, backward-of-reverse l=1..D-1 (operand W_l), backward-of-forward l=D-1..1 (operand W_l^T)
//              (bf16 2-term split: adjoints have no bounded range; unscaled bf16 panels)
// Contractions over POINTS (weight gradients, per-channel sums) are not done here: every per-point quantity they
// need is left as an fp32 slab [32 channel-quads][128 points] float4 per tile (oi_wgrad.cuh) and contracted by
// wgrad_tc_kernel.  Warp roles: 0-7 / 8-15 epilogue of slot 0 / 1, 16 TMA producer, 17 MMA issuer.
#include "oi_internal.cuh"
#include "oi_render_common.cuh"
#include "oi_tc.cuh"
#include "oi_wgrad.cuh"

#ifndef OI_BWD_TIMING_FLAGS
#define OI_BWD_TIMING_FLAGS 0
#endif

namespace oi {

namespace {

constexpr int kTcThreads = 576;
constexpr int kEpiThreadsPerSlot = 256;
constexpr int kProducerWarp = 16, kMmaWarp = 17;
constexpr int kTcStages = 3;
constexpr int kPanelBytes = 65536;
constexpr int kSubPanelBytes = 16384;
constexpr float kWScale = 256.0f;
constexpr float kInvWScale = 1.0f / 256.0f;
constexpr uint32_t kIdescF16 = tc::make_idesc_f16(128, 128);
constexpr uint32_t kIdescBf16 = tc::make_idesc_f16(128, 128) | (1u << 7) | (1u << 10);

// per-CTA scratch slabs of one slot ([32 quads][128 points] float4 each)
constexpr int kCtaG = 0;    // G[l] -> index l-1 (l = 1..7); later holds c_bar_{l-1}
constexpr int kCtaUC = 7;   // 2^8 * W_cf h_D
constexpr int kCtaHB = 8;   // h_bar_D
constexpr int kCtaSlabs = 9;

struct BwdTcArgs {
  RenderKArgs r;
  const float* adj;        // [N][8]
  float* scratch;
  size_t scratch_stride;   // floats per CTA
  float* slabs;            // [tiles of this launch][kSlabsPerTile][32][128] float4
  float* aux;              // [tiles of this launch][16][128]
  int tile_begin, tile_end;
  OiNetGrads g;            // per-channel sums over points are reduced here (warp butterflies + atomics)
  float* d_film;           // [n_inst][9][2][128] (dgamma, db)
};

struct __align__(1024) BwdTcSmem {
  unsigned char w[kTcStages][kPanelBytes];
  float2 film[2][kFilm][kW];   // per slot: (gamma', delta) in the pair layout of film_kernel
  float4 w0[kW];               // (W_0[n][0..2], 0)
  float4 head[kW];             // 2^8 * (w_sigma[n], wc_grad[0..2][n])
  float4 rgbw[kW];             // (W_rgb[0..2][n], 0)
  unsigned long long w_full[kTcStages], w_empty[kTcStages];
  float xch[2][128][8];
  unsigned long long acc_full[2], a_ready[2];
  uint32_t tmem_base;
};
static_assert(sizeof(BwdTcSmem) <= 227 * 1024, "BwdTcSmem exceeds the 227 KB per-CTA limit");

using tc::issue_split_layer_mmas;
using tc::named_bar_sync;
using tc::split2_bf16;

// round-to-nearest TF32 (the contraction kernel feeds these values to kind::tf32 MMAs, which ignore the low bits)
__device__ __forceinline__ float tf32r(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// Sum over the 32 lanes of a warp (= 32 sample points) of 16 per-lane values (= 16 channels), by recursive halving:
// after the four exchange steps lane L holds channel 8 b4 + 4 b3 + 2 b2 + b1 (bits of L) summed over 16 lanes; one
// more exchange completes the sum, and the even lanes add it to dst[channel * stride].  16 shuffles per call.
__device__ __forceinline__ void colsum16_impl(const float (&v)[16], float* dst, int stride, int lane) {
  float a8[8], a4[4], a2[2];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float keep = up ? v[i + 8] : v[i], send = up ? v[i] : v[i + 8];
      a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float keep = up ? a8[i + 4] : a8[i], send = up ? a8[i] : a8[i + 4];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float keep = up ? a4[i + 2] : a4[i], send = up ? a4[i] : a4[i + 2];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool up = (lane & 2) != 0;
  float a1 = (up ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  if ((lane & 1) == 0) atomicAdd(dst + (size_t)ch * stride, a1);
}

#define colsum16(v, dst, stride, lane)                  \
  do {                                                  \
    if (!skip_cols) colsum16_impl(v, dst, stride, lane); \
  } while (0)

__global__ void __launch_bounds__(kTcThreads, 1) bwd_tc_kernel(const BwdTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  BwdTcSmem& sm = *reinterpret_cast<BwdTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.r.D;
  const BlobLayout L = blob_layout(D);
  const float* cst = a.r.blob + L.const_off;
  const unsigned char* panels_f16 = reinterpret_cast<const unsigned char*>(a.r.blob + L.tc_off);
  const unsigned char* panels_bf16 = reinterpret_cast<const unsigned char*>(a.r.blob + L.tcb_off);
  const int NR = 2 * (D - 1) + 1;   // recompute panels (fp16)
  const int NP = 4 * D - 2;         // all MMA layers per tile
  const int n_tiles = a.tile_end - a.tile_begin;
  const int n_pairs = (n_tiles + 1) / 2;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sm.a_ready[t], kEpiThreadsPerSlot);
      mbar_init(&sm.acc_full[t], 1);
    }
    mbar_fence_init();
  }
  if (warp == kProducerWarp) {
    tc::tmem_alloc(&sm.tmem_base, 512);
    tc::tmem_relinquish();
  }
  for (int n = tid; n < kW; n += kTcThreads) {
    sm.w0[n] = make_float4(cst[BlobLayout::kW0t + n], cst[BlobLayout::kW0t + kW + n], cst[BlobLayout::kW0t + 2 * kW + n], 0.f);
    sm.head[n] = make_float4(kWScale * cst[BlobLayout::kWsig + n], kWScale * cst[BlobLayout::kWcg + n],
                             kWScale * cst[BlobLayout::kWcg + kW + n], kWScale * cst[BlobLayout::kWcg + 2 * kW + n]);
    sm.rgbw[n] = make_float4(cst[BlobLayout::kWrgb + n], cst[BlobLayout::kWrgb + kW + n],
                             cst[BlobLayout::kWrgb + 2 * kW + n], 0.f);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
        for (int p = 0; p < NP; ++p, ++it) {
          const int stage = it % kTcStages;
          if (it >= kTcStages) mbar_wait_sleep(&sm.w_empty[stage], ((it / kTcStages) - 1) & 1);
          mbar_expect_tx(&sm.w_full[stage], kPanelBytes);
          // adjoint panels: [colour^T | forward orientation l=1..D-1 | reverse orientation l=D-1..1]
          const unsigned char* src = (p < NR) ? panels_f16 + (size_t)p * kPanelBytes
                                              : panels_bf16 + (size_t)(p - NR) * kPanelBytes;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            tma_bulk_g2s(sm.w[stage] + q * kSubPanelBytes, src + q * kSubPanelBytes, kSubPanelBytes, &sm.w_full[stage]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int it = 0;
      uint32_t ar_phase[2] = {0u, 0u};
      for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
        const int n_active = (2 * pi + 1 < n_tiles) ? 2 : 1;
        for (int p = 0; p < NP; ++p, ++it) {
          const int stage = it % kTcStages;
          mbar_wait_sleep(&sm.w_full[stage], (it / kTcStages) & 1);
          const uint32_t wbase = smem_u32(sm.w[stage]);
          const uint32_t idesc = (p < NR) ? kIdescF16 : kIdescBf16;
          for (int t = 0; t < n_active; ++t) {
            mbar_wait_sleep(&sm.a_ready[t], ar_phase[t], 2000u);
            ar_phase[t] ^= 1u;
            tc::fence_after_thread_sync();
            const uint32_t acc = tmem_base + t * 256;
            issue_split_layer_mmas(acc, acc + 128, acc + 192, wbase, idesc);
            tc::mma_commit(&sm.acc_full[t]);
          }
          tc::mma_commit(&sm.w_empty[stage]);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int t = warp >> 3;
    const int h = (warp >> 2) & 1;
    const int m = (warp & 3) * 32 + lane;
    const int sub = tid & (kEpiThreadsPerSlot - 1);
    const int n0 = h * 64;
    const int Q0 = h * 16;   // first channel quad of this thread
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t acc = tmem_base + t * 256 + lane_field + n0;
    const uint32_t a_hi = tmem_base + t * 256 + lane_field + 128 + h * 32;
    const uint32_t a_lo = a_hi + 64;
    float4* scr4 = reinterpret_cast<float4*>(a.scratch + (size_t)blockIdx.x * a.scratch_stride +
                                             (size_t)t * kCtaSlabs * kSlabFloats) + m;
    uint32_t af_phase = 0u;
    // timing experiments only (results invalid; build with -DOI_BWD_TIMING_FLAGS=1): OiRenderBwdDesc.flags bit 3 =
    // skip the operand-slab stores, bit 4 = skip the column-sum butterflies (profiles/r01_summary.md)
#if OI_BWD_TIMING_FLAGS
    const bool skip_ops = (a.r.flags & 8) != 0, skip_cols = (a.r.flags & 16) != 0;
#else
    constexpr bool skip_ops = false, skip_cols = false;
#endif
#define OI_CTA(slab, quad) scr4[((size_t)(slab) * 32 + (quad)) * 128]
#define OI_GS(slab, quad) gs4[((size_t)(slab) * 32 + (quad)) * 128]   /* ARG slabs: [quad][128 points] float4 */
/* operand slabs: K-major SWIZZLE_128B tf32 image [32-point block][channel][32 points], 16-byte chunks XOR (ch & 7) */
#define OI_OP(slab, j, v)                                                                                  \
  do {                                                                                                    \
    if (!skip_ops) gso[(size_t)(slab) * kSlabFloats + (n0 + (j)) * 32 + ((mc ^ ((j) & 7)) << 2)] = tf32r(v); \
  } while (0)
#define OI_OP4(slab, j0, a0, a1, a2, a3) \
  do {                                   \
    OI_OP(slab, (j0) + 0, a0);           \
    OI_OP(slab, (j0) + 1, a1);           \
    OI_OP(slab, (j0) + 2, a2);           \
    OI_OP(slab, (j0) + 3, a3);           \
  } while (0)
#define OI_A_READY()               \
  do {                             \
    tc::wait_st();                 \
    tc::fence_before_thread_sync(); \
    mbar_arrive(&sm.a_ready[t]);   \
  } while (0)
#define OI_WAIT_ACC()                                \
  do {                                               \
    mbar_wait_sleep(&sm.acc_full[t], af_phase);      \
    af_phase ^= 1u;                                  \
    tc::fence_after_thread_sync();                   \
  } while (0)

    for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
      const int lt = 2 * pi + t;   // tile index inside this launch
      if (lt >= n_tiles) continue;
      const int tile = a.tile_begin + lt;
      const int inst = tile / a.r.tiles_per_inst;
      const int tin = tile - inst * a.r.tiles_per_inst;
      float* slab_tile = a.slabs + (size_t)lt * kSlabsPerTile * kSlabFloats;
      float4* gs4 = reinterpret_cast<float4*>(slab_tile) + m;
      float* gso = slab_tile + (m >> 5) * 4096 + (m & 3);
      const int mc = (m & 31) >> 2;
      // aux operand (N = 16, rows 0..3 used): [32-point block][4 rows][32 points], same swizzle
      float* auxo = a.aux + (size_t)lt * 512 + (m >> 5) * 128 + (m & 3);
      float* dfilm = a.d_film + (size_t)inst * kFilm * 2 * kW;   // [slot][dgamma | db][128]
      {
        const float2* src = reinterpret_cast<const float2*>(a.r.film_tc) + (size_t)inst * kFilm * kW;
        float2* dst = &sm.film[t][0][0];
        for (int i = sub; i < kFilm * kW; i += kEpiThreadsPerSlot) dst[i] = src[i];
      }
      float px, py, pz, sdf_bar, nb0, nb1, nb2, zb0, zb1, zb2;
      {
        const PointCtx pc = point_prologue(a.r, inst, tin, m, false);
        px = pc.px;
        py = pc.py;
        pz = pc.pz;
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
        if (pc.valid) {
          const float4* ap = reinterpret_cast<const float4*>(a.adj + ((size_t)pc.ray * a.r.S + pc.si) * 8);
          q0 = ap[0];
          q1 = ap[1];
        }
        sdf_bar = q0.x; nb0 = q0.y; nb1 = q0.z; nb2 = q0.w;
        zb0 = q1.x; zb1 = q1.y; zb2 = q1.z;
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);

      // =========================== recompute ===========================
      // ---------------- layer 0 (K = 3) on the FMA pipe ----------------
      {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][0]) + n0 * 2;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s[4], ar[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              const float4 w0 = sm.w0[n0 + j];
              const float u = fmaf(w0.z, pz, fmaf(w0.y, py, w0.x * px));
              ar[e] = fmaf(flf[(j >> 1) * 4 + (j & 1)], u, flf[(j >> 1) * 4 + 2 + (j & 1)]);
              s[e] = __sinf(ar[e]);
            }
            OI_GS(kSlabArg + 0, Q0 + c * 4 + q) = make_float4(ar[0], ar[1], ar[2], ar[3]);
            OI_OP4(kSlabH + 1, c * 16 + q * 4, s[0], s[1], s[2], s[3]);
            tc::split2(s[0], s[1], hi[2 * q], lo[2 * q]);
            tc::split2(s[2], s[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(a_hi + c * 8, hi);
          tc::tmem_st8(a_lo + c * 8, lo);
        }
        OI_A_READY();
      }
      // ---------------- forward layers 1..D-1 ----------------
      for (int l = 1; l < D; ++l) {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][l]) + n0 * 2;
        OI_WAIT_ACC();
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s[4], ar[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              ar[e] = fmaf(flf[(j >> 1) * 4 + (j & 1)], __uint_as_float(u[q * 4 + e]), flf[(j >> 1) * 4 + 2 + (j & 1)]);
              s[e] = __sinf(ar[e]);
            }
            OI_GS(kSlabArg + l, Q0 + c * 4 + q) = make_float4(ar[0], ar[1], ar[2], ar[3]);
            OI_OP4(kSlabH + l + 1, c * 16 + q * 4, s[0], s[1], s[2], s[3]);
            tc::split2(s[0], s[1], hi[2 * q], lo[2 * q]);
            tc::split2(s[2], s[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(a_hi + c * 8, hi);
          tc::tmem_st8(a_lo + c * 8, lo);
        }
        OI_A_READY();
      }
      // ---------------- colour features -> slot UC; t_{D-1} = w_sigma gamma cos(a_{D-1}) ----------------
      {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][D - 1]) + n0 * 2;
        OI_WAIT_ACC();
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int quad = Q0 + c * 4 + q;
            OI_CTA(kCtaUC, quad) = make_float4(__uint_as_float(u[q * 4]), __uint_as_float(u[q * 4 + 1]),
                                               __uint_as_float(u[q * 4 + 2]), __uint_as_float(u[q * 4 + 3]));
            const float4 ar4 = OI_GS(kSlabArg + D - 1, quad);
            const float ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
            float tv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              tv[e] = sm.head[n0 + j].x * flf[(j >> 1) * 4 + (j & 1)] * __cosf(ar[e]);  // 2^8 w_s * gamma/2^8 * cos
            }
            OI_OP4(kSlabT + D - 1, c * 16 + q * 4, tv[0], tv[1], tv[2], tv[3]);
            tc::split2(tv[0], tv[1], hi[2 * q], lo[2 * q]);
            tc::split2(tv[2], tv[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(a_hi + c * 8, hi);
          tc::tmem_st8(a_lo + c * 8, lo);
        }
        OI_A_READY();
      }
      // ---------------- reverse sweep l = D-1 .. 1: g_l (slot G[l]), t_{l-1} (slab T[l-1]) ----------------
      float gx = 0.f, gy = 0.f, gz = 0.f;
      for (int l = D - 1; l >= 1; --l) {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][l - 1]) + n0 * 2;
        const float gscale = (l - 1 == 0) ? kInvWScale : 1.0f;   // gamma'_0 is unscaled
        OI_WAIT_ACC();
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int quad = Q0 + c * 4 + q;
            const float4 ar4 = OI_GS(kSlabArg + l - 1, quad);
            const float ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
            float gv[4], tv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              const float accv = __uint_as_float(u[q * 4 + e]);           // 2^8 g_l
              gv[e] = accv * kInvWScale;
              tv[e] = accv * (flf[(j >> 1) * 4 + (j & 1)] * gscale) * __cosf(ar[e]);  // g_l gamma cos = t_{l-1}
            }
            OI_CTA(kCtaG + l - 1, quad) = make_float4(gv[0], gv[1], gv[2], gv[3]);
            if (l > 1) {
              OI_OP4(kSlabT + l - 1, c * 16 + q * 4, tv[0], tv[1], tv[2], tv[3]);
              tc::split2(tv[0], tv[1], hi[2 * q], lo[2 * q]);
              tc::split2(tv[2], tv[3], hi[2 * q + 1], lo[2 * q + 1]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 w = sm.w0[n0 + c * 16 + q * 4 + e];
                gx = fmaf(w.x, tv[e], gx);
                gy = fmaf(w.y, tv[e], gy);
                gz = fmaf(w.z, tv[e], gz);
              }
            }
          }
          if (l > 1) {
            tc::tmem_st8(a_hi + c * 8, hi);
            tc::tmem_st8(a_lo + c * 8, lo);
          }
        }
        if (l > 1) OI_A_READY();
      }
      // ---------------- combine the two column halves: normal ----------------
      float* xch = &sm.xch[t][m][0];
      xch[h * 4 + 1] = gx;
      xch[h * 4 + 2] = gy;
      xch[h * 4 + 3] = gz;
      named_bar_sync(1 + t, kEpiThreadsPerSlot);
      {
        const int o = (h ^ 1) * 4;
        gx += xch[o + 1];
        gy += xch[o + 2];
        gz += xch[o + 3];
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);   // exchange buffer free again

      // =========================== adjoint sweeps ===========================
      // ---------------- colour layer: recompute + backward; A <- u_bar_c (bf16) ----------------
      float nc0 = 0.f, nc1 = 0.f, nc2 = 0.f;   // W_cg^T u_bar_c, this thread's channels
      {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][OI_MAX_DEPTH]) + n0 * 2;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[8], lo[8];
          float dgs[16], sns[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int quad = Q0 + c * 4 + q;
            const float4 uc4 = OI_CTA(kCtaUC, quad);
            const float uc[4] = {uc4.x, uc4.y, uc4.z, uc4.w};
            float ar[4], ub_[4], dg[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              const float4 hd = sm.head[n0 + j];
              float pre = fmaf(hd.y, gx, uc[e]);
              pre = fmaf(hd.z, gy, pre);
              pre = fmaf(hd.w, gz, pre);
              const float gp = flf[(j >> 1) * 4 + (j & 1)];
              ar[e] = fmaf(gp, pre, flf[(j >> 1) * 4 + 2 + (j & 1)]);
              const float cs = __cosf(ar[e]);
              sns[q * 4 + e] = __sinf(ar[e]);
              const float4 rw = sm.rgbw[n0 + j];
              const float hb = fmaf(rw.x, zb0, fmaf(rw.y, zb1, rw.z * zb2));
              const float ab = hb * cs;
              dg[e] = ab * ar[e];                              // a_bar * a_c  (see finalize_bwd_tc_kernel)
              dgs[q * 4 + e] = dg[e];
              ub_[e] = ab * (gp * kWScale);                    // u_bar_c = a_bar * gamma
              nc0 = fmaf(hd.y * kInvWScale, ub_[e], nc0);
              nc1 = fmaf(hd.z * kInvWScale, ub_[e], nc1);
              nc2 = fmaf(hd.w * kInvWScale, ub_[e], nc2);
            }
            OI_OP4(kSlabUBC, c * 16 + q * 4, ub_[0], ub_[1], ub_[2], ub_[3]);
            split2_bf16(ub_[0], ub_[1], hi[2 * q], lo[2 * q]);
            split2_bf16(ub_[2], ub_[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(a_hi + c * 8, hi);
          tc::tmem_st8(a_lo + c * 8, lo);
          if (c == 3) OI_A_READY();   // the tensor core starts on W_cf^T u_bar_c while the column sums are reduced
          const int nc = n0 + c * 16;
          colsum16(dgs, dfilm + (size_t)OI_MAX_DEPTH * 2 * kW + nc, 1, lane);          // V_c = sum a_bar a
          {
            float tmp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = zb0 * sns[i];
            colsum16(tmp, a.g.rgb_weight + 0 * kW + nc, 1, lane);                       // dW_rgb = z_bar (x) h_c
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = zb1 * sns[i];
            colsum16(tmp, a.g.rgb_weight + 1 * kW + nc, 1, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = zb2 * sns[i];
            colsum16(tmp, a.g.rgb_weight + 2 * kW + nc, 1, lane);
          }
        }
      }
      // normal_bar = direct + W_cg^T u_bar_c (both halves)
      xch[h * 4 + 1] = nc0;
      xch[h * 4 + 2] = nc1;
      xch[h * 4 + 3] = nc2;
      named_bar_sync(1 + t, kEpiThreadsPerSlot);
      {
        const int o = (h ^ 1) * 4;
        nb0 += nc0 + xch[o + 1];
        nb1 += nc1 + xch[o + 2];
        nb2 += nc2 + xch[o + 3];
      }
      if (h == 0) {
        auxo[0 * 32 + ((mc ^ 0) << 2)] = tf32r(gx);
        auxo[1 * 32 + ((mc ^ 1) << 2)] = tf32r(gy);
        auxo[2 * 32 + ((mc ^ 2) << 2)] = tf32r(gz);
        auxo[3 * 32 + ((mc ^ 3) << 2)] = 1.0f;
      }
      // ---------------- h_bar_D = W_cf^T u_bar_c + sdf_bar w_s -> slot HB;
      //                  backward of the reverse sweep, l = 0 (K = 3): A <- g_bar_1 ----------------
      {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][0]) + n0 * 2;
        OI_WAIT_ACC();
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
          float t0s[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int quad = Q0 + c * 4 + q;
            float hb[4], cb[4], gb[4];
            const float4 ar4 = OI_GS(kSlabArg + 0, quad);
            const float ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
            const float4 g14 = OI_CTA(kCtaG + 0, quad);   // g_1
            const float g1[4] = {g14.x, g14.y, g14.z, g14.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              hb[e] = fmaf(sdf_bar, sm.head[n0 + j].x * kInvWScale, __uint_as_float(u[q * 4 + e]));
              const float4 w = sm.w0[n0 + j];
              const float tb = fmaf(w.x, nb0, fmaf(w.y, nb1, w.z * nb2));   // t_bar_0 = W_0 normal_bar
              const float c0 = flf[(j >> 1) * 4 + (j & 1)] * __cosf(ar[e]);  // gamma_0 cos a_0
              cb[e] = tb * g1[e];
              gb[e] = tb * c0;
              t0s[q * 4 + e] = g1[e] * c0;   // t_0
            }
            OI_CTA(kCtaHB, quad) = make_float4(hb[0], hb[1], hb[2], hb[3]);
            OI_CTA(kCtaG + 0, quad) = make_float4(cb[0], cb[1], cb[2], cb[3]);   // c_bar_0
            OI_OP4(kSlabGB + 1, c * 16 + q * 4, gb[0], gb[1], gb[2], gb[3]);
            split2_bf16(gb[0], gb[1], hi[2 * q], lo[2 * q]);
            split2_bf16(gb[2], gb[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(a_hi + c * 8, hi);
          tc::tmem_st8(a_lo + c * 8, lo);
          if (c == 3) OI_A_READY();
          {  // dW_0 += t_0 (x) normal_bar
            float* dst = a.g.pts_weight[0] + (size_t)(n0 + c * 16) * 3;
            float tmp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = nb0 * t0s[i];
            colsum16(tmp, dst + 0, 3, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = nb1 * t0s[i];
            colsum16(tmp, dst + 1, 3, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = nb2 * t0s[i];
            colsum16(tmp, dst + 2, 3, lane);
          }
        }
      }
      // ---------------- backward of the reverse sweep, l = 1..D-1: t_bar_l = W_l g_bar_l ----------------
      for (int l = 1; l < D; ++l) {
        const float* flf = reinterpret_cast<const float*>(sm.film[t][l]) + n0 * 2;
        const int gslab = (l < D - 1) ? kCtaG + l : kCtaHB;   // g_{l+1}, or h_bar_D at the top
        OI_WAIT_ACC();
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
          float dgs[16], dwss[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int quad = Q0 + c * 4 + q;
            const float4 ar4 = OI_GS(kSlabArg + l, quad);
            const float ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
            float o[4];
            if (l < D - 1) {
              const float4 g4 = OI_CTA(gslab, quad);   // g_{l+1}
              const float gn[4] = {g4.x, g4.y, g4.z, g4.w};
              float cb[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = c * 16 + q * 4 + e;
                const float tb = __uint_as_float(u[q * 4 + e]);
                cb[e] = tb * gn[e];                                                     // c_bar_l
                o[e] = tb * (flf[(j >> 1) * 4 + (j & 1)] * kWScale) * __cosf(ar[e]);    // g_bar_{l+1}
              }
              OI_CTA(kCtaG + l, quad) = make_float4(cb[0], cb[1], cb[2], cb[3]);
              OI_OP4(kSlabGB + l + 1, c * 16 + q * 4, o[0], o[1], o[2], o[3]);
            } else {
              // top: t_{D-1} = w_s c_{D-1}; then the backward of the forward sweep for layer D-1
              const float4 hb4 = OI_CTA(gslab, quad);   // h_bar_D
              const float hb[4] = {hb4.x, hb4.y, hb4.z, hb4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = c * 16 + q * 4 + e;
                const float tb = __uint_as_float(u[q * 4 + e]);
                const float gam = flf[(j >> 1) * 4 + (j & 1)] * kWScale;
                const float sn = __sinf(ar[e]), cs = __cosf(ar[e]);
                dwss[q * 4 + e] = fmaf(sdf_bar, sn, tb * gam * cs);   // d w_s = sdf_bar h_D + t_bar_{D-1} c_{D-1}
                const float cbar = tb * (sm.head[n0 + j].x * kInvWScale);
                const float ab = hb[e] * cs - cbar * gam * sn;
                dgs[q * 4 + e] = fmaf(ab, ar[e], gam * (cbar * cs));
                o[e] = ab * gam;   // u_bar_{D-1}
              }
              OI_OP4(kSlabUB + l, c * 16 + q * 4, o[0], o[1], o[2], o[3]);
            }
            split2_bf16(o[0], o[1], hi[2 * q], lo[2 * q]);
            split2_bf16(o[2], o[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(a_hi + c * 8, hi);
          tc::tmem_st8(a_lo + c * 8, lo);
          if (c == 3) OI_A_READY();
          if (l == D - 1) {
            colsum16(dwss, a.g.sigma_weight + n0 + c * 16, 1, lane);
            colsum16(dgs, dfilm + (size_t)l * 2 * kW + n0 + c * 16, 1, lane);   // V_{D-1}
          }
        }
      }
      // ---------------- backward of the forward sweep: h_bar_l = W_l^T u_bar_l, then layer k = l-1 ----------------
      for (int l = D - 1; l >= 1; --l) {
        const int k = l - 1;
        const float* flf = reinterpret_cast<const float*>(sm.film[t][k]) + n0 * 2;
        const float gsc = (k == 0) ? 1.0f : kWScale;
        OI_WAIT_ACC();
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
          float dgs[16], ubs[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int quad = Q0 + c * 4 + q;
            const float4 ar4 = OI_GS(kSlabArg + k, quad);
            const float ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
            const float4 cb4 = OI_CTA(kCtaG + k, quad);   // c_bar_k
            const float cb[4] = {cb4.x, cb4.y, cb4.z, cb4.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = c * 16 + q * 4 + e;
              const float gam = flf[(j >> 1) * 4 + (j & 1)] * gsc;
              const float sn = __sinf(ar[e]), cs = __cosf(ar[e]);
              const float ab = __uint_as_float(u[q * 4 + e]) * cs - cb[e] * gam * sn;
              dgs[q * 4 + e] = fmaf(ab, ar[e], gam * (cb[e] * cs));
              o[e] = ab * gam;   // u_bar_k
              ubs[q * 4 + e] = o[e];
            }
            if (k >= 1) {
              OI_OP4(kSlabUB + k, c * 16 + q * 4, o[0], o[1], o[2], o[3]);
              split2_bf16(o[0], o[1], hi[2 * q], lo[2 * q]);
              split2_bf16(o[2], o[3], hi[2 * q + 1], lo[2 * q + 1]);
            }
          }
          if (k >= 1) {
            tc::tmem_st8(a_hi + c * 8, hi);
            tc::tmem_st8(a_lo + c * 8, lo);
            if (c == 3) OI_A_READY();
          }
          const int nc = n0 + c * 16;
          colsum16(dgs, dfilm + (size_t)k * 2 * kW + nc, 1, lane);   // V_k = sum a_bar a + gamma c_bar cos a
          if (k == 0) {
            colsum16(ubs, dfilm + kW + nc, 1, lane);                  // d b_0 = sum u_bar_0
            float* dst = a.g.pts_weight[0] + (size_t)nc * 3;          // dW_0 += u_bar_0 (x) x
            float tmp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = px * ubs[i];
            colsum16(tmp, dst + 0, 3, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = py * ubs[i];
            colsum16(tmp, dst + 1, 3, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) tmp[i] = pz * ubs[i];
            colsum16(tmp, dst + 2, 3, lane);
          }
        }
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);   // film table / exchange buffer of this slot may be reused now
    }
#undef OI_CTA
#undef OI_GS
#undef OI_OP
#undef OI_OP4
#undef OI_A_READY
#undef OI_WAIT_ACC
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == kProducerWarp) tc::tmem_dealloc(tmem_base, 512);
}

// TC variant of the finalize step.  The sweep kernel reduces V = sum_m a_bar a + gamma c_bar cos a and the
// contraction kernel db = sum_m u_bar = gamma sum_m a_bar per instance; with a = gamma u + beta:
//   dL/dbeta = db / gamma,   dL/dgamma = sum_m a_bar u + c_bar cos a = (V - beta db / gamma) / gamma,
// db is summed over instances into the bias gradient; plus the variance.
__global__ void finalize_bwd_tc_kernel(int D, int n_inst, int R, const float* __restrict__ film,
                                       const float* __restrict__ d_film, const float* __restrict__ invs_partial,
                                       const float* __restrict__ blob, OiNetGrads g) {
  const int n = threadIdx.x;
  const int l = blockIdx.x;
  if (l <= D) {
    const int slot = (l < D) ? l : OI_MAX_DEPTH;
    float s = 0.f;
    for (int i = 0; i < n_inst; ++i) {
      const size_t o = ((size_t)i * kFilm + slot) * 2 * kW;
      const float db = d_film[o + kW + n], V = d_film[o + n];
      const float gam = film[o + n], bet = film[o + kW + n];
      const float dbeta = db / gam;
      s += db;
      g.film_gamma[((size_t)i * kFilm + slot) * kW + n] += (V - bet * dbeta) / gam;
      g.film_beta[((size_t)i * kFilm + slot) * kW + n] += dbeta;
    }
    float* dst = (l < D) ? g.pts_bias[l] : g.views_bias;
    dst[n] += s;
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
      const bool inside = inv_s > 1e-6f && inv_s < 1e6f;
      if (inside) g.variance[0] += (float)(tot * 10.0 * (double)inv_s);
    }
  }
}

}  // namespace

int render_bwd_tc_ctas(int n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_pairs = (n_tiles + 1) / 2;
  int ctas = sms < n_pairs ? sms : n_pairs;
  return ctas < 1 ? 1 : ctas;
}
size_t render_bwd_tc_scratch_floats() { return (size_t)2 * kCtaSlabs * kSlabFloats; }
size_t render_bwd_tc_slab_floats_per_tile() { return (size_t)kSlabsPerTile * kSlabFloats + 512; }

// Runs the two tensor-core kernels over tiles [0, n_tiles) in chunks of at most `chunk_tiles`.
int launch_render_bwd_tc(const OiRenderBwdDesc& d, const RenderKArgs& geo, const float* adj, const float* invs_partial,
                         float* d_film, float* scratch, float* slabs, float* aux, int chunk_tiles,
                         int n_ctas, cudaStream_t st) {
  const int n_inst = geo.n_inst, D = geo.D;
  OI_CHECK_CUDA(cudaFuncSetAttribute(bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(BwdTcSmem)));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  if (d.evt_core_start) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d.evt_core_start), st));
  for (int t0 = 0; t0 < geo.n_tiles; t0 += chunk_tiles) {
    const int t1 = (t0 + chunk_tiles < geo.n_tiles) ? t0 + chunk_tiles : geo.n_tiles;
    BwdTcArgs a;
    a.r = geo;
    a.adj = adj;
    a.scratch = scratch;
    a.scratch_stride = render_bwd_tc_scratch_floats();
    a.slabs = slabs;
    a.aux = aux;
    a.tile_begin = t0;
    a.tile_end = t1;
    a.g = d.grads;
    a.d_film = d_film;
    const int ctas = render_bwd_tc_ctas(t1 - t0) < n_ctas ? render_bwd_tc_ctas(t1 - t0) : n_ctas;
    bwd_tc_kernel<<<ctas, kTcThreads, sizeof(BwdTcSmem), st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());

    // ---- contraction over the points of this chunk
    WgArgs w;
    memset(&w, 0, sizeof(w));
    w.n_tiles = t1 - t0;
    w.tiles_per_inst = geo.tiles_per_inst;
    w.slabs_per_tile = kSlabsPerTile;
    w.slabs = slabs;
    w.aux = aux;
    w.tile0 = t0;
    float* dfilm0 = d_film;
    auto dbia = [&](int slot) { return dfilm0 + (size_t)slot * 2 * kW + kW; };
    const int inst_stride = kFilm * 2 * kW;
    int ng = 0;
    for (int l = 1; l < D; ++l) {
      WgGroup& g = w.groups[ng++];
      g.n_pairs = 2;
      g.pairs[0] = WgPair{kSlabUB + l, kSlabH + l};     // u_bar_l (x) h_l
      g.pairs[1] = WgPair{kSlabT + l, kSlabGB + l};     // t_l (x) g_bar_l
      g.use_aux = 1;
      g.aux_out[3] = dbia(l);                           // d b_l = sum u_bar_l  (aux column 3 = 1)
      g.aux_inst_stride[3] = inst_stride;
      g.aux_ch_stride[3] = 1;
      g.out = d.grads.pts_weight[l];
      g.out_ld = kW;
      g.weight = 4;
    }
    {  // colour layer: u_bar_c (x) h_D, u_bar_c (x) normal, sum u_bar_c
      WgGroup& g = w.groups[ng++];
      g.n_pairs = 1;
      g.pairs[0] = WgPair{kSlabUBC, kSlabH + D};
      g.use_aux = 1;
      for (int j = 0; j < 3; ++j) {
        g.aux_out[j] = d.grads.views_weight + kW + j;
        g.aux_inst_stride[j] = 0;
        g.aux_ch_stride[j] = kW + 3;
      }
      g.aux_out[3] = dbia(OI_MAX_DEPTH);
      g.aux_inst_stride[3] = inst_stride;
      g.aux_ch_stride[3] = 1;
      g.out = d.grads.views_weight;
      g.out_ld = kW + 3;
      g.weight = 2;
    }
    w.n_groups = ng;
    w.n_ctas = sms;
    int rc = launch_wgrad_tc(w, st);
    if (rc) return rc;
  }
  if (d.evt_core_stop) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d.evt_core_stop), st));

  finalize_bwd_tc_kernel<<<D + 2, kW, 0, st>>>(D, n_inst, d.n_rays, geo.film, d_film, invs_partial, geo.blob, d.grads);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
