// tcgen05 (5th-gen tensor core) core of the fused SDF render kernel (OI_IMPL_TCGEN05).
//
// Same mathematics as oi_render_ffma.cu, but every 128x128 FiLM-SIREN contraction runs on the tensor cores:
//   * a CTA (one per SM, persistent) owns TWO tiles of 128 sample points; row m of a tile lives in TMEM lane m;
//   * the fp32 activation is split into two fp16 terms (h = hi + lo, |err| <= 2^-22 |h|), the fp32 weight
//     likewise (pre-split, pre-scaled by 2^8 and pre-swizzled by oi_pack_weights), and one layer is three
//     chains of tcgen05.mma.kind::f16 with fp32 accumulation in TMEM:  hi*Whi + lo*Whi + hi*Wlo;
//   * the A operand (activations) is written straight into TMEM by the epilogue warps (tcgen05.st) and
//     consumed from TMEM (.ts form) -- activations never touch shared memory;
//   * B panels (64 KB per layer: {hi,lo} x 2 k-blocks, K-major SWIZZLE_128B images) are streamed from L2 by
//     TMA bulk copies through a 3-stage mbarrier ring and are shared by both tiles;
//   * per TMEM lane, two threads (64 channels each) do the FiLM epilogue in four 16-channel chunks: sincos, split,
//     tcgen05.st of the next operand, gamma*cos to the reverse-sweep scratch;
//   * a tile slot owns two 128-column TMEM buffers used as a ping-pong: the epilogue of stage s reads the
//     accumulator from buffer s&1 and writes the split operand of the next layer IN PLACE over the 16 columns it has
//     just consumed (8 packed hi + 8 packed lo); the MMAs of stage s read that operand and accumulate into the other
//     buffer.  The k-blocks of a chunk are issued as soon as the chunk is stored, so a slot's tensor work runs
//     behind its own epilogue (only the last 6 of the 24 MMAs of a layer are exposed) and the other slot fills the
//     rest.
// Warp roles: 0-7 epilogue of tile slot 0, 8-15 of slot 1 (112 registers after setmaxnreg), 16 TMA producer,
// 17 / 18 MMA issuer of slot 0 / 1, 19 idle (32 registers).
#include "oi_internal.cuh"
#include "oi_render_common.cuh"
#include "oi_tc.cuh"

namespace oi {

namespace {

constexpr int kTcThreads = 640;            // 5 warpgroups: 16 epilogue warps + {TMA producer, 2 MMA issuers, 1 idle warp}
constexpr int kEpiThreadsPerSlot = 256;     // 8 warps per tile slot
constexpr int kEpiWarpsPerSlot = 8;
constexpr int kProducerWarp = 16, kMmaWarp = 17;   // MMA issuer of slot t = warp kMmaWarp + t
constexpr int kTcStages = 2;
constexpr int kStageBytes = 2048;          // one warp's chunk of one layer: [4 channel quads][32 points] float4
constexpr int kPanelBytes = 65536;
constexpr int kSubPanelBytes = 16384;
constexpr float kWScale = 256.0f;        // weights are packed as 2^8 * W (see pack_weights_kernel)
constexpr float kInvWScale = 1.0f / 256.0f;
constexpr uint32_t kIdesc = tc::make_idesc_f16(128, 128);

struct __align__(1024) TcSmem {
  unsigned char w[kTcStages][kPanelBytes];  // weight panels (128 KB)
  float4 stg[16][2][128];                   // per epilogue warp: two 2 KB staging buffers of the scratch traffic (64 KB)
  float2 film[2][kFilm][kW];                // per tile: (gamma', delta)
  float4 w0[kW];                            // (W_0[n][0..2], 0)
  float4 head[kW];                          // 2^8 * (w_sigma[n], wc_grad[0..2][n])
  float4 rgbw[kW];                          // (W_rgb[0..2][n], 0)
  unsigned long long w_full[kTcStages], w_empty[kTcStages];
  float xch[2][128][8];                     // per slot / point: partial sums exchanged between the column halves
  unsigned long long acc_full[2], a_ready[2][4];   // a_ready[slot][chunk]: one arrival per epilogue warp
  unsigned long long ld_full[16][2];               // per epilogue warp / staging buffer: scratch re-read landed
  float comp[2][4][26];                            // in-kernel compositing: per slot / warp {product, 8 + 17 partial sums}
  double red[3][20];                               // last CTA: reduction of the per-ray partials of the global scalars
  int is_last;
  uint32_t tmem_base;
};
static_assert(sizeof(TcSmem) <= 227 * 1024, "TcSmem exceeds the 227 KB per-CTA limit");

// sin / cos of a FiLM pre-activation for the tensor-core epilogue.  OI_TC_SINCOS_REDUCE = 1: two-constant
// Cody-Waite reduction before the MUFU approximations (as in the FFMA core).  0 (default): feed the MUFU unit
// directly -- its input scaling x * (1/2pi) is an fp32 multiply, i.e. an absolute phase error of |x| * 6e-8 rad,
// the same order as the fp32 rounding of the pre-activation itself (ulp(32) = 3.8e-6), and it saves four
// instructions per element on a path that is bound by issue slots.
#ifndef OI_TC_SINCOS_REDUCE
#define OI_TC_SINCOS_REDUCE 0
#endif
__device__ __forceinline__ void sin_film(float x, float* s) {
#if OI_TC_SINCOS_REDUCE
  const float kInv2Pi = 0.15915494309189535f;
  const float k2PiHi = 6.2831854820251465f;
  const float k2PiLo = -1.7484555314695172e-07f;
  float t = fmaf(x, kInv2Pi, 12582912.0f);
  float k = t - 12582912.0f;
  float r = fmaf(k, -k2PiHi, x);
  r = fmaf(k, -k2PiLo, r);
  *s = __sinf(r);
#else
  *s = __sinf(x);
#endif
}
__device__ __forceinline__ void cos_film(float x, float* c) {
#if OI_TC_SINCOS_REDUCE
  float s;
  sincos_film(x, &s, c);
#else
  *c = __cosf(x);
#endif
}
__device__ __forceinline__ void sincos_tc(float x, float* s, float* c) {
#if OI_TC_SINCOS_REDUCE
  sincos_film(x, s, c);
#else
  *s = __sinf(x);
  *c = __cosf(x);
#endif
}

using tc::named_bar_sync;


// nanosleep back-off of the single-lane service warps between mbarrier probes (ns)
#ifndef OI_TC_SLEEP_PRODUCER
#define OI_TC_SLEEP_PRODUCER 400u
#endif
#ifndef OI_TC_CHUNK_UNROLL   // unroll factor of the four-chunk loops of the epilogue stages (code size vs I-cache)
#define OI_TC_CHUNK_UNROLL 4
#endif
constexpr int kChunkUnroll = OI_TC_CHUNK_UNROLL;
#ifndef OI_TC_SLEEP_MMA
#define OI_TC_SLEEP_MMA 100u
#endif

// 24 MMAs of one layer of one tile: acc = hi*Whi + lo*Whi + hi*Wlo over K = 128 (fp16 split operands).
__device__ __forceinline__ void issue_layer_mmas(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t wbase) {
  tc::issue_split_layer_mmas(acc, a_hi, a_lo, wbase, kIdesc);
}

// The 6 MMAs that consume epilogue chunk c of both column halves: k-blocks c (channels 16c..16c+15) and 4 + c
// (channels 64+16c..).  Operand of k-block k in the in-place layout: hi = abuf + 16k (8 packed columns), lo = +8.
__device__ __forceinline__ void issue_chunk_mmas(uint32_t acc, uint32_t abuf, uint32_t wbase, int c) {
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    const uint64_t bhi = tc::make_desc_k_sw128(wbase + kb * kSubPanelBytes + c * 32);
    const uint64_t blo = tc::make_desc_k_sw128(wbase + 2 * kSubPanelBytes + kb * kSubPanelBytes + c * 32);
    const uint32_t a_hi = abuf + 16 * (kb * 4 + c), a_lo = a_hi + 8;
    tc::mma_ts(acc, a_hi, bhi, kIdesc, (c > 0 || kb > 0) ? 1u : 0u);
    tc::mma_ts(acc, a_lo, bhi, kIdesc, 1u);
    tc::mma_ts(acc, a_hi, blo, kIdesc, 1u);
  }
}

// Contract B: Phong shading of one sample and per-ray compositing of the 12 maps of Generator.render_maps
// (generator.py:80-174, lighting.py:126-225) in the tail of a tile, same arithmetic as render_maps_kernel.  Called by
// the 128 tail threads of a tile slot (one per sample point); rays = aligned runs of S points.  Inlined into the kMaps
// instantiation only (`a` is the __grid_constant__ kernel parameter: every field is a constant-bank operand; behind
// a call it became a generic pointer and every store forced the next pointer to be re-loaded -- 30 k cycles per tile).
__device__ __forceinline__ void maps_tail(const RenderKArgs& a, float (*comp)[26], int bar_id, int inst, int ray, bool valid,
                                       float px, float py, float pz, float mid, float gx, float gy, float gz,
                                       float rgb0, float rgb1, float rgb2, float wgt, float ws, float col0, float col1,
                                       float col2, int lane, int wq) {
  const int S = a.S, seg = S < 32 ? S : 32, lis = lane & (seg - 1);
  const int wpr = S > 32 ? (S >> 5) : 1, wir = wq & (wpr - 1);
  const float rgbv[3] = {rgb0, rgb1, rgb2};
  // ---- contract B: Phong shading of this sample and compositing of the 12 maps of Generator.render_maps
  //      (generator.py:80-174, lighting.py:126-225), same arithmetic as render_maps_kernel
  const MapsKArgs& mp = a.maps;
  const float colf[3] = {col0, col1, col2};
  float lpar[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) lpar[i] = mp.light_params ? mp.light_params[i] : mp.lp[i];
  float Lx = mp.light_dir[inst * 3], Ly = mp.light_dir[inst * 3 + 1], Lz = mp.light_dir[inst * 3 + 2];
  {
    const float ln = fmaxf(sqrtf(Lx * Lx + Ly * Ly + Lz * Lz), 1e-6f);
    Lx /= ln;
    Ly /= ln;
    Lz /= ln;
  }
  const float nn = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-6f);
  const float ux = gx / nn, uy = gy / nn, uz = gz / nn;
  const float cosv = ux * Lx + uy * Ly + uz * Lz;
  const float ang = fmaxf(cosv, 0.f);
  float vx = a.rays_o[ray * 3] - px, vy = a.rays_o[ray * 3 + 1] - py,
        vz = a.rays_o[ray * 3 + 2] - pz;
  const float vn = fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-6f);
  vx /= vn;
  vy /= vn;
  vz /= vn;
  const float rx = -Lx + 2.0f * (cosv * ux), ry = -Ly + 2.0f * (cosv * uy), rz = -Lz + 2.0f * (cosv * uz);
  const float al = fmaxf(vx * rx + vy * ry + vz * rz, 0.f) * (cosv > 0.f ? 1.f : 0.f);
  const float spw = powf(al, lpar[9]) * wgt;
  float q[17];   // sh[3], ns[3], sp[3], df[3], n[3], z, zmin
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float diffc = lpar[3 + c] * ang, shade = lpar[c] + diffc;
    q[c] = shade * wgt;
    q[3 + c] = shade * rgbv[c] * wgt;
    q[6 + c] = lpar[6 + c] * spw;
    q[9 + c] = diffc * wgt;
  }
  q[12] = gx * wgt;
  q[13] = gy * wgt;
  q[14] = gz * wgt;
  q[15] = mid * wgt;
  q[16] = valid ? mid : 3.0e38f;
  for (int dd = seg >> 1; dd >= 1; dd >>= 1) {
#pragma unroll
    for (int i = 0; i < 17; ++i) {
      const float o = __shfl_xor_sync(0xffffffffu, q[i], dd);
      q[i] = (i == 16) ? fminf(q[i], o) : q[i] + o;
    }
  }
  if (S > 32) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 17; ++i) comp[wq][9 + i] = q[i];
    }
    named_bar_sync(bar_id, 128);
    if (wir == 0 && lane == 0) {
      for (int j = 1; j < wpr; ++j) {
#pragma unroll
        for (int i = 0; i < 17; ++i) {
          const float o = comp[wq + j][9 + i];
          q[i] = (i == 16) ? fminf(q[i], o) : q[i] + o;
        }
      }
    }
  }
  if (valid && lis == 0 && wir == 0) {
    const int PP = mp.rays_per_image, bimg = ray / PP, pix = ray - bimg * PP;
    auto put3 = [&](float* dst, float x0, float x1, float x2) {
      if (!dst) return;
      dst[((size_t)bimg * 3 + 0) * PP + pix] = x0;
      dst[((size_t)bimg * 3 + 1) * PP + pix] = x1;
      dst[((size_t)bimg * 3 + 2) * PP + pix] = x2;
    };
    auto put1 = [&](float* dst, float x) {
      if (dst) dst[(size_t)bimg * PP + pix] = x;
    };
    const float r0 = q[3] + q[6], r1 = q[4] + q[7], r2 = q[5] + q[8];
    const float* bg = mp.bg_color + bimg * 3;
    put3(mp.image, r0 + bg[0] * (1.0f - ws), r1 + bg[1] * (1.0f - ws), r2 + bg[2] * (1.0f - ws));
    put3(mp.image_no_bg, r0, r1, r2);
    put3(mp.shading_map, q[0], q[1], q[2]);
    put3(mp.color_map, colf[0], colf[1], colf[2]);
    put1(mp.weight_sum_map, ws);
    put1(mp.mask, fminf(fmaxf(ws, 1e-3f), 1.0f - 1e-3f));
    put3(mp.amb_shading_map, lpar[0] * ws, lpar[1] * ws, lpar[2] * ws);
    put3(mp.diff_shading_map, q[9], q[10], q[11]);
    put3(mp.normal_map, q[12], q[13], q[14]);
    put3(mp.no_specular_map, q[3], q[4], q[5]);
    put3(mp.specular_map, q[6], q[7], q[8]);
    put1(mp.z_map, q[15]);
    if (mp.z_min_per_ray) mp.z_min_per_ray[ray] = q[16];
  }
}

// Panel index (in the packed blob: fwd l=1..D-1 | colour-feature | reverse l=D-1..1) of MMA stage p of a tile.
// Fine pass order: forward layers, reverse layers, colour-feature layer LAST (its operand h_D is re-loaded from the
// scratch, its result is consumed straight from TMEM by the colour epilogue).
__device__ __forceinline__ int stage_panel(int p, int D, int coarse) {
  if (coarse || p < D - 1) return p;
  return (p < 2 * D - 2) ? p + 1 : D - 1;
}

// kMaps: the contract-B variant (shading maps composited in the tile tail); the plain variant carries none of its code.
template <bool kMaps>
__global__ void __launch_bounds__(kTcThreads, 1) render_tc_kernel(const __grid_constant__ RenderKArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const BlobLayout L = blob_layout(D);
  const float* cst = a.blob + L.const_off;
  const unsigned char* panels = reinterpret_cast<const unsigned char*>(a.blob + L.tc_off);
  const int NP = a.coarse ? (D - 1) : (2 * (D - 1) + 1);  // MMA stages per tile
  const int n_pairs = (a.n_tiles + 1) / 2;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&sm.w_full[s], 1);
      mbar_init(&sm.w_empty[s], 2);     // released by both MMA issuers
    }
    for (int t = 0; t < 2; ++t) {
      for (int c = 0; c < 4; ++c) mbar_init(&sm.a_ready[t][c], kEpiWarpsPerSlot);
      mbar_init(&sm.acc_full[t], 1);
    }
    for (int w = 0; w < 16; ++w)
      for (int b = 0; b < 2; ++b) mbar_init(&sm.ld_full[w][b], 1);
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

  // the service warpgroup hands its registers to the four epilogue warpgroups: 20 x 96 = 16 x 112 + 4 x 32
  if (warp >= 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp == kProducerWarp) {
    // ===================== TMA producer (weight panels) =====================
    if (lane == 0) {
      int it = 0;
      for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
        for (int p = 0; p < NP; ++p, ++it) {
          const int stage = it % kTcStages;
          if (it >= kTcStages) mbar_wait_backoff(&sm.w_empty[stage], ((it / kTcStages) - 1) & 1, OI_TC_SLEEP_PRODUCER);
          mbar_expect_tx(&sm.w_full[stage], kPanelBytes);
          const unsigned char* src = panels + (size_t)stage_panel(p, D, a.coarse) * kPanelBytes;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            tma_bulk_g2s(sm.w[stage] + q * kSubPanelBytes, src + q * kSubPanelBytes, kSubPanelBytes, &sm.w_full[stage]);
        }
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
    // ===================== MMA issuer of tile slot t =====================
    if (lane == 0) {
      const int t = warp - kMmaWarp;
      const uint32_t buf = tmem_base + t * 256;
      int it = 0;
      uint32_t ar_phase = 0u;
      for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
        const bool active = 2 * pi + t < a.n_tiles;
        for (int p = 0; p < NP; ++p, ++it) {
          const int stage = it % kTcStages;
          mbar_wait_backoff(&sm.w_full[stage], (it / kTcStages) & 1, OI_TC_SLEEP_MMA);
          if (!active) {   // odd tile count: the idle slot only releases the panel
            mbar_arrive(&sm.w_empty[stage]);
            continue;
          }
          const uint32_t wbase = smem_u32(sm.w[stage]);
          const uint32_t abuf = buf + (p & 1) * 128, acc = buf + ((p + 1) & 1) * 128;
#pragma unroll kChunkUnroll
          for (int c = 0; c < 4; ++c) {
            mbar_wait_backoff(&sm.a_ready[t][c], ar_phase, OI_TC_SLEEP_MMA);
            tc::fence_after_thread_sync();
            issue_chunk_mmas(acc, abuf, wbase, c);
          }
          ar_phase ^= 1u;
          tc::mma_commit(&sm.acc_full[t]);
          tc::mma_commit(&sm.w_empty[stage]);
        }
      }
    }
  } else if (warp < 16) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ===================== epilogue warps =====================
    // 16 warps: slot t = warp / 8 (tile of the pair), column half h = (warp / 4) % 2, TMEM lane quarter = warp % 4.
    // Thread (m, h) owns channels [64h, 64h+64) of sample point m of its tile: four 16-column chunks per layer.
    const int t = warp >> 3;
    const int h = (warp >> 2) & 1;
    const int m = (warp & 3) * 32 + lane;
    const int sub = tid & (kEpiThreadsPerSlot - 1);
    const int n0 = h * 64;
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    // my 64 columns of the slot's two ping-pong buffers: stage s reads its accumulator from buf[s & 1] and writes
    // the next operand in place (chunk c: 8 packed hi columns at +16c, 8 packed lo columns at +16c+8)
    const uint32_t buf0 = tmem_base + t * 256 + lane_field + n0;
    // Reverse-sweep scratch of this warp: per saved layer (slots 1..D-2: gamma' cos(arg_l); slot D-1: the split
    // h_D operand of the colour layer) four 2 KB chunks [quad q][lane] float4, written and re-read by this warp
    // only, through its two shared-memory staging buffers and TMA bulk copies (the LSU never sees this traffic).
    unsigned char* gscr = reinterpret_cast<unsigned char*>(a.scratch + (size_t)blockIdx.x * a.scratch_stride) +
                          (size_t)t * (D + 1) * 65536 + (size_t)(warp & 7) * kStageBytes;
    float4* stg = &sm.stg[warp][0][0];
    unsigned long long* ld_bar = &sm.ld_full[warp][0];
    const bool discard = (a.flags & 1) != 0 && lane < 16;
    const int n_loads = 4 * (D - 1);
    uint32_t af_phase = 0u;
    uint32_t ld_n = 0u;   // bulk loads consumed by this warp (buffer = ld_n & 1, parity = (ld_n >> 1) & 1)
    int film_inst = -1;
#ifdef OI_TC_PROFILE   // developer build: cycles per phase of one epilogue warp per slot (block 0), printed at exit
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#define OI_PROF(i)                 \
  do {                             \
    const long long now = clock64(); \
    prof[i] += now - prof_t;       \
    prof_t = now;                  \
  } while (0)
#else
#define OI_PROF(i)
#endif
#ifdef OI_TC_PROFILE2   // cycles per step of a reverse-stage chunk
    long long prof2[6] = {0, 0, 0, 0, 0, 0};
    long long prof2_t = clock64();
#define OI_PROF2(i)                  \
  do {                               \
    const long long now = clock64(); \
    prof2[i] += now - prof2_t;       \
    prof2_t = now;                   \
  } while (0)
#else
#define OI_PROF2(i)
#endif
#define OI_CHUNK_READY(c)                             \
  do {                                                \
    tc::wait_st();                                    \
    tc::fence_before_thread_sync();                   \
    __syncwarp();                                     \
    if (lane == 0) mbar_arrive(&sm.a_ready[t][(c)]);  \
  } while (0)
#define OI_WAIT_ACC()                                \
  do {                                               \
    mbar_wait_sleep(&sm.acc_full[t], af_phase);      \
    af_phase ^= 1u;                                  \
    tc::fence_after_thread_sync();                   \
  } while (0)
    // global address of chunk c of scratch slot `slot`
    auto gchunk = [&](int slot, int c) { return gscr + (size_t)slot * 65536 + (size_t)c * (8 * kStageBytes); };
    // k-th scratch re-read of a tile: slots D-2, D-3, .., 1 (reverse stages l = D-1 .. 2), then slot D-1 (h_D)
    auto gload = [&](int k) {
      const int kk = k >> 2;
      return gchunk(kk < D - 2 ? D - 2 - kk : D - 1, k & 3);
    };
    auto load_issue = [&](int k, uint32_t seq) {   // seq = ld_n the load will have when consumed
      if (lane == 0) {
        unsigned long long* bar = ld_bar + (seq & 1u);
        mbar_expect_tx(bar, kStageBytes);
        tma_bulk_g2s(stg + (seq & 1u) * 128, gload(k), kStageBytes, bar);
      }
    };
    auto load_wait = [&]() -> const float4* {
#ifdef OI_TC_PROFILE
      const long long t0 = clock64();
#endif
      mbar_wait(ld_bar + (ld_n & 1u), (ld_n >> 1) & 1u);
#ifdef OI_TC_PROFILE
      prof[4] += clock64() - t0;   // (the colour wait is reported together with the re-read waits)
#endif
      return stg + (ld_n & 1u) * 128 + lane;
    };
    // after the chunk has consumed re-read k (and a __syncwarp()): refill its buffer two re-reads ahead, drop the
    // dead lines from L2
    auto load_done = [&](int k) {
      if (k + 2 < n_loads) {
        load_issue(k + 2, ld_n + 2);
      }
      if (discard) l2_discard_128(gload(k) + lane * 128);
      ++ld_n;
    };

    for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
      const int tile = 2 * pi + t;
      if (tile >= a.n_tiles) continue;
      const int inst = tile / a.tiles_per_inst;
      const int tin = tile - inst * a.tiles_per_inst;
      if (inst != film_inst) {  // FiLM table of this tile's instance (consecutive tiles mostly share it)
        const float2* src = reinterpret_cast<const float2*>(a.film_tc) + (size_t)inst * kFilm * kW;
        float2* dst = &sm.film[t][0][0];
        for (int i = sub; i < kFilm * kW; i += kEpiThreadsPerSlot) dst[i] = src[i];
        film_inst = inst;
      }
      float px, py, pz;
      {
        const PointCtx pc = point_prologue(a, inst, tin, m, h == 0);
        px = pc.px;
        py = pc.py;
        pz = pc.pz;
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);
      OI_PROF(0);

      float sdf_acc = 0.f;
      // ---------------- layer 0 (K = 3) on the FMA pipe; its gamma' cos is recomputed in the last reverse stage ----
      {
        const float4* fl = reinterpret_cast<const float4*>(sm.film[t][0]) + n0 / 2;   // (g0, g1, d0, d1) per pair
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s[4];
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int j = c * 16 + q * 4 + e;
              const float4 w0 = sm.w0[n0 + j], w1 = sm.w0[n0 + j + 1];
              const float4 f = fl[j >> 1];
              const float2 u = make_float2(fmaf(w0.z, pz, fmaf(w0.y, py, w0.x * px)),
                                           fmaf(w1.z, pz, fmaf(w1.y, py, w1.x * px)));
              const float2 arg = tc::fma2(make_float2(f.x, f.y), u, make_float2(f.z, f.w));
              sin_film(arg.x, &s[e]);
              sin_film(arg.y, &s[e + 1]);
            }
            tc::split2(s[0], s[1], hi[2 * q], lo[2 * q]);
            tc::split2(s[2], s[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(buf0 + c * 16, hi);
          tc::tmem_st8(buf0 + c * 16 + 8, lo);
          OI_CHUNK_READY(c);
        }
      }
      OI_PROF(1);
      // ---------------- forward layers 1..D-2: accumulator chunk c+1 is in flight while chunk c is processed ----
      const int l_last = D - 1;
      for (int l = 1; l < l_last; ++l) {
        const float4* fl = reinterpret_cast<const float4*>(sm.film[t][l]) + n0 / 2;
        const uint32_t acc = buf0 + (l & 1) * 128;
        OI_WAIT_ACC();
        OI_PROF(2);
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll kChunkUnroll
        for (int c = 0; c < 4; ++c) {
          float4* sb = reinterpret_cast<float4*>(gchunk(l, c)) + lane;   // my 16 bytes of each of the chunk's 4 quads
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s[4], cv[4];
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int j = c * 16 + q * 4 + e;
              const float4 f = fl[j >> 1];
              const float2 arg = tc::fma2(make_float2(f.x, f.y),
                                          make_float2(__uint_as_float(u[q * 4 + e]), __uint_as_float(u[q * 4 + e + 1])),
                                          make_float2(f.z, f.w));
              float c0, c1;
              sincos_tc(arg.x, &s[e], &c0);
              sincos_tc(arg.y, &s[e + 1], &c1);
              const float2 cvp = tc::mul2(make_float2(f.x, f.y), make_float2(c0, c1));
              cv[e] = cvp.x;
              cv[e + 1] = cvp.y;
            }
            if (!a.coarse) sb[q * 32] = make_float4(cv[0], cv[1], cv[2], cv[3]);
            tc::split2(s[0], s[1], hi[2 * q], lo[2 * q]);
            tc::split2(s[2], s[3], hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(acc + c * 16, hi);
          tc::tmem_st8(acc + c * 16 + 8, lo);
          OI_CHUNK_READY(c);
        }
        OI_PROF(3);
      }
      // ---------------- last forward layer D-1: sdf head; the operand written in place is the START of the reverse
      //                  sweep, t_{D-1} = w_sigma * gamma cos(arg_{D-1}); h_D (split) goes to scratch slot D-1 -------
      if (l_last >= 1) {
        const int l = l_last;
        const float4* fl = reinterpret_cast<const float4*>(sm.film[t][l]) + n0 / 2;
        const uint32_t acc = buf0 + (l & 1) * 128;
        OI_WAIT_ACC();
        OI_PROF(2);
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll kChunkUnroll
        for (int c = 0; c < 4; ++c) {
          uint4* sb = reinterpret_cast<uint4*>(gchunk(D - 1, c)) + lane;
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s[4], tv[4];
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int j = c * 16 + q * 4 + e;
              const float4 f = fl[j >> 1];
              const float2 arg = tc::fma2(make_float2(f.x, f.y),
                                          make_float2(__uint_as_float(u[q * 4 + e]), __uint_as_float(u[q * 4 + e + 1])),
                                          make_float2(f.z, f.w));
              const float ws0 = sm.head[n0 + j].x, ws1 = sm.head[n0 + j + 1].x;
              if (a.coarse) {
                sin_film(arg.x, &s[e]);
                sin_film(arg.y, &s[e + 1]);
              } else {
                float c0, c1;
                sincos_tc(arg.x, &s[e], &c0);
                sincos_tc(arg.y, &s[e + 1], &c1);
                const float2 tp = tc::mul2(tc::mul2(make_float2(f.x, f.y), make_float2(ws0, ws1)), make_float2(c0, c1));
                tv[e] = tp.x;
                tv[e + 1] = tp.y;
              }
              sdf_acc = fmaf(ws0, s[e], sdf_acc);
              sdf_acc = fmaf(ws1, s[e + 1], sdf_acc);
            }
            if (!a.coarse) {
              tc::split2(tv[0], tv[1], hi[2 * q], lo[2 * q]);
              tc::split2(tv[2], tv[3], hi[2 * q + 1], lo[2 * q + 1]);
              uint4 hs;   // split h_D of this channel quad: (hi01, hi23, lo01, lo23)
              tc::split2(s[0], s[1], hs.x, hs.z);
              tc::split2(s[2], s[3], hs.y, hs.w);
              sb[q * 32] = hs;
            }
          }
          if (!a.coarse) {
            tc::tmem_st8(acc + c * 16, hi);
            tc::tmem_st8(acc + c * 16 + 8, lo);
            OI_CHUNK_READY(c);
          }
        }
        OI_PROF(3);
      }
      float* xch = &sm.xch[t][m][0];
      if (a.coarse) {
        // sdf = sum of the two column halves
        if (h == 1) xch[0] = sdf_acc;
        named_bar_sync(1 + t, kEpiThreadsPerSlot);
        if (h == 0) {
          const PointCtx pc = point_prologue(a, inst, tin, m, false);
          if (pc.valid)
            a.sdf_coarse[(size_t)pc.ray * a.S + pc.si] = (sdf_acc + xch[0]) * kInvWScale + cst[BlobLayout::kScalars + 0];
        }
        named_bar_sync(1 + t, kEpiThreadsPerSlot);
        continue;
      }
      // ---------------- the staging buffers change direction: every store has left shared memory, the saved
      //                  gamma' cos of layers <= D-2 have landed (the four h_D stores may still be in flight) ---------
      fence_proxy_async_global();
      __syncwarp();
      load_issue(0, ld_n);
      load_issue(1, ld_n + 1);
      OI_PROF(5);
      // ---------------- reverse sweep l = D-1 .. 2: t_{l-1} = g_l * gamma' cos(arg_{l-1}) ----------------
      int k = 0;   // scratch re-reads consumed in this tile
      for (int l = D - 1; l >= 2; --l) {
        const uint32_t acc = buf0 + ((l + 1) & 1) * 128;   // stage 2D - 1 - l
        OI_WAIT_ACC();
        OI_PROF(6);
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll kChunkUnroll
        for (int c = 0; c < 4; ++c, ++k) {
          OI_PROF2(0);
          const float4* lb = load_wait();
          float4 csc[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) csc[q] = lb[q * 32];
          OI_PROF2(1);
          tc::wait_ld();
          OI_PROF2(2);
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 t01 = tc::mul2(make_float2(__uint_as_float(u[q * 4]), __uint_as_float(u[q * 4 + 1])),
                                        make_float2(csc[q].x, csc[q].y));
            const float2 t23 = tc::mul2(make_float2(__uint_as_float(u[q * 4 + 2]), __uint_as_float(u[q * 4 + 3])),
                                        make_float2(csc[q].z, csc[q].w));
            tc::split2(t01.x, t01.y, hi[2 * q], lo[2 * q]);
            tc::split2(t23.x, t23.y, hi[2 * q + 1], lo[2 * q + 1]);
          }
          tc::tmem_st8(acc + c * 16, hi);
          tc::tmem_st8(acc + c * 16 + 8, lo);
          OI_PROF2(3);
          OI_CHUNK_READY(c);
          OI_PROF2(4);
          load_done(k);
          OI_PROF2(5);
        }
        OI_PROF(7);
      }
      // ---------------- last reverse stage l = 1: grad_x sdf = W_0^T (g_1 * gamma' cos(arg_0)), layer 0 recomputed;
      //                  the operand written in place is h_D (from scratch) for the colour-feature MMA ----------------
      float gx = 0.f, gy = 0.f, gz = 0.f;
      {
        const float4* fl = reinterpret_cast<const float4*>(sm.film[t][0]) + n0 / 2;
        const uint32_t acc = buf0;   // stage 2D - 2
        OI_WAIT_ACC();
        OI_PROF(6);
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll kChunkUnroll
        for (int c = 0; c < 4; ++c, ++k) {
          const uint4* lb = reinterpret_cast<const uint4*>(load_wait());
          uint4 hs[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) hs[q] = lb[q * 32];
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int j = c * 16 + q * 4 + e;
              const float4 w0 = sm.w0[n0 + j], w1 = sm.w0[n0 + j + 1];
              const float4 f = fl[j >> 1];
              const float2 uu = make_float2(fmaf(w0.z, pz, fmaf(w0.y, py, w0.x * px)),
                                            fmaf(w1.z, pz, fmaf(w1.y, py, w1.x * px)));
              const float2 arg = tc::fma2(make_float2(f.x, f.y), uu, make_float2(f.z, f.w));
              float c0, c1;
              cos_film(arg.x, &c0);
              cos_film(arg.y, &c1);
              const float2 cv = tc::mul2(make_float2(f.x * kInvWScale, f.y * kInvWScale), make_float2(c0, c1));
              const float2 tv = tc::mul2(make_float2(__uint_as_float(u[q * 4 + e]), __uint_as_float(u[q * 4 + e + 1])), cv);
              gx = fmaf(w0.x, tv.x, gx);
              gy = fmaf(w0.y, tv.x, gy);
              gz = fmaf(w0.z, tv.x, gz);
              gx = fmaf(w1.x, tv.y, gx);
              gy = fmaf(w1.y, tv.y, gy);
              gz = fmaf(w1.z, tv.y, gz);
            }
          }
          const uint32_t hi[8] = {hs[0].x, hs[0].y, hs[1].x, hs[1].y, hs[2].x, hs[2].y, hs[3].x, hs[3].y};
          const uint32_t lo[8] = {hs[0].z, hs[0].w, hs[1].z, hs[1].w, hs[2].z, hs[2].w, hs[3].z, hs[3].w};
          tc::tmem_st8(acc + c * 16, hi);
          tc::tmem_st8(acc + c * 16 + 8, lo);
          OI_CHUNK_READY(c);
          load_done(k);
        }
        OI_PROF(7);
      }
      // ---------------- combine the two column halves: sdf and grad_x sdf ----------------
      xch[h * 4 + 0] = sdf_acc;
      xch[h * 4 + 1] = gx;
      xch[h * 4 + 2] = gy;
      xch[h * 4 + 3] = gz;
      named_bar_sync(1 + t, kEpiThreadsPerSlot);
      {
        const int o = (h ^ 1) * 4;
        sdf_acc += xch[o + 0];
        gx += xch[o + 1];
        gy += xch[o + 2];
        gz += xch[o + 3];
      }
      const float sdf = sdf_acc * kInvWScale + cst[BlobLayout::kScalars + 0];
      // ---------------- colour layer: feature part 2^8 W_c[:, :128] h_D straight from the accumulator, normal part
      //                  in registers; rgb head (this thread's 64 channels) ----------------
      float rgb[3] = {0.f, 0.f, 0.f};
      {
        const float* flc = reinterpret_cast<const float*>(sm.film[t][OI_MAX_DEPTH]) + n0 * 2;   // pair layout
        const uint32_t acc = buf0 + 128;   // stage 2D - 1
        OI_WAIT_ACC();
        OI_PROF(4);
        uint32_t ub[2][16];
        tc::tmem_ld16_async(acc, ub[0]);
#pragma unroll kChunkUnroll
        for (int c = 0; c < 4; ++c) {
          tc::wait_ld();
          if (c < 3) tc::tmem_ld16_async(acc + (c + 1) * 16, ub[(c + 1) & 1]);
          const uint32_t(&u)[16] = ub[c & 1];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int j = c * 16 + i;
            const float4 hd = sm.head[n0 + j];
            const float2 f = make_float2(flc[(j >> 1) * 4 + (j & 1)], flc[(j >> 1) * 4 + 2 + (j & 1)]);
            float pre = fmaf(hd.y, gx, __uint_as_float(u[i]));
            pre = fmaf(hd.z, gy, pre);
            pre = fmaf(hd.w, gz, pre);
            float s;
            sin_film(fmaf(f.x, pre, f.y), &s);
            const float4 rw = sm.rgbw[n0 + j];
            rgb[0] = fmaf(rw.x, s, rgb[0]);
            rgb[1] = fmaf(rw.y, s, rgb[1]);
            rgb[2] = fmaf(rw.z, s, rgb[2]);
          }
        }
        tc::fence_before_thread_sync();   // orders these TMEM reads before the next tile's operand stores / MMAs
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);   // everybody has read the first exchange
      if (h == 1) {
        xch[0] = rgb[0];
        xch[1] = rgb[1];
        xch[2] = rgb[2];
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);
      if (h == 0) {
        rgb[0] += xch[0];
        rgb[1] += xch[1];
        rgb[2] += xch[2];
        const PointCtx pc = point_prologue(a, inst, tin, m, false);
        if (!a.fuse_composite) {
          point_tail(a, pc, cst, sdf, gx, gy, gz, rgb);
        } else {
          // ---- per-ray compositing (renderer.py:300-338), rays = aligned runs of S of this tile's 128 points:
          //      weights = alpha * exclusive cumprod(1 - alpha + 1e-7), weight sum / max, colour, partial sums of the
          //      two global scalars.  Same association as composite_kernel (carry * exclusive scan inside a warp).
          float to[4] = {0.f, 0.f, 0.f, 0.f};
          point_tail(a, pc, cst, sdf, gx, gy, gz, rgb, to);
          const int S = a.S, seg = S < 32 ? S : 32, lis = lane & (seg - 1), wq = warp & 3;
          const float alpha = pc.valid ? to[0] : 0.f;
          float incl = pc.valid ? (1.0f - alpha + 1e-7f) : 1.0f;
          for (int dd = 1; dd < seg; dd <<= 1) {
            const float o = __shfl_up_sync(0xffffffffu, incl, dd);
            if (lis >= dd) incl *= o;
          }
          float excl = __shfl_up_sync(0xffffffffu, incl, 1);
          if (lis == 0) excl = 1.0f;
          float carry = 1.0f;
          const int wpr = S > 32 ? (S >> 5) : 1, wir = wq & (wpr - 1);   // warps per ray, my warp's index in its ray
          if (S > 32) {
            if (lane == 31) sm.comp[t][wq][0] = incl;
            named_bar_sync(3 + t, 128);
            for (int j = 0; j < wir; ++j) carry *= sm.comp[t][wq - wir + j][0];
          }
          const float wgt = alpha * (carry * excl);
          float v[8];
          {
            const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
            const float relax = sqrtf(pc.px * pc.px + pc.py * pc.py + pc.pz * pc.pz) < 1.2f ? 1.f : 0.f;
            const bool ok = pc.valid;
            v[0] = wgt;
            v[1] = ok ? wgt : -1e30f;
            v[2] = to[1] * wgt;
            v[3] = to[2] * wgt;
            v[4] = to[3] * wgt;
            v[5] = ok ? relax * e * e : 0.f;
            v[6] = ok ? relax : 0.f;
            v[7] = ok ? expf(-100.0f * fabsf(sdf)) : 0.f;
          }
          for (int dd = seg >> 1; dd >= 1; dd >>= 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float o = __shfl_xor_sync(0xffffffffu, v[i], dd);
              v[i] = (i == 1) ? fmaxf(v[i], o) : v[i] + o;
            }
          }
          if (S > 32) {
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) sm.comp[t][wq][1 + i] = v[i];
            }
            named_bar_sync(3 + t, 128);
            if (wir == 0 && lane == 0) {
              for (int j = 1; j < wpr; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float o = sm.comp[t][wq + j][1 + i];
                  v[i] = (i == 1) ? fmaxf(v[i], o) : v[i] + o;
                }
              }
            }
          }
          if (pc.valid) {
            if (a.alpha) a.alpha[(size_t)pc.ray * S + pc.si] = wgt;   // the `weights` output
            if (lis == 0 && wir == 0) {                  // first point of the ray: per-ray outputs
              const int ray = pc.ray;
              if (a.weight_sum) a.weight_sum[ray] = v[0];
              if (a.weight_max) a.weight_max[ray] = v[1];
              if (a.color_fine) {
                a.color_fine[ray * 3 + 0] = v[2];
                a.color_fine[ray * 3 + 1] = v[3];
                a.color_fine[ray * 3 + 2] = v[4];
              }
              if (a.s_val) a.s_val[ray] = cst[BlobLayout::kScalars + 5];
              a.partials[(size_t)ray * 3 + 0] = v[5];
              a.partials[(size_t)ray * 3 + 1] = v[6];
              a.partials[(size_t)ray * 3 + 2] = v[7];
            }
          }
          if (kMaps)
            maps_tail(a, sm.comp[t], 3 + t, inst, pc.ray, pc.valid, pc.px, pc.py, pc.pz, pc.mid, gx, gy, gz, to[1], to[2],
                      to[3], wgt, v[0], v[2], v[3], v[4], lane, warp & 3);
        }
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);  // film table / exchange buffer of this slot may be reused now
      OI_PROF(8);
    }
#ifdef OI_TC_PROFILE
    if (blockIdx.x == 0 && lane == 0 && (warp & 7) == 0)
      printf("tcprof slot %d: setup %lld l0 %lld | fwd wait %lld work %lld | colour wait %lld dir-switch %lld | rev wait "
             "%lld work %lld | colour+tail %lld | store-buffer wait %lld (inside fwd work), re-read wait %lld (inside rev work)\n", t, prof[0], prof[1], prof[2], prof[3], 0ll, prof[5], prof[6],
             prof[7], prof[8], prof[9], prof[4]);
#endif
#ifdef OI_TC_PROFILE2
    if (blockIdx.x == 0 && lane == 0 && (warp & 7) == 0)
      printf("tcprof2 slot %d reverse chunks: other %lld | re-read wait+LDS %lld | wait_ld %lld | mul/split/STTM issue %lld | "
             "wait_st+fence+syncwarp+arrive %lld | refill+discard %lld\n", t, prof2[0], prof2[1], prof2[2], prof2[3],
             prof2[4], prof2[5]);
#endif
#undef OI_CHUNK_READY
#undef OI_WAIT_ACC
#undef OI_PROF
#undef OI_PROF2
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == kProducerWarp) tc::tmem_dealloc(tmem_base, 512);
  if (!a.fuse_composite) return;
  // ---- the last CTA to finish reduces the per-ray partials of gradient_error (renderer.py:309-311) and surface_loss
  //      (:338) in a fixed order (deterministic), as composite_kernel does for the unfused path
  __threadfence();
  __syncthreads();
  if (tid == 0) sm.is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!sm.is_last) return;
  __threadfence();
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int r = tid; r < a.R; r += kTcThreads) {
    s0 += (double)__ldcg(a.partials + (size_t)r * 3 + 0);
    s1 += (double)__ldcg(a.partials + (size_t)r * 3 + 1);
    s2 += (double)__ldcg(a.partials + (size_t)r * 3 + 2);
  }
#pragma unroll
  for (int dd = 16; dd >= 1; dd >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, dd);
    s1 += __shfl_xor_sync(0xffffffffu, s1, dd);
    s2 += __shfl_xor_sync(0xffffffffu, s2, dd);
  }
  if (lane == 0) {
    sm.red[0][warp] = s0;
    sm.red[1][warp] = s1;
    sm.red[2][warp] = s2;
  }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < kTcThreads / 32; ++i) {
      t0 += sm.red[0][i];
      t1 += sm.red[1][i];
      t2 += sm.red[2][i];
    }
    if (a.gradient_error) *a.gradient_error = (float)(t0 / (t1 + 1e-5));
    if (a.surface_loss) *a.surface_loss = (float)(t2 / ((double)a.R * (double)a.S));
    *a.ticket = 0;  // self-reset so that the workspace can be reused without a memset
  }
}

// ------------------------------------------------------------------------------------------------
// Self-test of the UMMA building blocks: D[128,128] = A[128,128] * B[128,128]^T through the same
// split / TMEM-operand / SWIZZLE_128B-panel path as the render kernel.  panel != NULL: B comes from a
// packed 64 KB panel image by TMA (checks oi_pack_weights); panel == NULL: B is split and swizzled in-kernel.
// ------------------------------------------------------------------------------------------------
struct __align__(1024) SelfTestSmem {
  unsigned char w[kPanelBytes];
  unsigned long long w_full, acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             const unsigned char* __restrict__ panel,
                                                             float* __restrict__ Dout) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SelfTestSmem& sm = *reinterpret_cast<SelfTestSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&sm.w_full, 1);
    mbar_init(&sm.acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tc::tmem_alloc(&sm.tmem_base, 256);
    tc::tmem_relinquish();
  }
  if (panel == nullptr) {
    __half* wh = reinterpret_cast<__half*>(sm.w);
    for (int i = tid; i < kW * kW; i += blockDim.x) {
      const int n = i / kW, k = i % kW;
      const float v = B[n * kW + k];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const int kb = k >> 6, kk = k & 63;
      const int idx = n * 64 + (((kk >> 3) ^ (n & 7)) << 3) + (kk & 7);
      wh[kb * 8192 + idx] = hi;
      wh[2 * 8192 + kb * 8192 + idx] = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_base;
  if (panel != nullptr && tid == 128) {
    mbar_expect_tx(&sm.w_full, kPanelBytes);
    for (int q = 0; q < 4; ++q)
      tma_bulk_g2s(sm.w + q * kSubPanelBytes, panel + q * kSubPanelBytes, kSubPanelBytes, &sm.w_full);
  }
  if (warp < 4) {
    const int m = tid;
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    const uint32_t acc = tmem_base + lane_field;
    for (int c = 0; c < 4; ++c) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        tc::split2(A[m * kW + c * 32 + j], A[m * kW + c * 32 + j + 1], hi[j >> 1], lo[j >> 1]);
      tc::tmem_st16(acc + 128 + c * 16, hi);
      tc::tmem_st16(acc + 192 + c * 16, lo);
    }
    tc::wait_st();
    tc::fence_before_thread_sync();
  }
  __syncthreads();
  if (tid == 128) {
    tc::fence_after_thread_sync();
    if (panel != nullptr) mbar_wait(&sm.w_full, 0);
    issue_layer_mmas(tmem_base, tmem_base + 128, tmem_base + 192, smem_u32(sm.w));
    tc::mma_commit(&sm.acc_full);
  }
  if (warp < 4) {
    const int m = tid;
    const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16);
    mbar_wait(&sm.acc_full, 0);
    tc::fence_after_thread_sync();
    for (int c = 0; c < 4; ++c) {
      float u[32];
      tc::tmem_ld32(acc + c * 32, u);
#pragma unroll
      for (int j = 0; j < 32; ++j) Dout[m * kW + c * 32 + j] = u[j];
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace

size_t render_tc_scratch_floats(int depth, int* n_ctas, int n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_pairs = (n_tiles + 1) / 2;
  int ctas = sms < n_pairs ? sms : n_pairs;
  if (ctas < 1) ctas = 1;
  if (n_ctas) *n_ctas = ctas;
  return (size_t)2 * (depth + 1) * kW * 128;
}

int launch_render_tc(const RenderKArgs& a, cudaStream_t st) {
  if (a.D < 2) return set_error(OI_ERR_UNSUPPORTED, "the tcgen05 core needs depth >= 2 (use OI_IMPL_FFMA)");
  int n_ctas = 0;
  render_tc_scratch_floats(a.D, &n_ctas, a.n_tiles);
  if (a.maps.enabled) {
    OI_CHECK_CUDA(cudaFuncSetAttribute(render_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(TcSmem)));
    render_tc_kernel<true><<<n_ctas, kTcThreads, sizeof(TcSmem), st>>>(a);
  } else {
    OI_CHECK_CUDA(cudaFuncSetAttribute(render_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(TcSmem)));
    render_tc_kernel<false><<<n_ctas, kTcThreads, sizeof(TcSmem), st>>>(a);
  }
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

int launch_tc_selftest(const float* A, const float* B, const void* panel, float* D, cudaStream_t st) {
  OI_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(SelfTestSmem)));
  tc_selftest_kernel<<<1, 160, sizeof(SelfTestSmem), st>>>(A, B, static_cast<const unsigned char*>(panel), D);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
