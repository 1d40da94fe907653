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
// Contractions over POINTS (weight gradients, bias sums) are not done here: every per-point operand they need is left
// as a slab per tile (oi_wgrad.cuh) and contracted by wgrad_tc_kernel.
//
// What the sweep leaves in memory is 16 bits wide (it was HBM-bound on fp32 / TF32 slabs):
//   * contraction operands as fp16 in the UMMA image, with exact power-of-two scales: the point's upstream adjoints are
//     multiplied by 2^-e_m when the tile starts (so every adjoint of the sweep carries the scale), forward-type
//     operands by 2^(e_m - e_ref); template parameter F16, selected per call on the device (bwd_mode, oi_wgrad.cuh),
//     F16 = false keeps the TF32 slabs;
//   * pre-activations as 16-bit phases (a mod 2 pi; later stages only take sin / cos of them), OI_BWD_ARG16;
//   * g_l / c_bar_l as fp16 with F16, OI_BWD_G16.
//
// Round-2 structure of the epilogues:
//   * every stage walks its 64 channels in eight OCTS of 8 columns: tcgen05.ld.x8 (double-buffered), the two
//     scratch words it re-reads (pre-activations a_l, g_{l+1} / c_bar_l) loaded one oct ahead -- the first before the
//     wait on the accumulator barrier, after an L2 prefetch of the whole stage's re-reads -- and no per-stage arrays
//     that outlive an oct; the oct loop is unrolled in groups of 4 (fully unrolled: instruction-fetch stalls);
//   * dL/dgamma needs NO per-point column sums: with a = gamma u + beta, u = W h + b,
//         sum_m a_bar u + c_bar cos a = (1/gamma) [ sum_k W[j][k] dW[j][k] + b[j] db[j] ]
//     (dW = both parts of the layer's weight gradient, per instance; db = sum_m u_bar), because
//     sum_m a_bar_j h_k = dW^fwd_jk / gamma_j and sum_m c_bar_j cos a_j = sum_m t_bar_j t_j / gamma_j =
//     sum_k W_jk dW^rev_jk / gamma_j.  finalize_bwd_tc_kernel evaluates it from the per-instance weight gradients;
//     tests/test_backward_math.py checks the identity in fp64.  That removes 10 of the 21 warp butterflies per
//     16-channel chunk and every "a_bar * a" product from the sweep;
//   * 640 threads = 5 warpgroups: the producer / issuer warpgroup drops to 32 registers (setmaxnreg), the four
//     epilogue warpgroups grow to 112.
// Warp roles: 0-7 / 8-15 epilogue of slot 0 / 1, 16 TMA producer, 17 MMA issuer, 18-19 idle.
#include "oi_internal.cuh"
#include "oi_render_common.cuh"
#include "oi_tc.cuh"
#include "oi_wgrad.cuh"

#ifndef OI_BWD_PF2
#define OI_BWD_PF2 1   // octs of look-ahead of the scratch re-reads in stages that re-read one slab
#endif
#ifndef OI_BWD_L2PF
#define OI_BWD_L2PF 1   // prefetch a stage's scratch re-reads into L2 at the start of the stage
#endif
#ifndef OI_BWD_PFOCT
#define OI_BWD_PFOCT 3  // with OI_BWD_L2PF == 2: the NEXT stage's re-reads are prefetched after this oct of a stage
#endif
#ifndef OI_BWD_STHINT
#define OI_BWD_STHINT 0  // experiment: L2 eviction hints on the slab stores (1: operand slabs evict-first; 2: + per-CTA
#endif                   // slabs evict-last; 3: + ARG slabs evict-last)
#ifndef OI_BWD_ARG16
#define OI_BWD_ARG16 1   // pre-activations kept for later stages as 16-bit phases (a mod 2 pi in 2^-16 turns), not fp32
#endif
#ifndef OI_BWD_G16
#define OI_BWD_G16 1     // with fp16 operand slabs: the per-CTA g / c_bar slabs as fp16 too
#endif
#ifndef OI_BWD_EPI_REGS
#define OI_BWD_EPI_REGS 112   // registers of the epilogue warps after setmaxnreg (16 x this + 4 x 32 <= 2048)
#endif
#ifndef OI_BWD_RANGE_STATS
#define OI_BWD_RANGE_STATS 0   // diagnostic build: max |fp16 operand| (forward- / adjoint-type) into ctl[6] / ctl[7]
#endif
#ifndef OI_BWD_OCT_UNROLL
#define OI_BWD_OCT_UNROLL 4   // octs per unrolled group of a stage's oct loop (2, 4 or 8)
#endif
#ifndef OI_BWD_PF4
#define OI_BWD_PF4 1   // ... in stages that re-read two slabs
#endif

namespace oi {

namespace {

constexpr int kTcThreads = 640;
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
  OiNetGrads g;            // per-channel sums over points of the narrow heads (warp butterflies + atomics)
  float* d_film;           // [n_inst][9][2][128] (unused | db)
  float* dw0;              // [n_inst][128][3] per-instance dW_0 (both parts)
  const unsigned int* ctl; // control block of the call (oi_wgrad.cuh: bwd_mode)
  unsigned int* sticky;    // fp16 overflow guard word of the workspace (oi_wgrad.cuh)
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

using tc::named_bar_sync;

// The 24 MMAs of one layer of one tile (acc = hi*Whi + lo*Whi + hi*Wlo over K = 128), written for the 32-register
// issuer thread: two 64-bit descriptors advanced by immediates instead of 16 hoisted ones.
__device__ __forceinline__ void issue_layer_lean(uint32_t acc, uint32_t wbase, uint32_t idesc) {
  uint64_t dhi = tc::make_desc_k_sw128(wbase);
  const uint32_t a_hi = acc + 128, a_lo = acc + 192;
#pragma unroll 1
  for (int kb = 0; kb < 2; ++kb) {
    const uint64_t dlo = dhi + (2 * 16384 >> 4);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t k8 = (uint32_t)(kb * 4 + ks) * 8;
      tc::mma_ts(acc, a_hi + k8, dhi + ks * 2, idesc, (kb | ks) ? 1u : 0u);
      tc::mma_ts(acc, a_lo + k8, dhi + ks * 2, idesc, 1u);
      tc::mma_ts(acc, a_hi + k8, dlo + ks * 2, idesc, 1u);
    }
    dhi += 16384 >> 4;
  }
}

// ---- TMEM accessors at oct granularity (8 accumulator columns / 4 packed operand columns) ----
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}

// fp32 pair -> bf16 {hi, lo}: hi = truncation to the upper 16 bits (one PRMT packs two of them), lo = rn(v - hi);
// v = hi + lo + O(2^-16 |v|), full fp32 exponent range.  5 instructions per pair.
__device__ __forceinline__ void split2_bf16t(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const uint32_t b0 = __float_as_uint(v0), b1 = __float_as_uint(v1);
  hi = __byte_perm(b0, b1, 0x7632);
  const float2 d = tc::sub2(make_float2(v0, v1),
                            make_float2(__uint_as_float(b0 & 0xFFFF0000u), __uint_as_float(b1 & 0xFFFF0000u)));
  const __nv_bfloat162 l = __floats2bfloat162_rn(d.x, d.y);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// value stored into a TF32 operand slab: round-to-nearest (ties away) is the +2^12 of the bit pattern; the tensor core
// drops the 13 low mantissa bits itself
__device__ __forceinline__ float tf32_bias(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

__device__ __forceinline__ uint64_t l2_policy(bool last) {
  uint64_t pol;
  if (last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_hint(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint4(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}

// ---- 16-bit forms of what later stages re-read (8 values of an oct <-> one uint4) ----
// Pre-activation a -> phase: a mod 2 pi in units of 2^-16 turn.  Later stages only take sin / cos of it, for gradient
// terms: |error| <= 2 pi 2^-17 = 4.8e-5 rad, below the 2^-11 rounding of the tensor-core operands it feeds.
__device__ __forceinline__ uint32_t phase16(float a) {
  // round(a * 65536 / 2 pi) as a 32-bit integer (saturating); the caller keeps its low 16 bits = the phase mod 2 pi.
  // Exact to one unit while |a| < 1600 rad, where fp32 itself still resolves 2^-16 turn.
  return (uint32_t)__float2int_rn(a * 10430.378350470453f);
}
__device__ __forceinline__ uint4 pack_phase8(const float (&a)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = __byte_perm(phase16(a[2 * i]), phase16(a[2 * i + 1]), 0x5410);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void unpack_phase8(const float4& b, float (&a)[8]) {
  const uint32_t w[4] = {__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)};
  const float kStep = 9.587379924285257e-05f, kOff = -8388608.0f * 9.587379924285257e-05f;   // 2 pi / 65536
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // (2^23 + n) * step - 2^23 * step, one rounding: angle in [0, 2 pi)
    a[2 * i] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7610)), kStep, kOff);
    a[2 * i + 1] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7632)), kStep, kOff);
  }
}
__device__ __forceinline__ uint4 pack_half8(const float (&v)[8], float sc) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(v[2 * i + 1] * sc), "f"(v[2 * i] * sc));
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void unpack_half8(const float4& b, float inv_sc, float (&v)[8]) {
  const uint32_t w[4] = {__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    v[2 * i] = f.x * inv_sc;
    v[2 * i + 1] = f.y * inv_sc;
  }
}

// Sum over the 32 lanes of a warp (= 32 sample points) of 8 per-lane values (= 8 channels) by recursive halving;
// lane L ends up with channel 4 b4 + 2 b3 + b2 (bits of L) and lanes with (L & 3) == 0 add it to dst[ch * stride].
__device__ __forceinline__ void colsum8(const float (&v)[8], float* dst, int stride, int lane) {
  float a4[4], a2[2];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float keep = up ? v[i + 4] : v[i], send = up ? v[i] : v[i + 4];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float keep = up ? a4[i + 2] : a4[i], send = up ? a4[i] : a4[i + 2];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  const bool up = (lane & 4) != 0;
  float a1 = (up ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 4);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  const int ch = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  if ((lane & 3) == 0) atomicAdd(dst + (size_t)ch * stride, a1);
}

// One stage of a tile slot.  For o = 0..7: u = accumulator columns [8o, 8o+8) of this thread (if WAIT), buf = the
// stage's scratch float4s of oct o.  `load(o, buf)` issues the global loads of oct o, PF octs ahead of their use;
// the first PF octs are issued BEFORE the wait on the accumulator barrier (their latency sits under the MMA).
template <int NL, bool WAIT, int PF, class Load, class Body, class WaitAcc, class Next>
__device__ __forceinline__ void run_stage(uint32_t acc, Load load, Body body, WaitAcc wait_acc, Next next) {
  float4 buf[PF + 1][NL > 0 ? NL : 1];
#pragma unroll
  for (int o = 0; o < PF; ++o) load(o, buf[o]);
  uint32_t ub[2][8];
  if (WAIT) {
    wait_acc();
    tmem_ld8_async(acc, ub[0]);
  }
  // the oct loop is unrolled in groups of kOctUnroll (a multiple of the buffer depths, so that every buffer index is
  // static): fully unrolled the kernel is ~17 000 instructions and stalls on instruction fetch
  constexpr int U = OI_BWD_OCT_UNROLL;
  static_assert(8 % U == 0 && U % 2 == 0 && U % (PF + 1) == 0, "oct unroll must divide 8 and cover the buffer rotation");
#pragma unroll 1
  for (int oo = 0; oo < 8; oo += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int o = oo + u;
      if (o + PF < 8) load(o + PF, buf[(u + PF) % (PF + 1)]);
      if (WAIT) {
        tc::wait_ld();
        if (o < 7) tmem_ld8_async(acc + (o + 1) * 8, ub[(u + 1) & 1]);
      }
      body(o, ub[u & 1], buf[u % (PF + 1)]);
      if (o == OI_BWD_PFOCT) next();
    }
  }
}

// F16: the operand slabs are written as scaled fp16 (oi_wgrad.cuh) instead of TF32; both variants are launched and the
// one bwd_mode() does not select exits at once.
template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1) bwd_tc_kernel(const BwdTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  BwdTcSmem& sm = *reinterpret_cast<BwdTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the call's snapshot of the workspace's overflow guard (adj_stats_kernel): the same for every kernel of the call
  const unsigned int guard_state = a.ctl[kCtlGuardSnapshot];
  const BwdMode mode = bwd_mode(a.ctl, a.r.flags, guard_state);
  if (mode.f16 != F16) return;
  const bool probing = !F16 && guard_state != kF16Unsafe && !(a.r.flags & OI_BWD_FLAG_FORCE_TF32);
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

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
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
              tma_bulk_g2s(sm.w[stage] + q * kSubPanelBytes, src + q * kSubPanelBytes, kSubPanelBytes,
                           &sm.w_full[stage]);
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
              issue_layer_lean(tmem_base + t * 256, wbase, idesc);
              tc::mma_commit(&sm.acc_full[t]);
            }
            tc::mma_commit(&sm.w_empty[stage]);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(OI_BWD_EPI_REGS));
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
    const uint64_t pol_first = (OI_BWD_STHINT >= 1) ? l2_policy(false) : 0ull;
    const uint64_t pol_last = (OI_BWD_STHINT >= 2) ? l2_policy(true) : 0ull;
    // operand slabs: K-major SWIZZLE_128B tf32 image [32-point block][channel][32 points]; the 16-byte chunk of this
    // thread's point is XOR-permuted by (channel & 7) = the position e inside an oct
    const int mc = (m & 31) >> 2;
    const int mc4 = mc << 2;   // element e of an oct lands at e * 32 + ((mc ^ e) << 2) = e * 32 + (mc4 ^ (e << 2))

    auto wait_acc = [&]() {
      mbar_wait_sleep(&sm.acc_full[t], af_phase);
      af_phase ^= 1u;
      tc::fence_after_thread_sync();
    };
    auto a_ready = [&]() {
      tc::wait_st();
      tc::fence_before_thread_sync();
      mbar_arrive(&sm.a_ready[t]);
    };
    auto no_load = [](int, float4 (&)[1]) {};
    // L2 prefetch of everything this warp re-reads from one slab in the coming stage: 16 channel quads x 512 B
    // (the slabs were written tens of microseconds ago and have left the L2); p_q0 = this thread's float4 of quad Q0
    auto pf_raw = [&](const float4* p_q0) {
      const float4* base = p_q0 - lane + (lane & 3) * 8;
      l2_prefetch(base + (size_t)(lane >> 2) * 128);
      l2_prefetch(base + (size_t)((lane >> 2) + 8) * 128);
    };
    // OI_BWD_L2PF == 1: a stage prefetches its own re-reads when it starts; == 2: the re-reads of the NEXT stage are
    // prefetched from the middle of the stage before (pf_nxt, called through run_stage's `next` hook)
    auto pf_slab = [&](const float4* p_q0) {
      if (OI_BWD_L2PF == 1) pf_raw(p_q0);
    };
    auto pf_nxt = [&](const float4* p_q0) {
      if (OI_BWD_L2PF == 2) pf_raw(p_q0);
    };
    auto no_next = []() {};

    for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
      const int lt = 2 * pi + t;   // tile index inside this launch
      if (lt >= n_tiles) continue;
      const int tile = a.tile_begin + lt;
      const int inst = tile / a.r.tiles_per_inst;
      const int tin = tile - inst * a.r.tiles_per_inst;
      float* slab_tile = a.slabs + (size_t)lt * kSlabsPerTile * kSlabFloats;
      float4* gs4 = reinterpret_cast<float4*>(slab_tile) + m;                  // ARG slabs [quad][128 points] float4
      float* gso = slab_tile + (m >> 5) * 4096 + (m & 3) + n0 * 32;           // operand slabs, this thread's channels
      // aux operand (N = 16, rows 0..3 used): [32-point block][4 rows][32 points], same swizzle
      float* auxo = a.aux + (size_t)lt * 512 + (m >> 5) * 128 + (m & 3);
      // fp16 operand slabs: 64-point block m >> 6, row = channel (128 B), 16-byte chunk ((m & 63) >> 3) ^ (channel & 7),
      // this lane PAIR's 4 bytes (points m & ~1, m | 1) at ((m & 6) * 2); channel & 7 = the position e inside an oct
      unsigned char* gso16 = reinterpret_cast<unsigned char*>(slab_tile) + (m >> 6) * 16384 + n0 * 128 + (m & 6) * 2;
      const int mc16 = ((m & 63) >> 3) << 4;   // byte offset of the chunk before the XOR: (chunk ^ e) << 4 = mc16 ^ (e << 4)
      const uint32_t pair_sel = (lane & 1) ? 0x3276u : 0x5410u;
      unsigned char* pair_ptr[4];   // + row and chunk of channel e = 2i + (lane & 1) inside an oct
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = 2 * i + (lane & 1);
        pair_ptr[i] = gso16 + e * 128 + (mc16 ^ (e << 4));
      }
      unsigned char* auxo16 = reinterpret_cast<unsigned char*>(a.aux + (size_t)lt * 512) + (m >> 6) * 512 + (m & 7) * 2;
      float* dfilm = a.d_film + (size_t)inst * kFilm * 2 * kW;   // [slot][unused | db][128]
      float* dw0 = a.dw0 + (size_t)inst * kW * 3;
      {
        const float2* src = reinterpret_cast<const float2*>(a.r.film_tc) + (size_t)inst * kFilm * kW;
        float2* dst = &sm.film[t][0][0];
        for (int i = sub; i < kFilm * kW; i += kEpiThreadsPerSlot) dst[i] = src[i];
      }
      float px, py, pz, sdf_bar, nb0, nb1, nb2, zb0, zb1, zb2;
      // fp16 slabs: the point's upstream adjoints are scaled by sc_adj = 2^-e_m right here, so that EVERY adjoint of the
      // sweep carries the scale (exact: a power of two; the bf16-split MMA operands have fp32's exponent range) and the
      // adjoint-type operands / c_bar need no multiply; sums over points taken on the FMA pipe un-scale first
      // (sc_adj_inv, the un-scaled z_bar / normal_bar below).  Forward-type operands are multiplied by sc_fwd.
      float sc_fwd = 1.f, sc_adj_inv = 1.f;
      float zu0, zu1, zu2;   // un-scaled z_bar
      float rs_fwd = 0.f, rs_adj = 0.f;
      uint32_t guard = 0u;
      __half2 guard2 = __float2half2_rn(0.f);
      float probe_adj = 0.f, probe_fwd = 0.f, probe_max = 0.f;
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
        zu0 = zb0; zu1 = zb1; zu2 = zb2;
        if (F16) {
          const int e_m = adj_exponent(q0, q1);
          const float sc_adj = pow2i(-e_m - kF16AdjShift);
          sc_adj_inv = pow2i(e_m + kF16AdjShift);
          sc_fwd = pow2i(e_m - mode.e_ref);
          sdf_bar *= sc_adj; nb0 *= sc_adj; nb1 *= sc_adj; nb2 *= sc_adj;
          zb0 *= sc_adj; zb1 *= sc_adj; zb2 *= sc_adj;
        } else if (probing) {   // what the fp16 path would multiply the two operand types by
          const int e_m = adj_exponent(q0, q1);
          probe_adj = pow2i(-e_m - kF16AdjShift);
          probe_fwd = pow2i(e_m - mode.e_ref);
        }
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);

#define OI_ARG(l, quad) gs4[((size_t)(kSlabArg + (l)) * 32 + (quad)) * 128]
#define OI_CTA(slab, quad) scr4[((size_t)(slab) * 32 + (quad)) * 128]
#define OI_FILM4(l) (reinterpret_cast<const float4*>(sm.film[t][(l)]) + n0 / 2)   /* (g0, g1, d0, d1) per pair */
      // the 8 values of an oct -> operand slab `slab` (rounded to TF32)
      auto op8 = [&](int slab, int o, const float (&v)[8], bool adjoint) {
        if (F16) {
          // lanes 2j / 2j+1 swap one value per channel pair: the even lane then holds channel 2i of both points, the odd
          // lane channel 2i+1, and each stores ONE packed fp16x2 (4 stores per oct instead of 8)
          const float sc = adjoint ? 1.0f : sc_fwd;
          if (OI_BWD_RANGE_STATS) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (adjoint) rs_adj = fmaxf(rs_adj, fabsf(v[e]));
              else rs_fwd = fmaxf(rs_fwd, fabsf(v[e] * sc));
            }
          }
          const uint32_t uoff = (uint32_t)slab16_offset(slab) + (uint32_t)o * 1024u;   // warp-uniform part of the address
          uint32_t pks[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t own;   // (channel 2i+1, channel 2i) of this lane's point
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(own) : "f"(v[2 * i + 1] * sc), "f"(v[2 * i] * sc));
            const uint32_t nbr = __shfl_xor_sync(0xffffffffu, own, 1);
            // even lane: channel 2i of points (m, m+1) = (own.lo, nbr.lo); odd lane: channel 2i+1 of (m-1, m) = (nbr.hi, own.hi)
            const uint32_t pk = __byte_perm(own, nbr, pair_sel);
            *reinterpret_cast<uint32_t*>(pair_ptr[i] + uoff) = pk;
            pks[i] = pk;
          }
          // overflow guard: running max |x| over every operand this thread writes (one packed-half max per word)
#pragma unroll
          for (int i = 0; i < 4; ++i) guard2 = __hmax2(guard2, __habs2(*reinterpret_cast<const __half2*>(&pks[i])));
          return;
        }
        float* p = gso + (size_t)slab * kSlabFloats + o * 256;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (OI_BWD_STHINT >= 1) st_hint(p + e * 32 + (mc4 ^ (e << 2)), tf32_bias(v[e]), pol_first);
          else p[e * 32 + (mc4 ^ (e << 2))] = tf32_bias(v[e]);
        }
        if (probing) {   // magnitude of the fp16 operands these values would be
          const float sc = adjoint ? probe_adj : probe_fwd;
          float mx = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) mx = fmaxf(mx, fabsf(v[e]));
          probe_max = fmaxf(probe_max, mx * sc);
        }
      };
      auto st_arg = [&](float4* p, float4 v) {
        if (OI_BWD_STHINT >= 3) st_hint4(p, v, pol_last);
        else *p = v;
      };
      auto st_cta = [&](float4* p, float4 v) {
        if (OI_BWD_STHINT >= 2) st_hint4(p, v, pol_last);
        else *p = v;
      };
      // ---- what later stages re-read: pre-activations (ARG slabs) and the per-CTA g / c_bar slabs ----
      constexpr bool kArg16 = OI_BWD_ARG16 != 0;
      constexpr bool kG16 = F16 && (OI_BWD_G16 != 0);
      // 16-bit layout of a slab: [oct 0..15][128 points] uint4 (32 KB, the first half of the slab's 64 KB)
      uint4* arg16 = reinterpret_cast<uint4*>(slab_tile) + (size_t)(h * 8) * 128 + m;
      uint4* cta16 = reinterpret_cast<uint4*>(scr4 - m) + (size_t)(h * 8) * 128 + m;
      auto arg_st = [&](int l, int o, const float (&ar)[8]) {
        if (kArg16) {
          arg16[(size_t)l * (kSlabFloats / 4) + o * 128] = pack_phase8(ar);
        } else {
          st_arg(&OI_ARG(l, Q0 + 2 * o), make_float4(ar[0], ar[1], ar[2], ar[3]));
          st_arg(&OI_ARG(l, Q0 + 2 * o + 1), make_float4(ar[4], ar[5], ar[6], ar[7]));
        }
      };
      auto arg_ld = [&](int l, int o, float4& b0, float4& b1) {
        if (kArg16) {
          b0 = *reinterpret_cast<const float4*>(&arg16[(size_t)l * (kSlabFloats / 4) + o * 128]);
        } else {
          b0 = OI_ARG(l, Q0 + 2 * o);
          b1 = OI_ARG(l, Q0 + 2 * o + 1);
        }
      };
      auto arg_get = [&](const float4& b0, const float4& b1, float (&ar)[8]) {
        if (kArg16) {
          unpack_phase8(b0, ar);
        } else {
          ar[0] = b0.x; ar[1] = b0.y; ar[2] = b0.z; ar[3] = b0.w;
          ar[4] = b1.x; ar[5] = b1.y; ar[6] = b1.z; ar[7] = b1.w;
        }
      };
      // g (forward-type, stored as is) / c_bar (adjoint-type, scaled like the adjoint operands)
      auto g_st = [&](int slab, int o, const float (&v)[8], bool adjoint) {
        if (kG16) {
          cta16[(size_t)slab * (kSlabFloats / 4) + o * 128] = pack_half8(v, 1.0f);   // c_bar carries the adjoint scale already
        } else {
          st_cta(&OI_CTA(slab, Q0 + 2 * o), make_float4(v[0], v[1], v[2], v[3]));
          st_cta(&OI_CTA(slab, Q0 + 2 * o + 1), make_float4(v[4], v[5], v[6], v[7]));
        }
      };
      auto g_ld = [&](int slab, int o, float4& b0, float4& b1) {
        if (kG16) {
          b0 = *reinterpret_cast<const float4*>(&cta16[(size_t)slab * (kSlabFloats / 4) + o * 128]);
        } else {
          b0 = OI_CTA(slab, Q0 + 2 * o);
          b1 = OI_CTA(slab, Q0 + 2 * o + 1);
        }
      };
      auto g_get = [&](const float4& b0, const float4& b1, float (&v)[8], bool adjoint) {
        if (kG16) {
          unpack_half8(b0, 1.0f, v);
        } else {
          v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w;
          v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
        }
      };
      // L2 prefetch of this warp's part of a 16-bit slab: 8 octs x 512 B = 32 lines, one per lane
      auto pf16 = [&](const uint4* p_oct0) { l2_prefetch(p_oct0 - lane + (size_t)(lane >> 2) * 128 + (lane & 3) * 8); };
      auto arg_pf = [&](int l, bool next) {
        if (OI_BWD_L2PF != (next ? 2 : 1)) return;
        if (kArg16) pf16(arg16 + (size_t)l * (kSlabFloats / 4));
        else pf_raw(&OI_ARG(l, Q0));
      };
      auto g_pf = [&](int slab, bool next) {
        if (OI_BWD_L2PF != (next ? 2 : 1)) return;
        if (kG16) pf16(cta16 + (size_t)slab * (kSlabFloats / 4));
        else pf_raw(&OI_CTA(slab, Q0));
      };
      // the 8 values of an oct -> next A operand (TMEM), fp16 or bf16 two-term split
      auto a8_f16 = [&](int o, const float (&v)[8]) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tc::split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
        tmem_st4(a_hi + o * 4, hi);
        tmem_st4(a_lo + o * 4, lo);
      };
      auto a8_bf16 = [&](int o, const float (&v)[8]) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2_bf16t(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
        tmem_st4(a_hi + o * 4, hi);
        tmem_st4(a_lo + o * 4, lo);
      };

      // =========================== recompute ===========================
      // ---------------- layer 0 (K = 3) on the FMA pipe ----------------
      {
        const float4* fl = OI_FILM4(0);
        run_stage<0, false, 1>(acc, no_load, [&](int o, const uint32_t (&)[8], const float4 (&)[1]) {
          float ar[8], s[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 wa = sm.w0[n0 + o * 8 + 2 * i], wb = sm.w0[n0 + o * 8 + 2 * i + 1];
            const float4 f = fl[o * 4 + i];
            const float2 u = make_float2(fmaf(wa.z, pz, fmaf(wa.y, py, wa.x * px)),
                                         fmaf(wb.z, pz, fmaf(wb.y, py, wb.x * px)));
            const float2 arg = tc::fma2(make_float2(f.x, f.y), u, make_float2(f.z, f.w));
            ar[2 * i] = arg.x;
            ar[2 * i + 1] = arg.y;
            s[2 * i] = __sinf(arg.x);
            s[2 * i + 1] = __sinf(arg.y);
          }
          arg_st(0, o, ar);
          op8(kSlabH + 1, o, s, false);
          a8_f16(o, s);
        }, wait_acc, no_next);
        a_ready();
      }
      // ---------------- forward layers 1..D-1 ----------------
      for (int l = 1; l < D; ++l) {
        const float4* fl = OI_FILM4(l);
        run_stage<0, true, 1>(acc, no_load, [&](int o, const uint32_t (&u)[8], const float4 (&)[1]) {
          float ar[8], s[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float2 arg = tc::fma2(make_float2(f.x, f.y),
                                        make_float2(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1])),
                                        make_float2(f.z, f.w));
            ar[2 * i] = arg.x;
            ar[2 * i + 1] = arg.y;
            s[2 * i] = __sinf(arg.x);
            s[2 * i + 1] = __sinf(arg.y);
          }
          arg_st(l, o, ar);
          op8(kSlabH + l + 1, o, s, false);
          a8_f16(o, s);
        }, wait_acc, no_next);
        a_ready();
      }
      // ---------------- colour features -> slot UC; t_{D-1} = w_sigma gamma cos(a_{D-1}) ----------------
      {
        const float4* fl = OI_FILM4(D - 1);
        arg_pf(D - 1, false);
        run_stage<2, true, OI_BWD_PF2>(acc, [&](int o, float4 (&b)[2]) {
          arg_ld(D - 1, o, b[0], b[1]);
        }, [&](int o, const uint32_t (&u)[8], const float4 (&b)[2]) {
          st_cta(&OI_CTA(kCtaUC, Q0 + 2 * o), make_float4(__uint_as_float(u[0]), __uint_as_float(u[1]), __uint_as_float(u[2]),
                                                   __uint_as_float(u[3])));
          st_cta(&OI_CTA(kCtaUC, Q0 + 2 * o + 1), make_float4(__uint_as_float(u[4]), __uint_as_float(u[5]),
                                                       __uint_as_float(u[6]), __uint_as_float(u[7])));
          float ar[8];
          arg_get(b[0], b[1], ar);
          float tv[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const int n = n0 + o * 8 + 2 * i;
            // 2^8 w_s * gamma / 2^8 * cos
            tv[2 * i] = sm.head[n].x * f.x * __cosf(ar[2 * i]);
            tv[2 * i + 1] = sm.head[n + 1].x * f.y * __cosf(ar[2 * i + 1]);
          }
          op8(kSlabT + D - 1, o, tv, false);
          a8_f16(o, tv);
        }, wait_acc, [&]() { if (D >= 2) arg_pf(D - 2, true); });
        a_ready();
      }
      // ---------------- reverse sweep l = D-1 .. 1: g_l (slot G[l]), t_{l-1} (slab T[l-1]) ----------------
      float gx = 0.f, gy = 0.f, gz = 0.f;
      for (int l = D - 1; l >= 1; --l) {
        const float4* fl = OI_FILM4(l - 1);
        const float gscale = (l - 1 == 0) ? kInvWScale : 1.0f;   // gamma'_0 is unscaled
        arg_pf(l - 1, false);
        run_stage<2, true, OI_BWD_PF2>(acc, [&](int o, float4 (&b)[2]) {
          arg_ld(l - 1, o, b[0], b[1]);
        }, [&](int o, const uint32_t (&u)[8], const float4 (&b)[2]) {
          float ar[8];
          arg_get(b[0], b[1], ar);
          float gv[8], tv[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float a0 = __uint_as_float(u[2 * i]), a1 = __uint_as_float(u[2 * i + 1]);   // 2^8 g_l
            gv[2 * i] = a0 * kInvWScale;
            gv[2 * i + 1] = a1 * kInvWScale;
            tv[2 * i] = a0 * (f.x * gscale) * __cosf(ar[2 * i]);           // g_l gamma cos = t_{l-1}
            tv[2 * i + 1] = a1 * (f.y * gscale) * __cosf(ar[2 * i + 1]);
          }
          g_st(kCtaG + l - 1, o, gv, false);
          if (l > 1) {
            op8(kSlabT + l - 1, o, tv, false);
            a8_f16(o, tv);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 w = sm.w0[n0 + o * 8 + e];
              gx = fmaf(w.x, tv[e], gx);
              gy = fmaf(w.y, tv[e], gy);
              gz = fmaf(w.z, tv[e], gz);
            }
          }
        }, wait_acc, [&]() { if (l >= 2) arg_pf(l - 2, true); else pf_nxt(&OI_CTA(kCtaUC, Q0)); });
        if (l > 1) a_ready();
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
        const float4* fl = OI_FILM4(OI_MAX_DEPTH);
        pf_slab(&OI_CTA(kCtaUC, Q0));
        run_stage<2, false, OI_BWD_PF2>(acc, [&](int o, float4 (&b)[2]) {
          b[0] = OI_CTA(kCtaUC, Q0 + 2 * o);
          b[1] = OI_CTA(kCtaUC, Q0 + 2 * o + 1);
        }, [&](int o, const uint32_t (&)[8], const float4 (&b)[2]) {
          const float uc[8] = {b[0].x, b[0].y, b[0].z, b[0].w, b[1].x, b[1].y, b[1].z, b[1].w};
          float ub_[8], sn[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float gp[2] = {f.x, f.y}, dl[2] = {f.z, f.w};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = 2 * i + e2;
              const float4 hd = sm.head[n0 + o * 8 + e];
              float pre = fmaf(hd.y, gx, uc[e]);
              pre = fmaf(hd.z, gy, pre);
              pre = fmaf(hd.w, gz, pre);
              const float arg = fmaf(gp[e2], pre, dl[e2]);
              const float cs = __cosf(arg);
              sn[e] = __sinf(arg);
              const float4 rw = sm.rgbw[n0 + o * 8 + e];
              const float hb = fmaf(rw.x, zb0, fmaf(rw.y, zb1, rw.z * zb2));
              ub_[e] = hb * cs * (gp[e2] * kWScale);               // u_bar_c = a_bar * gamma
              nc0 = fmaf(hd.y * kInvWScale, ub_[e], nc0);
              nc1 = fmaf(hd.z * kInvWScale, ub_[e], nc1);
              nc2 = fmaf(hd.w * kInvWScale, ub_[e], nc2);
            }
          }
          op8(kSlabUBC, o, ub_, true);
          a8_bf16(o, ub_);
          if (o == 7) a_ready();   // the tensor core starts on W_cf^T u_bar_c while the column sums are reduced
          {
            float tmp[8];
            const int nc = n0 + o * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = zu0 * sn[i];
            colsum8(tmp, a.g.rgb_weight + 0 * kW + nc, 1, lane);                       // dW_rgb = z_bar (x) h_c
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = zu1 * sn[i];
            colsum8(tmp, a.g.rgb_weight + 1 * kW + nc, 1, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = zu2 * sn[i];
            colsum8(tmp, a.g.rgb_weight + 2 * kW + nc, 1, lane);
          }
        }, wait_acc, [&]() { arg_pf(0, true); g_pf(kCtaG + 0, true); });
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
      const float nu0 = F16 ? nb0 * sc_adj_inv : nb0, nu1 = F16 ? nb1 * sc_adj_inv : nb1,
                  nu2 = F16 ? nb2 * sc_adj_inv : nb2;   // un-scaled normal_bar
      if (h == 0) {
        if (F16) {   // rows of 128 B = 64 points, chunk XOR row
          const float sc_n = sc_fwd * (1.0f / (float)(1 << kF16NormalShift));
          const float av[4] = {gx * sc_n, gy * sc_n, gz * sc_n, sc_fwd};
          // overflow guard: the normal is the largest forward-type operand when the SDF head grows (every tile: 3 values)
          if (fmaxf(fmaxf(fabsf(av[0]), fabsf(av[1])), fabsf(av[2])) >= kF16GuardLimit) guard |= 0x80000000u;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            unsigned short hbits;
            asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hbits) : "f"(av[r]));
            *reinterpret_cast<unsigned short*>(auxo16 + r * 128 + (mc16 ^ (r << 4))) = hbits;
          }
        } else {
          if (probing)
            probe_max = fmaxf(probe_max, fmaxf(fmaxf(fabsf(gx), fabsf(gy)), fabsf(gz)) * probe_fwd *
                                             (1.0f / (float)(1 << kF16NormalShift)));
          auxo[0 * 32 + ((mc ^ 0) << 2)] = tf32_bias(gx);
          auxo[1 * 32 + ((mc ^ 1) << 2)] = tf32_bias(gy);
          auxo[2 * 32 + ((mc ^ 2) << 2)] = tf32_bias(gz);
          auxo[3 * 32 + ((mc ^ 3) << 2)] = 1.0f;
        }
      }
      // ---------------- h_bar_D = W_cf^T u_bar_c + sdf_bar w_s -> slot HB;
      //                  backward of the reverse sweep, l = 0 (K = 3): A <- g_bar_1 ----------------
      {
        const float4* fl = OI_FILM4(0);
        arg_pf(0, false);
        g_pf(kCtaG + 0, false);
        run_stage<4, true, OI_BWD_PF4>(acc, [&](int o, float4 (&b)[4]) {
          arg_ld(0, o, b[0], b[1]);
          g_ld(kCtaG + 0, o, b[2], b[3]);
        }, [&](int o, const uint32_t (&u)[8], const float4 (&b)[4]) {
          float ar[8];
          arg_get(b[0], b[1], ar);
          float g1[8];
          g_get(b[2], b[3], g1, false);
          float hb[8], cb[8], gb[8], t0[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float gp[2] = {f.x, f.y};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = 2 * i + e2;
              hb[e] = fmaf(sdf_bar, sm.head[n0 + o * 8 + e].x * kInvWScale, __uint_as_float(u[e]));
              const float4 w = sm.w0[n0 + o * 8 + e];
              const float tb = fmaf(w.x, nb0, fmaf(w.y, nb1, w.z * nb2));   // t_bar_0 = W_0 normal_bar
              const float c0 = gp[e2] * __cosf(ar[e]);                      // gamma_0 cos a_0
              cb[e] = tb * g1[e];
              gb[e] = tb * c0;
              t0[e] = g1[e] * c0;   // t_0
            }
          }
          st_cta(&OI_CTA(kCtaHB, Q0 + 2 * o), make_float4(hb[0], hb[1], hb[2], hb[3]));
          st_cta(&OI_CTA(kCtaHB, Q0 + 2 * o + 1), make_float4(hb[4], hb[5], hb[6], hb[7]));
          g_st(kCtaG + 0, o, cb, true);
          op8(kSlabGB + 1, o, gb, true);
          a8_bf16(o, gb);
          if (o == 7) a_ready();
          {  // dW_0 += t_0 (x) normal_bar (per instance)
            float* dst = dw0 + (size_t)(n0 + o * 8) * 3;
            float tmp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = nu0 * t0[i];
            colsum8(tmp, dst + 0, 3, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = nu1 * t0[i];
            colsum8(tmp, dst + 1, 3, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = nu2 * t0[i];
            colsum8(tmp, dst + 2, 3, lane);
          }
        }, wait_acc, [&]() { if (1 < D - 1) { arg_pf(1, true); g_pf(kCtaG + 1, true); } else { arg_pf(D - 1, true); pf_nxt(&OI_CTA(kCtaHB, Q0)); } });
      }
      // ---------------- backward of the reverse sweep, l = 1..D-2: t_bar_l = W_l g_bar_l ----------------
      for (int l = 1; l < D - 1; ++l) {
        const float4* fl = OI_FILM4(l);
        arg_pf(l, false);
        g_pf(kCtaG + l, false);
        run_stage<4, true, OI_BWD_PF4>(acc, [&](int o, float4 (&b)[4]) {
          arg_ld(l, o, b[0], b[1]);
          g_ld(kCtaG + l, o, b[2], b[3]);
        }, [&](int o, const uint32_t (&u)[8], const float4 (&b)[4]) {
          float ar[8];
          arg_get(b[0], b[1], ar);
          float gn[8];
          g_get(b[2], b[3], gn, false);
          float cb[8], gb[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float2 gam = tc::mul2(make_float2(f.x, f.y), make_float2(kWScale, kWScale));
            const float tb0 = __uint_as_float(u[2 * i]), tb1 = __uint_as_float(u[2 * i + 1]);
            cb[2 * i] = tb0 * gn[2 * i];                                  // c_bar_l
            cb[2 * i + 1] = tb1 * gn[2 * i + 1];
            gb[2 * i] = tb0 * gam.x * __cosf(ar[2 * i]);                 // g_bar_{l+1}
            gb[2 * i + 1] = tb1 * gam.y * __cosf(ar[2 * i + 1]);
          }
          g_st(kCtaG + l, o, cb, true);
          op8(kSlabGB + l + 1, o, gb, true);
          a8_bf16(o, gb);
        }, wait_acc, [&]() { if (l + 1 < D - 1) { arg_pf(l + 1, true); g_pf(kCtaG + l + 1, true); } else { arg_pf(D - 1, true); pf_nxt(&OI_CTA(kCtaHB, Q0)); } });
        a_ready();
      }
      // ---------------- top, l = D-1: t_{D-1} = w_s c_{D-1}; then the backward of the forward sweep for layer D-1
      {
        const int l = D - 1;
        const float4* fl = OI_FILM4(l);
        arg_pf(l, false);
        pf_slab(&OI_CTA(kCtaHB, Q0));
        run_stage<4, true, OI_BWD_PF4>(acc, [&](int o, float4 (&b)[4]) {
          arg_ld(l, o, b[0], b[1]);
          b[2] = OI_CTA(kCtaHB, Q0 + 2 * o);          // h_bar_D
          b[3] = OI_CTA(kCtaHB, Q0 + 2 * o + 1);
        }, [&](int o, const uint32_t (&u)[8], const float4 (&b)[4]) {
          float ar[8];
          arg_get(b[0], b[1], ar);
          const float hb[8] = {b[2].x, b[2].y, b[2].z, b[2].w, b[3].x, b[3].y, b[3].z, b[3].w};
          float ubv[8], dws[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float gp[2] = {f.x * kWScale, f.y * kWScale};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = 2 * i + e2;
              const float tb = __uint_as_float(u[e]);
              const float gam = gp[e2];
              const float sn = __sinf(ar[e]), cs = __cosf(ar[e]);
              dws[e] = fmaf(sdf_bar, sn, tb * gam * cs);   // d w_s = sdf_bar h_D + t_bar_{D-1} c_{D-1}
              const float cbar = tb * (sm.head[n0 + o * 8 + e].x * kInvWScale);
              const float ab = hb[e] * cs - cbar * gam * sn;
              ubv[e] = ab * gam;   // u_bar_{D-1}
            }
          }
          op8(kSlabUB + l, o, ubv, true);
          a8_bf16(o, ubv);
          if (o == 7) a_ready();
          if (F16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dws[i] *= sc_adj_inv;
          }
          colsum8(dws, a.g.sigma_weight + n0 + o * 8, 1, lane);
        }, wait_acc, [&]() { arg_pf(D - 2, true); g_pf(kCtaG + D - 2, true); });
      }
      // ---------------- backward of the forward sweep: h_bar_l = W_l^T u_bar_l, then layer k = l-1 ----------------
      for (int l = D - 1; l >= 1; --l) {
        const int k = l - 1;
        const float4* fl = OI_FILM4(k);
        const float gsc = (k == 0) ? 1.0f : kWScale;
        arg_pf(k, false);
        g_pf(kCtaG + k, false);
        run_stage<4, true, OI_BWD_PF4>(acc, [&](int o, float4 (&b)[4]) {
          arg_ld(k, o, b[0], b[1]);
          g_ld(kCtaG + k, o, b[2], b[3]);
        }, [&](int o, const uint32_t (&u)[8], const float4 (&b)[4]) {
          float ar[8];
          arg_get(b[0], b[1], ar);
          float cb[8];
          g_get(b[2], b[3], cb, true);
          float ubv[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 f = fl[o * 4 + i];
            const float gp[2] = {f.x * gsc, f.y * gsc};
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const int e = 2 * i + e2;
              const float gam = gp[e2];
              const float sn = __sinf(ar[e]), cs = __cosf(ar[e]);
              const float ab = __uint_as_float(u[e]) * cs - cb[e] * gam * sn;
              ubv[e] = ab * gam;   // u_bar_k
            }
          }
          if (k >= 1) {
            op8(kSlabUB + k, o, ubv, true);
            a8_bf16(o, ubv);
          } else {
            const int nc = n0 + o * 8;
            if (F16) {
#pragma unroll
              for (int i = 0; i < 8; ++i) ubv[i] *= sc_adj_inv;
            }
            colsum8(ubv, dfilm + kW + nc, 1, lane);                  // d b_0 = sum u_bar_0
            float* dst = dw0 + (size_t)nc * 3;                       // dW_0 += u_bar_0 (x) x (per instance)
            float tmp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = px * ubv[i];
            colsum8(tmp, dst + 0, 3, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = py * ubv[i];
            colsum8(tmp, dst + 1, 3, lane);
#pragma unroll
            for (int i = 0; i < 8; ++i) tmp[i] = pz * ubv[i];
            colsum8(tmp, dst + 2, 3, lane);
          }
        }, wait_acc, [&]() { if (l >= 2) { arg_pf(l - 2, true); g_pf(kCtaG + l - 2, true); } });
        if (k >= 1) a_ready();
      }
      if ((F16 && ((guard & 0x80000000u) || fmaxf(__low2float(guard2), __high2float(guard2)) >= kF16GuardLimit)) ||
          (!F16 && probe_max >= kF16GuardLimit))
        *a.sticky = kF16Unsafe;
      if (OI_BWD_RANGE_STATS) {
        atomicMax(const_cast<unsigned int*>(a.ctl) + 6, __float_as_uint(rs_fwd));
        atomicMax(const_cast<unsigned int*>(a.ctl) + 7, __float_as_uint(rs_adj));
      }
      named_bar_sync(1 + t, kEpiThreadsPerSlot);   // film table / exchange buffer of this slot may be reused now
    }
#undef OI_ARG
#undef OI_CTA
#undef OI_FILM4
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == kProducerWarp) tc::tmem_dealloc(tmem_base, 512);
}

// Finalize of the tensor-core backward.  Inputs: the per-instance weight gradients dW_l (both parts) left by the
// contraction kernel (l >= 1, colour) and by the sweep kernel's butterflies (l = 0), and db = sum_m u_bar per
// instance.  With a = gamma u + beta, u = W h + b (see the header of this file):
//   dL/dbeta = db / gamma,    dL/dgamma = ( sum_k W[j][k] dW[j][k] + b[j] db[j] ) / gamma,
// db and dW are summed over the instances into the bias / weight gradients; plus the variance.
// Block l = 0..D-1: pts_linears[l]; block D: views_linears; block D+1: variance.  Thread = output channel j.
__global__ void finalize_bwd_tc_kernel(int D, int n_inst, int R, const float* __restrict__ film,
                                       const float* __restrict__ d_film, const float* __restrict__ invs_partial,
                                       const float* __restrict__ blob, const float* __restrict__ dwi,
                                       const float* __restrict__ dwc, const float* __restrict__ dw0, OiNetGrads g,
                                       unsigned int* guard) {
  const int n = threadIdx.x;
  const int l = blockIdx.x;
  const BlobLayout L = blob_layout(D);
  const float* cst = blob + L.const_off;
  const float* stream = blob + L.stream_off;
  if (l <= D) {
    const int slot = (l < D) ? l : OI_MAX_DEPTH;
    const float bias = cst[BlobLayout::kBias + slot * kW + n];
    float s = 0.f;
    for (int i = 0; i < n_inst; ++i) {
      float dot = 0.f;
      if (l == 0) {
        const float* row = dw0 + ((size_t)i * kW + n) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float v = row[k];
          dot = fmaf(cst[BlobLayout::kW0t + k * kW + n], v, dot);
          g.pts_weight[0][n * 3 + k] += v;
        }
      } else if (l < D) {
        const float4* row = reinterpret_cast<const float4*>(dwi + (((size_t)i * (D - 1) + (l - 1)) * kW + n) * kW);
        const float4* wr = reinterpret_cast<const float4*>(stream + (size_t)(1 + 8 * (2 * D - 1 - l)) * kChunkFloats +
                                                           (size_t)n * kW);   // W_l as stored: row = output channel
        float4* out = reinterpret_cast<float4*>(g.pts_weight[l] + (size_t)n * kW);
        for (int k = 0; k < kW / 4; ++k) {
          const float4 v = row[k], w = wr[k];
          dot = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, dot))));
          float4 o = out[k];
          o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
          out[k] = o;
        }
      } else {
        const float* row = dwc + ((size_t)i * kW + n) * (kW + 3);
        const float* wr = stream + (size_t)(1 + 8 * (2 * D - 1)) * kChunkFloats + (size_t)n * kW;   // W_c[:, :128] as stored
        float* out = g.views_weight + (size_t)n * (kW + 3);
        for (int k = 0; k < kW; ++k) {
          const float v = row[k];
          dot = fmaf(wr[k], v, dot);
          out[k] += v;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float v = row[kW + k];
          dot = fmaf(cst[BlobLayout::kWcg + k * kW + n], v, dot);
          out[kW + k] += v;
        }
      }
      const size_t o = ((size_t)i * kFilm + slot) * 2 * kW;
      const float db = d_film[o + kW + n];
      const float gam = film[o + n];
      s += db;
      g.film_gamma[((size_t)i * kFilm + slot) * kW + n] += fmaf(bias, db, dot) / gam;
      g.film_beta[((size_t)i * kFilm + slot) * kW + n] += db / gam;
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
      const float inv_s = cst[BlobLayout::kScalars + 4];
      const double tot = red[0] + red[1] + red[2] + red[3];
      const bool inside = inv_s > 1e-6f && inv_s < 1e6f;
      if (inside) g.variance[0] += (float)(tot * 10.0 * (double)inv_s);
      // fp16 overflow guard (oi_wgrad.cuh): a call that ran (probed) without tripping it marks the workspace safe
      if (guard != nullptr && *guard != kF16Unsafe) *guard = kF16Safe;
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
// per-instance weight gradients: [n_inst][D-1][128][128] | [n_inst][128][131] | [n_inst][128][3]
size_t render_bwd_tc_dw_floats(int n_inst, int depth) {
  return (size_t)n_inst * ((size_t)(depth - 1) * kW * kW + (size_t)kW * (kW + 3) + (size_t)kW * 3) + 16;
}

// Runs the two tensor-core kernels over tiles [0, n_tiles) in chunks of at most `chunk_tiles`.
int launch_render_bwd_tc(const OiRenderBwdDesc& d, const RenderKArgs& geo, const float* adj, const float* invs_partial,
                         float* d_film, float* scratch, float* slabs, float* aux, float* dw_inst,
                         const unsigned int* ctl, unsigned int* sticky, int chunk_tiles, int n_ctas, cudaStream_t st) {
  const int n_inst = geo.n_inst, D = geo.D;
  OI_CHECK_CUDA(cudaFuncSetAttribute(bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(BwdTcSmem)));
  OI_CHECK_CUDA(cudaFuncSetAttribute(bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(BwdTcSmem)));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* dwi = dw_inst;
  float* dwc = dwi + (((size_t)n_inst * (D - 1) * kW * kW + 3) & ~(size_t)3);
  float* dw0 = dwc + (((size_t)n_inst * kW * (kW + 3) + 3) & ~(size_t)3);
  OI_CHECK_CUDA(cudaMemsetAsync(dw_inst, 0, render_bwd_tc_dw_floats(n_inst, D) * sizeof(float), st));

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
    a.dw0 = dw0;
    const int ctas = render_bwd_tc_ctas(t1 - t0) < n_ctas ? render_bwd_tc_ctas(t1 - t0) : n_ctas;
    a.ctl = ctl;
    a.sticky = sticky;
    // both operand formats are launched; bwd_mode() (adjoint statistics, on the device) lets one of them exit at once
    // (also with a forced format: bwd_mode() overrides a forced fp16 when the adjoints are not finite)
    bwd_tc_kernel<false><<<ctas, kTcThreads, sizeof(BwdTcSmem), st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
    bwd_tc_kernel<true><<<ctas, kTcThreads, sizeof(BwdTcSmem), st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());

    // ---- contraction over the points of this chunk (per-instance outputs: the finalize kernel needs them)
    WgArgs w;
    memset(&w, 0, sizeof(w));
    w.n_tiles = t1 - t0;
    w.tiles_per_inst = geo.tiles_per_inst;
    w.slabs_per_tile = kSlabsPerTile;
    w.slabs = slabs;
    w.aux = aux;
    w.ctl = ctl;
    w.flags = geo.flags;
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
      g.out = dwi + (size_t)(l - 1) * kW * kW;
      g.out_inst_stride = (D - 1) * kW * kW;
      g.out_ld = kW;
      g.weight = 4;
    }
    {  // colour layer: u_bar_c (x) h_D, u_bar_c (x) normal, sum u_bar_c
      WgGroup& g = w.groups[ng++];
      g.n_pairs = 1;
      g.pairs[0] = WgPair{kSlabUBC, kSlabH + D};
      g.use_aux = 1;
      for (int j = 0; j < 3; ++j) {
        g.aux_out[j] = dwc + kW + j;
        g.aux_inst_stride[j] = kW * (kW + 3);
        g.aux_ch_stride[j] = kW + 3;
      }
      g.aux_out[3] = dbia(OI_MAX_DEPTH);
      g.aux_inst_stride[3] = inst_stride;
      g.aux_ch_stride[3] = 1;
      g.out = dwc;
      g.out_inst_stride = kW * (kW + 3);
      g.out_ld = kW + 3;
      g.weight = 2;
    }
    w.n_groups = ng;
    w.n_ctas = sms;
    int rc = launch_wgrad_tc(w, st);
    if (rc) return rc;
  }
  if (d.evt_core_stop) OI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(d.evt_core_stop), st));

  finalize_bwd_tc_kernel<<<D + 2, kW, 0, st>>>(D, n_inst, d.n_rays, geo.film, d_film, invs_partial, geo.blob, dwi, dwc,
                                               dw0, d.grads,
                                               (geo.flags & OI_BWD_FLAG_FORCE_TF32) ? nullptr : sticky);   // no probe ran
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
