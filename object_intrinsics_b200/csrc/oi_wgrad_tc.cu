// Point-contraction kernel of the tensor-core backward (second kernel of oi_render_backward, OI_IMPL_TCGEN05):
//     dW[i][j] = sum over sample points m of X[m][i] * Y[m][j]          (128 x 128 weight gradients)
//     dv[i][c] = sum over sample points m of X[m][i] * aux[m][c]        (c < 4: bias / narrow-column gradients)
// The first kernel (oi_render_bwd_tc.cu) leaves every operand as an fp32 "slab" per 128-point tile whose memory
// layout IS the canonical K-major SWIZZLE_128B UMMA operand image for kind::tf32 (the contraction index K = points:
// blocks of 32 points, each [128 channels][32 points] with 128-byte rows whose 16-byte chunks are XOR-permuted by
// channel & 7): 64 points of an operand are one contiguous 32 KB TMA bulk copy straight into shared memory, consumed
// by tcgen05.mma.kind::tf32 as it lands -- no register staging, no conversion instructions.  (On this part the
// MN-major form of kind::tf32 returned zeros in the selftest, hence the point-contiguous layout.)  (The producer
// rounds the values to TF32 with round-to-nearest when it stores them, so the tensor core's truncation of the low
// mantissa bits is exact.)  The accumulators stay in TMEM for the whole point range of the CTA and are flushed
// once per instance segment with vector reductions (red.global.add.v4.f32).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-5 accumulator flush.  Grid: CTAs are shared out over the groups
// (one group = one weight matrix) in proportion to their work per tile; a CTA owns a contiguous tile range.
#include "oi_internal.cuh"
#include "oi_tc.cuh"
#include "oi_wgrad.cuh"

namespace oi {

namespace {

constexpr int kWgThreads = 192;
constexpr int kHalfBytes = 32768;            // one operand, 64 points: 2 blocks of [128 channels][32 points] fp32
constexpr int kAuxBytes = 4096;              // N = 16 operand: 2 blocks of [16 rows][32 points]; rows 0..3 = aux
constexpr int kWgStages = 3;
// instruction descriptors: D fp32, A = B = tf32, both K-major; M = 128, N = 128 / 16
__host__ __device__ constexpr uint32_t idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

struct __align__(1024) WgSmem {
  unsigned char x[kWgStages][kHalfBytes];
  unsigned char y[kWgStages][kHalfBytes];
  unsigned char aux[kWgStages][kAuxBytes];
  unsigned long long full[kWgStages], empty[kWgStages], acc_done, acc_free;
  uint32_t tmem_base;
};
static_assert(sizeof(WgSmem) <= 227 * 1024, "WgSmem exceeds the per-CTA limit");

__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_ss_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// F16 = false: TF32 slabs, one pipeline stage = 64 points of an operand pair (two per tile);
// F16 = true : fp16 slabs (oi_wgrad.cuh), one stage = all 128 points of a tile; same 32 KB per operand and stage, same
//              eight MMAs per stage (K = 8 tf32 / K = 16 fp16 elements = 32 bytes of a 128-byte row each).
template <bool F16>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float out_scale = 1.0f;
  if (a.ctl != nullptr) {
    const BwdMode mode = bwd_mode(a.ctl, a.flags, a.ctl[kCtlGuardSnapshot]);
    if (mode.f16 != F16) return;   // the other variant of this launch pair does the work
    if (F16) out_scale = pow2i(mode.e_ref + kF16AdjShift);
  } else if (F16) {
    return;
  }
  constexpr int kHalves = F16 ? 1 : 2;
  constexpr uint32_t kIdMain = F16 ? tc::make_idesc_f16(128, 128) : idesc_tf32(128);
  constexpr uint32_t kIdAux = F16 ? tc::make_idesc_f16(128, 16) : idesc_tf32(16);
  int group = 0;
  while (group + 1 < a.n_groups && (int)blockIdx.x >= a.groups[group + 1].cta0) ++group;
  const WgGroup& G = a.groups[group];
  const int split = (int)blockIdx.x - G.cta0;
  const int t_begin = (int)((long long)a.n_tiles * split / G.n_splits);
  const int t_end = (int)((long long)a.n_tiles * (split + 1) / G.n_splits);
  const int n_pairs = G.n_pairs;
  const bool use_aux = G.use_aux != 0;

  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_done, 1);
    mbar_init(&sm.acc_free, 128);
    mbar_fence_init();
  }
  for (int i = tid; i < kWgStages * kAuxBytes / 16; i += kWgThreads)
    reinterpret_cast<float4*>(&sm.aux[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    tc::tmem_alloc(&sm.tmem_base, 256);
    tc::tmem_relinquish();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const unsigned char* slabs = reinterpret_cast<const unsigned char*>(a.slabs) +
                                     (size_t)tile * a.slabs_per_tile * (2 * kHalfBytes);
        // aux: TF32 4 x 512 B per tile (32-point blocks of [4 rows][32 points] fp32); fp16 2 x 512 B (64-point blocks
        // of [4 rows][64 points] fp16) in the first half of the same 2 KB
        const unsigned char* aux = reinterpret_cast<const unsigned char*>(a.aux) + (size_t)tile * 2048;
        for (int half = 0; half < kHalves; ++half) {
          for (int p = 0; p < n_pairs; ++p, ++it) {
            const int stage = it % kWgStages;
            if (it >= kWgStages) mbar_wait_sleep(&sm.empty[stage], ((it / kWgStages) - 1) & 1, 1000u);
            const bool with_aux = (p == 0) && use_aux;
            mbar_expect_tx(&sm.full[stage], 2 * kHalfBytes + (with_aux ? 1024 : 0));
            const size_t xo = F16 ? slab16_offset(G.pairs[p].x_slab) : ((size_t)G.pairs[p].x_slab * 2 + half) * kHalfBytes;
            const size_t yo = F16 ? slab16_offset(G.pairs[p].y_slab) : ((size_t)G.pairs[p].y_slab * 2 + half) * kHalfBytes;
            tma_bulk_g2s(sm.x[stage], slabs + xo, kHalfBytes, &sm.full[stage]);
            tma_bulk_g2s(sm.y[stage], slabs + yo, kHalfBytes, &sm.full[stage]);
            if (with_aux) {   // rows 0..3 of each block; rows 4..15 stay zero
              tma_bulk_g2s(sm.aux[stage], aux + (half * 2) * 512, 512, &sm.full[stage]);
              tma_bulk_g2s(sm.aux[stage] + 2048, aux + (half * 2 + 1) * 512, 512, &sm.full[stage]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int it = 0, seg = 0;
      int cur_inst = (t_begin < t_end) ? (a.tile0 + t_begin) / a.tiles_per_inst : 0;
      bool fresh = true;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int inst = (a.tile0 + tile) / a.tiles_per_inst;
        if (inst != cur_inst) {
          // instance boundary: hand the accumulators to the flush warps, wait until they have been read
          tc::mma_commit(&sm.acc_done);
          mbar_wait_sleep(&sm.acc_free, seg & 1, 1000u);
          tc::fence_after_thread_sync();
          ++seg;
          cur_inst = inst;
          fresh = true;
        }
        for (int half = 0; half < kHalves; ++half) {
          for (int p = 0; p < n_pairs; ++p, ++it) {
            const int stage = it % kWgStages;
            mbar_wait_sleep(&sm.full[stage], (it / kWgStages) & 1, 1000u);
            tc::fence_after_thread_sync();
            const uint32_t xb = smem_u32(sm.x[stage]), yb = smem_u32(sm.y[stage]), ab = smem_u32(sm.aux[stage]);
            const bool with_aux = (p == 0) && use_aux;
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 8 (tf32) / 16 (fp16) points per instruction: block k / 4, 32-byte step k % 4
              const uint32_t xo = (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u;
              const uint64_t xd = tc::make_desc_k_sw128(xb + xo);
              const uint64_t yd = tc::make_desc_k_sw128(yb + xo);
              const uint64_t ad = tc::make_desc_k_sw128(ab + (uint32_t)(k >> 2) * 2048u + (uint32_t)(k & 3) * 32u);
              const uint32_t acc = (fresh && k == 0) ? 0u : 1u;
              if (F16) {
                mma_ss_f16(tmem_base, xd, yd, kIdMain, acc);
                if (with_aux) mma_ss_f16(tmem_base + 128, xd, ad, kIdAux, acc);
              } else {
                mma_ss_tf32(tmem_base, xd, yd, kIdMain, acc);
                if (with_aux) mma_ss_tf32(tmem_base + 128, xd, ad, kIdAux, acc);
              }
            }
            fresh = false;
            tc::mma_commit(&sm.empty[stage]);
          }
        }
      }
      if (t_begin < t_end) tc::mma_commit(&sm.acc_done);
    }
  } else if (t_begin < t_end && n_pairs > 0) {
    // ===================== accumulator flush: warp w reads TMEM lanes 32 (w % 4) .. +31 =====================
    const int lq = warp & 3;
    const int i = lq * 32 + lane;   // row of the accumulator = channel of X
    const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16);
    int seg = 0;
    int cur_inst = (a.tile0 + t_begin) / a.tiles_per_inst;
    for (int tile = t_begin; tile <= t_end; ++tile) {
      const int inst = (tile < t_end) ? (a.tile0 + tile) / a.tiles_per_inst : -1;
      if (inst == cur_inst) continue;
      mbar_wait_sleep(&sm.acc_done, seg & 1, 1000u);
      tc::fence_after_thread_sync();
      float* row = G.out + (size_t)cur_inst * G.out_inst_stride + (size_t)i * G.out_ld;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float u[32];
        tc::tmem_ld32(taddr + c * 32, u);
        if ((G.out_ld & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + c * 32 + j),
                         "f"(u[j] * out_scale), "f"(u[j + 1] * out_scale), "f"(u[j + 2] * out_scale),
                         "f"(u[j + 3] * out_scale)
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(row + c * 32 + j, u[j] * out_scale);
        }
      }
      if (use_aux) {
        uint32_t r[16];
        tc::tmem_ld16_async(taddr + 128, r);
        tc::wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float* dst = G.aux_out[c];
          if (dst != nullptr)
            atomicAdd(dst + (size_t)cur_inst * G.aux_inst_stride[c] + (size_t)i * G.aux_ch_stride[c],
                      __uint_as_float(r[c]) * ((F16 && c < 3) ? out_scale * (float)(1 << kF16NormalShift) : out_scale));
        }
      }
      tc::fence_before_thread_sync();
      mbar_arrive(&sm.acc_free);
      ++seg;
      cur_inst = inst;
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace

int launch_wgrad_tc(const WgArgs& a_in, cudaStream_t st) {
  WgArgs a = a_in;
  if (a.n_groups < 1 || a.n_groups > WG_MAX_GROUPS || a.n_ctas < a.n_groups)
    return set_error(OI_ERR_INVALID_ARGUMENT, "wgrad: bad grouping (%d groups, %d CTAs)", a.n_groups, a.n_ctas);
  // share the CTAs out in proportion to the groups' work per tile; never more splits than tiles
  int wsum = 0;
  for (int g = 0; g < a.n_groups; ++g) wsum += a.groups[g].weight > 0 ? a.groups[g].weight : 1;
  int cta = 0;
  for (int g = 0; g < a.n_groups; ++g) {
    const int wg = a.groups[g].weight > 0 ? a.groups[g].weight : 1;
    int n = a.n_ctas * wg / wsum;
    if (n < 1) n = 1;
    if (n > a.n_tiles) n = a.n_tiles;
    a.groups[g].cta0 = cta;
    a.groups[g].n_splits = n;
    cta += n;
  }
  OI_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(WgSmem)));
  wgrad_tc_kernel<false><<<cta, kWgThreads, sizeof(WgSmem), st>>>(a);
  OI_CHECK_CUDA(cudaGetLastError());
  if (a.ctl != nullptr) {   // the fp16-slab variant; one of the two exits at once
    OI_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(WgSmem)));
    wgrad_tc_kernel<true><<<cta, kWgThreads, sizeof(WgSmem), st>>>(a);
    OI_CHECK_CUDA(cudaGetLastError());
  }
  return OI_OK;
}

}  // namespace oi
