// Point-contraction kernel of the tensor-core backward (second kernel of oi_render_backward, OI_IMPL_TCGEN05):
//     dW[i][j] = sum over sample points m of X[m][i] * Y[m][j]
// for the 128x128 weight matrices of the SDF / colour networks, and the per-channel sums over points that give
// the bias, FiLM (gamma, beta), head and layer-0 gradients.  The first kernel (oi_render_bwd_tc.cu) leaves, per
// 128-point tile, its per-point quantities as fp32 "slabs" [32 channel-quads][128 points] float4 in global
// scratch; this kernel streams them once:
//   * 8 loader warps read a 64-point step of an (X, Y) slab pair (coalesced 16-byte loads), apply the operand
//     transform (identity or sin), split every value into two bf16 terms (x = hi + lo, |err| <= 2^-17 |x|; bf16
//     because adjoints have no bounded range) and write the four images {X,Y} x {hi,lo} into shared memory in the
//     canonical MN-major SWIZZLE_128B layout ([point][64 channels] rows of 128 bytes, two channel blocks);
//   * one thread issues tcgen05.mma.kind::f16 (bf16 inputs, fp32 accumulation in TMEM) with BOTH operands MN-major:
//     D[i][j] += Xhi^T Yhi + Xlo^T Yhi + Xhi^T Ylo, K = 16 points per instruction; the accumulator stays in TMEM for
//     the whole point range of the CTA and is flushed once with vector reductions (red.global.add.v4.f32);
//   * the loader threads own fixed channels, so the per-channel sums are plain register accumulations.
// Grid: n_groups x n_splits CTAs; group = which matrix / which set of column sums, split = contiguous tile range.
#include <cuda_bf16.h>

#include "oi_internal.cuh"
#include "oi_tc.cuh"
#include "oi_wgrad.cuh"

namespace oi {

namespace {

constexpr int kWgThreads = 288;        // 8 loader/epilogue warps + 1 MMA warp
constexpr int kLoaders = 256;
constexpr int kStepPts = 64;           // points per pipeline stage
constexpr int kImgBytes = 16384;       // one operand image: [2 channel blocks][64 points][128 B]
constexpr int kStageBytes = 4 * kImgBytes;  // Xhi, Xlo, Yhi, Ylo
constexpr int kWgStages = 3;
// instruction descriptor: D fp32, A = B = bf16, both MN-major, M = N = 128
constexpr uint32_t kIdescWg = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) |
                              ((128u >> 4) << 24);

struct __align__(1024) WgSmem {
  unsigned char img[kWgStages][kStageBytes];
  unsigned long long full[kWgStages], empty[kWgStages], acc_done;
  uint32_t tmem_base;
};
static_assert(sizeof(WgSmem) <= 227 * 1024, "WgSmem exceeds the per-CTA limit");

// MN-major SWIZZLE_128B operand: 64 channels (128 B) contiguous per point row, 8-row groups 1024 B apart
// (stride byte offset), the second 64-channel block 8192 B further (leading byte offset).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Writes the 4 channels of quad q at point row r (0..63 inside the stage) into a hi and a lo image.
__device__ __forceinline__ void store_quad(unsigned char* img_hi, unsigned char* img_lo, int q, int r, float4 v) {
  uint32_t h0, l0, h1, l1;
  split_bf16x2(v.x, v.y, h0, l0);
  split_bf16x2(v.z, v.w, h1, l1);
  const int block = q >> 4, chunk = ((q & 15) >> 1) ^ (r & 7);
  const int off = block * 8192 + r * 128 + chunk * 16 + (q & 1) * 8;
  *reinterpret_cast<uint2*>(img_hi + off) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(img_lo + off) = make_uint2(l0, l1);
}

__device__ __forceinline__ float4 tf_apply(float4 v, int tf) {
  if (tf == WG_TF_SIN) return make_float4(__sinf(v.x), __sinf(v.y), __sinf(v.z), __sinf(v.w));
  return v;
}

struct ColAcc {
  float v[4][4];  // [local quad][channel in quad]
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) v[a][b] = 0.f;
  }
  __device__ __forceinline__ void add(int lq, float4 x, float m) {
    v[lq][0] = fmaf(x.x, m, v[lq][0]);
    v[lq][1] = fmaf(x.y, m, v[lq][1]);
    v[lq][2] = fmaf(x.z, m, v[lq][2]);
    v[lq][3] = fmaf(x.w, m, v[lq][3]);
  }
  // sum over the 32 lanes (points) and add to dst[channel * stride]; this warp owns channels 16*warp..16*warp+15
  __device__ __forceinline__ void flush(float* dst, int stride, int warp, int lane) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float s = v[a][b];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0 && dst != nullptr) atomicAdd(dst + (size_t)(16 * warp + 4 * a + b) * stride, s);
        v[a][b] = 0.f;
      }
  }
};

constexpr int kMaxCol = 5;  // column-sum accumulators per thread (80 registers)

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int group = 0;
  while (group + 1 < a.n_groups && (int)blockIdx.x >= a.groups[group + 1].cta0) ++group;
  const WgGroup& G = a.groups[group];
  const int split = (int)blockIdx.x - G.cta0;
  const int t_begin = (int)((long long)a.n_tiles * split / G.n_splits);
  const int t_end = (int)((long long)a.n_tiles * (split + 1) / G.n_splits);
  const int n_pairs = G.n_pairs;                    // MMA operand pairs per 64-point step (0, 1 or 2)
  const int steps_total = (t_end - t_begin) * 2 * n_pairs;

  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&sm.full[s], kLoaders);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 8) {
    tc::tmem_alloc(&sm.tmem_base, 128);
    tc::tmem_relinquish();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (int it = 0; it < steps_total; ++it) {
        const int stage = it % kWgStages;
        mbar_wait_sleep(&sm.full[stage], (it / kWgStages) & 1, 2000u);
        tc::fence_after_thread_sync();
        const uint32_t base = smem_u32(sm.img[stage]);
#pragma unroll
        for (int k = 0; k < kStepPts / 16; ++k) {
          const uint32_t ko = (uint32_t)k * 16u * 128u;  // 16 point rows
          const uint64_t xhi = make_desc_mn_sw128(base + 0 * kImgBytes + ko);
          const uint64_t xlo = make_desc_mn_sw128(base + 1 * kImgBytes + ko);
          const uint64_t yhi = make_desc_mn_sw128(base + 2 * kImgBytes + ko);
          const uint64_t ylo = make_desc_mn_sw128(base + 3 * kImgBytes + ko);
          tc::mma_ss(tmem_base, xhi, yhi, kIdescWg, (it > 0 || k > 0) ? 1u : 0u);
          tc::mma_ss(tmem_base, xlo, yhi, kIdescWg, 1u);
          tc::mma_ss(tmem_base, xhi, ylo, kIdescWg, 1u);
        }
        tc::mma_commit(&sm.empty[stage]);
      }
      tc::mma_commit(&sm.acc_done);
    }
  } else {
    // ===================== loaders (and, at the end, the accumulator flush) =====================
    ColAcc col[kMaxCol];
#pragma unroll
    for (int c = 0; c < kMaxCol; ++c) col[c].zero();
    int it = 0;
    int cur_inst = (t_begin < t_end) ? (a.tile0 + t_begin) / a.tiles_per_inst : 0;

    auto flush_cols = [&](int inst) {
#pragma unroll
      for (int c = 0; c < kMaxCol; ++c) {
        if (c < G.n_cols) {
          const WgCol& C = G.cols[c];
          col[c].flush(C.out + (size_t)inst * C.inst_stride, C.ch_stride, warp, lane);
        }
      }
    };

    for (int tile = t_begin; tile < t_end; ++tile) {
      const int inst = (a.tile0 + tile) / a.tiles_per_inst;
      if (inst != cur_inst) {
        flush_cols(cur_inst);
        cur_inst = inst;
      }
      const float4* slabs = reinterpret_cast<const float4*>(a.slabs) + (size_t)tile * a.slabs_per_tile * 4096;
      const float* aux = a.aux + (size_t)tile * 16 * 128;
      for (int half = 0; half < 2; ++half) {
        const int m0 = half * 64 + lane, m1 = m0 + 32;
        // ---- MMA operand pairs (with the column sums that ride on the X / Y operand)
        for (int p = 0; p < n_pairs; ++p, ++it) {
          const WgPair& P = G.pairs[p];
          const int stage = it % kWgStages;
          if (it >= kWgStages) mbar_wait_sleep(&sm.empty[stage], ((it / kWgStages) - 1) & 1, 2000u);
          unsigned char* base = sm.img[stage];
          const float4* xs = slabs + (size_t)P.x_slab * 4096;
          const float4* ys = slabs + (size_t)P.y_slab * 4096;
          float4 xv[8], yv[8];
#pragma unroll
          for (int lq = 0; lq < 4; ++lq) {
            const int q = warp * 4 + lq;
            xv[2 * lq] = xs[q * 128 + m0];
            xv[2 * lq + 1] = xs[q * 128 + m1];
            yv[2 * lq] = ys[q * 128 + m0];
            yv[2 * lq + 1] = ys[q * 128 + m1];
          }
#pragma unroll
          for (int lq = 0; lq < 4; ++lq) {
            const int q = warp * 4 + lq;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = lane + 32 * e;
              const float4 x = tf_apply(xv[2 * lq + e], P.x_tf), y = tf_apply(yv[2 * lq + e], P.y_tf);
              store_quad(base + 0 * kImgBytes, base + 1 * kImgBytes, q, r, x);
              store_quad(base + 2 * kImgBytes, base + 3 * kImgBytes, q, r, y);
#pragma unroll
              for (int c = 0; c < kMaxCol; ++c) {
                if (c < G.n_cols) {
                  const WgCol& C = G.cols[c];
                  if (C.src == WG_SRC_PAIR_X + 2 * p || C.src == WG_SRC_PAIR_Y + 2 * p) {
                    const float mult = (C.mult < 0) ? 1.0f : aux[C.mult * 128 + half * 64 + r];
                    col[c].add(lq, (C.src == WG_SRC_PAIR_X + 2 * p) ? x : y, mult);
                  }
                }
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> UMMA (async proxy) reads
          mbar_arrive(&sm.full[stage]);
        }
        // ---- stand-alone column sums
#pragma unroll
        for (int c = 0; c < kMaxCol; ++c) {
          if (c < G.n_cols) {
            const WgCol& C = G.cols[c];
            if (C.src == WG_SRC_SLAB) {
              const float4* ss = slabs + (size_t)C.slab * 4096;
#pragma unroll
              for (int lq = 0; lq < 4; ++lq) {
                const int q = warp * 4 + lq;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int r = lane + 32 * e;
                  const float4 x = tf_apply(ss[q * 128 + half * 64 + r], C.tf);
                  const float mult = (C.mult < 0) ? 1.0f : aux[C.mult * 128 + half * 64 + r];
                  col[c].add(lq, x, mult);
                }
              }
            }
          }
        }
      }
    }
    if (t_begin < t_end) flush_cols(cur_inst);

    // ---- flush the TMEM accumulator: warp w reads lanes 32 (w % 4) .. +31, columns 64 (w / 4) .. +63
    if (n_pairs > 0 && steps_total > 0) {
      mbar_wait_sleep(&sm.acc_done, 0u, 2000u);
      tc::fence_after_thread_sync();
      const int i = (warp & 3) * 32 + lane;
      const int j0 = (warp >> 2) * 64;
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + j0;
      float* row = G.out + (size_t)i * G.out_ld + j0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float u[32];
        tc::tmem_ld32(taddr + c * 32, u);
        if ((G.out_ld & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + c * 32 + j), "f"(u[j]),
                         "f"(u[j + 1]), "f"(u[j + 2]), "f"(u[j + 3])
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(row + c * 32 + j, u[j]);
        }
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_base, 128);
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
  OI_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(WgSmem)));
  wgrad_tc_kernel<<<cta, kWgThreads, sizeof(WgSmem), st>>>(a);
  OI_CHECK_CUDA(cudaGetLastError());
  return OI_OK;
}

}  // namespace oi
