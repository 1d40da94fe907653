// Interface between the two kernels of the tensor-core backward: the per-tile slab scratch written by
// bwd_tc_kernel (oi_render_bwd_tc.cu) and the job tables of wgrad_tc_kernel (oi_wgrad_tc.cu).
#pragma once
#include "oi_internal.cuh"

namespace oi {

// ---- slab ids: 64 KB per slab and 128-point tile.
//      ARG slabs (read back by the sweep kernel): [32 channel-quads][128 points] float4.
//      Operand slabs (H, T, GB, UB, UBC; values rounded to TF32 by the producer): four 32-point blocks, each the
//      canonical K-major SWIZZLE_128B tf32 UMMA image [128 channels][32 points] (128-byte rows, 16-byte chunks
//      XOR-permuted by channel & 7) -- see oi_wgrad_tc.cu. ----
constexpr int kSlabArg = 0;    // ARG[l], l = 0..7 : FiLM pre-activation a_l = gamma_l u_l + beta_l  (h_{l+1} = sin a_l)
constexpr int kSlabT = 7;      // T[l],   l = 1..7 (7 + l) : t_l of the reverse sweep
constexpr int kSlabGB = 14;    // GB[l],  l = 1..7 (14 + l): g_bar_l
constexpr int kSlabUB = 21;    // UB[l],  l = 1..7 (21 + l): u_bar_l
constexpr int kSlabUBC = 29;   // u_bar of the colour layer
constexpr int kSlabH = 29;     // H[l],   l = 1..8 (29 + l): h_l = sin a_{l-1}
constexpr int kSlabsPerTile = 38;
constexpr int kSlabFloats = 128 * 128;
// ---- aux: per tile four 32-point blocks of [4 rows][32 points] fp32 (same swizzle): rows = normal.xyz, 1 ----

// ---- 16-bit operand slabs (OI_BWD_FLAG_* / automatic): the 30 operand slabs are stored as fp16 instead -- 32 KB per
//      slab: two 64-point blocks of the K-major SWIZZLE_128B fp16 image [128 channels][64 points]; the ARG slabs stay
//      fp32; slab s >= kSlabOperand0 lives at 8 * 64 KB + (s - 8) * 32 KB inside the tile's (unchanged) region.
//      fp16 has TF32's 10 mantissa bits but 5 exponent bits, so every operand is scaled by an exact power of two:
//        adjoint-type operands (UB, GB, UBC)   x 2^-e_m              e_m = exponent of max |adj[m][0..6]| of the point
//        forward-type operands (H, T, aux)     x 2^(e_m - e_ref)     e_ref = (exponent of the global max |adj|) - 8
//      => every product carries 2^-e_ref, undone when the contraction kernel flushes its accumulators.  The adjoint
//      operands are then bounded by the network's Jacobians alone, the forward operands by 2^8 |value|; points whose
//      adjoint is more than 2^18 below the global maximum lose precision gradually.  bwd_mode() (below) selects the
//      path per call ON THE DEVICE from the adjoint statistics adj_stats_kernel leaves in the control block: fp16 when
//      those points carry < 2^-12 of the total adjoint mass, TF32 otherwise; both kernel variants are launched and the
//      one not selected exits at once (no host round trip).
constexpr int kSlabOperand0 = 8;
constexpr int kSlabBytes = 65536, kSlab16Bytes = 32768;
__host__ __device__ constexpr size_t slab16_offset(int s) {
  return s < kSlabOperand0 ? (size_t)s * kSlabBytes : (size_t)kSlabOperand0 * kSlabBytes + (size_t)(s - kSlabOperand0) * kSlab16Bytes;
}
constexpr int OI_BWD_FLAG_FORCE_TF32 = 32, OI_BWD_FLAG_FORCE_F16 = 64;
// placement inside fp16's range (measured with -DOI_BWD_RANGE_STATS=1, tools_bwd_range.py: at shifts 10 / 0 the largest
// forward-type operand was 1 000 - 5 200 and the largest adjoint-type operand 180 - 2 000 over the bench networks, a
// random initialisation and two losses): forward-type x 2^(e_m - e_max + kF16RefShift), adjoint-type x 2^(-e_m -
// kF16AdjShift); the contraction kernel multiplies its result by 2^(e_max - kF16RefShift + kF16AdjShift).  The adjoint
// side is tight at its LOW end (small Jacobians from the colour adjoint to the SDF layers put entries near fp16's
// subnormals: with kF16AdjShift = 2 a single-ray loss lost 2.5x in accuracy), so it stays 0 (headroom 32x - 350x).
constexpr int kF16RefShift = 8, kF16AdjShift = 0, kF16LowShift = 18, kF16MassShift = 12;

// control block of one backward call (32-bit words, zeroed by launch_bwd_tail):
//   [0] relax count   [1] bits of max |adj|   [2,3] u64 total adjoint mass   [4,5] u64 mass of the low points
struct BwdMode {
  bool f16;
  int e_ref;
};
// Guard against fp16 overflow, one word of the workspace that no call resets (the owner zeroes the workspace once):
//   0 (unknown)   the call keeps the TF32 operands and PROBES: the sweep tracks the largest |x| of what the fp16 operands
//                 would be (every operand of every tile); if
//                 none reaches half of fp16's largest number (|x| >= 32768) the finalize kernel marks the workspace
//                 kF16Safe
//   kF16Safe      bwd_mode() may choose fp16; the fp16 sweep keeps a running max |x| of the operands it writes
//   kF16Unsafe    set by any sweep that sees |x| >= 8192 (8x below saturation): TF32 from the next
//                 call on, for the life of the workspace
// Every kernel of a call decides on the snapshot adj_stats_kernel takes of that word (control word kCtlGuardSnapshot;
// 0 = unknown when the format is forced and no snapshot is taken -- forced formats ignore the guard).
constexpr unsigned int kF16Safe = 0xF16C0DE5u, kF16Unsafe = 0xF16D15ABu;
constexpr int kF16NormalShift = 4;   // the normal columns of the aux operand carry another 2^-4 (|grad sdf| grows with the
                                     // SDF head: it is the first forward-type operand to reach the guard limit)
constexpr float kF16GuardLimit = 8192.0f;   // an eighth of fp16's largest number; measured operands stay below 2 100
constexpr int kCtlGuardSnapshot = 9;
__host__ __device__ __forceinline__ BwdMode bwd_mode(const unsigned int* ctl, int flags, unsigned int guard = kF16Safe) {
  const unsigned int mb = ctl[1];
  const int ex = (int)((mb >> 23) & 0xFFu);
  const unsigned long long tot = (unsigned long long)ctl[2] | ((unsigned long long)ctl[3] << 32);
  const unsigned long long low = (unsigned long long)ctl[4] | ((unsigned long long)ctl[5] << 32);
  BwdMode m;
  m.f16 = ex > 0 && ex < 255 && low <= (tot >> kF16MassShift) && guard == kF16Safe;
  if (flags & OI_BWD_FLAG_FORCE_TF32) m.f16 = false;
  if ((flags & OI_BWD_FLAG_FORCE_F16) && ex > 0 && ex < 255) m.f16 = true;
  m.e_ref = ex - 127 - kF16RefShift;
  return m;
}
// 2^e as a float, e clamped to the normal range (e < -126 -> 0)
__device__ __forceinline__ float pow2i(int e) {
  return e < -126 ? 0.f : __uint_as_float((unsigned int)((e > 127 ? 127 : e) + 127) << 23);
}
// exponent of max |adj| of a point (its first 7 adjoint components); zero / denormal -> -126
__device__ __forceinline__ int adj_exponent(float4 q0, float4 q1) {
  const float a = fmaxf(fmaxf(fmaxf(fabsf(q0.x), fabsf(q0.y)), fmaxf(fabsf(q0.z), fabsf(q0.w))),
                        fmaxf(fmaxf(fabsf(q1.x), fabsf(q1.y)), fabsf(q1.z)));
  const int ex = (int)((__float_as_uint(a) >> 23) & 0xFFu);
  return ex == 0 ? -126 : ex - 127;
}

constexpr int WG_MAX_GROUPS = 12;

struct WgPair {
  int x_slab, y_slab;
};
struct WgGroup {
  int n_pairs;      // operand pairs accumulated into the same matrix (1 or 2)
  WgPair pairs[2];
  float* out;       // [128][out_ld] per instance (instance i at out + i * out_inst_stride), accumulated with reductions
  int out_ld;
  int out_inst_stride;   // floats; 0 = one matrix shared by all instances
  // narrow products of the X operand of pair 0 with the aux columns: aux_out[c][inst * inst_stride + i * ch_stride]
  // += sum_m X[m][i] aux[m][c]; NULL entries are skipped, aux_out == all NULL disables the extra MMA
  float* aux_out[4];
  int aux_inst_stride[4], aux_ch_stride[4];
  int use_aux;
  int weight;       // relative work per tile (slab reads); CTAs are shared out in proportion
  int cta0, n_splits;  // filled by launch_wgrad_tc: CTAs [cta0, cta0 + n_splits) work on this group
};
struct WgArgs {
  int n_tiles, tiles_per_inst, n_groups, n_ctas, slabs_per_tile;
  int tile0;           // global index of slab tile 0 (instance of slab tile t = (tile0 + t) / tiles_per_inst)
  const float* slabs;  // [n_tiles][slabs_per_tile][4 blocks][128 channels][32 points]
  const float* aux;    // [n_tiles][4 blocks][4 rows][32 points]
  const unsigned int* ctl;   // control block of the backward call (bwd_mode); NULL = TF32 slabs unconditionally
  int flags;
  WgGroup groups[WG_MAX_GROUPS];
};

int launch_wgrad_tc(const WgArgs& a, cudaStream_t st);

}  // namespace oi
