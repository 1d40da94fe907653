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
  WgGroup groups[WG_MAX_GROUPS];
};

int launch_wgrad_tc(const WgArgs& a, cudaStream_t st);

}  // namespace oi
