// Interface between the two kernels of the tensor-core backward: the per-tile slab scratch written by
// bwd_tc_kernel (oi_render_bwd_tc.cu) and the job tables of wgrad_tc_kernel (oi_wgrad_tc.cu).
#pragma once
#include "oi_internal.cuh"

namespace oi {

// ---- slab ids: every slab is [32 channel-quads][128 points] float4 (64 KB) of one 128-point tile ----
constexpr int kSlabArg = 0;    // ARG[l], l = 0..7 : FiLM pre-activation a_l = gamma_l u_l + beta_l  (h_{l+1} = sin a_l)
constexpr int kSlabT = 7;      // T[l],   l = 1..7 (7 + l) : t_l of the reverse sweep
constexpr int kSlabGB = 14;    // GB[l],  l = 1..7 (14 + l): g_bar_l
constexpr int kSlabUB = 21;    // UB[l],  l = 1..7 (21 + l): u_bar_l
constexpr int kSlabUBC = 29;   // u_bar of the colour layer
constexpr int kSlabsPerTile = 30;
constexpr int kSlabFloats = 128 * 128;
// ---- aux rows: [16][128 points] floats per tile ----
constexpr int kAuxN = 0;       // normal (3)
constexpr int kAuxSB = 3;      // sdf_bar

enum { WG_TF_RAW = 0, WG_TF_SIN = 1 };
enum { WG_SRC_SLAB = 0, WG_SRC_PAIR_X = 1, WG_SRC_PAIR_Y = 2 };  // pair p: X = 1 + 2p, Y = 2 + 2p
constexpr int WG_MAX_GROUPS = 12;

struct WgPair {
  int x_slab, y_slab, x_tf, y_tf;
};
struct WgCol {
  int src;          // WG_SRC_SLAB, or the X / Y operand of pair p (already loaded and transformed)
  int slab, tf;     // for WG_SRC_SLAB
  int mult;         // aux row multiplying every point, or -1
  float* out;       // out[inst * inst_stride + channel * ch_stride] += sum over points
  int inst_stride, ch_stride;
};
struct WgGroup {
  int n_pairs, n_cols;
  WgPair pairs[2];
  WgCol cols[5];
  float* out;       // [128][out_ld] accumulated with reductions; NULL when n_pairs == 0
  int out_ld;
  int weight;       // relative work per tile (slab reads); CTAs are shared out in proportion
  int cta0, n_splits;  // filled by launch_wgrad_tc: CTAs [cta0, cta0 + n_splits) work on this group
};
struct WgArgs {
  int n_tiles, tiles_per_inst, n_groups, n_ctas, slabs_per_tile;
  int tile0;           // global index of slab tile 0 (instance of slab tile t = (tile0 + t) / tiles_per_inst)
  const float* slabs;  // [n_tiles][slabs_per_tile][32][128] float4
  const float* aux;    // [n_tiles][16][128]
  WgGroup groups[WG_MAX_GROUPS];
};

int launch_wgrad_tc(const WgArgs& a, cudaStream_t st);

}  // namespace oi
