// Interface between the two kernels of the tensor-core backward: the per-tile slab scratch written by
// bwd_tc_kernel (oi_render_bwd_tc.cu) and the job tables of wgrad_tc_kernel (oi_wgrad_tc.cu).
#pragma once
#include "oi_internal.cuh"

namespace oi {

// ---- slab ids: every slab is [32 channel-quads][128 points] float4 (64 KB) of one 128-point tile ----
constexpr int kSlabArg = 0;    // ARG[l], l = 0..7 : FiLM pre-activation a_l = gamma_l u_l + beta_l  (h_{l+1} = sin a_l)
constexpr int kSlabT = 8;      // T[l],   l = 0..7 : t_l of the reverse sweep (t_0 feeds dW_0)
constexpr int kSlabGB = 16;    // GB[l],  l = 1..7 : g_bar_l
constexpr int kSlabUB = 24;    // UB[l],  l = 0..7 : u_bar_l
constexpr int kSlabDG = 32;    // DG[l],  l = 0..7 : a_bar_l u_l + c_bar_l cos a_l  (column sum = dL/dgamma_l)
constexpr int kSlabUBC = 40;   // u_bar of the colour layer
constexpr int kSlabArgC = 41;  // pre-activation of the colour layer
constexpr int kSlabDGC = 42;   // a_bar_c u_c
constexpr int kSlabDWS = 43;   // t_bar_{D-1} c_{D-1}   (column sum = second part of d w_sigma)
constexpr int kSlabsPerTile = 44;
constexpr int kSlabFloats = 128 * 128;
// ---- aux rows: [16][128 points] floats per tile ----
constexpr int kAuxX = 0;       // sample position (3)
constexpr int kAuxN = 3;       // normal (3)
constexpr int kAuxNB = 6;      // normal_bar, total (3)
constexpr int kAuxZB = 9;      // adjoint of the rgb pre-activation (3)
constexpr int kAuxSB = 12;     // sdf_bar

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
};
struct WgArgs {
  int n_tiles, tiles_per_inst, n_groups, n_splits, slabs_per_tile;
  int tile0;           // global index of slab tile 0 (instance of slab tile t = (tile0 + t) / tiles_per_inst)
  const float* slabs;  // [n_tiles][slabs_per_tile][32][128] float4
  const float* aux;    // [n_tiles][16][128]
  WgGroup groups[WG_MAX_GROUPS];
};

int launch_wgrad_tc(const WgArgs& a, cudaStream_t st);

}  // namespace oi
