// tcgen05 (5th-gen tensor core) core of the fused SDF render kernel (OI_IMPL_TCGEN05) -- placeholder until
// the UMMA path lands; reports OI_ERR_UNSUPPORTED so that callers fail loudly instead of silently degrading.
#include "oi_internal.cuh"

namespace oi {

size_t render_tc_scratch_floats(int depth, int* n_ctas, int n_tiles) {
  return render_ffma_scratch_floats(depth, n_ctas, n_tiles);
}

int launch_render_tc(const RenderKArgs&, cudaStream_t) {
  return set_error(OI_ERR_UNSUPPORTED, "OI_IMPL_TCGEN05 is not built into this library");
}

}  // namespace oi
