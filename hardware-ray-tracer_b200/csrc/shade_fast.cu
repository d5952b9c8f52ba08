// BRT_RENDER_FAST_SHADING (opt-in): the shade kernels of shade_kernels.cuh compiled WITHOUT the arithmetic contract — FMA contraction,
// approximate division and (reciprocal) square root (csrc/Makefile builds this file with -fmad=true --use_fast_math). Traversal, ray
// generation and accumulation stay exact, so primary ids are unchanged; radiance differs from the oracle in the last bits and stays
// inside the 1e-3 relative RMSE bar of the north star (tests/test_gpu_parity.py::test_fast_shading). Measured in profiles/r2_shade.md.
#define BRT_SHADE_PRIMARY k_shade_primary_fast
#define BRT_SHADE_BOUNCE k_shade_fast
#include "shade_kernels.cuh"

namespace brt {

void launch_shade_fast(const ShadeParams& sp, uint32_t grid, bool primary, cudaStream_t stream) {
  if (primary) k_shade_primary_fast<<<grid, 128, 0, stream>>>(sp);
  else k_shade_fast<BRT_SHADE_WINDOW><<<grid, 128, 0, stream>>>(sp);
}

}  // namespace brt
