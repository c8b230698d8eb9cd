// b2sv: tile executor, complex64 instantiations, swizzled layout (kernel: tile_kernel.cuh).
#include "tile_kernel.cuh"

namespace b2sv {

void launch_tile_pass_c64(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                          cudaStream_t stream, int max_ctas) {
    launch_tile_pass_v<float, 13, 5, false>(state, pp, n_eff, rank_bits, stream, max_ctas);
}

} // namespace b2sv
