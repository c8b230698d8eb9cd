// b2sv: tile executor, complex64 instantiations for plain-layout passes (bulk async tile loads).
#include "tile_kernel.cuh"

namespace b2sv {

void launch_tile_pass_c64_bulk(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                               cudaStream_t stream, int max_ctas) {
    launch_tile_pass_v<float, 13, 5, true>(state, pp, n_eff, rank_bits, stream, max_ctas);
}

} // namespace b2sv
