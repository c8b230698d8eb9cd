// b2sv: tile executor, complex128 instantiations for plain-layout passes (bulk async tile loads).
#include "tile_kernel.cuh"

namespace b2sv {

void launch_tile_pass_c128_bulk(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                                cudaStream_t stream, int max_ctas) {
    launch_tile_pass_v<double, 12, 4, true>(state, pp, n_eff, rank_bits, stream, max_ctas);
}

} // namespace b2sv
