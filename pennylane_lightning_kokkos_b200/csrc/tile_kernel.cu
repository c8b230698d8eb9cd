// b2sv: tile executor, complex128 instantiations (swizzled layout) + dtype dispatch (kernel: tile_kernel.cuh).
#include "tile_kernel.cuh"

namespace b2sv {

void launch_tile_pass_c64(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                          cudaStream_t stream, int max_ctas); // tile_kernel_c64.cu
void launch_tile_pass_c64_bulk(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                               cudaStream_t stream, int max_ctas); // tile_kernel_c64_bulk.cu
void launch_tile_pass_c128_bulk(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                                cudaStream_t stream, int max_ctas); // tile_kernel_bulk.cu

// Tile geometry per dtype: complex128 -> 2^12 amps (64 KiB), complex64 -> 2^13 amps (64 KiB);
// three buffers per CTA (192 KiB of the 227 KiB an sm_100 CTA may use).
void tile_config(int dtype, int *B, int *R) {
    *B = dtype == 1 ? 12 : 13;
    *R = dtype == 1 ? 4 : 5; // 16 double2 / 32 float2 amplitudes per thread = 64 registers
}

void launch_tile_pass(int dtype, void *state, const PassParams &pp, int n_eff,
                      uint64_t rank_bits, cudaStream_t stream, int max_ctas) {
    const bool bulk = pp.hdr.plain_layout != 0 && !tile_prof();
    B2_ABORT_IF(pp.hdr.plain_layout != 0 && tile_prof(),
                "B2SV_TILE_PROF needs B2SV_BULK=0 (the profiling variant has the swizzled loader only)");
    if (dtype == 1) {
        if (bulk)
            launch_tile_pass_c128_bulk(state, pp, n_eff, rank_bits, stream, max_ctas);
        else
            launch_tile_pass_v<double, 12, 4, false>(state, pp, n_eff, rank_bits, stream, max_ctas);
    } else {
        if (bulk)
            launch_tile_pass_c64_bulk(state, pp, n_eff, rank_bits, stream, max_ctas);
        else
            launch_tile_pass_c64(state, pp, n_eff, rank_bits, stream, max_ctas);
    }
}

// Reads and clears the phase timers (zeros unless B2SV_TILE_PROF=1; complex128 kernels only).
void tile_prof_read(unsigned long long out[16]) {
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpyFromSymbol(out, g_tile_prof, sizeof(unsigned long long) * 16));
    unsigned long long z[16] = {0};
    CUDA_CHECK(cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z)));
}

} // namespace b2sv
