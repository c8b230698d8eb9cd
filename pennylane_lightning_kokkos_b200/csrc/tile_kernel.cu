// b2sv: tile executor, complex128 instantiations + dtype dispatch (kernel: tile_kernel.cuh).
#include "tile_kernel.cuh"

namespace b2sv {

void launch_tile_pass_c64(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                          cudaStream_t stream, int max_ctas); // tile_kernel_c64.cu

// Tile geometry per dtype: complex128 -> 2^12 amps (64 KiB), complex64 -> 2^13 amps (64 KiB);
// three buffers per CTA (192 KiB of the 227 KiB an sm_100 CTA may use).
void tile_config(int dtype, int *B, int *R) {
    *B = dtype == 1 ? 12 : 13;
    *R = dtype == 1 ? 4 : 5; // 16 double2 / 32 float2 amplitudes per thread = 64 registers
}

void launch_tile_pass(int dtype, void *state, const PassParams &pp, int n_eff,
                      uint64_t rank_bits, cudaStream_t stream, int max_ctas) {
    if (dtype == 1)
        launch_tile_pass_t<double, 12, 4>(state, pp, n_eff, rank_bits, stream, max_ctas);
    else
        launch_tile_pass_c64(state, pp, n_eff, rank_bits, stream, max_ctas);
}

// Reads and clears the phase timers (zeros unless B2SV_TILE_PROF=1; complex128 kernels only).
void tile_prof_read(unsigned long long out[16]) {
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpyFromSymbol(out, g_tile_prof, sizeof(unsigned long long) * 16));
    unsigned long long z[16] = {0};
    CUDA_CHECK(cudaMemcpyToSymbol(g_tile_prof, z, sizeof(z)));
}

} // namespace b2sv
