// Micro-benchmark: what does HBM deliver for an IN-PLACE read-modify-write stream (the tile
// executor's traffic pattern) compared with an out-of-place copy?  2^30 complex128 amplitudes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o inplace_stream inplace_stream.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int U>
__global__ void __launch_bounds__(1024) inplace_scale(double2 *a, size_t n, double s) {
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) v[u] = a[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) { v[u].x *= s; v[u].y *= s; a[i + u * stride] = v[u]; }
    }
}
template <int U>
__global__ void __launch_bounds__(1024) copy_scale(const double2 *__restrict__ a, double2 *__restrict__ b, size_t n, double s) {
    size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) v[u] = a[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; u++) if (i + u * stride < n) { v[u].x *= s; v[u].y *= s; b[i + u * stride] = v[u]; }
    }
}
// tile-ordered in-place: CTA b walks 64 KiB tiles b, b+grid, ... (the executor's assignment)
__global__ void __launch_bounds__(1024) inplace_tiles(double2 *a, size_t n_tiles, double s) {
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        double2 *p = a + t * 4096;
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = p[u * 1024 + threadIdx.x];
#pragma unroll
        for (int u = 0; u < 4; u++) { v[u].x *= s; v[u].y *= s; p[u * 1024 + threadIdx.x] = v[u]; }
    }
}
int main() {
    const size_t n = size_t(1) << 30;
    double2 *a, *b;
    CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16 / 2));
    CK(cudaMemset(a, 0, n * 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char *name, float ms, double bytes) { printf("{\"kernel\": \"%s\", \"ms\": %.3f, \"gbs\": %.1f}\n", name, ms, bytes / ms / 1e6); };
    for (int grid : {148, 296, 592, 1184}) {
        float best = 1e9;
        for (int r = 0; r < 5; r++) { cudaEventRecord(e0); inplace_scale<4><<<grid, 1024>>>(a, n, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        char nm[64]; snprintf(nm, 64, "inplace_scale<4> grid %d", grid); report(nm, best, 2.0 * n * 16);
    }
    for (int grid : {148, 296}) {
        float best = 1e9;
        for (int r = 0; r < 5; r++) { cudaEventRecord(e0); inplace_tiles<<<grid, 1024>>>(a, n / 4096, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        char nm[64]; snprintf(nm, 64, "inplace_tiles grid %d", grid); report(nm, best, 2.0 * n * 16);
    }
    {
        float best = 1e9; const size_t h = n / 2;
        for (int r = 0; r < 5; r++) { cudaEventRecord(e0); copy_scale<4><<<592, 1024>>>(a, b, h, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        report("copy_scale<4> grid 592 (half size)", best, 2.0 * h * 16);
    }
    CK(cudaGetLastError());
    return 0;
}
