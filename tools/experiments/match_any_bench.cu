// Micro-benchmark: throughput of __match_any_sync vs the 14-ballot peer search, 32 warps per SM (as in bgzf_compress_kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/match_any_bench tools/experiments/match_any_bench.cu && build/match_any_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned peers_ballot(uint32_t h)
{
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 14; k++) {
        const uint32_t bit = (h >> k) & 1u;
        const unsigned b = __ballot_sync(0xffffffffu, bit != 0);
        peers &= b ^ (bit - 1u);
    }
    return peers;
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t *out, int iters, uint32_t spread, long long *cyc)
{
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        const uint32_t h = (x >> 10) % spread;          // spread: number of distinct values (small = many equal lanes)
        acc += MODE == 0 ? __match_any_sync(0xffffffffu, h) : peers_ballot(h);
    }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    uint32_t *out; long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (uint32_t spread : { 16384u, 64u, 4u, 1u }) {
        k<0><<<148, 1024>>>(out, iters, spread, cyc); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        const double a = (double)h[0] / iters;
        k<1><<<148, 1024>>>(out, iters, spread, cyc); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        const double b = (double)h[0] / iters;
        printf("distinct values %5u: match.any %.1f cycles per call-round of 32 warps (%.2f per warp-call), 14 ballots %.1f (%.2f)\n", spread, a, a / 32, b, b / 32);
    }
    return 0;
}
