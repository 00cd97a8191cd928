// probe_dmma_lat.cu — latency / issue interval of mma.sync.m8n8k4.f64 on this GPU: one warp per SM, C independent
// accumulator chains, clock64 around a long dependent loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_dmma_lat probe_dmma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int C>
__global__ void k(double *out, long long *cyc, int iters, int warps) {
    if ((threadIdx.x >> 5) >= warps) return;
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001;
    double c[2 * C];
    for (int q = 0; q < 2 * C; ++q) c[q] = q;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < C; ++q) dmma884(c[2 * q], c[2 * q + 1], a, b);
    }
    const long long t1 = clock64();
    double r = 0;
    for (int q = 0; q < 2 * C; ++q) r += c[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C>
void run(double *d, long long *dc, int warps) {
    const int iters = 4096;
    k<C><<<148, 512>>>(d, dc, 64, warps);
    k<C><<<148, 512>>>(d, dc, iters, warps);
    long long c = 0;
    cudaMemcpy(&c, dc, sizeof c, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d chains %d: %.1f cycles per DMMA per warp (%.1f per loop of %d)\n", warps, C, (double)c / iters / C, (double)c / iters, C);
}

int main() {
    double *d; cudaMalloc(&d, sizeof(double) * 148 * 512);
    long long *dc; cudaMalloc(&dc, 8);
    for (int warps : {1, 2, 4, 8, 16}) {
        run<1>(d, dc, warps); run<2>(d, dc, warps); run<4>(d, dc, warps); run<8>(d, dc, warps);
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
