// probe_dmma.cu — microbenchmark: FP64 DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 f64) throughput on this GPU,
// alone and concurrently (do they share a pipe?).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_dmma probe_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// mode 0: DFMA only; 1: DMMA m8n8k4 only; 2: DMMA m16n8k8 only; 3: even warps DFMA, odd warps m8n8k4; 4: even DFMA, odd m16n8k8
__global__ void k(double *out, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool fma_role = (mode == 0) || (mode >= 3 && (warp & 1) == 0);
    const bool big = (mode == 2 || mode == 4);
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001;
    double r = 0;
    if (fma_role) {
        double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
        for (int i = 0; i < iters; ++i) {
            x0 = fma(x0, b, 1e-9); x1 = fma(x1, b, 1e-9); x2 = fma(x2, b, 1e-9); x3 = fma(x3, b, 1e-9);
            x4 = fma(x4, b, 1e-9); x5 = fma(x5, b, 1e-9); x6 = fma(x6, b, 1e-9); x7 = fma(x7, b, 1e-9);
        }
        r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    } else if (!big) {
        double c[16];
        for (int q = 0; q < 16; ++q) c[q] = q;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 8; ++q) dmma884(c[2 * q], c[2 * q + 1], a, b);
        }
        for (int q = 0; q < 16; ++q) r += c[q];
    } else {
        double c[16], av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
        for (int q = 0; q < 16; ++q) c[q] = q;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) dmma1688(c + 4 * q, av, bv);
        }
        for (int q = 0; q < 16; ++q) r += c[q];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 512, iters = 1 << 14;
    double *d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 5; ++mode) {
        k<<<blocks, threads>>>(d, 256, mode);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); k<<<blocks, threads>>>(d, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
        // per iteration: FMA thread 8 fma = 16 flop; m8n8k4 warp: 8 mma x 512 flop; m16n8k8 warp: 4 mma x 2048 flop
        double fma_tf = 0, mma_tf = 0;
        if (mode == 0) fma_tf = nthreads * 16.0 * iters;
        if (mode == 1) mma_tf = nwarps * 8 * 512.0 * iters;
        if (mode == 2) mma_tf = nwarps * 4 * 2048.0 * iters;
        if (mode == 3) { fma_tf = nthreads / 2 * 16.0 * iters; mma_tf = nwarps / 2 * 8 * 512.0 * iters; }
        if (mode == 4) { fma_tf = nthreads / 2 * 16.0 * iters; mma_tf = nwarps / 2 * 4 * 2048.0 * iters; }
        printf("mode %d: %.3f ms  DFMA %.2f TF  DMMA %.2f TF  (sum %.2f)\n", mode, best, fma_tf / best / 1e9, mma_tf / best / 1e9,
               (fma_tf + mma_tf) / best / 1e9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
