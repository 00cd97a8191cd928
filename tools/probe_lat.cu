// probe_lat.cu — dependent-issue latency of FP64 ops on this GPU (one warp, one chain), and of rsqrt / division / shared-memory loads.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // dependent DADD chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = a + c; a = a + c; a = a + c; a = a + c; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // dependent rsqrt chain
    double r = a + 2.0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { r = rsqrt(r) + 1.0; r = rsqrt(r) + 1.0; r = rsqrt(r) + 1.0; r = rsqrt(r) + 1.0; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // dependent division chain
    double d = a + 3.0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { d = 1.0 / d + 1.0; d = 1.0 / d + 1.0; d = 1.0 / d + 1.0; d = 1.0 / d + 1.0; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // dependent FP32 FMA chain
    float f = (float)a, g = 1.0000001f, h = 1e-9f;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { f = fmaf(f, g, h); f = fmaf(f, g, h); f = fmaf(f, g, h); f = fmaf(f, g, h); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 4 independent DFMA chains
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    out[threadIdx.x] = a + r + d + f + x0 + x1 + x2 + x3;
}
int main() {
    double *o; long long *c, h[6];
    cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
    const int iters = 4096;
    for (int threads : {32, 128, 512}) {
        k<<<1, threads>>>(o, c, 16);
        k<<<1, threads>>>(o, c, iters);
        cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
        printf("threads %4d: cycles per op  DFMA dep %.1f  DADD dep %.1f  rsqrt+add dep %.1f  div+add dep %.1f  FFMA dep %.1f  DFMA 4 chains %.1f (per op)\n", threads,
               h[0] / (4.0 * iters), h[1] / (4.0 * iters), h[2] / (4.0 * iters), h[3] / (4.0 * iters), h[4] / (4.0 * iters), h[5] / (4.0 * iters));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
