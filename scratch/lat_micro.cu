#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double *out, long long *cyc, int iters, double a, double b) {
    double x = threadIdx.x, y = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fma(x, a, b);
    }
    long long t1 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) y = y * a;
    }
    long long t2 = clock64();
    double z = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) z = __shfl_up_sync(0xffffffffu, z, 1) + 1.0;
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    if (x + y + z == 12345.678) out[0] = x;
}
int main() {
    double *out; long long *cyc, h[3];
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    const int iters = 4096;
    for (int warps : {1, 2, 4, 8}) {
        k_lat<<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
        cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
        printf("warps/CTA %d (1 CTA): dependent DFMA %.1f cyc, DMUL %.1f cyc, shfl+dadd %.1f cyc per op\n", warps,
               (double)h[0] / (iters * 16), (double)h[1] / (iters * 16), (double)h[2] / (iters * 16));
    }
    return 0;
}
