// Dependent-chain latency of DADD on sm_100a, alone and next to sibling warps that keep the FP64 pipe busy.
// (mc.cu's warp 0 adds 991 terms one by one: this is the floor of a Monte-Carlo frame.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int n, double a, int busy_warps, int mode) {
    const int wid = threadIdx.x >> 5;
    if (wid == 0) {
        double s = threadIdx.x, t = threadIdx.x + 1.0;
        long long t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < n; i++) { s = __dadd_rn(s, a); }
        long long t1 = clock64();
#pragma unroll 16
        for (int i = 0; i < n; i++) { s = __dadd_rn(s, a); t = __dadd_rn(t, a); }
        long long t2 = clock64();
        out[threadIdx.x] = s + t;
        if (threadIdx.x == 0)
            printf("busy sibling warps %d (mode %d): DADD chain %.2f cycles per step; two interleaved chains %.2f per step\n", busy_warps, mode,
                   (double)(t1 - t0) / n, (double)(t2 - t1) / n);
    } else if (wid <= busy_warps) {
        double v[8];
        for (int q = 0; q < 8; q++) v[q] = threadIdx.x + q;
        if (mode == 0) {
            for (int i = 0; i < 6 * n; i++)
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = fma(v[q], a, 1e-9);
        } else {
            for (int i = 0; i < n / 2; i++)
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = sqrt(v[q] + 1.0) / (v[q] + 3.0);
        }
        double r = 0;
        for (int q = 0; q < 8; q++) r += v[q];
        out[threadIdx.x] = r;
    }
}
int main() {
    double *o;
    cudaMalloc(&o, 8 * 1024);
    for (int mode = 0; mode < 2; mode++)
        for (int b : {0, 1, 3, 7}) { k<<<1, 32 * (b + 1)>>>(o, 1 << 15, 1.0000001, b, mode); cudaDeviceSynchronize(); }
    return 0;
}
