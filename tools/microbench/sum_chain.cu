// mc.cu's in-order sum of n double2 terms from shared memory: cycles per term for several loop shapes.
// (one warp, every lane adds the same terms; DADD dependent latency is 8.2 cycles on sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
template <int B>
__device__ __forceinline__ void batch_sum(const double2 *terms, int n, double &se, double &sv) {
    // double buffer of B terms: batch i + 1 is loaded (B loads back to back) before batch i is added
    const unsigned t0 = (unsigned)__cvta_generic_to_shared(terms);
    double2 u[B], w[B];
#pragma unroll
    for (int q = 0; q < B; q++) u[q] = lds_f64x2(t0 + 16u * q);
    int k = B;
    for (; k + B <= n; k += 2 * B) {
#pragma unroll
        for (int q = 0; q < B; q++) w[q] = lds_f64x2(t0 + 16u * (unsigned)(k + q));
#pragma unroll
        for (int q = 0; q < B; q++) { se = se + u[q].x; sv = sv + u[q].y; }
        if (k + 2 * B <= n) {
#pragma unroll
            for (int q = 0; q < B; q++) u[q] = lds_f64x2(t0 + 16u * (unsigned)(k + B + q));
        }
#pragma unroll
        for (int q = 0; q < B; q++) { se = se + w[q].x; sv = sv + w[q].y; }
    }
    if (k - B + B <= n && ((n / B) & 1)) {
#pragma unroll
        for (int q = 0; q < B; q++) { se = se + u[q].x; sv = sv + u[q].y; }
    }
}
__global__ void k(double *out, int n, int mode) {
    extern __shared__ double2 terms[];
    const int wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n; i += blockDim.x) terms[i] = make_double2(1e-3 * i, 1e-4 * i);
    __syncthreads();
    double se = 0.0, sv = 0.0;
    long long t0 = clock64();
    if (mode == 0 && wid == 0) batch_sum<8>(terms, n, se, sv);
    if (mode == 1 && wid == 0) batch_sum<16>(terms, n, se, sv);
    if (mode == 2) {       // warp 0 adds the x parts, warp 1 the y parts
        const double *t = (const double *)terms + wid;
        double s = 0.0;
#pragma unroll 8
        for (int i = 0; i < n; i++) s = s + t[2 * i];
        se = s;
    }
    if (mode == 3 && wid == 0) {   // terms in registers of the lanes, broadcast by shuffles (2 per double)
        for (int b = 0; b < n; b += 32) {
            const double2 mine = terms[b + (threadIdx.x & 31)];
#pragma unroll
            for (int l = 0; l < 32; l++) { se = se + __shfl_sync(0xffffffffu, mine.x, l); sv = sv + __shfl_sync(0xffffffffu, mine.y, l); }
        }
    }
    if (mode == 4 && wid == 0) {   // only the x chain, rotating buffer: is it the second chain or the loads?
        const double *t = (const double *)terms;
        double s = 0.0;
#pragma unroll 16
        for (int i = 0; i < n; i++) s = s + t[2 * i];
        se = s;
    }
    long long t1 = clock64();
    out[threadIdx.x] = se + sv;
    const char *names[] = {"double buffer of 8", "double buffer of 16", "two warps, one chain each", "lane-held terms + shuffles", "one chain only, plain loop"};
    if (threadIdx.x == 0) printf("%-28s %d terms: %.2f cycles per term\n", names[mode], n, (double)(t1 - t0) / n);
}
int main() {
    double *o;
    cudaMalloc(&o, 8 * 1024);
    for (int m = 0; m < 5; m++) { k<<<1, 64, 1024 * 16>>>(o, 1024, m); cudaDeviceSynchronize(); }
    return 0;
}
