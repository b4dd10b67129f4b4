// packed.cu -- issue-rate micro-benchmarks for the packed FP32 (f32x2) forms of sm_100a: FFMA2 / FMUL2 / FADD2,
// alone and in the instruction mix of the packed pair loop of direct_fp32.cu (2 pairs per packed op).
// Prints warp-instructions per clock per SM and, for the mixes, pairs per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define UNROLL 16

enum { V_FFMA, V_FFMA2_8, V_FFMA2_16, V_FMUL2, V_FADD2, V_FFMA2_FMNMX, V_FFMA2_FFMA, V_PAIR_SCALAR, V_PAIR_PACKED,
       V_PAIR_PACKED_LDS, V_COUNT };
static const char *names[] = {"FFMA r,r,r x8", "FFMA2 x8 chains", "FFMA2 x16 chains", "FMUL2 x8", "FADD2 x8",
                              "FFMA2+2 FMNMX", "FFMA2+FFMA 1:1", "scalar pair mix (20)", "packed pair mix (15+4+2)/2",
                              "packed pair mix + LDS"};

__device__ __forceinline__ float rsq(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int V>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(a + threadIdx.x, b, a, b);
    __syncthreads();
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (V == V_FFMA) v[c].x = fmaf(v[c].x, a, b);
                if (V == V_FFMA2_8) v[c] = __ffma2_rn(v[c], A, B);
                if (V == V_FFMA2_16) { v[c] = __ffma2_rn(v[c], A, B); v[c + 8] = __ffma2_rn(v[c + 8], B, A); }
                if (V == V_FMUL2) v[c] = __fmul2_rn(v[c], A);
                if (V == V_FADD2) v[c] = __fadd2_rn(v[c], A);
                if (V == V_FFMA2_FMNMX) {
                    v[c] = __ffma2_rn(v[c], A, B);
                    v[c + 8].x = fmaxf(v[c + 8].x, v[c].x);
                    v[c + 8].y = fmaxf(v[c + 8].y, v[c].y);
                }
                if (V == V_FFMA2_FFMA) { v[c] = __ffma2_rn(v[c], A, B); v[c + 8].x = fmaf(v[c + 8].x, a, b); }
                if (V == V_PAIR_SCALAR && c < 4) {     // two scalar pairs per c (same work as one packed step)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        float x = h ? v[c].y : v[c].x;
                        float dx = x - a, dy = x - b, dz = x + a;
                        float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        float r2c = fmaxf(r2, 9.0f);
                        float ri = rsq(r2c);
                        float s = ri * ri, s3 = s * s * s;
                        float vv = fmaf(a * b, s3, -(b * a));
                        float e = fmaf(vv, s3, (a * b) * ri);
                        float uu = __saturatef(fmaf(r2, -1.0f / 144.0f, 1.0f));
                        float &acc = h ? v[c + 8].y : v[c + 8].x;
                        acc = fmaf(uu * uu, e, acc);
                    }
                }
                if ((V == V_PAIR_PACKED || V == V_PAIR_PACKED_LDS) && c < 4) {
                    float2 X = v[c], Y = v[c], Z = v[c], Q = A;
                    if (V == V_PAIR_PACKED_LDS) {
                        const float4 l0 = sm[(u * 8 + c * 2) & 255], l1 = sm[(u * 8 + c * 2 + 1) & 255];
                        X = make_float2(l0.x + v[c].x, l0.y); Y = make_float2(l0.z, l0.w);
                        Z = make_float2(l1.x, l1.y); Q = make_float2(l1.z, l1.w);
                    }
                    float2 dx = __fadd2_rn(X, A), dy = __fadd2_rn(Y, B), dz = __fadd2_rn(Z, A);
                    float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    float2 r2c = make_float2(fminf(fmaxf(r2.x, 9.0f), 144.0f), fminf(fmaxf(r2.y, 9.0f), 144.0f));
                    float2 ri = make_float2(rsq(r2c.x), rsq(r2c.y));
                    float2 s = __fmul2_rn(ri, ri);
                    float2 s3 = __fmul2_rn(__fmul2_rn(s, s), s);
                    float2 vv = __ffma2_rn(A, s3, B);
                    float2 e = __ffma2_rn(vv, s3, __fmul2_rn(Q, ri));
                    float2 up = __fadd2_rn(make_float2(144.f, 144.f), make_float2(-r2c.x, -r2c.y));
                    v[c + 8] = __ffma2_rn(__fmul2_rn(up, up), e, v[c + 8]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(float *d, int sms, double clk_hz, double per, double pairs_per) {
    int blocks = sms * 8, iters = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<V><<<blocks, 256>>>(d, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    double units = (double)UNROLL * iters * (double)blocks * 8;          // per-warp inner bodies, all warps
    double winst = per * units / (best * 1e-3) / clk_hz / sms;
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM", names[V], best, winst);
    if (pairs_per > 0) printf("  %6.3f warp-pairs/clk/SM = %5.2f cycles/pair/SMSP", pairs_per * units / (best * 1e-3) / clk_hz / sms,
                              4.0 / (pairs_per * units / (best * 1e-3) / clk_hz / sms));
    printf("\n");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clk = clk_khz * 1e3;
    printf("%s, %d SMs, clock attr %.0f MHz\n", p.name, p.multiProcessorCount, clk / 1e6);
    float *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    int s = p.multiProcessorCount;
    run<V_FFMA>(d, s, clk, 8, 0); run<V_FFMA2_8>(d, s, clk, 8, 0); run<V_FFMA2_16>(d, s, clk, 16, 0);
    run<V_FMUL2>(d, s, clk, 8, 0); run<V_FADD2>(d, s, clk, 8, 0); run<V_FFMA2_FMNMX>(d, s, clk, 24, 0);
    run<V_FFMA2_FFMA>(d, s, clk, 16, 0);
    run<V_PAIR_SCALAR>(d, s, clk, 4 * 2 * 20, 8); run<V_PAIR_PACKED>(d, s, clk, 4 * 21, 8);
    run<V_PAIR_PACKED_LDS>(d, s, clk, 4 * 23, 8);
    return 0;
}
