// pipes.cu -- issue-rate micro-benchmarks for the instruction mix of the direct pair kernel.
// Prints warp-instructions per clock per SM for several FP32 instruction forms (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

enum { V_FFMA_RRR, V_FFMA_RRI, V_FFMA_SAME, V_FMUL_RR, V_FADD_RR, V_FFMA_FMNMX, V_FFMA_MUFU8, V_MUFU, V_FMNMX,
       V_FFMA_RRR16, V_MIX_PAIR, V_COUNT };
static const char *names[] = {"FFMA r,r,r", "FFMA r,r,imm", "FFMA v,v,v", "FMUL r,r", "FADD r,r", "FFMA+FMNMX 1:1",
                              "FFMA:MUFU 8:1", "MUFU.RSQ", "FMNMX", "FFMA r,r,r 16 chains", "pair-like mix"};

template <int V>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 0.001f + i;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (V == V_FFMA_RRR) v[c] = fmaf(v[c], a, b);
                if (V == V_FFMA_RRI) v[c] = fmaf(v[c], a, 1e-7f);
                if (V == V_FFMA_SAME) v[c] = fmaf(v[c], v[c], v[c]);
                if (V == V_FMUL_RR) v[c] = v[c] * a;
                if (V == V_FADD_RR) v[c] = v[c] + a;
                if (V == V_FFMA_FMNMX) { v[c] = fmaf(v[c], a, b); v[c + 8] = fmaxf(v[c + 8], v[c]); }
                if (V == V_FFMA_MUFU8) { v[c] = fmaf(v[c], a, b); if (c == 0) asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[8 + (u & 7)])); }
                if (V == V_MUFU) asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
                if (V == V_FMNMX) v[c] = fmaxf(v[c], a);
                if (V == V_FFMA_RRR16) { v[c] = fmaf(v[c], a, b); v[c + 8] = fmaf(v[c + 8], b, a); }
                if (V == V_MIX_PAIR) {   // 3 FADD, 1 FMUL, 2 FFMA, 1 FMNMX, 1 MUFU, 8 FMUL, 4 FFMA, 1 FMNMX  (~21)
                    float dx = v[c] - a, dy = v[c] - b, dz = v[c] + a;
                    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    float r2c = fmaxf(r2, 9.0f), ri;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ri) : "f"(r2c));
                    float s = ri * ri, s3 = s * s * s;
                    float vv = fmaf(a * b, s3, -(b * a));
                    float e = fmaf(vv, s3, (a * b) * ri);
                    float uu = fmaxf(fmaf(r2c, -1.0f / 144.0f, 1.0f), 0.0f);
                    v[c + 8] = fmaf(uu * uu, e, v[c + 8]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(float *d, int sms, double clk_hz, int per) {
    int blocks = sms * 8, iters = 512;
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
    double winst = (double)per * CHAINS * UNROLL * iters * (double)blocks * 8;   // warp instructions
    double per_clk_sm = winst / (best * 1e-3) / clk_hz / sms;
    printf("%-24s %8.3f ms  %6.3f warp-instr/clk/SM  (%.1f%% of 4/clk)\n", names[V], best, per_clk_sm, 100 * per_clk_sm / 4);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clk = clk_khz * 1e3;
    printf("%s, %d SMs, clock attr %.0f MHz\n", p.name, p.multiProcessorCount, clk / 1e6);
    float *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    int s = p.multiProcessorCount;
    run<V_FFMA_RRR>(d, s, clk, 1); run<V_FFMA_RRI>(d, s, clk, 1); run<V_FFMA_SAME>(d, s, clk, 1);
    run<V_FMUL_RR>(d, s, clk, 1); run<V_FADD_RR>(d, s, clk, 1); run<V_FFMA_FMNMX>(d, s, clk, 2);
    run<V_FFMA_MUFU8>(d, s, clk, 1); run<V_MUFU>(d, s, clk, 1); run<V_FMNMX>(d, s, clk, 1);
    run<V_FFMA_RRR16>(d, s, clk, 2); run<V_MIX_PAIR>(d, s, clk, 21);
    return 0;
}
