// runtime.cu -- device selection, error plumbing, timers, raw buffers, roofline micro-benchmarks
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <new>
#include <stdexcept>

namespace mmo {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return MMO_ECUDA;
}
int on_exception() {
    try { throw; }
    catch (const std::bad_alloc &) { set_error("out of host memory"); return MMO_ENOMEM; }
    catch (const std::exception &e) { set_error("internal error: %s", e.what()); return MMO_EINVAL; }
    catch (...) { set_error("internal error (unknown exception)"); return MMO_EINVAL; }
}
Runtime &rt() {
    static Runtime r;
    return r;
}
int stage_buffer(void **p) {
    Runtime &R = rt();
    if (!R.stage) MMO_CUDA(cudaHostAlloc(&R.stage, kStageHalf + kStageReadback, cudaHostAllocDefault));
    *p = R.stage;
    return MMO_OK;
}
// ---- per-kernel timing ---------------------------------------------------------------------------
struct KTimer {
    bool enabled = false;
    std::vector<cudaEvent_t> pool;
    struct Rec { int id; cudaEvent_t e0, e1; };
    std::vector<Rec> pending;
    double ms[K_COUNT] = {0};
    long long launches[K_COUNT] = {0};
};
static KTimer g_kt;
static cudaEvent_t kt_event() {
    if (!g_kt.pool.empty()) { cudaEvent_t e = g_kt.pool.back(); g_kt.pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
static void kt_resolve() {
    for (auto &r : g_kt.pending) {
        cudaEventSynchronize(r.e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        g_kt.ms[r.id] += ms;
        g_kt.launches[r.id]++;
        g_kt.pool.push_back(r.e0);
        g_kt.pool.push_back(r.e1);
    }
    g_kt.pending.clear();
}
KernelScope::KernelScope(int kid) : id(kid), on(g_kt.enabled) {
    if (on) {
        KTimer::Rec r{kid, kt_event(), kt_event()};
        cudaEventRecord(r.e0, rt().stream);
        g_kt.pending.push_back(r);
    }
}
KernelScope::~KernelScope() {
    if (on) {
        cudaEventRecord(g_kt.pending.back().e1, rt().stream);
        if (g_kt.pending.size() > 4096) kt_resolve();
    }
}

// ---- caching device allocator ------------------------------------------------------------------------
struct PoolBlock { void *p; size_t bytes; int epoch; };   // epoch = the mmo_init (device) the block was allocated under
static std::vector<PoolBlock> g_pool_free;          // cached, not in use
static std::vector<PoolBlock> g_pool_live;          // handed out
static size_t g_pool_cached = 0;
static const size_t kPoolMaxCached = (size_t)2 << 30;

int pool_alloc(void **out, size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    int best = -1;
    for (int i = 0; i < (int)g_pool_free.size(); i++)
        if (g_pool_free[i].epoch == rt().epoch && g_pool_free[i].bytes >= bytes && g_pool_free[i].bytes <= 2 * bytes + 4096 &&
            (best < 0 || g_pool_free[i].bytes < g_pool_free[best].bytes))
            best = i;
    if (best >= 0) {
        PoolBlock b = g_pool_free[best];
        g_pool_free.erase(g_pool_free.begin() + best);
        g_pool_cached -= b.bytes;
        g_pool_live.push_back(b);
        *out = b.p;
        return MMO_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {           // give the cache back to the driver and retry once
        cudaGetLastError();
        pool_trim();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
    g_pool_live.push_back({p, bytes, rt().epoch});
    *out = p;
    return MMO_OK;
}
void pool_free(void *p) {
    for (size_t i = 0; i < g_pool_live.size(); i++)
        if (g_pool_live[i].p == p) {
            PoolBlock b = g_pool_live[i];
            g_pool_live.erase(g_pool_live.begin() + i);
            if (rt().ready && b.epoch == rt().epoch && g_pool_cached + b.bytes <= kPoolMaxCached) {     // a block of an earlier device is never cached
                // work queued on the library stream may still use the block: the pool is only handed
                // out again to work on the same stream, which is ordered after it
                g_pool_free.push_back(b);
                g_pool_cached += b.bytes;
            } else {
                cudaFree(b.p);
            }
            return;
        }
    cudaFree(p);     // not ours (e.g. allocated before a re-init)
}
void pool_trim() {
    for (auto &b : g_pool_free) cudaFree(b.p);
    g_pool_free.clear();
    g_pool_cached = 0;
}

int require_ready() {
    if (!rt().ready) {
        set_error("mmo_init() has not been called (or failed): there is no CPU fallback");
        return MMO_ESTATE;
    }
    return MMO_OK;
}

// ---- micro-benchmarks: the measured denominators of the roofline --------------------------------
// 8 independent FMA chains per thread, enough to saturate the FMA pipes at full occupancy.
template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T *out, int iters, T a, T b) {
    T v0 = (T)threadIdx.x, v1 = v0 + (T)1, v2 = v0 + (T)2, v3 = v0 + (T)3;
    T v4 = v0 + (T)4, v5 = v0 + (T)5, v6 = v0 + (T)6, v7 = v0 + (T)7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            // explicit fma(): this file is compiled with -fmad=false
            v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
            v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

__global__ void copy_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

// Pure trilinear-gather pattern of K4 without its arithmetic: every thread reads the 8 corner floats of pseudo-random
// cells of T type-major maps (G3D.t layout, index i + j*x_dim + k*xy_dim) and adds them up.  The maps are read once
// before the timed launches, so they sit in L2 when they fit (C3: 46.8 MB of 126 MB).
__global__ void __launch_bounds__(256)
gather_kernel(const float *__restrict__ maps, size_t nvox, int d0, int d1, int d2, int T, int iters, float *__restrict__ out) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int xy = d0 * d1;
    float acc = 0.f;
    unsigned long long z = tid * 0x9e3779b97f4a7c15ull + 0x1234567ull;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        z += 0x9e3779b97f4a7c15ull;
        unsigned long long h = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        h = (h ^ (h >> 27)) * 0x94d049bb133111ebull;
        h ^= h >> 31;
        const int t = (int)((h & 0xffff) % (unsigned)T);
        const int i = (int)(((h >> 16) & 0xffff) % (unsigned)(d0 - 1)), j = (int)(((h >> 32) & 0xffff) % (unsigned)(d1 - 1));
        const int k = (int)((h >> 48) % (unsigned)(d2 - 1));
        const float *a = maps + (size_t)t * nvox + (size_t)i + (size_t)j * d0 + (size_t)k * xy;
        acc += (__ldg(a) + __ldg(a + 1)) + (__ldg(a + d0) + __ldg(a + d0 + 1)) +
               (__ldg(a + xy) + __ldg(a + xy + 1)) + (__ldg(a + xy + d0) + __ldg(a + xy + d0 + 1));
    }
    out[tid] = acc;
}

// the same for the z-pair copy the look-ups read (mmo_grid::zpair): four 8-byte loads, two rows of x
__global__ void __launch_bounds__(256)
gather_zpair_kernel(const float2 *__restrict__ zp, size_t zvox, int d0, int d1, int d2, int T, int iters, float *__restrict__ out) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int xy = d0 * d1;
    float acc = 0.f;
    unsigned long long z = tid * 0x9e3779b97f4a7c15ull + 0x1234567ull;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        z += 0x9e3779b97f4a7c15ull;
        unsigned long long h = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        h = (h ^ (h >> 27)) * 0x94d049bb133111ebull;
        h ^= h >> 31;
        const int t = (int)((h & 0xffff) % (unsigned)T);
        const int i = (int)(((h >> 16) & 0xffff) % (unsigned)(d0 - 1)), j = (int)(((h >> 32) & 0xffff) % (unsigned)(d1 - 1));
        const int k = (int)((h >> 48) % (unsigned)(d2 - 1));
        const float2 *a = zp + (size_t)t * zvox + (size_t)i + (size_t)j * d0 + (size_t)k * xy;
        const float2 p = __ldg(a), q = __ldg(a + 1), r = __ldg(a + d0), s = __ldg(a + d0 + 1);
        acc += ((p.x + p.y) + (q.x + q.y)) + ((r.x + r.y) + (s.x + s.y));
    }
    out[tid] = acc;
}

template <typename T>
static int measure_fma(double *tflops) {
    MMO_TRY(require_ready());
    Runtime &R = rt();
    const int blocks = R.sm_count * 8, threads = 256, iters = 4096;
    DevBuf<T> out;
    MMO_TRY(out.alloc((size_t)blocks * threads));
    cudaEvent_t e0, e1;
    MMO_CUDA(cudaEventCreate(&e0));
    MMO_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        MMO_CUDA(cudaEventRecord(e0, R.stream));
        fma_chain_kernel<T><<<blocks, threads, 0, R.stream>>>(out.p, iters, (T)1.0000001, (T)1e-7);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cudaEventRecord(e1, R.stream));
        MMO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MMO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return MMO_OK;
}

}  // namespace mmo

using namespace mmo;

extern "C" {

const char *mmo_last_error(void) { return g_err; }

const char *mmo_build_info(void) {
    return "libmmo_b200 sm_100a CUDA " MMO_STR(CUDART_VERSION) " built " __DATE__ " " __TIME__;
}

int mmo_device_count(int *n) try {
    MMO_REQUIRE(n != nullptr, "mmo_device_count: null pointer");
    MMO_CUDA(cudaGetDeviceCount(n));
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_init(int device) try {
    Runtime &R = rt();
    if (R.ready && R.device == device) return MMO_OK;
    if (R.ready) mmo_shutdown();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("mmo_init: no CUDA device available (%s); libmmo_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return MMO_ECUDA;
    }
    MMO_REQUIRE(device >= 0 && device < n, "mmo_init: device %d out of range [0,%d)", device, n);
    MMO_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MMO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("mmo_init: device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor);
        return MMO_ECUDA;
    }
    R.device = device;
    R.sm_count = prop.multiProcessorCount;
    MMO_CUDA(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking));
    MMO_CUDA(cudaStreamCreateWithFlags(&R.copy_stream, cudaStreamNonBlocking));
    MMO_CUDA(cudaEventCreate(&R.ev0));
    MMO_CUDA(cudaEventCreate(&R.ev1));
    R.launches = 0;
    R.epoch++;
    R.ready = true;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_shutdown(void) try {
    Runtime &R = rt();
    if (!R.ready) return MMO_OK;
    cudaStreamSynchronize(R.stream);
    scan_drop_caches();
    direct_drop_caches();
    mc_drop_caches();
    pool_trim();
    if (R.l2_scratch) cudaFree(R.l2_scratch);
    if (R.stage) cudaFreeHost(R.stage);
    R.stage = nullptr;
    R.l2_scratch = nullptr;
    R.l2_scratch_bytes = 0;
    cudaEventDestroy(R.ev0);
    cudaEventDestroy(R.ev1);
    cudaStreamDestroy(R.stream);
    R.stream = nullptr;
    if (R.copy_stream) { cudaStreamSynchronize(R.copy_stream); cudaStreamDestroy(R.copy_stream); }
    R.copy_stream = nullptr;
    R.ready = false;
    return MMO_OK;
} MMO_CATCH_ALL

int64_t mmo_launch_count(void) { return rt().launches; }

int mmo_kernel_timing(int on) try {
    MMO_TRY(require_ready());
    kt_resolve();
    g_kt.enabled = on != 0;
    for (int i = 0; i < K_COUNT; i++) { g_kt.ms[i] = 0.0; g_kt.launches[i] = 0; }
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_kernel_time_get(int kernel_id, double *ms, int64_t *launches) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(kernel_id >= 0 && kernel_id < K_COUNT, "mmo_kernel_time_get: bad kernel id %d", kernel_id);
    kt_resolve();
    if (ms) *ms = g_kt.ms[kernel_id];
    if (launches) *launches = g_kt.launches[kernel_id];
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_timer_start(void) try {
    MMO_TRY(require_ready());
    MMO_CUDA(cudaEventRecord(rt().ev0, rt().stream));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_timer_stop(float *ms) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(ms != nullptr, "mmo_timer_stop: null pointer");
    MMO_CUDA(cudaEventRecord(rt().ev1, rt().stream));
    MMO_CUDA(cudaEventSynchronize(rt().ev1));
    MMO_CUDA(cudaEventElapsedTime(ms, rt().ev0, rt().ev1));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_sync(void) try {
    MMO_TRY(require_ready());
    MMO_CUDA(cudaStreamSynchronize(rt().stream));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_l2_flush(void) try {
    MMO_TRY(require_ready());
    Runtime &R = rt();
    const size_t bytes = (size_t)256 << 20;   // > 126 MB L2
    if (!R.l2_scratch) {
        MMO_CUDA(cudaMalloc(&R.l2_scratch, bytes));
        R.l2_scratch_bytes = bytes;
    }
    MMO_CUDA(cudaMemsetAsync(R.l2_scratch, 0x5a, R.l2_scratch_bytes, R.stream));
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_measure_fp32_peak(double *tflops) try {
    MMO_REQUIRE(tflops != nullptr, "null pointer");
    return measure_fma<float>(tflops);
} MMO_CATCH_ALL
int mmo_measure_fp64_peak(double *tflops) try {
    MMO_REQUIRE(tflops != nullptr, "null pointer");
    return measure_fma<double>(tflops);
} MMO_CATCH_ALL
int mmo_measure_hbm_copy(double *gbs) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(gbs != nullptr, "null pointer");
    Runtime &R = rt();
    const size_t n = (size_t)1 << 26;   // 64 Mi float4 = 1 GiB in, 1 GiB out
    DevBuf<float4> a, b;
    MMO_TRY(a.alloc(n));
    MMO_TRY(b.alloc(n));
    MMO_CUDA(cudaMemsetAsync(a.p, 0, n * sizeof(float4), R.stream));
    cudaEvent_t e0, e1;
    MMO_CUDA(cudaEventCreate(&e0));
    MMO_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        MMO_CUDA(cudaEventRecord(e0, R.stream));
        copy_kernel<<<R.sm_count * 16, 512, 0, R.stream>>>(a.p, b.p, n);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cudaEventRecord(e1, R.stream));
        MMO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MMO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double g = 2.0 * (double)n * sizeof(float4) / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gbs = best;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_measure_l2_gather(const int32_t dims[3], int32_t T, double *lookups_per_s) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(dims && lookups_per_s && T > 0 && dims[0] > 1 && dims[1] > 1 && dims[2] > 1, "mmo_measure_l2_gather: bad arguments");
    Runtime &R = rt();
    const size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
    DevBuf<float> maps, out;
    MMO_TRY(maps.alloc(nvox * (size_t)T));
    const int blocks = R.sm_count * 32, threads = 256, iters = 64;
    MMO_TRY(out.alloc((size_t)blocks * threads));
    MMO_CUDA(cudaMemsetAsync(maps.p, 0, nvox * (size_t)T * sizeof(float), R.stream));
    cudaEvent_t e0, e1;
    MMO_CUDA(cudaEventCreate(&e0));
    MMO_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        MMO_CUDA(cudaEventRecord(e0, R.stream));
        gather_kernel<<<blocks, threads, 0, R.stream>>>(maps.p, nvox, dims[0], dims[1], dims[2], T, iters, out.p);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cudaEventRecord(e1, R.stream));
        MMO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MMO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double lps = (double)blocks * threads * iters / (ms * 1e-3);
        if (rep > 0 && lps > best) best = lps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lookups_per_s = best;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_measure_l2_gather_zpair(const int32_t dims[3], int32_t T, double *lookups_per_s) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(dims && lookups_per_s && T > 0 && dims[0] > 1 && dims[1] > 1 && dims[2] > 1, "mmo_measure_l2_gather_zpair: bad arguments");
    Runtime &R = rt();
    const size_t zvox = (size_t)dims[0] * dims[1] * (dims[2] - 1);
    DevBuf<float2> zp;
    DevBuf<float> out;
    MMO_TRY(zp.alloc(zvox * (size_t)T));
    const int blocks = R.sm_count * 32, threads = 256, iters = 64;
    MMO_TRY(out.alloc((size_t)blocks * threads));
    MMO_CUDA(cudaMemsetAsync(zp.p, 0, zvox * (size_t)T * sizeof(float2), R.stream));
    cudaEvent_t e0, e1;
    MMO_CUDA(cudaEventCreate(&e0));
    MMO_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        MMO_CUDA(cudaEventRecord(e0, R.stream));
        gather_zpair_kernel<<<blocks, threads, 0, R.stream>>>(zp.p, zvox, dims[0], dims[1], dims[2], T, iters, out.p);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cudaEventRecord(e1, R.stream));
        MMO_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MMO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double lps = (double)blocks * threads * iters / (ms * 1e-3);
        if (rep > 0 && lps > best) best = lps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lookups_per_s = best;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_dev_alloc(size_t bytes, void **dptr) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(dptr != nullptr, "mmo_dev_alloc: null pointer");
    MMO_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_dev_free(void *dptr) try {
    MMO_TRY(require_ready());
    if (dptr) MMO_CUDA(cudaFree(dptr));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_h2d(void *dptr, const void *host, size_t bytes) try {
    MMO_TRY(require_ready());
    MMO_CUDA(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, rt().stream));
    MMO_CUDA(cudaStreamSynchronize(rt().stream));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_d2h(void *host, const void *dptr, size_t bytes) try {
    MMO_TRY(require_ready());
    MMO_CUDA(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, rt().stream));
    MMO_CUDA(cudaStreamSynchronize(rt().stream));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_host_alloc(size_t bytes, void **hptr) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(hptr != nullptr, "mmo_host_alloc: null pointer");
    MMO_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_host_free(void *hptr) try {
    MMO_TRY(require_ready());
    if (hptr) MMO_CUDA(cudaFreeHost(hptr));
    return MMO_OK;
} MMO_CATCH_ALL

}  // extern "C"
