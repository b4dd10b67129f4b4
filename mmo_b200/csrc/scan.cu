// scan.cu -- exhaustive rigid-body scan driver and K7 top-k / argmin.
//   Lds.exhaustive_rigid_ligand_docking   src/lds.ml:1040-1114
//   ROI.get_bounds / ROI.is_inside        src/ROI.ml:65-82
//   Grid.from_box (translation lattice)   src/grid.ml:40-52
//   G3D.vdW_clash_AND                     src/G3D.ml:189-213
//   Mol.is_ligand_center_vdW_occuppied    src/mol.ml:1209-1218
//
// The reference's loop nest (z, y, x, rotation) becomes: host enumerates the in-ROI lattice points
// (cheap: <= a few 1e4), the device handles slabs of points x all rotations:
//   prefilter (bitmask clash)  ->  survivor frame list  ->  score kernel  ->  argmin + top-k filter.
// frame = rot_i + n_rot*(i + j*x_dim + k*xy_dim) grows monotonically along the reference's loop
// order, so "first pose wins ties" (strict <, lds.ml:1099) == "smallest frame wins ties".
#include "common.cuh"
#include "pose.cuh"
#include <math.h>
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <string.h>
#include <memory>

namespace mmo {

struct ScoreFrame { double s; long long f; };

__device__ __forceinline__ bool sf_less(double s1, long long f1, double s2, long long f2) {
    return (s1 < s2) || (s1 == s2 && f1 < f2);
}

// score = e_intra_const +. ene_inter (lds.ml:1324-1325); per-block argmin; candidates with
// score <= thr are appended to the top-k candidate buffer
__global__ void __launch_bounds__(256)
scan_reduce_kernel(const double *__restrict__ energies, const int64_t *__restrict__ frames, int64_t id_base, int64_t n,
                   double e_intra, const double *__restrict__ thr, double *__restrict__ cand_s,
                   long long *__restrict__ cand_f, unsigned long long *__restrict__ cand_n,
                   unsigned long long cand_cap, ScoreFrame *__restrict__ block_best) {
    __shared__ double sh_s[8];
    __shared__ long long sh_f[8];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = INFINITY;
    long long f = 0x7fffffffffffffffLL;
    if (p < n) {
        s = e_intra + energies[p];
        f = frames ? frames[p] : id_base + p;
        if (cand_cap && s <= *thr) {
            unsigned long long k = atomicAdd(cand_n, 1ull);
            if (k < cand_cap) { cand_s[k] = s; cand_f[k] = f; }
        }
        if (!(s == s)) s = INFINITY;     // NaN never wins a strict '<' in the reference
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double s2 = __shfl_xor_sync(0xffffffffu, s, o);
        long long f2 = __shfl_xor_sync(0xffffffffu, f, o);
        if (sf_less(s2, f2, s, f)) { s = s2; f = f2; }
    }
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_f[threadIdx.x >> 5] = f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            if (sf_less(sh_s[w], sh_f[w], s, f)) { s = sh_s[w]; f = sh_f[w]; }
        block_best[blockIdx.x].s = s;
        block_best[blockIdx.x].f = f;
    }
}

// argmin over the per-block minima of scan_reduce_kernel (ties to the smaller frame): one ScoreFrame travels to the host
__global__ void __launch_bounds__(1024)
scan_best_kernel(const ScoreFrame *__restrict__ block_best, unsigned n, ScoreFrame *__restrict__ out) {
    __shared__ double sh_s[32];
    __shared__ long long sh_f[32];
    double s = INFINITY;
    long long f = 0x7fffffffffffffffLL;
    for (unsigned b = threadIdx.x; b < n; b += blockDim.x) {
        const double s2 = block_best[b].s;
        const long long f2 = block_best[b].f;
        if (sf_less(s2, f2, s, f)) { s = s2; f = f2; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double s2 = __shfl_xor_sync(0xffffffffu, s, o);
        long long f2 = __shfl_xor_sync(0xffffffffu, f, o);
        if (sf_less(s2, f2, s, f)) { s = s2; f = f2; }
    }
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_f[threadIdx.x >> 5] = f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; w++)
            if (sf_less(sh_s[w], sh_f[w], s, f)) { s = sh_s[w]; f = sh_f[w]; }
        out->s = s;
        out->f = f;
    }
}

// k-th smallest of the per-block minima (n <= 4096), written to *thr as the first slab's candidate threshold: it bounds
// the k-th best score from above.  One block, bitonic sort in shared memory; two_stage adds the fp32 sweep's 2 delta.
__global__ void __launch_bounds__(1024)
scan_kth_kernel(const ScoreFrame *__restrict__ block_best, unsigned n, unsigned k, int two_stage, double *__restrict__ thr) {
    __shared__ double v[4096];
    for (unsigned i = threadIdx.x; i < 4096; i += 1024) v[i] = i < n ? block_best[i].s : INFINITY;
    __syncthreads();
    for (unsigned size = 2; size <= 4096; size <<= 1)
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < 2048; t += 1024) {
                const unsigned lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const double a = v[lo], b = v[hi];
                if ((a > b) == up) { v[lo] = b; v[hi] = a; }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        double est = v[k - 1];
        if (two_stage && est < INFINITY) est = __dadd_rn(est, __dmul_rn(2.0, __dmul_rn(8.0, fmax(__dmul_rn(1e-6, fabs(est)), 1e-4))));
        *thr = est;
    }
}

struct RotSet {
    DevBuf<double> rot;
    DevBuf<int32_t> perm;
    std::vector<double> host;        // the set as it was handed over: compared bytewise on every call
    int n = -1;
    int epoch = -1;
};

struct ScanJob {
    mmo_scan_params P;
    std::shared_ptr<RotSet> rs;
    int dims[3];
    double mins[3], q[3];
    std::vector<int64_t> points;           // active lattice points of the requested sub-range
    DevBuf<int64_t> d_points, d_frames;
    DevBuf<double> d_E, d_thr, d_cand_s;
    DevBuf<long long> d_cand_f;
    DevBuf<unsigned long long> d_counters;  // [0] survivors, [1] candidates
    DevBuf<ScoreFrame> d_block_best;
    std::vector<ScoreFrame> top;            // running top-k (ascending)
    int64_t slab_cap = 0;
    int64_t n_candidates = 0, n_scored = 0, pairs_eval = 0, pairs_in = 0;
    double best_s = INFINITY;
    long long best_f = -1;
    float device_ms = 0.f;
    bool collect_stats = false;
    // MMO_PREC_FP64 on the direct path: fp32 sweep that keeps every pose that could be in the exact
    // top-k (score <= tau + 2*delta), then bit-exact re-scoring of those few in the strict kernel
    bool two_stage = false;
    int k_eff = 0;
    std::vector<ScoreFrame> exact_top;       // filled by scan_finalize
    double exact_best_s = INFINITY;
    long long exact_best_f = -1;
    bool exact_ready = false;
    // one-shot scans (mmo_scan): the caller's rotation bytes are uploaded and compared with the resident set on the copy
    // stream WHILE the scan runs on the resident set; the verdict is read when the scan is done (a different set: redo)
    bool speculative = false;
    bool rot_check_wanted = false;           // issue_rot_check still to be called (after the first slab's kernels are queued)
    bool rot_check_pending = false;
    DevBuf<double> rot_fresh;
    DevBuf<int> rot_flag;
    ~ScanJob() { if (rot_check_pending && rt().copy_stream) cudaStreamSynchronize(rt().copy_stream); }
};

// bound used for |E_fp32 - E_ref|: 8x the accuracy contract of MMO_PREC_FP32 (which the parity tests
// check at 1x on clashing and non-clashing poses alike)
static inline double fp32_delta(double e) { return 8.0 * std::max(1e-6 * fabs(e), 1e-4); }

// Bitv.get raises outside the vector; here a voxel outside the mask box reads as "not occupied" (mask.cu bit_ijk)
static bool host_bit(const mmo_mask *m, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= m->dims[0] || j >= m->dims[1] || k >= m->dims[2]) return false;
    const size_t idx = (size_t)i + (size_t)j * m->dims[0] + (size_t)k * m->dims[0] * m->dims[1];
    return (m->hwords[idx >> 5] >> (idx & 31)) & 1u;
}

// G3D.vdW_clash_AND on the host copy of the mask
static bool host_clash_and(const mmo_mask *m, double x, double y, double z) {
    double inv = 1.0 / m->step;
    int i0 = (int)(x * inv), j0 = (int)(y * inv), k0 = (int)(z * inv);
    int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    return host_bit(m, i0, j0, k0) && host_bit(m, i1, j0, k0) && host_bit(m, i1, j1, k0) && host_bit(m, i0, j1, k0) &&
           host_bit(m, i0, j0, k1) && host_bit(m, i1, j0, k1) && host_bit(m, i1, j1, k1) && host_bit(m, i0, j1, k1);
}

// Visiting order of the rotations: k-d leaves of 32 over the stereographic image of the unit
// quaternions, so that the 32 poses of a warp are similar rotations (tight per-atom boxes in the
// direct kernel).  Cached on the content of rot9: lds builds the rotation set once per run.
static void rotation_visit_order(int n_rot, const double *rot9, std::vector<int32_t> &cached) {
    std::vector<double> gx(n_rot), gy(n_rot), gz(n_rot);
    for (int r = 0; r < n_rot; r++) {
        const double *m = rot9 + 9 * (size_t)r;
        double q[4];     // (w, x, y, z), numerically safe branch on the largest diagonal term
        double tr = m[0] + m[4] + m[8];
        if (tr > 0.0) {
            double s = sqrt(tr + 1.0) * 2.0;
            q[0] = 0.25 * s; q[1] = (m[7] - m[5]) / s; q[2] = (m[2] - m[6]) / s; q[3] = (m[3] - m[1]) / s;
        } else if (m[0] > m[4] && m[0] > m[8]) {
            double s = sqrt(1.0 + m[0] - m[4] - m[8]) * 2.0;
            q[0] = (m[7] - m[5]) / s; q[1] = 0.25 * s; q[2] = (m[1] + m[3]) / s; q[3] = (m[2] + m[6]) / s;
        } else if (m[4] > m[8]) {
            double s = sqrt(1.0 + m[4] - m[0] - m[8]) * 2.0;
            q[0] = (m[2] - m[6]) / s; q[1] = (m[1] + m[3]) / s; q[2] = 0.25 * s; q[3] = (m[5] + m[7]) / s;
        } else {
            double s = sqrt(1.0 + m[8] - m[0] - m[4]) * 2.0;
            q[0] = (m[3] - m[1]) / s; q[1] = (m[2] + m[6]) / s; q[2] = (m[5] + m[7]) / s; q[3] = 0.25 * s;
        }
        if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
        double den = 1.0 + q[0];
        if (!(den > 1e-12)) den = 1e-12;
        gx[r] = q[1] / den; gy[r] = q[2] / den; gz[r] = q[3] / den;
        if (!(gx[r] == gx[r]) || !(gy[r] == gy[r]) || !(gz[r] == gz[r])) { gx[r] = gy[r] = gz[r] = 0.0; }
    }
    std::vector<int> order;
    kd_order(n_rot, gx.data(), gy.data(), gz.data(), 32, order);
    cached.assign(order.begin(), order.end());
}

// The rotation set of a run (lds builds it once, SO3.rotations, and scans every ligand with it) stays
// resident on the device together with its visiting order; a call that passes the same rotations again
// (same count, same bytes: one memcmp against the host copy) skips the 72 n_rot bytes upload and the k-d sort.
// leaked on purpose: must not run a destructor after the CUDA context / the allocator are gone
static std::shared_ptr<RotSet> &g_rotset = *new std::shared_ptr<RotSet>();
static long long g_rot_rescans = 0;    // one-shot scans that had to be redone because the rotation set was not the resident one
static int g_rot_cache_mode = 1;      // mmo_scan_set_rot_cache: 1 = upload + compare on the device (default), 0 = memcmp on the host
void scan_drop_caches() { g_rotset.reset(); }
// device-side comparison of a freshly uploaded rotation set with the resident one (mode 1: the bytes move on every call,
// the k-d visiting order is reused when they turn out to be the same set)
__global__ void __launch_bounds__(256)
rot_compare_kernel(const unsigned long long *__restrict__ a, const unsigned long long *__restrict__ b, size_t n, int *__restrict__ differ) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool d = false;
    for (; i < n; i += stride) d |= a[i] != b[i];
    if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0) atomicOr(differ, 1);
}

constexpr int kPinRotFlag = 6;       // slot of the rotation-set verdict in the pinned read-back area (as unsigned long long)
// The check of a one-shot scan, issued by scan_run_points once the first slab's kernels are queued: upload + compare on the
// copy stream.  Behind the kernels on purpose -- a pageable source (an OCaml float array) is staged by the driver while
// this call blocks the host (~1 ms for 7.2 MB), and by then the GPU has 10 ms of work.
static int issue_rot_check(ScanJob &J) {
    Runtime &R = rt();
    const size_t nw = (size_t)J.P.n_rot * 9;
    void *stage = nullptr;
    MMO_TRY(stage_buffer(&stage));
    int *pin_flag = (int *)((unsigned long long *)((char *)stage + kStageHalf) + kPinRotFlag);
    // (the two buffers were taken from the pool in scan_setup, before any kernel of this call was queued: the pool hands
    //  blocks out in stream order of the LIBRARY stream, so a block freed by a launcher a moment ago may still be in use
    //  there and must not be written from the copy stream)
    J.rot_check_pending = true;          // from here on the job waits for the copy stream before it frees anything
    J.rot_check_wanted = false;
    MMO_CUDA(cudaMemsetAsync(J.rot_flag.p, 0, sizeof(int), R.copy_stream));
    MMO_CUDA(cudaMemcpyAsync(J.rot_fresh.p, J.P.rot9, nw * sizeof(double), cudaMemcpyHostToDevice, R.copy_stream));
    rot_compare_kernel<<<R.sm_count * 4, 256, 0, R.copy_stream>>>((const unsigned long long *)J.rot_fresh.p, (const unsigned long long *)J.rs->rot.p, nw, J.rot_flag.p);
    MMO_LAUNCH_CHECK();
    MMO_CUDA(cudaMemcpyAsync(pin_flag, J.rot_flag.p, sizeof(int), cudaMemcpyDeviceToHost, R.copy_stream));
    return MMO_OK;
}

static int get_rotset(int n_rot, const double *rot9, std::shared_ptr<RotSet> &out, ScanJob *spec) {
    if (spec && g_rot_cache_mode == 1 && g_rotset && g_rotset->n == n_rot && g_rotset->epoch == rt().epoch) {
        MMO_TRY(spec->rot_fresh.alloc((size_t)n_rot * 9));
        MMO_TRY(spec->rot_flag.alloc(1));
        spec->rot_check_wanted = true;       // scan on the resident set now, verdict on the caller's bytes when the scan is done
        out = g_rotset;
        return MMO_OK;
    }
    if (g_rot_cache_mode == 1 && g_rotset && g_rotset->n == n_rot && g_rotset->epoch == rt().epoch) {
        Runtime &R = rt();
        const size_t nw = (size_t)n_rot * 9;
        DevBuf<double> fresh;
        DevBuf<int> flag;
        MMO_TRY(fresh.alloc(nw));
        MMO_TRY(flag.alloc(1));
        MMO_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), R.stream));
        MMO_CUDA(cudaMemcpyAsync(fresh.p, rot9, nw * sizeof(double), cudaMemcpyHostToDevice, R.stream));
        rot_compare_kernel<<<R.sm_count * 4, 256, 0, R.stream>>>((const unsigned long long *)fresh.p, (const unsigned long long *)g_rotset->rot.p, nw, flag.p);
        MMO_LAUNCH_CHECK();
        int differ = 0;
        MMO_CUDA(cudaMemcpyAsync(&differ, flag.p, sizeof(int), cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));
        if (!differ) { out = g_rotset; return MMO_OK; }
    } else if (g_rotset && g_rotset->n == n_rot && g_rotset->epoch == rt().epoch &&
               memcmp(g_rotset->host.data(), rot9, (size_t)n_rot * 9 * sizeof(double)) == 0) {
        out = g_rotset;
        return MMO_OK;
    }
    g_rotset.reset();
    std::shared_ptr<RotSet> rs = std::make_shared<RotSet>();
    rs->n = n_rot; rs->epoch = rt().epoch;
    rs->host.assign(rot9, rot9 + (size_t)n_rot * 9);
    std::vector<int32_t> perm;
    rotation_visit_order(n_rot, rot9, perm);
    MMO_TRY(rs->rot.upload(rot9, (size_t)n_rot * 9));
    MMO_TRY(rs->perm.upload(perm));
    g_rotset = rs;
    out = rs;
    return MMO_OK;
}


static int scan_setup(ScanJob &J) {
    const mmo_scan_params &P = J.P;
    // ROI.get_bounds (ROI.ml:76-82) -> Bbox.create_6f -> Grid.from_box trans_step (lds.ml:1065-1069)
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = P.roi_c[d] - P.roi_r; hi[d] = P.roi_c[d] + P.roi_r; J.mins[d] = lo[d]; }
    for (int d = 0; d < 3; d++) {
        J.dims[d] = grid_num_steps(P.trans_step, hi[d] - lo[d]) + 1;
        int np = J.dims[d] - 1;
        J.q[d] = np > 0 ? (P.trans_step * (double)np) / (double)np : 0.0;
    }
    const int64_t nvox = (int64_t)J.dims[0] * J.dims[1] * J.dims[2];
    int64_t p0 = std::max<int64_t>(0, P.first_point);
    int64_t p1 = (P.n_points < 0) ? nvox : std::min(nvox, p0 + P.n_points);
    const mmo_ligand *lig = P.lig;
    // lds.ml:1044-1052: AND-prefilter on the lattice point only if the ligand's centre is inside one
    // of its own atoms (mol.ml:1209-1218, V3.dist c xyz < radius)
    bool center_filter = false;
    if (P.vdw_mask) {
        MMO_REQUIRE(lig->has_r, "mmo_scan: the vdW prefilter needs ligand radii");
        for (int i = 0; i < lig->n && !center_filter; i++) {
            double dx = 0.0 - lig->hx[i], dy = 0.0 - lig->hy[i], dz = 0.0 - lig->hz[i];
            if (sqrt(dx * dx + dy * dy + dz * dz) < lig->hr[i]) center_filter = true;
        }
    }
    const double r2 = P.roi_r * P.roi_r;       // ROI.ml:19-20
    const int x_dim = J.dims[0], xy_dim = J.dims[0] * J.dims[1];
    J.points.clear();
    for (int64_t p = p0; p < p1; p++) {
        int k = (int)(p / xy_dim);
        int j = (int)((p - (int64_t)k * xy_dim) / x_dim);
        int i = (int)(p - ((int64_t)k * xy_dim + (int64_t)j * x_dim));
        double pos[3] = {J.mins[0] + (double)i * J.q[0], J.mins[1] + (double)j * J.q[1], J.mins[2] + (double)k * J.q[2]};
        double dx = P.roi_c[0] - pos[0], dy = P.roi_c[1] - pos[1], dz = P.roi_c[2] - pos[2];
        if (!(dx * dx + dy * dy + dz * dz < r2)) continue;           // ROI.is_inside, strict
        if (center_filter && host_clash_and(P.vdw_mask, pos[0], pos[1], pos[2])) continue;
        J.points.push_back(p);
    }
    MMO_TRY(get_rotset(P.n_rot, P.rot9, J.rs, J.speculative ? &J : nullptr));
    MMO_TRY(J.d_points.upload(J.points));
    // slab: as many lattice points as fit ~4M candidate poses
    int64_t pts_per_slab = std::max<int64_t>(1, (int64_t)(4 << 20) / std::max(1, P.n_rot));
    pts_per_slab = std::max<int64_t>(1, std::min<int64_t>(pts_per_slab, (int64_t)J.points.size()));
    J.slab_cap = pts_per_slab * P.n_rot;
    MMO_TRY(J.d_frames.alloc((size_t)J.slab_cap));
    MMO_TRY(J.d_E.alloc((size_t)J.slab_cap));
    MMO_TRY(J.d_thr.alloc(1));
    MMO_TRY(J.d_counters.alloc(2));
    J.two_stage = (P.prec == MMO_PREC_FP64 && P.rec != nullptr);
    J.k_eff = J.two_stage ? std::max(P.topk, 1) : P.topk;
    if (J.k_eff > 0) {
        MMO_TRY(J.d_cand_s.alloc((size_t)J.slab_cap));
        MMO_TRY(J.d_cand_f.alloc((size_t)J.slab_cap));
    }
    MMO_TRY(J.d_block_best.alloc((size_t)(J.slab_cap + 255) / 256 + 1));     // + the slab's best (scan_best_kernel)
    J.top.clear();
    return MMO_OK;
}

static int scan_run_points(ScanJob &J, int64_t a0, int64_t a1) {
    const mmo_scan_params &P = J.P;
    Runtime &R = rt();
    const int64_t pts_per_slab = J.slab_cap / P.n_rot;
    PoseSrc src = {};
    src.kind = 2;
    src.rot9 = J.rs->rot.p;
    src.frames = J.d_frames.p;
    src.n_rot = P.n_rot;
    for (int d = 0; d < 3; d++) { src.lat_dims[d] = J.dims[d]; src.lat_min[d] = J.mins[d]; src.lat_q[d] = J.q[d]; }
    std::vector<ScoreFrame> hbest;
    std::vector<double> hs;
    std::vector<long long> hf;
    // pinned read-back area (second half of the library's staging buffer): counters, the slab's best, candidate prefix
    constexpr int64_t kCandPrefix = 8192;
    void *stage = nullptr;
    MMO_TRY(stage_buffer(&stage));
    unsigned long long *pin_u64 = (unsigned long long *)((char *)stage + kStageHalf);
    ScoreFrame *pin_best = (ScoreFrame *)(pin_u64 + 2);
    double *pin_cs = (double *)(pin_u64 + 8);
    long long *pin_cf = (long long *)(pin_cs + kCandPrefix);
    static_assert(64 + 2 * kCandPrefix * 8 <= (int64_t)kStageReadback, "read-back area exceeds the staging buffer");
    for (int64_t s0 = a0; s0 < a1; s0 += pts_per_slab) {
        const int64_t npts = std::min(pts_per_slab, a1 - s0);
        const int64_t n_cand = npts * P.n_rot;
        J.n_candidates += n_cand;
        double thr = (J.k_eff > 0 && (int)J.top.size() >= J.k_eff) ? J.top[J.k_eff - 1].s : INFINITY;
        if (J.two_stage && thr < INFINITY) thr = thr + 2.0 * fp32_delta(thr);
        J.exact_ready = false;
        MMO_CUDA(cudaMemsetAsync(J.d_counters.p, 0, 2 * sizeof(unsigned long long), R.stream));
        MMO_CUDA(cudaMemcpyAsync(J.d_thr.p, &thr, sizeof(double), cudaMemcpyHostToDevice, R.stream));
        MMO_TRY(launch_scan_prefilter(P.vdw_mask, P.lig, src, J.d_points.p + s0, J.rs->perm.p, n_cand, J.d_frames.p, J.d_counters.p));
        // without a mask the prefilter only writes the frames, in order: nothing to wait for
        unsigned long long n_surv = (unsigned long long)n_cand;
        if (P.vdw_mask) {
            MMO_CUDA(cudaMemcpyAsync(pin_u64, J.d_counters.p, sizeof n_surv, cudaMemcpyDeviceToHost, R.stream));
            MMO_CUDA(cudaStreamSynchronize(R.stream));
            n_surv = pin_u64[0];
        }
        J.n_scored += (int64_t)n_surv;
        if (n_surv == 0) continue;
        if (P.grid) {
            MMO_TRY(launch_interp(P.grid, P.lig, src, (int64_t)n_surv, J.d_E.p));
        } else if (P.prec == MMO_PREC_FP64 && !J.two_stage) {
            MMO_TRY(launch_direct_fp64(P.rec, P.lig, P.variant, src, (int64_t)n_surv, J.d_E.p));
        } else {
            MMO_TRY(launch_direct_fp32(P.rec, P.lig, P.variant, src, (int64_t)n_surv, J.d_E.p, J.collect_stats));
            if (J.collect_stats) { J.pairs_eval += R.stat_pairs; J.pairs_in += R.stat_inside; }
        }
        if (J.rot_check_wanted) MMO_TRY(issue_rot_check(J));
        const unsigned blocks = (unsigned)((n_surv + 255) / 256);
        auto reduce_pass = [&](bool with_candidates) -> int {
            KernelScope ks(K_REDUCE);
            scan_reduce_kernel<<<blocks, 256, 0, R.stream>>>(J.d_E.p, J.d_frames.p, 0, (int64_t)n_surv, P.e_intra_const,
                                                             J.d_thr.p, J.d_cand_s.p, J.d_cand_f.p, J.d_counters.p + 1,
                                                             with_candidates ? (unsigned long long)J.slab_cap : 0ull,
                                                             J.d_block_best.p);
            MMO_LAUNCH_CHECK();
            return MMO_OK;
        };
        hbest.resize(blocks);
        unsigned long long n_cnd = 0;
        if (J.k_eff > 0 && thr == INFINITY && blocks >= (unsigned)J.k_eff && blocks <= 4096u) {
            // No running threshold yet (first slab of a scan -- every call of a one-shot scan): the k-th smallest
            // per-block minimum bounds the k-th best score from above, so only the few poses below it travel to the
            // host.  Selected on the device: no round trip before the candidate pass.
            MMO_TRY(reduce_pass(false));
            scan_kth_kernel<<<1, 1024, 0, R.stream>>>(J.d_block_best.p, blocks, (unsigned)J.k_eff, J.two_stage ? 1 : 0, J.d_thr.p);
            MMO_LAUNCH_CHECK();
        } else if (J.k_eff > 0 && thr == INFINITY && blocks >= (unsigned)J.k_eff) {
            MMO_TRY(reduce_pass(false));
            MMO_CUDA(cudaMemcpyAsync(hbest.data(), J.d_block_best.p, blocks * sizeof(ScoreFrame), cudaMemcpyDeviceToHost, R.stream));
            MMO_CUDA(cudaStreamSynchronize(R.stream));
            std::vector<double> mins(blocks);
            for (unsigned b = 0; b < blocks; b++) mins[b] = hbest[b].s;
            std::nth_element(mins.begin(), mins.begin() + (J.k_eff - 1), mins.end());
            double est = mins[J.k_eff - 1];
            if (J.two_stage && est < INFINITY) est = est + 2.0 * fp32_delta(est);
            MMO_CUDA(cudaMemcpyAsync(J.d_thr.p, &est, sizeof(double), cudaMemcpyHostToDevice, R.stream));
        }
        MMO_TRY(reduce_pass(J.k_eff > 0));
        // ONE synchronisation per slab: the slab's best (reduced on the device), the candidate count and the first
        // kCandPrefix candidates travel together into pinned memory; only a slab with more candidates (the first ones
        // of a scan, before the running threshold bites) pays a second round trip
        scan_best_kernel<<<1, 1024, 0, R.stream>>>(J.d_block_best.p, blocks, J.d_block_best.p + blocks);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cudaMemcpyAsync(pin_best, J.d_block_best.p + blocks, sizeof(ScoreFrame), cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaMemcpyAsync(pin_u64 + 1, J.d_counters.p + 1, sizeof n_cnd, cudaMemcpyDeviceToHost, R.stream));
        if (J.k_eff > 0) {
            const size_t pre = (size_t)std::min<int64_t>(kCandPrefix, J.slab_cap);
            MMO_CUDA(cudaMemcpyAsync(pin_cs, J.d_cand_s.p, pre * sizeof(double), cudaMemcpyDeviceToHost, R.stream));
            MMO_CUDA(cudaMemcpyAsync(pin_cf, J.d_cand_f.p, pre * sizeof(long long), cudaMemcpyDeviceToHost, R.stream));
        }
        MMO_CUDA(cudaStreamSynchronize(R.stream));
        n_cnd = pin_u64[1];
        if (getenv("MMO_DEBUG_TIMING")) fprintf(stderr, "[scan slab] %lld poses, %llu candidates (threshold %s)\n", (long long)n_surv, n_cnd, thr == INFINITY ? "estimated" : "running");
        {
            const ScoreFrame b = *pin_best;
            if (b.s < J.best_s || (b.s == J.best_s && b.f < J.best_f && b.s != INFINITY)) { J.best_s = b.s; J.best_f = b.f; }
        }
        if (J.k_eff > 0 && n_cnd > 0) {
            hs.resize(n_cnd); hf.resize(n_cnd);
            const size_t pre = (size_t)std::min<unsigned long long>(n_cnd, (unsigned long long)kCandPrefix);
            memcpy(hs.data(), pin_cs, pre * sizeof(double));
            memcpy(hf.data(), pin_cf, pre * sizeof(long long));
            if (n_cnd > pre) {
                MMO_CUDA(cudaMemcpyAsync(hs.data() + pre, J.d_cand_s.p + pre, (n_cnd - pre) * sizeof(double), cudaMemcpyDeviceToHost, R.stream));
                MMO_CUDA(cudaMemcpyAsync(hf.data() + pre, J.d_cand_f.p + pre, (n_cnd - pre) * sizeof(long long), cudaMemcpyDeviceToHost, R.stream));
                MMO_CUDA(cudaStreamSynchronize(R.stream));
            }
            size_t old = J.top.size();
            J.top.resize(old + n_cnd);
            for (size_t i = 0; i < n_cnd; i++) { J.top[old + i].s = hs[i]; J.top[old + i].f = hf[i]; }
            auto less = [](const ScoreFrame &a, const ScoreFrame &b) {
                // NaN scores sort last; ties to the smaller frame
                bool an = a.s != a.s, bn = b.s != b.s;
                if (an != bn) return bn;
                return (a.s < b.s) || (a.s == b.s && a.f < b.f);
            };
            size_t keep = std::min<size_t>(J.top.size(), (size_t)J.k_eff);
            if (!J.two_stage) {
                // select, then sort the k kept: O(n + k log k) -- a one-shot call sees ~1e4 candidates (estimated threshold)
                if (keep < J.top.size()) std::nth_element(J.top.begin(), J.top.begin() + keep, J.top.end(), less);
                std::sort(J.top.begin(), J.top.begin() + keep, less);
                J.top.resize(keep);
            } else {
                // keep everything within 2*delta of the k-th best fp32 score: a superset of the exact top-k
                std::sort(J.top.begin(), J.top.end(), less);
                if (J.top.size() > keep) {
                    const double tau = J.top[keep - 1].s;
                    const double lim = tau + 2.0 * fp32_delta(tau);
                    size_t m = keep;
                    while (m < J.top.size() && J.top[m].s <= lim) m++;
                    J.top.resize(m);
                }
            }
        }
    }
    return MMO_OK;
}

// second stage of the fp64 scan: the surviving candidates are re-scored by the strict kernel
// (reference arithmetic and summation order), sorted exactly, ties to the smaller frame
static int scan_finalize(ScanJob &J) {
    if (!J.two_stage || J.exact_ready) return MMO_OK;
    const mmo_scan_params &P = J.P;
    Runtime &R = rt();
    J.exact_top.clear();
    J.exact_best_s = INFINITY;
    J.exact_best_f = -1;
    const size_t n = J.top.size();
    if (n > 0) {
        std::vector<int64_t> fr(n);
        for (size_t i = 0; i < n; i++) fr[i] = J.top[i].f;
        DevBuf<int64_t> d_fr;
        DevBuf<double> d_e;
        MMO_TRY(d_fr.upload(fr));
        MMO_TRY(d_e.alloc(n));
        PoseSrc src = {};
        src.kind = 2;
        src.rot9 = J.rs->rot.p;
        src.frames = d_fr.p;
        src.n_rot = P.n_rot;
        for (int d = 0; d < 3; d++) { src.lat_dims[d] = J.dims[d]; src.lat_min[d] = J.mins[d]; src.lat_q[d] = J.q[d]; }
        MMO_TRY(launch_direct_fp64(P.rec, P.lig, P.variant, src, (int64_t)n, d_e.p));
        std::vector<double> e(n);
        MMO_CUDA(cudaMemcpyAsync(e.data(), d_e.p, n * sizeof(double), cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));
        J.exact_top.resize(n);
        for (size_t i = 0; i < n; i++) { J.exact_top[i].s = P.e_intra_const + e[i]; J.exact_top[i].f = fr[i]; }
        std::sort(J.exact_top.begin(), J.exact_top.end(), [](const ScoreFrame &a, const ScoreFrame &b) {
            bool an = a.s != a.s, bn = b.s != b.s;
            if (an != bn) return bn;
            return (a.s < b.s) || (a.s == b.s && a.f < b.f);
        });
        if (J.exact_top[0].s < INFINITY) { J.exact_best_s = J.exact_top[0].s; J.exact_best_f = J.exact_top[0].f; }
        J.exact_top.resize(std::min<size_t>(n, (size_t)std::max(P.topk, 0)));
    }
    J.exact_ready = true;
    return MMO_OK;
}

static void scan_fill_result(const ScanJob &Jc, double *top_scores, int64_t *top_frames, mmo_scan_result *res) {
    const ScanJob &J = Jc;
    const mmo_scan_params &P = J.P;
    const std::vector<ScoreFrame> &top = J.two_stage ? J.exact_top : J.top;
    int ntop = (int)top.size();
    for (int i = 0; i < ntop; i++) {
        if (top_scores) top_scores[i] = top[i].s;
        if (top_frames) top_frames[i] = top[i].f;
    }
    res->n_candidates = J.n_candidates;
    res->n_scored = J.n_scored;
    const double best_s = J.two_stage ? J.exact_best_s : J.best_s;
    const long long best_f = J.two_stage ? J.exact_best_f : J.best_f;
    res->best_score = best_s;
    res->best_frame = best_f;
    res->n_top = ntop;
    for (int d = 0; d < 3; d++) res->lattice_dims[d] = J.dims[d];
    res->best_rot_i = 0;
    res->best_pos[0] = res->best_pos[1] = res->best_pos[2] = 0.0;   // V3.origin when nothing was scored
    if (best_f >= 0) {
        int64_t pt = best_f / P.n_rot;
        res->best_rot_i = (int32_t)(best_f - pt * P.n_rot);
        int xy = J.dims[0] * J.dims[1];
        int k = (int)(pt / xy);
        int j = (int)((pt - (int64_t)k * xy) / J.dims[0]);
        int i = (int)(pt - ((int64_t)k * xy + (int64_t)j * J.dims[0]));
        res->best_pos[0] = J.mins[0] + (double)i * J.q[0];
        res->best_pos[1] = J.mins[1] + (double)j * J.q[1];
        res->best_pos[2] = J.mins[2] + (double)k * J.q[2];
    }
    res->pairs_evaluated = J.pairs_eval;
    res->pairs_inside = J.pairs_in;
    res->device_ms = J.device_ms;
}

static int scan_check(const mmo_scan_params *p) {
    MMO_REQUIRE(p != nullptr, "mmo_scan: null parameters");
    MMO_REQUIRE(p->lig != nullptr, "mmo_scan: null ligand");
    MMO_REQUIRE((p->rec != nullptr) != (p->grid != nullptr), "mmo_scan: give exactly one of rec (direct) or grid (interpolated)");
    MMO_REQUIRE(p->n_rot > 0 && p->rot9 != nullptr, "mmo_scan: no rotations");
    MMO_REQUIRE(p->trans_step > 0.0 && p->roi_r > 0.0, "mmo_scan: bad lattice step or ROI radius");
    MMO_REQUIRE(p->topk >= 0, "mmo_scan: negative top-k");
    MMO_REQUIRE(!p->grid || p->lig->has_typ, "mmo_scan: interpolated scoring needs ligand FF types");
    MMO_REQUIRE(p->variant == MMO_VARIANT_GLOBAL || p->variant == MMO_VARIANT_SHIFTED, "mmo_scan: bad variant");
    MMO_REQUIRE(p->prec == MMO_PREC_FP32 || p->prec == MMO_PREC_FP64, "mmo_scan: bad precision");
    return MMO_OK;
}

}  // namespace mmo

using namespace mmo;

struct mmo_scan_job { ScanJob J; };

// speculative: only for a call that holds the caller's rotation buffer until it returns (mmo_scan)
static int scan_create(const mmo_scan_params *p, int collect_stats, bool speculative, mmo_scan_job **out) {
    MMO_TRY(require_ready());
    MMO_REQUIRE(out != nullptr, "mmo_scan_create: null output pointer");
    *out = nullptr;
    MMO_TRY(scan_check(p));
    mmo_scan_job *h = new mmo_scan_job();
    h->J.P = *p;
    h->J.collect_stats = collect_stats != 0;
    h->J.speculative = speculative;
    int rc = scan_setup(h->J);
    if (rc != MMO_OK) { delete h; return rc; }
    *out = h;
    return MMO_OK;
}

extern "C" {

int mmo_scan_create(const mmo_scan_params *p, int collect_stats, mmo_scan_job **out) try {
    return scan_create(p, collect_stats, false, out);
} MMO_CATCH_ALL

int mmo_scan_num_points(const mmo_scan_job *job, int64_t *n_active_points) try {
    MMO_REQUIRE(job && n_active_points, "mmo_scan_num_points: null pointer");
    *n_active_points = (int64_t)job->J.points.size();
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_scan_run(mmo_scan_job *job, int64_t first_active, int64_t n_active) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(job != nullptr, "mmo_scan_run: null job");
    int64_t n = (int64_t)job->J.points.size();
    int64_t a0 = std::max<int64_t>(0, first_active);
    int64_t a1 = n_active < 0 ? n : std::min(n, a0 + n_active);
    if (a0 >= a1) return MMO_OK;
    Runtime &R = rt();
    cudaEvent_t e0, e1;
    MMO_CUDA(cudaEventCreate(&e0));
    MMO_CUDA(cudaEventCreate(&e1));
    MMO_CUDA(cudaEventRecord(e0, R.stream));
    int rc = scan_run_points(job->J, a0, a1);
    cudaEventRecord(e1, R.stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    job->J.device_ms += ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
} MMO_CATCH_ALL

int mmo_scan_result_get(const mmo_scan_job *job, double *top_scores, int64_t *top_frames, mmo_scan_result *res) try {
    MMO_REQUIRE(job && res, "mmo_scan_result_get: null pointer");
    MMO_TRY(scan_finalize(const_cast<mmo_scan_job *>(job)->J));
    scan_fill_result(job->J, top_scores, top_frames, res);
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_scan_destroy(mmo_scan_job *job) try {
    delete job;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_scan(const mmo_scan_params *p, double *top_scores, int64_t *top_frames, mmo_scan_result *res) try {
    MMO_REQUIRE(res != nullptr, "mmo_scan: null result pointer");
    const bool dbg = getenv("MMO_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    mmo_scan_job *job = nullptr;
    MMO_TRY(scan_create(p, rt().collect_stats ? 1 : 0, true, &job));
    double t1 = now();
    int rc = mmo_scan_run(job, 0, -1);
    if (rc == MMO_OK && job->J.rot_check_wanted) rc = issue_rot_check(job->J);      // nothing was scored: check all the same
    if (job->J.rot_check_pending) {
        // the verdict on the caller's rotation bytes, uploaded and compared behind the kernels
        cudaError_t e = cudaStreamSynchronize(rt().copy_stream);
        job->J.rot_check_pending = false;
        void *stage = nullptr;
        if (e != cudaSuccess) rc = cuda_fail(e, "rotation set check", __FILE__, __LINE__);
        else if (rc == MMO_OK) rc = stage_buffer(&stage);
        if (rc == MMO_OK && *(const int *)((const unsigned long long *)((const char *)stage + kStageHalf) + kPinRotFlag) != 0) {
            // not the resident set after all: drop it and scan again with the caller's rotations
            if (dbg) fprintf(stderr, "[mmo_scan] the rotation set is not the resident one: scanning again\n");
            g_rot_rescans++;
            mmo_scan_destroy(job);
            job = nullptr;
            g_rotset.reset();
            MMO_TRY(scan_create(p, rt().collect_stats ? 1 : 0, false, &job));
            rc = mmo_scan_run(job, 0, -1);
        }
    }
    double t2 = now();
    if (rc == MMO_OK) rc = scan_finalize(job->J);
    if (rc == MMO_OK) scan_fill_result(job->J, top_scores, top_frames, res);
    double t3 = now();
    mmo_scan_destroy(job);
    if (dbg) fprintf(stderr, "[mmo_scan] create %.1f ms, run %.1f ms, finalize %.1f ms, destroy %.1f ms\n", t1 - t0, t2 - t1, t3 - t2, now() - t3);
    return rc;
} MMO_CATCH_ALL

int mmo_scan_rot_rescans(int64_t *count) try {
    MMO_REQUIRE(count != nullptr, "mmo_scan_rot_rescans: null pointer");
    *count = (int64_t)g_rot_rescans;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_scan_set_rot_cache(int mode) try {
    MMO_REQUIRE(mode == 0 || mode == 1, "mmo_scan_set_rot_cache: mode must be 0 or 1");
    g_rot_cache_mode = mode;
    return MMO_OK;
} MMO_CATCH_ALL

// top-k of a device-resident energy list: per-block minima first (their k-th smallest bounds the k-th best from above),
// then only the entries below that bound travel to the host, where they are sorted with the reference's tie rule
int mmo_topk_select_dev(const double *d_E, int64_t n, int32_t k, int64_t id_base, double *out_scores,
                        int64_t *out_ids, int32_t *out_n) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(k > 0 && n >= 0 && out_scores && out_ids && out_n, "mmo_topk_select_dev: bad arguments");
    *out_n = 0;
    if (n == 0) return MMO_OK;
    MMO_REQUIRE(d_E != nullptr, "mmo_topk_select_dev: null energies");
    Runtime &R = rt();
    const unsigned blocks = (unsigned)((n + 255) / 256);
    DevBuf<ScoreFrame> d_bb;
    DevBuf<double> d_thr, d_cs;
    DevBuf<long long> d_cf;
    DevBuf<unsigned long long> d_cn;
    MMO_TRY(d_bb.alloc(blocks)); MMO_TRY(d_thr.alloc(1)); MMO_TRY(d_cn.alloc(1));
    double thr = INFINITY;
    std::vector<ScoreFrame> hb(blocks);
    if (blocks >= (unsigned)k) {
        KernelScope ks(K_REDUCE);
        scan_reduce_kernel<<<blocks, 256, 0, R.stream>>>(d_E, nullptr, id_base, n, 0.0, d_thr.p, nullptr, nullptr, nullptr, 0ull, d_bb.p);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cudaMemcpyAsync(hb.data(), d_bb.p, blocks * sizeof(ScoreFrame), cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));
        std::vector<double> mins(blocks);
        for (unsigned b = 0; b < blocks; b++) mins[b] = hb[b].s;
        std::nth_element(mins.begin(), mins.begin() + (k - 1), mins.end());
        thr = mins[k - 1];
    }
    // capacity: every entry <= thr; thr = +inf (few blocks, or fewer than k finite minima) keeps everything
    const size_t cap = (size_t)n;
    MMO_TRY(d_cs.alloc(cap)); MMO_TRY(d_cf.alloc(cap));
    MMO_CUDA(cudaMemsetAsync(d_cn.p, 0, sizeof(unsigned long long), R.stream));
    MMO_CUDA(cudaMemcpyAsync(d_thr.p, &thr, sizeof(double), cudaMemcpyHostToDevice, R.stream));
    {
        KernelScope ks(K_REDUCE);
        scan_reduce_kernel<<<blocks, 256, 0, R.stream>>>(d_E, nullptr, id_base, n, 0.0, d_thr.p, d_cs.p, d_cf.p, d_cn.p, (unsigned long long)cap, d_bb.p);
        MMO_LAUNCH_CHECK();
    }
    unsigned long long nc = 0;
    MMO_CUDA(cudaMemcpyAsync(&nc, d_cn.p, sizeof nc, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaStreamSynchronize(R.stream));
    std::vector<double> hs(nc);
    std::vector<long long> hf(nc);
    if (nc) {
        MMO_CUDA(cudaMemcpyAsync(hs.data(), d_cs.p, nc * sizeof(double), cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaMemcpyAsync(hf.data(), d_cf.p, nc * sizeof(long long), cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));
    }
    std::vector<ScoreFrame> all(nc);
    for (size_t i = 0; i < nc; i++) { all[i].s = hs[i]; all[i].f = hf[i]; }
    auto less = [](const ScoreFrame &a, const ScoreFrame &b) {
        bool an = a.s != a.s, bn = b.s != b.s;
        if (an != bn) return bn;
        return (a.s < b.s) || (a.s == b.s && a.f < b.f);
    };
    const size_t keep = std::min<size_t>(all.size(), (size_t)k);
    std::partial_sort(all.begin(), all.begin() + keep, all.end(), less);
    for (size_t i = 0; i < keep; i++) { out_scores[i] = all[i].s; out_ids[i] = all[i].f; }
    *out_n = (int32_t)keep;
    return MMO_OK;
} MMO_CATCH_ALL

// K-way merge of per-GPU lists; lists need not be sorted.  Order: score ascending, NaN last, ties to
// the smaller frame (= earlier in the reference's loop order).
int mmo_topk_merge(int32_t n_lists, int32_t k, const double *scores, const int64_t *frames,
                   const int32_t *counts, double *out_scores, int64_t *out_frames, int32_t *out_n) try {
    MMO_REQUIRE(n_lists >= 0 && k >= 0 && out_n, "mmo_topk_merge: bad arguments");
    std::vector<ScoreFrame> all;
    for (int l = 0; l < n_lists; l++) {
        MMO_REQUIRE(counts[l] >= 0 && counts[l] <= k, "mmo_topk_merge: list %d has %d entries (k = %d)", l, counts[l], k);
        for (int i = 0; i < counts[l]; i++) all.push_back({scores[(size_t)l * k + i], (long long)frames[(size_t)l * k + i]});
    }
    auto less = [](const ScoreFrame &a, const ScoreFrame &b) {
        bool an = a.s != a.s, bn = b.s != b.s;
        if (an != bn) return bn;
        return (a.s < b.s) || (a.s == b.s && a.f < b.f);
    };
    std::sort(all.begin(), all.end(), less);
    int n = (int)std::min<size_t>(all.size(), (size_t)k);
    for (int i = 0; i < n; i++) { out_scores[i] = all[i].s; out_frames[i] = all[i].f; }
    *out_n = n;
    return MMO_OK;
} MMO_CATCH_ALL

}  // extern "C"
