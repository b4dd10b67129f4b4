// direct_fp32.cu -- K1: batched UFF Lennard-Jones + Coulomb pair sum, fp32 pair arithmetic.
//
// Replaces the inner loops of Mol.ene_inter_UFF_shifted_brute / _global_brute (src/mol.ml:796-849)
// for many poses per launch.  Layout of the computation (no tensor cores: non-linear pair sum):
//
//   thread  = one pose; a chunk of LJ ligand atoms lives in registers (coordinates + vdW factors)
//   block   = 128 poses; receptor blobs (kBlob Morton-sorted atoms, fp32, relative to the receptor
//             origin) are staged tile by tile in shared memory and read as warp-wide broadcasts
//   cull    = per (warp, ligand chunk, blob): bounding-box distance >= 12 A  ->  blob skipped
//             (shifted variant only; a skipped pair has weight exactly 0 in the reference, mol.ml:836)
//   pair    = 19 issue slots: 3 FADD, FMUL+2 FFMA (r2), FMNMX clamp, MUFU.RSQ, ... see pair_energy()
//   sum     = the LJ pair terms of one receptor atom are summed in fp32 (<= LJ terms), then added to
//             a per-thread fp64 accumulator
//
// Accuracy contract (MMO_PREC_FP32): |E - E_ref| <= max(1e-6 |E_ref|, 1e-4 kcal/mol).  fp32 cannot
// deliver that for close contacts (r^-12), so the fast path clamps r^2 at H = x_max_rec*x_max_lig/kTau
// and a second, sparse kernel (hard_fix_kernel) adds  e64(r) - e64(sqrt(H))  in the reference's own
// double arithmetic for the few pairs with r^2 < H, found through the receptor's voxel lists.
#include "common.cuh"
#include "pose.cuh"
#include <math.h>

namespace mmo {

constexpr int LJ = 8;            // ligand atoms per register chunk
constexpr int TPB = 128;         // poses per block
constexpr int TILE_BLOBS = 32;   // blobs per shared-memory tile
constexpr int TILE_ATOMS = TILE_BLOBS * kBlob;

struct FastArgs {
    int n_blobs;
    int n_atoms;             // real receptor atoms (the last blob may be padded)
    const float4 *xyzq;
    const float2 *ab;
    const float *blob_box;
    double origin[3];
    int L;
    const double *lx, *ly, *lz;
    const float4 *lparam;
    float H;                 // clamp on r^2 (fast path) == close-contact threshold (fix pass)
    unsigned long long *stats;   // [0] pairs evaluated, [1] pairs inside the cut-off (STATS builds)
};

// one receptor atom against one ligand atom; returns w * (EW q_i q_j / r + d_ij (p6^2 - 2 p6))
template <int VARIANT>
__device__ __forceinline__ float pair_energy(float dx, float dy, float dz, float qi, float Ai, float Bi,
                                             float qj, float Aj, float Bj, float H, float acc) {
    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float r2c = fmaxf(r2, H);                       // close contacts are finished in fp64 elsewhere
    float rinv = rsqrtf(r2c);
    float s = rinv * rinv;
    float s3 = s * s * s;
    float v = fmaf(Ai * Aj, s3, -(Bi * Bj));        // (A_i A_j) s^3 - B_i B_j
    float er = (qi * qj) * rinv;                    // qi already carries 332.0637/4
    float e = fmaf(v, s3, er);
    if (VARIANT == MMO_VARIANT_SHIFTED) {
        float u = fmaxf(fmaf(r2c, -1.0f / 144.0f, 1.0f), 0.0f);   // 0 beyond the 12 A cut-off
        return fmaf(u * u, e, acc);
    } else {
        return acc + e;
    }
}

template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(TPB)
direct_fp32_kernel(FastArgs a, PoseSrc src, int64_t n_poses, double *__restrict__ out) {
    __shared__ float4 s_xyzq[TILE_ATOMS];
    __shared__ float2 s_ab[TILE_ATOMS];
    __shared__ float s_box[TILE_BLOBS * 6];
    extern __shared__ float4 s_lparam[];          // L entries {A_j, B_j, q_j, 0}

    const int tid = threadIdx.x;
    const int64_t p = (int64_t)blockIdx.x * TPB + tid;
    const bool valid = p < n_poses;
    const int64_t pp = valid ? p : n_poses - 1;   // idle lanes shadow the last pose, result discarded
    for (int j = tid; j < a.L; j += TPB) s_lparam[j] = a.lparam[j];

    double acc = 0.0;
    unsigned long long n_eval = 0, n_in = 0;
    const int n_chunks = (a.L + LJ - 1) / LJ;
    const int n_tiles = (a.n_blobs + TILE_BLOBS - 1) / TILE_BLOBS;

    for (int c = 0; c < n_chunks; c++) {
        // ---- this pose's chunk of ligand atoms: reference arithmetic in double, then fp32 -----
        float cx[LJ], cy[LJ], cz[LJ], cA[LJ], cB[LJ], cQ[LJ];
        __syncthreads();     // s_lparam visible (first pass); previous tile fully consumed
        {
            PoseRT P;
            if (src.kind != 1) load_pose_rt(src, pp, P);
#pragma unroll
            for (int jj = 0; jj < LJ; jj++) {
                int j = c * LJ + jj;
                if (j < a.L) {
                    double x, y, z;
                    if (src.kind == 1) {
                        x = src.xs[pp * a.L + j]; y = src.ys[pp * a.L + j]; z = src.zs[pp * a.L + j];
                    } else {
                        pose_atom_rt(P, __ldg(a.lx + j), __ldg(a.ly + j), __ldg(a.lz + j), x, y, z);
                    }
                    cx[jj] = (float)(x - a.origin[0]);
                    cy[jj] = (float)(y - a.origin[1]);
                    cz[jj] = (float)(z - a.origin[2]);
                    float4 lp = s_lparam[j];
                    cA[jj] = lp.x; cB[jj] = lp.y; cQ[jj] = lp.z;
                } else {          // padding atom: no charge, no vdW
                    cx[jj] = -1e6f; cy[jj] = -1e6f; cz[jj] = -1e6f;
                    cA[jj] = 0.f; cB[jj] = 0.f; cQ[jj] = 0.f;
                }
            }
        }
        // ---- warp bounding box of the chunk (real atoms only) ---------------------------------
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
        if (VARIANT == MMO_VARIANT_SHIFTED) {
#pragma unroll
            for (int jj = 0; jj < LJ; jj++) {
                if (c * LJ + jj < a.L) {
                    lo[0] = fminf(lo[0], cx[jj]); hi[0] = fmaxf(hi[0], cx[jj]);
                    lo[1] = fminf(lo[1], cy[jj]); hi[1] = fmaxf(hi[1], cy[jj]);
                    lo[2] = fminf(lo[2], cz[jj]); hi[2] = fmaxf(hi[2], cz[jj]);
                }
            }
#pragma unroll
            for (int d = 0; d < 3; d++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
                    hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
                }
            }
        }

        for (int t = 0; t < n_tiles; t++) {
            if (t > 0) __syncthreads();
            const int b0 = t * TILE_BLOBS;
            const int nb = min(TILE_BLOBS, a.n_blobs - b0);
            for (int k = tid; k < nb * kBlob; k += TPB) {
                s_xyzq[k] = __ldg(a.xyzq + (size_t)b0 * kBlob + k);
                s_ab[k] = __ldg(a.ab + (size_t)b0 * kBlob + k);
            }
            for (int k = tid; k < nb * 6; k += TPB) s_box[k] = __ldg(a.blob_box + (size_t)b0 * 6 + k);
            __syncthreads();

            for (int b = 0; b < nb; b++) {
                if (VARIANT == MMO_VARIANT_SHIFTED) {
                    const float *bx = s_box + b * 6;
                    float gx = fmaxf(0.f, fmaxf(bx[0] - hi[0], lo[0] - bx[3]));
                    float gy = fmaxf(0.f, fmaxf(bx[1] - hi[1], lo[1] - bx[4]));
                    float gz = fmaxf(0.f, fmaxf(bx[2] - hi[2], lo[2] - bx[5]));
                    if (fmaf(gz, gz, fmaf(gy, gy, gx * gx)) >= 144.0f) continue;   // warp-uniform
                }
                if (STATS) n_eval += (unsigned long long)min(kBlob, a.n_atoms - (b0 + b) * kBlob) * min(LJ, a.L - c * LJ);
#pragma unroll 2
                for (int i = 0; i < kBlob; i++) {
                    const float4 ra = s_xyzq[b * kBlob + i];
                    const float2 rp = s_ab[b * kBlob + i];
                    float f = 0.f;
#pragma unroll
                    for (int jj = 0; jj < LJ; jj++) {
                        float dx = ra.x - cx[jj], dy = ra.y - cy[jj], dz = ra.z - cz[jj];
                        f = pair_energy<VARIANT>(dx, dy, dz, ra.w, rp.x, rp.y, cQ[jj], cA[jj], cB[jj], a.H, f);
                        if (STATS) {
                            float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            if (r2 < 144.0f && ra.x < 1e5f && c * LJ + jj < a.L) n_in++;
                        }
                    }
                    acc += (double)f;
                }
            }
        }
    }
    if (valid) out[p] = acc;
    if (STATS && valid) {
        atomicAdd(a.stats + 0, n_eval);
        atomicAdd(a.stats + 1, n_in);
    }
}

// ---- close-contact correction in the reference's double arithmetic ----------------------------------
struct FixArgs {
    const double *px, *py, *pz, *pq;     // receptor, original order
    const int32_t *pelt;
    double vox_lo[3], vox_inv;
    int vox_dim[3];
    const int32_t *vox_off, *vox_idx;
    int L;
    const double *lx, *ly, *lz, *lq;
    const int32_t *lelt;
    double H;                            // exactly the fp32 clamp value
    const double *xij, *dij;             // kEltTab^2 tables (UFF.ml:32-51)
    unsigned long long *stats;           // [2] pairs re-evaluated
};

template <int VARIANT>
__device__ __forceinline__ double e64(double r2, double qq, double xij, double dij) {
    // mol.ml:811-815 / 838-845 for one pair
    double r = sqrt(r2);
    if (r < 0.01) r = 0.01;
    double t = xij / r;
    double t2 = t * t;
    double p6 = (t2 * t2) * t2;
    double e = kElecWeight * (qq / r) + dij * ((-2.0 * p6) + (p6 * p6));
    if (VARIANT == MMO_VARIANT_SHIFTED) {
        double w = 0.0;
        if (r < 12.0) { double u = 1.0 - (r / 12.0) * (r / 12.0); w = u * u; }
        e = w * e;
    }
    return e;
}

template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(128)
hard_fix_kernel(FixArgs a, PoseSrc src, int64_t n_poses, double *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_poses) return;
    PoseRT P;
    if (src.kind != 1) load_pose_rt(src, p, P);
    double corr = 0.0;
    unsigned long long n_fix = 0;
    for (int j = 0; j < a.L; j++) {
        double x, y, z;
        if (src.kind == 1) {
            x = src.xs[p * a.L + j]; y = src.ys[p * a.L + j]; z = src.zs[p * a.L + j];
        } else {
            pose_atom_rt(P, __ldg(a.lx + j), __ldg(a.ly + j), __ldg(a.lz + j), x, y, z);
        }
        double fx = (x - a.vox_lo[0]) * a.vox_inv, fy = (y - a.vox_lo[1]) * a.vox_inv, fz = (z - a.vox_lo[2]) * a.vox_inv;
        if (!(fx >= 0.0 && fy >= 0.0 && fz >= 0.0)) continue;
        int vi = (int)fx, vj = (int)fy, vk = (int)fz;
        if (vi >= a.vox_dim[0] || vj >= a.vox_dim[1] || vk >= a.vox_dim[2]) continue;
        size_t v = (size_t)vi + (size_t)vj * a.vox_dim[0] + (size_t)vk * a.vox_dim[0] * a.vox_dim[1];
        int k0 = __ldg(a.vox_off + v), k1 = __ldg(a.vox_off + v + 1);
        if (k0 == k1) continue;
        const double qj = __ldg(a.lq + j);
        const int ej = __ldg(a.lelt + j);
        for (int k = k0; k < k1; k++) {
            int i = __ldg(a.vox_idx + k);
            double dx = __ldg(a.px + i) - x, dy = __ldg(a.py + i) - y, dz = __ldg(a.pz + i) - z;
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < a.H) {
                int t = __ldg(a.pelt + i) * kEltTab + ej;
                double qq = __ldg(a.pq + i) * qj;
                double xij = __ldg(a.xij + t), dij = __ldg(a.dij + t);
                corr += e64<VARIANT>(r2, qq, xij, dij) - e64<VARIANT>(a.H, qq, xij, dij);
                if (STATS) n_fix++;
            }
        }
    }
    out[p] += corr;
    if (STATS) atomicAdd(a.stats + 2, n_fix);
}

// ---- host side -----------------------------------------------------------------------------------
static DevBuf<double> g_xij, g_dij;
static DevBuf<unsigned long long> g_stats;

static int ensure_fix_tables() {
    if (g_xij.p) return MMO_OK;
    std::vector<double> hx(kEltTab * kEltTab), hd(kEltTab * kEltTab);
    for (int a = 0; a < kEltTab; a++)
        for (int b = 0; b < kEltTab; b++) {
            bool ok = a < kNumElt && b < kNumElt;
            hx[a * kEltTab + b] = ok ? sqrt(kEltXi[a] * kEltXi[b]) : NAN;
            hd[a * kEltTab + b] = ok ? sqrt(kEltDi[a] * kEltDi[b]) : NAN;
        }
    MMO_TRY(g_xij.upload(hx));
    MMO_TRY(g_dij.upload(hd));
    MMO_TRY(g_stats.alloc(4));
    return MMO_OK;
}

int launch_direct_fp32(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src,
                       int64_t n_poses, double *d_out, bool collect_stats) {
    MMO_TRY(ensure_fix_tables());
    if (n_poses == 0) return MMO_OK;
    Runtime &R = rt();
    // clamp / close-contact threshold on r^2; a float so that both kernels see the same number
    const float H = (float)(std::max(rec->x_max, 1.0) * std::max(lig->x_max, 1.0) / kTau);
    FastArgs fa;
    fa.n_blobs = rec->n_blobs;
    fa.n_atoms = rec->n;
    fa.xyzq = rec->xyzq.p; fa.ab = rec->ab.p; fa.blob_box = rec->blob_box.p;
    for (int d = 0; d < 3; d++) fa.origin[d] = rec->origin[d];
    fa.L = lig->n;
    fa.lx = lig->x.p; fa.ly = lig->y.p; fa.lz = lig->z.p;
    fa.lparam = lig->fparam.p;
    fa.H = H;
    fa.stats = g_stats.p;
    FixArgs xa;
    xa.px = rec->x.p; xa.py = rec->y.p; xa.pz = rec->z.p; xa.pq = rec->q.p; xa.pelt = rec->elt.p;
    for (int d = 0; d < 3; d++) { xa.vox_lo[d] = rec->vox_lo[d]; xa.vox_dim[d] = rec->vox_dim[d]; }
    xa.vox_inv = 1.0 / rec->vox_edge;
    xa.vox_off = rec->vox_off.p; xa.vox_idx = rec->vox_idx.p;
    xa.L = lig->n;
    xa.lx = lig->x.p; xa.ly = lig->y.p; xa.lz = lig->z.p; xa.lq = lig->q.p; xa.lelt = lig->elt.p;
    xa.H = (double)H;
    xa.xij = g_xij.p; xa.dij = g_dij.p;
    xa.stats = g_stats.p;

    if (collect_stats) MMO_CUDA(cudaMemsetAsync(g_stats.p, 0, 4 * sizeof(unsigned long long), R.stream));
    const unsigned blocks = (unsigned)((n_poses + TPB - 1) / TPB);
    const size_t smem = (size_t)lig->n * sizeof(float4);
    const bool shifted = variant == MMO_VARIANT_SHIFTED;
    if (rec->n > 0) {
        {
        KernelScope ks(K_DIRECT_FP32);
        if (shifted && collect_stats) direct_fp32_kernel<MMO_VARIANT_SHIFTED, true><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, d_out);
        else if (shifted) direct_fp32_kernel<MMO_VARIANT_SHIFTED, false><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, d_out);
        else if (collect_stats) direct_fp32_kernel<MMO_VARIANT_GLOBAL, true><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, d_out);
        else direct_fp32_kernel<MMO_VARIANT_GLOBAL, false><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, d_out);
        }
        MMO_LAUNCH_CHECK();
        KernelScope ks2(K_HARD_FIX);
        const unsigned fblocks = (unsigned)((n_poses + 127) / 128);
        if (shifted && collect_stats) hard_fix_kernel<MMO_VARIANT_SHIFTED, true><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_out);
        else if (shifted) hard_fix_kernel<MMO_VARIANT_SHIFTED, false><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_out);
        else if (collect_stats) hard_fix_kernel<MMO_VARIANT_GLOBAL, true><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_out);
        else hard_fix_kernel<MMO_VARIANT_GLOBAL, false><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_out);
        MMO_LAUNCH_CHECK();
    } else {
        MMO_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n_poses * sizeof(double), R.stream));
    }
    if (collect_stats) {
        unsigned long long h[4];
        MMO_CUDA(cudaMemcpyAsync(h, g_stats.p, sizeof h, cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));
        R.stat_pairs = (int64_t)h[0]; R.stat_inside = (int64_t)h[1]; R.stat_fp64 = (int64_t)h[2];
    }
    return MMO_OK;
}

}  // namespace mmo
