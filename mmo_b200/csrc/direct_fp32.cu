// direct_fp32.cu -- K1: batched UFF Lennard-Jones + Coulomb pair sum, fp32 pair arithmetic.
//
// Replaces the inner loops of Mol.ene_inter_UFF_shifted_brute / _global_brute (src/mol.ml:796-849)
// for many poses per launch.  Layout of the computation (no tensor cores: non-linear pair sum):
//
//   thread  = one pose, block = 256 poses; the receptor (element-sorted k-d groups of 32 atoms, fp32,
//             relative to the receptor origin) is staged in shared memory, the whole ROI receptor at once
//             when it fits
//   cull    = per (warp, ligand atom): the atom's bounding box over the warp's 32 poses (6 CREDUX) is
//             tested first against the group boxes (one ballot per 32 groups), then against the
//             individual atoms of the near groups (one ballot per group); the survivors' coordinates and
//             charge products are compacted into a per-warp structure-of-arrays list.  Shifted variant
//             only: a culled pair has weight exactly 0 in the reference (mol.ml:836)
//   pair    = packed fp32 (FFMA2 / FMUL2 / FADD2 of sm_100a): one packed instruction serves two
//             receptor atoms of the list, 15 packed + 4 FMNMX + 2 MUFU.RSQ per two pairs; the list is
//             consumed 8 atoms at a time (8 LDS.128 with a warp-uniform address) = 4 independent
//             packed dependency chains.  A_i A_j and B_i B_j are loop invariants (one list per element)
//   sum     = fp32 inside a chain for kSumEvery steps, then F2F + DADD into a per-thread fp64 accumulator
//
// Accuracy contract (MMO_PREC_FP32): |E - E_ref| <= max(1e-6 |E_ref|, 1e-4 kcal/mol).  fp32 cannot
// deliver that for close contacts (r^-12), so the fast path clamps r^2 at H = x_max_rec*x_max_lig/kTau
// and a second, sparse kernel (hard_fix_kernel) adds  w(r) e64(r) - w(sqrt H) e64(sqrt H)  in the
// reference's own double arithmetic for the few pairs with r^2 < H, found through the receptor's voxel lists.
#include "common.cuh"
#include "pose.cuh"
#include <math.h>

namespace mmo {

constexpr int LJ = 8;            // ligand atoms per chunk (= one k-d leaf of the ligand)
constexpr int TPB = 256;         // poses per block
constexpr int LIST_CAP = 320;    // per-warp list of near receptor atoms (x, y, z, charge product)
constexpr int MAX_TILE_GROUPS = 64;
constexpr int kSumEvery = 2;     // list steps (of 8 atoms) summed in fp32 before the fp64 accumulation
static_assert(kBlob == 32, "one receptor group per warp-wide test");
static_assert(LIST_CAP % 8 == 0 && LIST_CAP > 128 + 8, "list must take 4 more groups before a flush");

struct FastArgs {
    int n_blobs;             // receptor groups of 32 atoms (k-d leaves, element-sorted)
    int n_types;             // receptor elements present
    int type_g0[kEltTab + 1];    // groups [type_g0[t], type_g0[t+1]) hold element t
    float type_A[kEltTab], type_B[kEltTab];
    const float4 *xyzq;
    const float4 *blob_box;
    double origin[3];
    int L;                   // real ligand atoms
    int n_fast;              // padded to a multiple of LJ
    const double *lx, *ly, *lz;      // template in fast-path order
    const int32_t *forder;           // fast-path position -> original atom index (explicit coordinates)
    const float4 *lparam;            // fast-path order
    float H;                 // clamp on r^2 (fast path) == close-contact threshold (fix pass)
    unsigned long long *stats;   // [0] pairs evaluated, [1] pairs inside the cut-off (STATS builds)
};

// MUFU.RSQ without the denormal-input fix-up sequence rsqrtf() expands to (the argument is >= H > 1)
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// warp-wide min / max of an fp32 value in one instruction (CREDUX, sm_100a), result is warp-uniform
__device__ __forceinline__ float warp_min(float x) {
    float y;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float warp_max(float x) {
    float y;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(y) : "f"(x));
    return y;
}

// Two receptor atoms (the halves of the packed operands) against one ligand atom.
//   SHIFTED: acc += (144 - r^2)^2 * e, with A_iA_j, B_iB_j and q_iq_j pre-divided by 144^2, so that the
//            weight is FF.shift_12A (FF.ml:17-20) and exactly 0 from 12 A on (r^2 is clamped to [H, 144])
//   GLOBAL : acc += e
// e = (A_iA_j s^3 - B_iB_j) s^3 + q_iq_j / r with s = 1/r^2   (= d_ij (p6^2 - 2 p6) + 83.0159 q_i q_j / r)
template <int VARIANT>
__device__ __forceinline__ float2 pair2(float2 X, float2 Y, float2 Z, float2 QQ, float2 nlx, float2 nly, float2 nlz,
                                        float2 AA, float2 nBB, float H, float2 acc, float2 &r2_out) {
    const float2 dx = __fadd2_rn(X, nlx), dy = __fadd2_rn(Y, nly), dz = __fadd2_rn(Z, nlz);
    const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    r2_out = r2;
    float2 r2c;                                      // close contacts are finished in fp64 elsewhere
    if (VARIANT == MMO_VARIANT_SHIFTED) {
        r2c.x = fminf(fmaxf(r2.x, H), 144.0f);
        r2c.y = fminf(fmaxf(r2.y, H), 144.0f);
    } else {
        r2c.x = fmaxf(r2.x, H);
        r2c.y = fmaxf(r2.y, H);
    }
    const float2 rinv = make_float2(rsqrt_fast(r2c.x), rsqrt_fast(r2c.y));
    const float2 s = __fmul2_rn(rinv, rinv);
    const float2 s3 = __fmul2_rn(__fmul2_rn(s, s), s);
    const float2 v = __ffma2_rn(AA, s3, nBB);
    const float2 e = __ffma2_rn(v, s3, __fmul2_rn(QQ, rinv));
    if (VARIANT == MMO_VARIANT_SHIFTED) {
        const float2 up = __ffma2_rn(r2c, make_float2(-1.0f, -1.0f), make_float2(144.0f, 144.0f));
        return __ffma2_rn(__fmul2_rn(up, up), e, acc);
    } else {
        return __fadd2_rn(acc, e);
    }
}

// Shared memory (dynamic): receptor tile float4 {x, y, z, 83.0159*q} [tile_atoms], group boxes
// [2*tile_groups], ligand parameters [n_fast], chunk coordinates x|y|z [LJ][TPB], per-warp lists
// x|y|z|qq [LIST_CAP] (structure of arrays: conflict-free compaction stores, LDS.128 = 4 atoms of one field).
template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(TPB, 2)
direct_fp32_kernel(FastArgs a, PoseSrc src, int64_t n_poses, int tile_groups, double *__restrict__ out) {
    extern __shared__ float4 smem4[];
    const int tile_atoms = tile_groups * kBlob;
    float4 *s_atom = smem4;                                   // tile_atoms
    float4 *s_box = s_atom + tile_atoms;                      // tile_groups * 2
    float4 *s_lparam = s_box + tile_groups * 2;               // n_fast
    float *s_c = (float *)(s_lparam + a.n_fast);              // 3 * LJ * TPB : field-major, then chunk atom, then pose
    float *s_list = s_c + 3 * LJ * TPB + (threadIdx.x >> 5) * (4 * LIST_CAP);
    float *s_lx = s_list, *s_ly = s_list + LIST_CAP, *s_lz = s_list + 2 * LIST_CAP, *s_lq = s_list + 3 * LIST_CAP;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t p = (int64_t)blockIdx.x * TPB + tid;
    const bool valid = p < n_poses;
    const int64_t pp = valid ? p : n_poses - 1;   // idle lanes shadow the last pose, result discarded
    for (int j = tid; j < a.n_fast; j += TPB) s_lparam[j] = a.lparam[j];
    // SHIFTED: the weight (144 - r^2)^2 / 144^2 is split between the pair and the invariants
    const float wscale = VARIANT == MMO_VARIANT_SHIFTED ? 1.0f / 20736.0f : 1.0f;

    double acc = 0.0;
    unsigned long long n_eval = 0, n_in = 0;
    const int n_chunks = a.n_fast / LJ;
    const int n_tiles = (a.n_blobs + tile_groups - 1) / tile_groups;

    for (int t = 0; t < n_tiles; t++) {
        // ---- stage a receptor tile (the whole ROI receptor when it fits: one tile, two barriers) ----
        __syncthreads();
        const int b0 = t * tile_groups;
        const int nb = min(tile_groups, a.n_blobs - b0);
        for (int k = tid; k < nb * kBlob; k += TPB) s_atom[k] = __ldg(a.xyzq + (size_t)b0 * kBlob + k);
        for (int k = tid; k < nb * 2; k += TPB) s_box[k] = __ldg(a.blob_box + (size_t)b0 * 2 + k);
        __syncthreads();

        for (int c = 0; c < n_chunks; c++) {
            // ---- this pose's chunk of ligand atoms: reference arithmetic in double, then fp32 -----
            // (own column of s_c only: no block barrier needed, __syncwarp orders the warp's accesses)
            {
                PoseRT P;
                if (src.kind != 1) load_pose_rt(src, pp, P);
#pragma unroll
                for (int jj = 0; jj < LJ; jj++) {
                    const int k = c * LJ + jj;
                    float vx = 0.f, vy = 0.f, vz = 0.f;
                    if (s_lparam[k].w != 0.f) {
                        double x, y, z;
                        if (src.kind == 1) {
                            const int j = __ldg(a.forder + k);
                            x = src.xs[pp * a.L + j]; y = src.ys[pp * a.L + j]; z = src.zs[pp * a.L + j];
                        } else {
                            pose_atom_rt(P, __ldg(a.lx + k), __ldg(a.ly + k), __ldg(a.lz + k), x, y, z);
                        }
                        vx = (float)(x - a.origin[0]);
                        vy = (float)(y - a.origin[1]);
                        vz = (float)(z - a.origin[2]);
                    }
                    s_c[(0 * LJ + jj) * TPB + tid] = vx;
                    s_c[(1 * LJ + jj) * TPB + tid] = vy;
                    s_c[(2 * LJ + jj) * TPB + tid] = vz;
                }
            }
#pragma unroll 1
            for (int jj = 0; jj < LJ; jj++) {
                const float4 lp = s_lparam[c * LJ + jj];
                if (lp.w == 0.f) continue;                              // padding atom (warp-uniform)
                const float lcx = s_c[(0 * LJ + jj) * TPB + tid], lcy = s_c[(1 * LJ + jj) * TPB + tid],
                            lcz = s_c[(2 * LJ + jj) * TPB + tid];
                const float2 nlx = make_float2(-lcx, -lcx), nly = make_float2(-lcy, -lcy), nlz = make_float2(-lcz, -lcz);
                // bounding box of this ligand atom over the warp's 32 poses
                float lo[3] = {lcx, lcy, lcz}, hi[3] = {lcx, lcy, lcz};
                unsigned long long gmask = ~0ull;                       // near mask of the tile's groups
                if (VARIANT == MMO_VARIANT_SHIFTED) {
                    lo[0] = warp_min(lcx); lo[1] = warp_min(lcy); lo[2] = warp_min(lcz);
                    hi[0] = warp_max(lcx); hi[1] = warp_max(lcy); hi[2] = warp_max(lcz);
                    // ---- level 1: which groups of 32 receptor atoms can be within 12 A of this atom? ----
                    unsigned gm[2];
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const int g = r * 32 + lane;
                        bool near = false;
                        if (g < nb) {
                            const float4 blo = s_box[g * 2], bhi = s_box[g * 2 + 1];
                            float gx = fmaxf(0.f, fmaxf(blo.x - hi[0], lo[0] - bhi.x));
                            float gy = fmaxf(0.f, fmaxf(blo.y - hi[1], lo[1] - bhi.y));
                            float gz = fmaxf(0.f, fmaxf(blo.z - hi[2], lo[2] - bhi.z));
                            near = fmaf(gz, gz, fmaf(gy, gy, gx * gx)) < 144.0f;
                        }
                        gm[r] = __ballot_sync(0xffffffffu, near);
                    }
                    gmask = ((unsigned long long)gm[1] << 32) | gm[0];
                }
                const float qjs = lp.z * wscale;
                // ---- one list per receptor element: A_i A_j, B_i B_j are invariants of the pair loop ----
#pragma unroll 1
                for (int ty = 0; ty < a.n_types; ty++) {
                    const int g_lo = max(a.type_g0[ty], b0) - b0, g_hi = min(a.type_g0[ty + 1], b0 + nb) - b0;
                    if (g_lo >= g_hi) continue;
                    const float AAs = a.type_A[ty] * lp.x * wscale, BBs = a.type_B[ty] * lp.y * wscale;
                    const float2 AA = make_float2(AAs, AAs), nBB = make_float2(-BBs, -BBs);
                    // ---- level 2: per-atom test, 4 groups per step (independent loads and tests), survivors
                    //      compacted into the warp's list; the list is consumed 8 atoms at a time ----
                    int n = 0;
                    for (int g4 = g_lo; g4 < g_hi || n > 0; g4 += 4) {
                        if (g4 < g_hi) {
                            const unsigned m4 = (unsigned)(gmask >> g4) & (0xfu >> max(0, 4 - (g_hi - g4)));
                            if (m4 != 0u) {
                                bool nr[4];
                                float4 pa[4];
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    // (slots past the tile are never selected: clamp the address, keep the predicate)
                                    pa[u] = s_atom[min((g4 + u) * kBlob + lane, tile_atoms - 1)];
                                    nr[u] = ((m4 >> u) & 1u) && pa[u].x < 0.5f * kFarAway;
                                    if (VARIANT == MMO_VARIANT_SHIFTED) {
                                        float gx = fmaxf(0.f, fmaxf(lo[0] - pa[u].x, pa[u].x - hi[0]));
                                        float gy = fmaxf(0.f, fmaxf(lo[1] - pa[u].y, pa[u].y - hi[1]));
                                        float gz = fmaxf(0.f, fmaxf(lo[2] - pa[u].z, pa[u].z - hi[2]));
                                        nr[u] = nr[u] && fmaf(gz, gz, fmaf(gy, gy, gx * gx)) < 144.0f;
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    const unsigned bm = __ballot_sync(0xffffffffu, nr[u]);
                                    if (nr[u]) {
                                        const int slot = n + __popc(bm & lt_mask);
                                        s_lx[slot] = pa[u].x; s_ly[slot] = pa[u].y; s_lz[slot] = pa[u].z;
                                        s_lq[slot] = pa[u].w * qjs;
                                    }
                                    n += __popc(bm);
                                }
                            }
                            if (n <= LIST_CAP - 128 - 8 && g4 + 4 < g_hi) continue;       // room for 4 more groups
                        }
                        if (n == 0) continue;
                        // ---- process the list: 4 packed chains = 8 pairs per step ----
                        if (STATS) n_eval += (unsigned long long)n;
                        const int n8 = (n + 7) & ~7;
                        if (lane < n8 - n) {                                 // pad with far-away, charge-free atoms
                            s_lx[n + lane] = kFarAway; s_ly[n + lane] = kFarAway; s_lz[n + lane] = kFarAway;
                            s_lq[n + lane] = 0.f;
                        }
                        __syncwarp();
                        float2 f[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) f[i] = make_float2(0.f, 0.f);
                        int since = 0;
#pragma unroll 1
                        for (int k = 0; k < n8; k += 8) {
                            const float4 X0 = *(const float4 *)(s_lx + k), X1 = *(const float4 *)(s_lx + k + 4);
                            const float4 Y0 = *(const float4 *)(s_ly + k), Y1 = *(const float4 *)(s_ly + k + 4);
                            const float4 Z0 = *(const float4 *)(s_lz + k), Z1 = *(const float4 *)(s_lz + k + 4);
                            const float4 Q0 = *(const float4 *)(s_lq + k), Q1 = *(const float4 *)(s_lq + k + 4);
                            float2 r2[4];
                            f[0] = pair2<VARIANT>(make_float2(X0.x, X0.y), make_float2(Y0.x, Y0.y), make_float2(Z0.x, Z0.y),
                                                  make_float2(Q0.x, Q0.y), nlx, nly, nlz, AA, nBB, a.H, f[0], r2[0]);
                            f[1] = pair2<VARIANT>(make_float2(X0.z, X0.w), make_float2(Y0.z, Y0.w), make_float2(Z0.z, Z0.w),
                                                  make_float2(Q0.z, Q0.w), nlx, nly, nlz, AA, nBB, a.H, f[1], r2[1]);
                            f[2] = pair2<VARIANT>(make_float2(X1.x, X1.y), make_float2(Y1.x, Y1.y), make_float2(Z1.x, Z1.y),
                                                  make_float2(Q1.x, Q1.y), nlx, nly, nlz, AA, nBB, a.H, f[2], r2[2]);
                            f[3] = pair2<VARIANT>(make_float2(X1.z, X1.w), make_float2(Y1.z, Y1.w), make_float2(Z1.z, Z1.w),
                                                  make_float2(Q1.z, Q1.w), nlx, nly, nlz, AA, nBB, a.H, f[3], r2[3]);
                            if (STATS) {
#pragma unroll
                                for (int i = 0; i < 4; i++) n_in += (r2[i].x < 144.0f) + (r2[i].y < 144.0f);
                            }
                            if (++since == kSumEvery) {
                                const float2 h = __fadd2_rn(__fadd2_rn(f[0], f[1]), __fadd2_rn(f[2], f[3]));
                                acc += (double)(h.x + h.y);
#pragma unroll
                                for (int i = 0; i < 4; i++) f[i] = make_float2(0.f, 0.f);
                                since = 0;
                            }
                        }
                        if (since != 0) {
                            const float2 h = __fadd2_rn(__fadd2_rn(f[0], f[1]), __fadd2_rn(f[2], f[3]));
                            acc += (double)(h.x + h.y);
                        }
                        n = 0;
                        __syncwarp();
                    }
                }
            }
            __syncwarp();    // the warp's s_c columns are rewritten by the next chunk
        }
    }
    if (valid) out[p] = acc;
    if (STATS && valid) {
        atomicAdd(a.stats + 0, n_eval);
        atomicAdd(a.stats + 1, n_in);
    }
}

// ---- close-contact correction (fp64) ---------------------------------------------------------------
// For every pair with r^2 < H the fast path evaluated w(sqrt(H)) e(sqrt(H)); this pass adds
// w(r) e(r) - w(sqrt(H)) e(sqrt(H)) in double.  Same formulas as mol.ml:811-815 / 838-845 (r clamped at 0.01, p6 = (x_ij/r)^6, shift weight),
// written with one reciprocal square root instead of sqrt + two divisions: the result only has to be
// accurate to ~1e-12 relative, not bit-identical (MMO_PREC_FP64 is the bit-identical mode).
struct FixArgs {
    const double4 *pxyzq;                // receptor {x, y, z, q}, original order
    const float4 *pxyz32;                // the same positions in fp32, relative to vox_lo (pre-test only)
    const int32_t *pelt;
    double vox_lo[3], vox_inv;
    int vox_dim[3];
    const int32_t *vox_off, *vox_idx;
    int L;
    const double *lx, *ly, *lz, *lq;
    const int32_t *lelt;
    double H;                            // exactly the fp32 clamp value
    double rinvH;                        // 1/sqrt(H)
    double wH;                           // shift weight at r^2 = H
    const double *xx, *dij, *vdwH;       // kEltTab^2 tables: x_i*x_j, d_ij, d_ij*(p6H^2 - 2 p6H)
    unsigned long long *stats;           // [2] pairs re-evaluated
};

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
        const double qj = kElecWeight * __ldg(a.lq + j);
        const int ej = __ldg(a.lelt + j);
        const float xf = (float)(x - a.vox_lo[0]), yf = (float)(y - a.vox_lo[1]), zf = (float)(z - a.vox_lo[2]);
        const float Hf = (float)a.H + 0.5f;
        for (int k = k0; k < k1; k++) {
            const int i = __ldg(a.vox_idx + k);
            // cheap fp32 pre-test (coordinates relative to the voxel grid corner, error << the 0.5 A^2 margin)
            const float4 rf = __ldg(a.pxyz32 + i);
            const float fdx = rf.x - xf, fdy = rf.y - yf, fdz = rf.z - zf;
            if (fdx * fdx + fdy * fdy + fdz * fdz >= Hf) continue;
            const double2 r01 = __ldg((const double2 *)(a.pxyzq + i));
            const double2 r23 = __ldg((const double2 *)(a.pxyzq + i) + 1);
            const double4 ra = make_double4(r01.x, r01.y, r23.x, r23.y);
            const double dx = ra.x - x, dy = ra.y - y, dz = ra.z - z;
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < a.H) {
                const int t = __ldg(a.pelt + i) * kEltTab + ej;
                const double qq = ra.w * qj;
                const double r2c = fmax(r2, 1e-4);                  // Math.non_zero_dist on r
                const double rinv = rsqrt(r2c);
                const double t2 = __ldg(a.xx + t) * (rinv * rinv);  // (x_ij / r)^2
                const double p6 = t2 * t2 * t2;
                const double e = qq * rinv + __ldg(a.dij + t) * (p6 * p6 - 2.0 * p6);
                const double eH = qq * a.rinvH + __ldg(a.vdwH + t);   // what the fast path evaluated (r clamped at sqrt(H))
                double d;
                if (VARIANT == MMO_VARIANT_SHIFTED) {
                    const double u = 1.0 - r2c * (1.0 / 144.0);
                    d = (u * u) * e - a.wH * eH;                       // the fast path clamped r^2 in the weight too
                } else {
                    d = e - eH;
                }
                corr += d;
                if (STATS) n_fix++;
            }
        }
    }
    out[p] += corr;
    if (STATS) atomicAdd(a.stats + 2, n_fix);
}

// ---- host side -----------------------------------------------------------------------------------
static DevBuf<double> g_xx, g_dij, g_vdwH;
static DevBuf<unsigned long long> g_stats;
static double g_vdwH_for = -1.0;

static int ensure_fix_tables(double H) {
    if (!g_xx.p) {
        std::vector<double> hx(kEltTab * kEltTab), hd(kEltTab * kEltTab);
        for (int a = 0; a < kEltTab; a++)
            for (int b = 0; b < kEltTab; b++) {
                bool ok = a < kNumElt && b < kNumElt;
                hx[a * kEltTab + b] = ok ? kEltXi[a] * kEltXi[b] : NAN;
                hd[a * kEltTab + b] = ok ? sqrt(kEltDi[a] * kEltDi[b]) : NAN;
            }
        MMO_TRY(g_xx.upload(hx));
        MMO_TRY(g_dij.upload(hd));
        MMO_TRY(g_stats.alloc(4));
    }
    if (g_vdwH_for != H) {
        std::vector<double> hv(kEltTab * kEltTab);
        for (int a = 0; a < kEltTab; a++)
            for (int b = 0; b < kEltTab; b++) {
                bool ok = a < kNumElt && b < kNumElt;
                double t2 = ok ? kEltXi[a] * kEltXi[b] / H : NAN;
                double p6 = t2 * t2 * t2;
                hv[a * kEltTab + b] = ok ? sqrt(kEltDi[a] * kEltDi[b]) * (p6 * p6 - 2.0 * p6) : NAN;
            }
        MMO_TRY(g_vdwH.upload(hv));
        g_vdwH_for = H;
    }
    return MMO_OK;
}

static int set_fast_smem(size_t smem) {
    static size_t done = 0;
    if (smem <= done) return MMO_OK;
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = smem;
    return MMO_OK;
}

int launch_direct_fp32(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src,
                       int64_t n_poses, double *d_out, bool collect_stats) {
    if (n_poses == 0) return MMO_OK;
    Runtime &R = rt();
    // clamp / close-contact threshold on r^2; a float so that both kernels see the same number
    const float H = (float)(std::max(rec->x_max, 1.0) * std::max(lig->x_max, 1.0) / kTau);
    MMO_TRY(ensure_fix_tables((double)H));
    FastArgs fa;
    fa.n_blobs = rec->n_blobs;
    fa.n_types = rec->n_types;
    for (int t = 0; t <= kEltTab; t++) fa.type_g0[t] = t <= rec->n_types ? rec->type_g0[t] : rec->n_blobs;
    for (int t = 0; t < kEltTab; t++) { fa.type_A[t] = rec->type_A[t]; fa.type_B[t] = rec->type_B[t]; }
    fa.xyzq = rec->xyzq.p; fa.blob_box = rec->blob_box.p;
    for (int d = 0; d < 3; d++) fa.origin[d] = rec->origin[d];
    fa.L = lig->n;
    fa.n_fast = lig->n_fast;
    fa.lx = lig->fx.p; fa.ly = lig->fy.p; fa.lz = lig->fz.p;
    fa.forder = lig->forder.p;
    fa.lparam = lig->fparam.p;
    fa.H = H;
    fa.stats = g_stats.p;
    FixArgs xa;
    xa.pxyzq = rec->xyzq64.p; xa.pxyz32 = rec->xyz32v.p; xa.pelt = rec->elt.p;
    for (int d = 0; d < 3; d++) { xa.vox_lo[d] = rec->vox_lo[d]; xa.vox_dim[d] = rec->vox_dim[d]; }
    xa.vox_inv = 1.0 / rec->vox_edge;
    xa.vox_off = rec->vox_off.p; xa.vox_idx = rec->vox_idx.p;
    xa.L = lig->n;
    xa.lx = lig->x.p; xa.ly = lig->y.p; xa.lz = lig->z.p; xa.lq = lig->q.p; xa.lelt = lig->elt.p;
    xa.H = (double)H;
    xa.rinvH = 1.0 / sqrt((double)H);
    xa.wH = (1.0 - (double)H / 144.0) * (1.0 - (double)H / 144.0);
    xa.xx = g_xx.p; xa.dij = g_dij.p; xa.vdwH = g_vdwH.p;
    xa.stats = g_stats.p;

    if (collect_stats) MMO_CUDA(cudaMemsetAsync(g_stats.p, 0, 4 * sizeof(unsigned long long), R.stream));
    const unsigned blocks = (unsigned)((n_poses + TPB - 1) / TPB);
    // receptor tile: everything when it fits (<= 64 groups = 2048 atoms), so that 2 blocks stay resident per SM
    const int tile_blobs = std::max(1, std::min(rec->n_blobs, MAX_TILE_GROUPS));
    const size_t smem = ((size_t)tile_blobs * kBlob + (size_t)tile_blobs * 2 + (size_t)lig->n_fast) * sizeof(float4) +
                        (size_t)3 * LJ * TPB * sizeof(float) + (size_t)(TPB / 32) * 4 * LIST_CAP * sizeof(float);
    const bool shifted = variant == MMO_VARIANT_SHIFTED;
    if (rec->n > 0) {
        MMO_TRY(set_fast_smem(smem));
        {
        KernelScope ks(K_DIRECT_FP32);
        if (shifted && collect_stats) direct_fp32_kernel<MMO_VARIANT_SHIFTED, true><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, tile_blobs, d_out);
        else if (shifted) direct_fp32_kernel<MMO_VARIANT_SHIFTED, false><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, tile_blobs, d_out);
        else if (collect_stats) direct_fp32_kernel<MMO_VARIANT_GLOBAL, true><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, tile_blobs, d_out);
        else direct_fp32_kernel<MMO_VARIANT_GLOBAL, false><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, tile_blobs, d_out);
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
