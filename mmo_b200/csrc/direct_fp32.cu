// direct_fp32.cu -- K1: batched UFF Lennard-Jones + Coulomb pair sum, fp32 pair arithmetic.
//
// Replaces the inner loops of Mol.ene_inter_UFF_shifted_brute / _global_brute (src/mol.ml:796-849)
// for many poses per launch.  Layout of the computation (no tensor cores: non-linear pair sum):
//
//   thread  = two poses (p and p + 32), warp = 64 consecutive poses, one persistent block of 16 warps per SM that
//             draws (64 poses, chunk split, receptor slice) work units from an atomic counter; the receptor
//             (k-d groups of 16 atoms, fp32, relative to the receptor origin) is staged in shared
//             memory when it fits (<= 2048 atoms), else read through L1 with a two-stage level 1 (one launch
//             either way); small batches (single-pose calls) also split the receptor into slices
//   cull    = per (warp, ligand atom): centre c and radius rho of the atom's positions over the warp's 64
//             poses (CREDUX); level 1: lane g tests the box of group g against the sphere (c, 12 + rho),
//             one ballot per 32 groups; level 2: the atoms of four near groups per step are tested two
//             per lane, and the survivors' centred coordinates x' = x - c, |x'|^2, charge product and
//             vdW products are compacted into a per-warp structure-of-arrays list.  Shifted variant
//             only: a culled pair has weight exactly 0 in the reference (mol.ml:836)
//   pair    = packed fp32 (FFMA2 / FMUL2 / FADD2 of sm_100a): one packed instruction serves two
//             receptor atoms of the list; r^2 = |x'|^2 + |l'|^2 - 2 x'.l' (l' = ligand atom - c, short
//             vectors: no harmful cancellation while rho is small, else the difference form is used);
//             12 packed + 2 FFMA.SAT + 2 FMNMX + 2 MUFU.RSQ per two pairs.  The list is consumed 4 atoms at a time
//             (7 LDS.128 with a warp-uniform address) for both poses of the thread = 4 independent packed chains
//   sum     = fp32 inside a chain for kSumEvery steps, then F2F + DADD into per-pose fp64 accumulators
//
// Accuracy contract (MMO_PREC_FP32): |E - E_ref| <= max(1e-6 |E_ref|, 1e-4 kcal/mol).  fp32 cannot
// deliver that for close contacts (r^-12), so the fast path clamps r^2 per ligand atom at H_j = x_j*x_max_rec/kTau
// and a second, sparse kernel (hard_fix_kernel: thread = pose; pose_fix_block_kernel: block = pose, for small batches)
// adds  w(r) e64(r) - w(sqrt H) e64(sqrt H)  in double for the few pairs with r^2 < H, found through the receptor's
// voxel lists.
//
// Item mode (further down): incoherent pose lists (conformer screens) are cut into (pose, ligand atom) items, sorted by
// Morton cell, and run through the same cull -> list -> packed pair loop with 64 spatially neighbouring items per warp.
#include "common.cuh"
#include "pose.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

namespace mmo {

constexpr int LJ = 4;            // ligand atoms per chunk (half a k-d leaf of the ligand)
constexpr int kFixLJ = LJ;
// hard_fix_kernel tunables (measured on C2, profiles/README.md): -D overrides are for tools/variants.sh experiments
#ifndef MMO_FIX_BLOCKS
#define MMO_FIX_BLOCKS 10
#endif
#ifndef MMO_FIX_UNROLL
#define MMO_FIX_UNROLL 4
#endif
#ifndef MMO_FIX_MARGIN
#define MMO_FIX_MARGIN 0.01f
#endif
constexpr int kFixUnroll = MMO_FIX_UNROLL;           // pre-test loop of hard_fix_kernel
constexpr int kFixBlocksPerSM = MMO_FIX_BLOCKS;   // hard_fix_kernel: 128-thread blocks resident per SM (caps the registers)
constexpr int TPB = 512;         // threads per block (one persistent block per SM: one copy of the receptor tile)
constexpr int kBlocksPerSM = 1;
constexpr int PPT = 2;           // poses per thread
constexpr int PPB = TPB * PPT;   // poses per block
constexpr int LIST_CAP = 256;    // per-warp list of near receptor atoms
constexpr int NF = 7;            // list fields: x', y', z', |x'|^2, q_i q_j, A_i A_j, -B_i B_j
constexpr int MAX_TILE_GROUPS = 128;
#ifndef MMO_SUM_EVERY
#define MMO_SUM_EVERY 32
#endif
#ifndef MMO_K1_UNROLL
#define MMO_K1_UNROLL 1
#endif
constexpr int kK1Unroll = MMO_K1_UNROLL;    // pair-loop steps per iteration (tuning)
constexpr int kSumEvery = MMO_SUM_EVERY;    // list steps (of 4 atoms) summed in fp32 before the fp64 accumulation
constexpr float kRhoExpand2 = 20.25f;  // the expanded form of r^2 is used while rho <= 4.5 A
static_assert(kBlob == 16, "two receptor groups per warp-wide test");
static_assert(LIST_CAP % 4 == 0 && LIST_CAP >= 128, "list must take one more step (64 atoms) before a flush");
constexpr int NEAR_CAP = MAX_TILE_GROUPS + 8;   // per-warp list of near group ids (bytes)

struct FastArgs {
    int n_blobs;             // receptor groups of 16 atoms (k-d leaves)
    float tab_A[kEltTab], tab_B[kEltTab];   // vdW factors by compact element index
    const float4 *xyzq;
    const uint8_t *gelt;
    const float4 *blob_box;
    const float4 *sup_box;       // boxes of 32 consecutive groups (untiled mode: receptors of more than 2048 atoms)
    int n_sup;
    double origin[3];
    int L;                   // real ligand atoms
    int n_fast;              // padded to a multiple of 8
    const double *lx, *ly, *lz;      // template in fast-path order
    const int32_t *forder;           // fast-path position -> original atom index (explicit coordinates)
    const float4 *lparam;            // fast-path order
    float hscale;            // x_max(receptor) / kTau: the clamp on r^2 (fast path) == close-contact threshold (fix pass) of
                             // ligand atom j is H_j = hscale * x_j (lparam.w); atoms with a pair below H_j + margin are flagged
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
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float y;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}

// Two receptor atoms (the halves of the packed operands) against one ligand atom of one pose.
//   EXPAND : r^2 = (|x'|^2 + |l'|^2) - 2 x'.l'   (m2 = -2 l', l2 = |l'|^2)         4 packed ops
//   else   : r^2 = |x' - l'|^2                   (m2 = -l')                        6 packed ops
//   SHIFTED: acc += sat(1 - r^2/144)^2 * e: FF.shift_12A (FF.ml:17-20), exactly 0 from 12 A on through the saturation,
//            so r^2 is only clamped from below (at H)
//   GLOBAL : acc += e
// e = (A_iA_j s^3 - B_iB_j) s^3 + q_iq_j / r with s = 1/r^2   (= d_ij (p6^2 - 2 p6) + 83.0159 q_i q_j / r)
template <int VARIANT, bool EXPAND>
__device__ __forceinline__ float2 pair2(float2 X, float2 Y, float2 Z, float2 S, float2 QQ, float2 AA, float2 nBB,
                                        float m2x, float m2y, float m2z, float l2, float H, float2 acc, float2 &r2_out) {
    float2 r2;
    if (EXPAND) {
        r2 = __ffma2_rn(X, bc2(m2x), __ffma2_rn(Y, bc2(m2y), __ffma2_rn(Z, bc2(m2z), __fadd2_rn(S, bc2(l2)))));
    } else {
        const float2 dx = __fadd2_rn(X, bc2(m2x)), dy = __fadd2_rn(Y, bc2(m2y)), dz = __fadd2_rn(Z, bc2(m2z));
        r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
    }
    r2_out = r2;
    float2 r2c;                                      // close contacts are finished in fp64 elsewhere
    r2c.x = fmaxf(r2.x, H);                          // no upper clamp: the shift weight saturates at 0
    r2c.y = fmaxf(r2.y, H);
    const float2 rinv = make_float2(rsqrt_fast(r2c.x), rsqrt_fast(r2c.y));
    const float2 s = __fmul2_rn(rinv, rinv);
    const float2 s3 = __fmul2_rn(__fmul2_rn(s, s), s);
    const float2 v = __ffma2_rn(AA, s3, nBB);
    const float2 e = __ffma2_rn(v, s3, __fmul2_rn(QQ, rinv));
    if (VARIANT == MMO_VARIANT_SHIFTED) {
        // FF.shift_12A as two scalar saturating FMAs: sat(1 - r^2/144) is exactly 0 from 12 A on, which makes the
        // upper clamp of r^2 (two FMNMX) unnecessary; same FMA-pipe time as one packed FMA
        const float2 up = make_float2(fma_sat(r2c.x, -1.0f / 144.0f, 1.0f), fma_sat(r2c.y, -1.0f / 144.0f, 1.0f));
        return __ffma2_rn(__fmul2_rn(up, up), e, acc);
    } else {
        return __fadd2_rn(acc, e);
    }
}

// consume the warp's list: n4 entries (a multiple of 4, padded), both poses of the thread
template <int VARIANT, bool EXPAND, bool STATS>
__device__ __forceinline__ void run_list(const float *s_l, int n, int n4, const float (&m2x)[PPT], const float (&m2y)[PPT],
                                         const float (&m2z)[PPT], const float (&l2)[PPT], float H, double (&acc)[PPT],
                                         float (&rmin)[PPT], unsigned long long (&n_in)[PPT]) {
    // blocks of kSumEvery steps: fp32 partial sums inside a block, one F2F + DADD per pose at its end
#pragma unroll 1
    for (int k0 = 0; k0 < n4; k0 += 4 * kSumEvery) {
        const int kend = min(n4, k0 + 4 * kSumEvery);
        float2 f[PPT][2];
#pragma unroll
        for (int h = 0; h < PPT; h++) f[h][0] = f[h][1] = make_float2(0.f, 0.f);
#pragma unroll kK1Unroll
        for (int k = k0; k < kend; k += 4) {
            const float4 X = *(const float4 *)(s_l + 0 * LIST_CAP + k), Y = *(const float4 *)(s_l + 1 * LIST_CAP + k);
            const float4 Z = *(const float4 *)(s_l + 2 * LIST_CAP + k), S = *(const float4 *)(s_l + 3 * LIST_CAP + k);
            const float4 Q = *(const float4 *)(s_l + 4 * LIST_CAP + k), A = *(const float4 *)(s_l + 5 * LIST_CAP + k);
            const float4 B = *(const float4 *)(s_l + 6 * LIST_CAP + k);
#pragma unroll
            for (int h = 0; h < PPT; h++) {
                float2 ra, rb;
                f[h][0] = pair2<VARIANT, EXPAND>(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y),
                                                 make_float2(S.x, S.y), make_float2(Q.x, Q.y), make_float2(A.x, A.y),
                                                 make_float2(B.x, B.y), m2x[h], m2y[h], m2z[h], l2[h], H, f[h][0], ra);
                f[h][1] = pair2<VARIANT, EXPAND>(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w),
                                                 make_float2(S.z, S.w), make_float2(Q.z, Q.w), make_float2(A.z, A.w),
                                                 make_float2(B.z, B.w), m2x[h], m2y[h], m2z[h], l2[h], H, f[h][1], rb);
                rmin[h] = fminf(fminf(rmin[h], fminf(ra.x, ra.y)), fminf(rb.x, rb.y));   // close contact seen? (fix pass)
                if (STATS) n_in[h] += (ra.x < 144.0f && k < n) + (ra.y < 144.0f && k + 1 < n) + (rb.x < 144.0f && k + 2 < n) +
                                      (rb.y < 144.0f && k + 3 < n);
            }
        }
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            const float2 t = __fadd2_rn(f[h][0], f[h][1]);
            acc[h] += (double)(t.x + t.y);
        }
    }
}

// Shared memory (dynamic): receptor tile float4 {x, y, z, 83.0159*q} [tile_atoms], group boxes
// [2*tile_groups], ligand parameters [n_fast], vdW table [16] float2, chunk coordinates x|y|z [LJ][PPB],
// per-warp lists [NF][LIST_CAP] (structure of arrays: conflict-free compaction stores, LDS.128 = 4 atoms
// of one field), element bytes [tile_atoms].
// TILED: the (ROI) receptor fits one shared-memory tile (<= 2048 atoms: C2).  Otherwise the groups are read through L1
// (read-only path) and culled in two stages over super-groups of 32, as in the item kernel: ONE launch for a receptor of
// any size instead of one per tile, each of which decoded the poses and re-derived the ligand coordinates again.
template <int VARIANT, bool STATS, bool TILED, bool SLICED>
__global__ void __launch_bounds__(TPB, kBlocksPerSM)
direct_fp32_kernel(FastArgs a, PoseSrc src, int64_t n_poses, int b0, int nb, int tile_groups, int n_split, int n_slices,
                   int near_cap, unsigned long long *__restrict__ work, double *__restrict__ out, uint8_t *__restrict__ flags) {
    // Persistent blocks (1 per SM), one launch for the receptor groups [b0, b0 + nb).  A work unit is
    // (64 consecutive poses, chunk split y, receptor slice sl): the warp that draws it sums the ligand chunks c = y,
    // y + n_split, ... of those poses against the groups of slice sl into out[(sl * n_split + y) * n_poses + p]; units are
    // handed out through one atomic counter, so warps never wait for each other and the tail of the launch is one unit
    // long.  The parts are added in a fixed order by the fix kernel.  n_slices > 1 only for small batches (a single
    // pose is 12 chunks x 8 slices = 96 warps instead of 12): the slices are what the fix kernels call tiles.
    extern __shared__ float4 smem4[];
    const int tile_atoms = tile_groups * kBlob;
    float4 *s_atom = smem4;                                   // tile_atoms + kBlob (a dummy group of far-away atoms)
    float4 *s_box = s_atom + tile_atoms + kBlob;              // tile_groups * 2
    float4 *s_lparam = s_box + tile_groups * 2;               // n_fast
    float2 *s_tab = (float2 *)(s_lparam + a.n_fast);          // 16
    float *s_c = (float *)(s_tab + 16);                       // 3 * LJ * PPB : field-major, then chunk atom, then pose
    float *s_l = s_c + 3 * LJ * PPB + (threadIdx.x >> 5) * (NF * LIST_CAP);
    uint8_t *s_elt = (uint8_t *)(s_c + 3 * LJ * PPB + (TPB / 32) * (NF * LIST_CAP));      // tile_atoms + kBlob
    uint16_t *s_near = (uint16_t *)(s_elt + ((tile_atoms + kBlob + 15) & ~15)) + (threadIdx.x >> 5) * near_cap;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int col0 = (tid >> 5) * (32 * PPT) + lane;          // this thread's columns of s_c: col0 + 32 h
    // ---- stage the receptor tile once per block ----
    for (int j = tid; j < a.n_fast; j += TPB) s_lparam[j] = a.lparam[j];
    if (tid < kEltTab) s_tab[tid] = make_float2(a.tab_A[tid], a.tab_B[tid]);
    if (TILED) {
        for (int k = tid; k < nb * kBlob; k += TPB) {
            s_atom[k] = __ldg(a.xyzq + (size_t)b0 * kBlob + k);
            s_elt[k] = __ldg(a.gelt + (size_t)b0 * kBlob + k);
        }
        for (int k = tid; k < nb * 2; k += TPB) s_box[k] = __ldg(a.blob_box + (size_t)b0 * 2 + k);
    }
    if (tid < kBlob) {
        s_atom[tile_atoms + tid] = make_float4(kFarAway, kFarAway, kFarAway, 0.f);
        s_elt[tile_atoms + tid] = 0;
    }
    __syncthreads();
    const int n_chunks = a.n_fast / LJ;
    const unsigned long long n_groups = (unsigned long long)((n_poses + 32 * PPT - 1) / (32 * PPT));
    const unsigned long long n_units = n_groups * (unsigned long long)n_split * (unsigned long long)n_slices;
    unsigned long long n_eval = 0, n_in_tot = 0;
#ifdef MMO_EXPERIMENT_NOCULL
    bool exp_have_list = false;
#endif

    for (;;) {
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(work, 1ull);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const int yz = (int)(u / n_groups);
        // SLICED is a template parameter so that the throughput path (one slice) keeps its code: measured, the run-time
        // form cost direct_fp32_kernel 2 % on C2
        const int sl = SLICED ? yz / n_split : 0, y = yz - sl * n_split;
        const int g_lo = SLICED ? nb * sl / n_slices : 0, g_hi = SLICED ? nb * (sl + 1) / n_slices : nb;    // this unit's groups
        const int64_t p0 = (int64_t)(u - (unsigned long long)yz * n_groups) * (32 * PPT) + lane;
        int64_t pp[PPT];
        bool valid[PPT];
        double acc[PPT];
        unsigned long long n_in[PPT];
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            const int64_t p = p0 + 32 * h;
            valid[h] = p < n_poses;
            pp[h] = valid[h] ? p : n_poses - 1;       // idle slots shadow the last pose, result discarded
            acc[h] = 0.0;
            n_in[h] = 0;
        }

        for (int c = y; c < n_chunks; c += n_split) {

            // ---- this thread's poses, chunk of ligand atoms: reference arithmetic in double, then fp32 ----
            // (own columns of s_c only: no block barrier needed, __syncwarp orders the warp's accesses)
#pragma unroll
            for (int h = 0; h < PPT; h++) {
                PoseRT P;
                if (src.kind != 1) load_pose_rt(src, pp[h], P);
#pragma unroll
                for (int jj = 0; jj < LJ; jj++) {
                    const int k = c * LJ + jj;
                    float vx = 0.f, vy = 0.f, vz = 0.f;
                    if (s_lparam[k].w != 0.f) {
                        double x, y, z;
                        if (src.kind == 1) {
                            const int j = __ldg(a.forder + k);
                            x = src.xs[pp[h] * a.L + j]; y = src.ys[pp[h] * a.L + j]; z = src.zs[pp[h] * a.L + j];
                        } else {
                            pose_atom_rt(P, __ldg(a.lx + k), __ldg(a.ly + k), __ldg(a.lz + k), x, y, z);
                        }
                        vx = (float)(x - a.origin[0]);
                        vy = (float)(y - a.origin[1]);
                        vz = (float)(z - a.origin[2]);
                    }
                    s_c[(0 * LJ + jj) * PPB + col0 + 32 * h] = vx;
                    s_c[(1 * LJ + jj) * PPB + col0 + 32 * h] = vy;
                    s_c[(2 * LJ + jj) * PPB + col0 + 32 * h] = vz;
                }
            }
            unsigned cbits[PPT];     // bit jj: atom jj of the chunk has a receptor atom within sqrt(H) (+ margin)
#pragma unroll
            for (int h = 0; h < PPT; h++) cbits[h] = 0u;
#pragma unroll 1
            for (int jj = 0; jj < LJ; jj++) {
                const float4 lp = s_lparam[c * LJ + jj];
                if (lp.w == 0.f) continue;                              // padding atom (warp-uniform)
                float rmin[PPT];
#pragma unroll
                for (int h = 0; h < PPT; h++) rmin[h] = 3e38f;
                float px[PPT], py[PPT], pz[PPT];
#pragma unroll
                for (int h = 0; h < PPT; h++) {
                    px[h] = s_c[(0 * LJ + jj) * PPB + col0 + 32 * h];
                    py[h] = s_c[(1 * LJ + jj) * PPB + col0 + 32 * h];
                    pz[h] = s_c[(2 * LJ + jj) * PPB + col0 + 32 * h];
                }
                // ---- centre and radius of this ligand atom's positions over the warp's 64 poses ----
                const float cx = 0.5f * (warp_min(fminf(px[0], px[1])) + warp_max(fmaxf(px[0], px[1])));
                const float cy = 0.5f * (warp_min(fminf(py[0], py[1])) + warp_max(fmaxf(py[0], py[1])));
                const float cz = 0.5f * (warp_min(fminf(pz[0], pz[1])) + warp_max(fmaxf(pz[0], pz[1])));
                float l2[PPT], m2x[PPT], m2y[PPT], m2z[PPT];
#pragma unroll
                for (int h = 0; h < PPT; h++) {
                    m2x[h] = px[h]; m2y[h] = py[h]; m2z[h] = pz[h];     // l (kept for the difference form)
                    px[h] -= cx; py[h] -= cy; pz[h] -= cz;              // l' = l - c
                    l2[h] = fmaf(pz[h], pz[h], fmaf(py[h], py[h], px[h] * px[h]));
                }
                const float rho2 = warp_max(fmaxf(l2[0], l2[1]));
                const bool expand = rho2 <= kRhoExpand2;
                const float reach = 12.0f + sqrtf(rho2) * 1.0001f + 1e-4f;
                const float reach2 = reach * reach;
#pragma unroll
                for (int h = 0; h < PPT; h++) {
                                        // expanded form: list and ligand atom relative to c; difference form (incoherent warps, c may be far
                    // from everything): both stay relative to the receptor origin, no second rounding
                    m2x[h] = expand ? -2.0f * px[h] : -m2x[h];
                    m2y[h] = expand ? -2.0f * py[h] : -m2y[h];
                    m2z[h] = expand ? -2.0f * pz[h] : -m2z[h];
                }
#ifdef MMO_EXPERIMENT_NOCULL
                // ceiling experiment (WRONG energies): after a warp's first list, every atom re-uses the stale list, 2 x 224
                // entries, i.e. the pair work of an average C2 atom without any culling -- what perfect overlap could reach
                if (exp_have_list) {
                    run_list<VARIANT, true, STATS>(s_l, 224, 224, m2x, m2y, m2z, l2, a.hscale * lp.w, acc, rmin, n_in);
                    run_list<VARIANT, true, STATS>(s_l, 224, 224, m2x, m2y, m2z, l2, a.hscale * lp.w, acc, rmin, n_in);
                    continue;
                }
                exp_have_list = true;
#endif
                // ---- level 1: which groups of 16 receptor atoms can be within 12 A of these positions?
                //      lane g tests group box g; the near group ids are compacted into s_near ----
                int ng = 0;
                __syncwarp();        // every lane is done reading the previous atom's s_near
                if (TILED) {
#pragma unroll
                    for (int r = 0; r < MAX_TILE_GROUPS / 32; r++) {
                        const int g = r * 32 + lane;
                        bool near = g >= g_lo && g < g_hi;
                        if (VARIANT == MMO_VARIANT_SHIFTED && near) {
                            const float4 blo = s_box[g * 2], bhi = s_box[g * 2 + 1];
                            const float gx = fmaxf(0.f, fmaxf(blo.x - cx, cx - bhi.x));
                            const float gy = fmaxf(0.f, fmaxf(blo.y - cy, cy - bhi.y));
                            const float gz = fmaxf(0.f, fmaxf(blo.z - cz, cz - bhi.z));
                            near = fmaf(gz, gz, fmaf(gy, gy, gx * gx)) < reach2;
                        }
                        const unsigned gm = __ballot_sync(0xffffffffu, near);
                        if (near) s_near[ng + __popc(gm & lt_mask)] = (uint16_t)g;
                        ng += __popc(gm);
                    }
                    if (lane < 4) s_near[ng + lane] = (uint16_t)tile_groups;        // pad with the dummy group
                } else {
                    auto box_near = [&](const float4 blo, const float4 bhi) {
                        const float gx = fmaxf(0.f, fmaxf(blo.x - cx, cx - bhi.x));
                        const float gy = fmaxf(0.f, fmaxf(blo.y - cy, cy - bhi.y));
                        const float gz = fmaxf(0.f, fmaxf(blo.z - cz, cz - bhi.z));
                        return fmaf(gz, gz, fmaf(gy, gy, gx * gx)) < reach2;
                    };
                    for (int s0 = 0; s0 < a.n_sup; s0 += 32) {
                        const int sidx = s0 + lane;
                        bool sn = sidx < a.n_sup;
                        if (VARIANT == MMO_VARIANT_SHIFTED && sn) sn = box_near(__ldg(a.sup_box + 2 * sidx), __ldg(a.sup_box + 2 * sidx + 1));
                        unsigned sm = __ballot_sync(0xffffffffu, sn);
                        while (sm) {
                            const int g = (s0 + __ffs(sm) - 1) * 32 + lane;
                            sm &= sm - 1u;
                            bool near = g >= g_lo && g < g_hi;
                            if (VARIANT == MMO_VARIANT_SHIFTED && near) near = box_near(__ldg(a.blob_box + 2 * g), __ldg(a.blob_box + 2 * g + 1));
                            const unsigned gm = __ballot_sync(0xffffffffu, near);
                            if (near) s_near[ng + __popc(gm & lt_mask)] = (uint16_t)g;
                            ng += __popc(gm);
                        }
                    }
                    if (lane < 4) s_near[ng + lane] = (uint16_t)nb;                  // the far-away group behind the last one
                }
                __syncwarp();
                const float qjs = lp.z, Ajs = lp.x, nBjs = -lp.y;
                const float Hj = a.hscale * lp.w;       // this ligand atom's clamp / close-contact threshold on r^2 (warp-uniform)
                // ---- level 2: the atoms of four near groups per step, two per lane (independent loads and
                //      tests); survivors are compacted into the warp's list, which is consumed whenever it
                //      cannot take another step ----
                int n = 0;
                for (int i = 0; i < ng || n > 0; i += 4) {
                    const bool more = i < ng;
                    if (more) {
                        const int atomA = s_near[i + (lane >> 4)] * kBlob + (lane & 15);
                        const int atomB = s_near[i + 2 + (lane >> 4)] * kBlob + (lane & 15);
                        const float4 pa = TILED ? s_atom[atomA] : __ldg(a.xyzq + atomA), pb = TILED ? s_atom[atomB] : __ldg(a.xyzq + atomB);
                        const int ea = TILED ? (int)s_elt[atomA] : (int)__ldg(a.gelt + atomA), eb = TILED ? (int)s_elt[atomB] : (int)__ldg(a.gelt + atomB);
                        const float Xa = pa.x - cx, Ya = pa.y - cy, Za = pa.z - cz;
                        const float Xb = pb.x - cx, Yb = pb.y - cy, Zb = pb.z - cz;
                        const float Sa = fmaf(Za, Za, fmaf(Ya, Ya, Xa * Xa)), Sb = fmaf(Zb, Zb, fmaf(Yb, Yb, Xb * Xb));
                        const bool na = VARIANT == MMO_VARIANT_SHIFTED ? Sa < reach2 : pa.x < 0.5f * kFarAway;
                        const bool nb_ = VARIANT == MMO_VARIANT_SHIFTED ? Sb < reach2 : pb.x < 0.5f * kFarAway;
                        const unsigned bma = __ballot_sync(0xffffffffu, na), bmb = __ballot_sync(0xffffffffu, nb_);
                        if (na) {
                            const float2 tab = s_tab[ea];
                            float *e = s_l + n + __popc(bma & lt_mask);
                            e[0 * LIST_CAP] = expand ? Xa : pa.x; e[1 * LIST_CAP] = expand ? Ya : pa.y; e[2 * LIST_CAP] = expand ? Za : pa.z;
                            e[3 * LIST_CAP] = Sa;
                            e[4 * LIST_CAP] = pa.w * qjs; e[5 * LIST_CAP] = tab.x * Ajs; e[6 * LIST_CAP] = tab.y * nBjs;
                        }
                        n += __popc(bma);
                        if (nb_) {
                            const float2 tab = s_tab[eb];
                            float *e = s_l + n + __popc(bmb & lt_mask);
                            e[0 * LIST_CAP] = expand ? Xb : pb.x; e[1 * LIST_CAP] = expand ? Yb : pb.y; e[2 * LIST_CAP] = expand ? Zb : pb.z;
                            e[3 * LIST_CAP] = Sb;
                            e[4 * LIST_CAP] = pb.w * qjs; e[5 * LIST_CAP] = tab.x * Ajs; e[6 * LIST_CAP] = tab.y * nBjs;
                        }
                        n += __popc(bmb);
                        if (n <= LIST_CAP - 64 && i + 4 < ng) continue;          // room for another step
                    }
                    if (n > 0) {

                        // ---- consume the list: 2 poses x 2 packed chains = 8 pairs per step ----
                        if (STATS) n_eval += (unsigned long long)n * (unsigned)(valid[0] + valid[1]);
                        const int n4 = (n + 3) & ~3;
                        if (lane < n4 - n) {                              // pad with far-away, inert atoms
                            float *e = s_l + n + lane;
                            e[0 * LIST_CAP] = kFarAway; e[1 * LIST_CAP] = kFarAway; e[2 * LIST_CAP] = kFarAway;
                            e[3 * LIST_CAP] = 3.0f * kFarAway * kFarAway;
                            e[4 * LIST_CAP] = 0.f; e[5 * LIST_CAP] = 0.f; e[6 * LIST_CAP] = 0.f;
                        }
                        __syncwarp();
                        if (expand) run_list<VARIANT, true, STATS>(s_l, n, n4, m2x, m2y, m2z, l2, Hj, acc, rmin, n_in);
                        else run_list<VARIANT, false, STATS>(s_l, n, n4, m2x, m2y, m2z, l2, Hj, acc, rmin, n_in);
                        n = 0;
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int h = 0; h < PPT; h++) cbits[h] |= (rmin[h] < Hj * 1.001f + 0.01f ? 1u : 0u) << jj;
            }
            // close-contact flags of this (slice, chunk): one byte per pose, read by the fix kernel
#pragma unroll
            for (int h = 0; h < PPT; h++) {
                if (valid[h])
                    flags[((int64_t)sl * n_chunks + c) * n_poses + p0 + 32 * h] = (uint8_t)cbits[h];
            }
            __syncwarp();    // the warp's s_c columns are rewritten by the next chunk
        }
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            if (valid[h]) out[(int64_t)yz * n_poses + p0 + 32 * h] = acc[h];
            if (STATS && valid[h]) n_in_tot += n_in[h];
        }
    }
    if (STATS) {
        atomicAdd(a.stats + 0, n_eval);
        atomicAdd(a.stats + 1, n_in_tot);
    }
}

// ---- item mode: incoherent pose lists (conformer screens, random poses) ---------------------------------
// The poses of such a list share nothing a warp could cull with.  The unit of work is therefore the ITEM
// (pose, ligand atom): item_prepare_kernel computes every item's position (reference arithmetic in double,
// then fp32) and its cell in a lattice of kItemCell-A cells over the receptor's surroundings; a stable radix
// sort by cell (cub::DeviceRadixSort, deterministic) makes 64 consecutive items spatial neighbours (rho <=
// ~2.2 A), and direct_items_kernel runs the same cull -> list -> packed pair loop on 64 items per warp.  Since
// the ligand atom now differs from lane to lane, the list carries the receptor factors only and the lanes keep
// three sums each,  E = A_j sum(w A_i s^6) - B_j sum(w B_i s^3) + q_j sum(w q_i / r):  15 packed ops per two pairs.
constexpr float kItemCell = 0.5f;
constexpr int kItemTPB = 256;
#ifndef MMO_ITEM_SUM_EVERY
#define MMO_ITEM_SUM_EVERY 8
#endif
#ifndef MMO_ITEM_MINB
#define MMO_ITEM_MINB 2
#endif
#ifndef MMO_ITEM_PREFETCH
#define MMO_ITEM_PREFETCH 0      /* requesting the next four groups' atoms one step ahead: measured, no gain (L1 hits) */
#endif
constexpr int kItemSumEvery = MMO_ITEM_SUM_EVERY;

struct ItemArgs {
    float tab_A[kEltTab], tab_B[kEltTab];
    const float4 *xyzq;          // receptor groups of 16 atoms (+ one dummy group of far-away atoms at index n_blobs)
    const uint8_t *gelt;
    const float4 *blob_box;      // {lo, hi} per group
    const float4 *sup_box;       // {lo, hi} per super-group of 32 consecutive groups (k-d order: spatially compact)
    int n_blobs, n_sup;
    const float4 *lparam;        // fast-path order {A_j, B_j, q_j, real?}
    int n_fast;
    const float4 *pos;           // item -> two float4: position relative to the receptor origin {x, y, z, 0} and the fp32
                                 // remainders of the double {lo_x, lo_y, lo_z, 0} (read by item_fix_kernel): one 32 B sector
    const uint32_t *perm;        // sorted rank -> item
    const unsigned long long *n_far;   // items beyond the lattice (energy exactly 0): sorted last
    unsigned long long n_items;
    float hscale;                // H of an item = hscale * lparam.w of its ligand atom (per lane)
    double *e_rank;              // = E of the item at sorted rank r (coalesced; item_fix_kernel scatters it to item order)
    uint8_t *f_rank;             // = its close-contact flag
    unsigned long long *stats;
};

template <int VARIANT, bool EXPAND, bool STATS>
__device__ __forceinline__ void run_list_items(const float *s_l, int n, int n4, const float (&m2x)[PPT], const float (&m2y)[PPT],
                                               const float (&m2z)[PPT], const float (&l2)[PPT], const float (&H)[PPT],
                                               double (&EA)[PPT], double (&EB)[PPT], double (&EQ)[PPT],
                                               float (&rmin)[PPT], unsigned long long (&n_in)[PPT]) {
    // blocks of kItemSumEvery steps in fp32, then F2F + DADD (nested loops: no per-step counter)
#pragma unroll 1
    for (int k0 = 0; k0 < n4; k0 += 4 * kItemSumEvery) {
        const int kend = min(n4, k0 + 4 * kItemSumEvery);
        float2 fA[PPT], fB[PPT], fQ[PPT];
#pragma unroll
        for (int h = 0; h < PPT; h++) fA[h] = fB[h] = fQ[h] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int k = k0; k < kend; k += 4) {
            const float4 X = *(const float4 *)(s_l + 0 * LIST_CAP + k), Y = *(const float4 *)(s_l + 1 * LIST_CAP + k);
            const float4 Z = *(const float4 *)(s_l + 2 * LIST_CAP + k), S = *(const float4 *)(s_l + 3 * LIST_CAP + k);
            const float4 Q = *(const float4 *)(s_l + 4 * LIST_CAP + k), A = *(const float4 *)(s_l + 5 * LIST_CAP + k);
            const float4 B = *(const float4 *)(s_l + 6 * LIST_CAP + k);
#pragma unroll
            for (int h = 0; h < PPT; h++) {
#pragma unroll
                for (int pk = 0; pk < 2; pk++) {
                    const float2 x2 = pk ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
                    const float2 y2 = pk ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
                    const float2 z2 = pk ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
                    const float2 s2 = pk ? make_float2(S.z, S.w) : make_float2(S.x, S.y);
                    const float2 q2 = pk ? make_float2(Q.z, Q.w) : make_float2(Q.x, Q.y);
                    const float2 a2 = pk ? make_float2(A.z, A.w) : make_float2(A.x, A.y);
                    const float2 b2 = pk ? make_float2(B.z, B.w) : make_float2(B.x, B.y);
                    float2 r2;
                    if (EXPAND) {
                        r2 = __ffma2_rn(x2, bc2(m2x[h]), __ffma2_rn(y2, bc2(m2y[h]), __ffma2_rn(z2, bc2(m2z[h]), __fadd2_rn(s2, bc2(l2[h])))));
                    } else {
                        const float2 dx = __fadd2_rn(x2, bc2(m2x[h])), dy = __fadd2_rn(y2, bc2(m2y[h])), dz = __fadd2_rn(z2, bc2(m2z[h]));
                        r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    }
                    rmin[h] = fminf(rmin[h], fminf(r2.x, r2.y));
                    const float2 r2c = make_float2(fmaxf(r2.x, H[h]), fmaxf(r2.y, H[h]));      // no upper clamp: the weight saturates
                    const float2 rinv = make_float2(rsqrt_fast(r2c.x), rsqrt_fast(r2c.y));
                    const float2 s = __fmul2_rn(rinv, rinv);
                    const float2 s3 = __fmul2_rn(__fmul2_rn(s, s), s);
                    float2 ws3 = s3, wr = rinv;
                    if (VARIANT == MMO_VARIANT_SHIFTED) {
                        // FF.shift_12A: sat(1 - r^2/144)^2, exactly 0 from 12 A on (see pair2)
                        const float2 up = make_float2(fma_sat(r2c.x, -1.0f / 144.0f, 1.0f), fma_sat(r2c.y, -1.0f / 144.0f, 1.0f));
                        const float2 w = __fmul2_rn(up, up);
                        ws3 = __fmul2_rn(w, s3);
                        wr = __fmul2_rn(w, rinv);
                    }
                    fB[h] = __ffma2_rn(ws3, b2, fB[h]);
                    fA[h] = __ffma2_rn(__fmul2_rn(ws3, s3), a2, fA[h]);
                    fQ[h] = __ffma2_rn(wr, q2, fQ[h]);
                    if (STATS) n_in[h] += (r2.x < 144.0f && k + 2 * pk < n) + (r2.y < 144.0f && k + 2 * pk + 1 < n);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            EA[h] += (double)(fA[h].x + fA[h].y);
            EB[h] += (double)(fB[h].x + fB[h].y);
            EQ[h] += (double)(fQ[h].x + fQ[h].y);
        }
    }
}

// 10 bits -> every third bit (Morton / Z-order interleave)
__device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// position (double, then fp32 relative to the receptor origin) and lattice cell of every item; consecutive threads take
// consecutive items, so that the 40 bytes written per item (and the coordinates read, for explicit conformers) are
// coalesced, and every thread takes kPrepIPT items a block-width apart with all their coordinate loads issued before the
// first use: the kernel waits on DRAM (three dependent-address loads per item), not on arithmetic
constexpr int kPrepIPT = 4;
__global__ void __launch_bounds__(256)
item_prepare_kernel(PoseSrc src, int64_t n_poses, int L, int n_fast, const double *__restrict__ lx, const double *__restrict__ ly,
                    const double *__restrict__ lz, const int32_t *__restrict__ forder, const float4 *__restrict__ lparam,
                    double ox, double oy, double oz, float cell_lo_x, float cell_lo_y, float cell_lo_z, float cell_inv,
                    int nx, int ny, int nz, uint32_t far_key, float4 *__restrict__ pos, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                    unsigned long long *__restrict__ n_far) {
    const int64_t n_items = n_poses * n_fast;
    const int64_t base = (int64_t)blockIdx.x * (256 * kPrepIPT) + threadIdx.x;
    double X[kPrepIPT], Y[kPrepIPT], Z[kPrepIPT];
    bool real[kPrepIPT];
#pragma unroll
    for (int r = 0; r < kPrepIPT; r++) {
        const int64_t it = base + r * 256;
        real[r] = false;
        X[r] = Y[r] = Z[r] = 0.0;
        if (it < n_items) {
            const uint32_t p = (uint32_t)it / (uint32_t)n_fast;         // a batch holds fewer than 2^32 items
            const int k = (int)((uint32_t)it - p * (uint32_t)n_fast);
            real[r] = __ldg(&lparam[k].w) != 0.f;
            if (real[r]) {
                if (src.kind == 1) {
                    const int64_t a = (int64_t)p * L + __ldg(forder + k);
                    X[r] = __ldg(src.xs + a); Y[r] = __ldg(src.ys + a); Z[r] = __ldg(src.zs + a);
                } else {
                    PoseRT P;
                    load_pose_rt(src, (int64_t)p, P);
                    pose_atom_rt(P, __ldg(lx + k), __ldg(ly + k), __ldg(lz + k), X[r], Y[r], Z[r]);
                }
            }
        }
    }
    unsigned n_far_lane = 0;
#pragma unroll
    for (int r = 0; r < kPrepIPT; r++) {
        const int64_t it = base + r * 256;
        if (it >= n_items) continue;
        uint32_t key = far_key;
        float4 v = make_float4(kFarAway, kFarAway, kFarAway, 0.f), vlo = make_float4(0.f, 0.f, 0.f, 0.f);
        if (real[r]) {
            const double Dx = X[r] - ox, Dy = Y[r] - oy, Dz = Z[r] - oz;
            v.x = (float)Dx; v.y = (float)Dy; v.z = (float)Dz;
            // hi + lo carries 48 bits of the double (< 2e-13 A at 50 A): what the fp64 close-contact pass works with
            vlo.x = (float)(Dx - (double)v.x); vlo.y = (float)(Dy - (double)v.y); vlo.z = (float)(Dz - (double)v.z);
            const float fx = (v.x - cell_lo_x) * cell_inv, fy = (v.y - cell_lo_y) * cell_inv, fz = (v.z - cell_lo_z) * cell_inv;
            if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)nx && fy < (float)ny && fz < (float)nz)
                key = spread3((uint32_t)fx) | (spread3((uint32_t)fy) << 1) | (spread3((uint32_t)fz) << 2);     // Z-order
        }
        pos[2 * it] = v;
        pos[2 * it + 1] = vlo;
        keys[it] = key;
        vals[it] = (uint32_t)it;
        n_far_lane += key == far_key;
    }
    // one atomic per warp
    for (int o = 16; o > 0; o >>= 1) n_far_lane += __shfl_xor_sync(0xffffffffu, n_far_lane, o);
    if ((threadIdx.x & 31) == 0 && n_far_lane) atomicAdd(n_far, (unsigned long long)n_far_lane);
}

// One launch for the whole receptor.  The receptor is NOT staged in shared memory: 64 cell-sorted items only touch the
// groups within 12 A + rho of their centre (a few dozen of, say, 625), consecutive units touch the same ones, and the
// read-only path keeps them in L1 -- while a shared-memory tile of 2048 atoms would mean five launches for a 10 000-atom
// receptor, each one gathering the items, culling and scattering the energies again.  Level 1 is two-staged: lane s
// tests the box of super-group s (32 consecutive k-d leaves), then lane g the box of group g of every near super-group.
// The next unit's items (two dependent gathers: rank -> item -> position) are requested before the current unit's pair
// loop and arrive behind it.
// Shared memory (dynamic): super-group boxes [2 * n_sup] float4, ligand parameters, vdW table, per-warp lists
// [NF][LIST_CAP], per-warp near-group ids (uint16).
struct ItemUnit { bool valid[PPT]; uint32_t item[PPT]; float4 v[PPT]; };

__device__ __forceinline__ void load_item_unit(const ItemArgs &a, unsigned long long u, unsigned long long n_near, int lane, ItemUnit &U) {
#pragma unroll
    for (int h = 0; h < PPT; h++) {
        const unsigned long long idx = u * (32 * PPT) + 32 * h + lane;
        U.valid[h] = idx < n_near;
        U.item[h] = __ldg(a.perm + (U.valid[h] ? idx : n_near - 1));    // idle slots shadow the last item
        U.v[h] = __ldg(a.pos + 2 * (size_t)U.item[h]);
    }
}

template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(kItemTPB, MMO_ITEM_MINB)
direct_items_kernel(ItemArgs a, int near_cap, unsigned long long *__restrict__ work) {
    extern __shared__ float4 smem4[];
    float4 *s_sup = smem4;                                    // 2 * n_sup
    float4 *s_lparam = s_sup + 2 * a.n_sup;
    float2 *s_tab = (float2 *)(s_lparam + a.n_fast);
    float *s_l = (float *)(s_tab + 16) + (threadIdx.x >> 5) * (NF * LIST_CAP);
    uint16_t *s_near = (uint16_t *)((float *)(s_tab + 16) + (kItemTPB / 32) * (NF * LIST_CAP)) + (threadIdx.x >> 5) * near_cap;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int j = tid; j < a.n_fast; j += kItemTPB) s_lparam[j] = a.lparam[j];
    if (tid < kEltTab) s_tab[tid] = make_float2(a.tab_A[tid], a.tab_B[tid]);
    for (int k = tid; k < 2 * a.n_sup; k += kItemTPB) s_sup[k] = __ldg(a.sup_box + k);
    __syncthreads();
    const unsigned long long n_near = a.n_items - *a.n_far;
    const unsigned long long n_units = (n_near + 32 * PPT - 1) / (32 * PPT);
    unsigned long long n_eval = 0, n_in_tot = 0;

    auto grab = [&]() {
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(work, 1ull);
        return __shfl_sync(0xffffffffu, u, 0);
    };
    unsigned long long u = grab();
    ItemUnit cur, nxt;
    if (u < n_units) load_item_unit(a, u, n_near, lane, cur);
    while (u < n_units) {
        const unsigned long long un = grab();
        if (un < n_units) load_item_unit(a, un, n_near, lane, nxt);          // in flight during this unit's pair loop
        float px[PPT], py[PPT], pz[PPT];
        float4 lp[PPT];
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            px[h] = cur.v[h].x; py[h] = cur.v[h].y; pz[h] = cur.v[h].z;
            lp[h] = s_lparam[cur.item[h] % (uint32_t)a.n_fast];
        }
        // ---- centre and radius of the warp's 64 items ----
        const float cx = 0.5f * (warp_min(fminf(px[0], px[1])) + warp_max(fmaxf(px[0], px[1])));
        const float cy = 0.5f * (warp_min(fminf(py[0], py[1])) + warp_max(fmaxf(py[0], py[1])));
        const float cz = 0.5f * (warp_min(fminf(pz[0], pz[1])) + warp_max(fmaxf(pz[0], pz[1])));
        float l2[PPT], rmin[PPT], m2x[PPT], m2y[PPT], m2z[PPT];
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            m2x[h] = px[h]; m2y[h] = py[h]; m2z[h] = pz[h];
            px[h] -= cx; py[h] -= cy; pz[h] -= cz;
            l2[h] = fmaf(pz[h], pz[h], fmaf(py[h], py[h], px[h] * px[h]));
            rmin[h] = 3e38f;
        }
        const float rho2 = warp_max(fmaxf(l2[0], l2[1]));
        const bool expand = rho2 <= kRhoExpand2;
        const float reach = 12.0f + sqrtf(rho2) * 1.0001f + 1e-4f;
        const float reach2 = reach * reach;
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            // expanded form: list and item relative to c; difference form (incoherent warps, c may be far from
            // everything): both stay relative to the receptor origin, no second rounding
            m2x[h] = expand ? -2.0f * px[h] : -m2x[h];
            m2y[h] = expand ? -2.0f * py[h] : -m2y[h];
            m2z[h] = expand ? -2.0f * pz[h] : -m2z[h];
        }
        double EA[PPT], EB[PPT], EQ[PPT];
        unsigned long long n_in[PPT];
        float Hl[PPT];
#pragma unroll
        for (int h = 0; h < PPT; h++) { EA[h] = EB[h] = EQ[h] = 0.0; n_in[h] = 0; Hl[h] = a.hscale * lp[h].w; }
        // ---- level 1: super-group boxes, then the group boxes of the near super-groups, against the sphere (c, 12 + rho) ----
        auto box_near = [&](const float4 blo, const float4 bhi) {
            const float gx = fmaxf(0.f, fmaxf(blo.x - cx, cx - bhi.x));
            const float gy = fmaxf(0.f, fmaxf(blo.y - cy, cy - bhi.y));
            const float gz = fmaxf(0.f, fmaxf(blo.z - cz, cz - bhi.z));
            return fmaf(gz, gz, fmaf(gy, gy, gx * gx)) < reach2;
        };
        int ng = 0;
        __syncwarp();
        for (int s0 = 0; s0 < a.n_sup; s0 += 32) {
            const int sidx = s0 + lane;
            bool sn = sidx < a.n_sup;
            if (VARIANT == MMO_VARIANT_SHIFTED && sn) sn = box_near(s_sup[2 * sidx], s_sup[2 * sidx + 1]);
            unsigned sm = __ballot_sync(0xffffffffu, sn);
            while (sm) {
                const int g = (s0 + __ffs(sm) - 1) * 32 + lane;
                sm &= sm - 1u;
                bool near = g < a.n_blobs;
                if (VARIANT == MMO_VARIANT_SHIFTED && near) near = box_near(__ldg(a.blob_box + 2 * g), __ldg(a.blob_box + 2 * g + 1));
                const unsigned gm = __ballot_sync(0xffffffffu, near);
                if (near) s_near[ng + __popc(gm & lt_mask)] = (uint16_t)g;
                ng += __popc(gm);
            }
        }
        if (lane < 4) s_near[ng + lane] = (uint16_t)a.n_blobs;        // pad with the dummy group
        __syncwarp();
        // ---- level 2 + list consumption (as in direct_fp32_kernel; the list carries receptor factors only).  The atoms
        //      of the next four groups are requested (L1 / L2) before the current four are tested ----
        int n = 0;
#if MMO_ITEM_PREFETCH
        float4 pa_n = make_float4(0.f, 0.f, 0.f, 0.f), pb_n = pa_n;
        int ea_n = 0, eb_n = 0;
        if (ng > 0) {
            const int atomA = s_near[(lane >> 4)] * kBlob + (lane & 15), atomB = s_near[2 + (lane >> 4)] * kBlob + (lane & 15);
            pa_n = __ldg(a.xyzq + atomA); pb_n = __ldg(a.xyzq + atomB);
            ea_n = __ldg(a.gelt + atomA); eb_n = __ldg(a.gelt + atomB);
        }
#endif
        for (int i = 0; i < ng || n > 0; i += 4) {
            if (i < ng) {
#if MMO_ITEM_PREFETCH
                const float4 pa = pa_n, pb = pb_n;
                const int ea = ea_n, eb = eb_n;
                if (i + 4 < ng) {
                    const int atomA = s_near[i + 4 + (lane >> 4)] * kBlob + (lane & 15);
                    const int atomB = s_near[i + 6 + (lane >> 4)] * kBlob + (lane & 15);
                    pa_n = __ldg(a.xyzq + atomA); pb_n = __ldg(a.xyzq + atomB);
                    ea_n = __ldg(a.gelt + atomA); eb_n = __ldg(a.gelt + atomB);
                }
#else
                const int atomA = s_near[i + (lane >> 4)] * kBlob + (lane & 15);
                const int atomB = s_near[i + 2 + (lane >> 4)] * kBlob + (lane & 15);
                const float4 pa = __ldg(a.xyzq + atomA), pb = __ldg(a.xyzq + atomB);
                const int ea = __ldg(a.gelt + atomA), eb = __ldg(a.gelt + atomB);
#endif
                const float Xa = pa.x - cx, Ya = pa.y - cy, Za = pa.z - cz;
                const float Xb = pb.x - cx, Yb = pb.y - cy, Zb = pb.z - cz;
                const float Sa = fmaf(Za, Za, fmaf(Ya, Ya, Xa * Xa)), Sb = fmaf(Zb, Zb, fmaf(Yb, Yb, Xb * Xb));
                const bool na = VARIANT == MMO_VARIANT_SHIFTED ? Sa < reach2 : pa.x < 0.5f * kFarAway;
                const bool nb_ = VARIANT == MMO_VARIANT_SHIFTED ? Sb < reach2 : pb.x < 0.5f * kFarAway;
                const unsigned bma = __ballot_sync(0xffffffffu, na), bmb = __ballot_sync(0xffffffffu, nb_);
                if (na) {
                    const float2 tab = s_tab[ea];
                    float *e = s_l + n + __popc(bma & lt_mask);
                    e[0 * LIST_CAP] = expand ? Xa : pa.x; e[1 * LIST_CAP] = expand ? Ya : pa.y; e[2 * LIST_CAP] = expand ? Za : pa.z;
                    e[3 * LIST_CAP] = Sa;
                    e[4 * LIST_CAP] = pa.w; e[5 * LIST_CAP] = tab.x; e[6 * LIST_CAP] = tab.y;
                }
                n += __popc(bma);
                if (nb_) {
                    const float2 tab = s_tab[eb];
                    float *e = s_l + n + __popc(bmb & lt_mask);
                    e[0 * LIST_CAP] = expand ? Xb : pb.x; e[1 * LIST_CAP] = expand ? Yb : pb.y; e[2 * LIST_CAP] = expand ? Zb : pb.z;
                    e[3 * LIST_CAP] = Sb;
                    e[4 * LIST_CAP] = pb.w; e[5 * LIST_CAP] = tab.x; e[6 * LIST_CAP] = tab.y;
                }
                n += __popc(bmb);
                if (n <= LIST_CAP - 64 && i + 4 < ng) continue;
            }
            if (n > 0) {
                if (STATS) n_eval += (unsigned long long)n * (unsigned)(cur.valid[0] + cur.valid[1]);
                const int n4 = (n + 3) & ~3;
                if (lane < n4 - n) {
                    float *e = s_l + n + lane;
                    e[0 * LIST_CAP] = kFarAway; e[1 * LIST_CAP] = kFarAway; e[2 * LIST_CAP] = kFarAway;
                    e[3 * LIST_CAP] = 3.0f * kFarAway * kFarAway;
                    e[4 * LIST_CAP] = 0.f; e[5 * LIST_CAP] = 0.f; e[6 * LIST_CAP] = 0.f;
                }
                __syncwarp();
                if (expand) run_list_items<VARIANT, true, STATS>(s_l, n, n4, m2x, m2y, m2z, l2, Hl, EA, EB, EQ, rmin, n_in);
                else run_list_items<VARIANT, false, STATS>(s_l, n, n4, m2x, m2y, m2z, l2, Hl, EA, EB, EQ, rmin, n_in);
                n = 0;
                __syncwarp();
            }
        }
#pragma unroll
        for (int h = 0; h < PPT; h++) {
            if (cur.valid[h]) {
                const unsigned long long r = u * (32 * PPT) + 32 * h + lane;       // sorted rank: coalesced stores
                a.e_rank[r] = (double)lp[h].x * EA[h] - (double)lp[h].y * EB[h] + (double)lp[h].z * EQ[h];
                a.f_rank[r] = rmin[h] < Hl[h] * 1.001f + 0.01f ? 1 : 0;
                if (STATS) n_in_tot += n_in[h];
            }
        }
        cur = nxt;
        u = un;
    }
    if (STATS) {
        atomicAdd(a.stats + 0, n_eval);
        atomicAdd(a.stats + 1, n_in_tot);
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
    const int32_t *forder;               // fast-path position -> original atom index
    const int32_t *lelt;
    double H[kEltTab];                   // by ligand element: exactly the fp32 clamp value hscale * x_j
    double wrH[kEltTab];                 // w(H) / sqrt(H): weight (1 for GLOBAL) x Coulomb factor of what the fast path evaluated
    const double4 *tab;                  // [receptor element * kEltTab + ligand element] {A = d_ij x_ij^12, B = 2 d_ij x_ij^6,
                                         //  w(H) d_ij (p6H^2 - 2 p6H) at the H of the ligand element, 0}: one 32 B sector per pair
    unsigned long long *stats;           // [2] pairs re-evaluated
};

// The close-contact correction of ONE ligand atom at (x, y, z) (original atom index j): sum over the receptor atoms
// with r^2 < H of  w(r) e(r) - w(sqrt H) e(sqrt H), found through the atom's voxel list.
template <int VARIANT, bool STATS>
__device__ __forceinline__ double close_contact_corr(const FixArgs &a, double x, double y, double z, int j,
                                                     unsigned long long &n_fix) {
    double corr = 0.0;
    const double fx = (x - a.vox_lo[0]) * a.vox_inv, fy = (y - a.vox_lo[1]) * a.vox_inv, fz = (z - a.vox_lo[2]) * a.vox_inv;
    if (!(fx >= 0.0 && fy >= 0.0 && fz >= 0.0)) return corr;
    const int vi = (int)fx, vj = (int)fy, vk = (int)fz;
    if (vi >= a.vox_dim[0] || vj >= a.vox_dim[1] || vk >= a.vox_dim[2]) return corr;
    const size_t v = (size_t)vi + (size_t)vj * a.vox_dim[0] + (size_t)vk * a.vox_dim[0] * a.vox_dim[1];
    const int k0 = __ldg(a.vox_off + v), k1 = __ldg(a.vox_off + v + 1);
    if (k0 == k1) return corr;
    const double qj = kElecWeight * __ldg(a.lq + j);
    const int ej = __ldg(a.lelt + j);
    const float xf = (float)(x - a.vox_lo[0]), yf = (float)(y - a.vox_lo[1]), zf = (float)(z - a.vox_lo[2]);
    const double H = a.H[ej], wrH = a.wrH[ej];
    const float Hf = (float)H + MMO_FIX_MARGIN;       // fp32 r^2 of coordinates below ~200 A: error < 1e-3 A^2
    const int32_t *__restrict__ idx = a.vox_idx + k0;
    const int nk = k1 - k0;
    // Two phases per window of 32 candidates, so that a warp whose lanes sit in different voxels pays
    // max(candidates) cheap tests + max(close pairs) fp64 evaluations, not their product: (1) the fp32 pre-test
    // (coordinates relative to the voxel grid corner, error << the margin) marks the survivors in a
    // 32-bit mask, (2) the survivors are evaluated in double, in list order.
    for (int kw = 0; kw < nk; kw += 32) {
        unsigned pass = 0u;
        const int kn = min(32, nk - kw);
#pragma unroll kFixUnroll
        for (int b = 0; b < kn; b++) {
            const float4 r4 = __ldg(a.pxyz32 + (__ldg(idx + kw + b) & 0xffffff));
            const float fdx = r4.x - xf, fdy = r4.y - yf, fdz = r4.z - zf;
            if (fdx * fdx + fdy * fdy + fdz * fdz < Hf) pass |= 1u << b;
        }
        while (pass != 0u) {
            const int b = __ffs((int)pass) - 1;
            pass &= pass - 1u;
            const int ie = __ldg(idx + kw + b);
            const int i = ie & 0xffffff;
            const double2 r01 = __ldg((const double2 *)(a.pxyzq + i));
            const double2 r23 = __ldg((const double2 *)(a.pxyzq + i) + 1);
            const double dx = r01.x - x, dy = r01.y - y, dz = r23.x - z;
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < H) {
                const double2 *t = (const double2 *)(a.tab + ((int)((unsigned)ie >> 24) * kEltTab + ej));
                const double2 AB = __ldg(t);
                const double wvH = __ldg((const double *)(t + 1));
                const double qq = r23.y * qj;
                const double r2c = r2 < 1e-4 ? 1e-4 : r2;           // Math.non_zero_dist on r (r2 is never NaN here)
                // 1/r: MUFU.RSQ in fp32 (relative error < 2e-7), one Newton step in double (-> < 1e-13)
                const double y0 = (double)rsqrt_fast((float)r2c);
                const double rinv = y0 * (1.5 - (0.5 * r2c) * (y0 * y0));
                const double s1 = rinv * rinv, s3 = (s1 * s1) * s1;
                // (the table stays in global memory / L1: staging it in shared memory per block was measured,
                //  0 % for hard_fix_kernel, +22 % time for item_fix_kernel, whose blocks mostly exit at once)
                const double ee = (AB.x * s3 - AB.y) * s3 + qq * rinv;     // (A s^3 - B) s^3 + qq / r
                const double eH = qq * wrH + wvH;     // w(H) x what the fast path evaluated (r clamped at sqrt(H), in the weight too)
                double d;
                if (VARIANT == MMO_VARIANT_SHIFTED) {
                    const double u = 1.0 - r2c * (1.0 / 144.0);
                    d = (u * u) * ee - eH;
                } else {
                    d = ee - eH;
                }
                corr += d;
                if (STATS) n_fix++;
            }
        }
    }
    return corr;
}

// ITEMS: part = per-item energies [pose][n_split = n_fast], corrections already added by item_fix_kernel: called with
//        n_chunks = 0, the kernel only sums the items of a pose in atom order;
// else  : part = per-split sums [n_split][pose], flags = per-(tile, chunk) bytes (direct_fp32_kernel)
template <int VARIANT, bool STATS, bool ITEMS>
__global__ void __launch_bounds__(128, kFixBlocksPerSM)
hard_fix_kernel(FixArgs a, PoseSrc src, int64_t n_poses, const double *part, int n_split,
                const uint8_t *__restrict__ flags, int n_tiles, int n_chunks, double *out) {   // part may alias out
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_poses) return;
    // the fast kernel's partial sums, added in a fixed order
    double e = 0.0;
    if (ITEMS) for (int k = 0; k < n_split; k++) e += part[p * n_split + k];
    else for (int k = 0; k < n_split; k++) e += part[(int64_t)k * n_poses + p];
    double corr = 0.0;
    unsigned long long n_fix = 0, n_flag = 0;
    bool have_pose = false;
    // the pose (rotation + translation, 12 doubles) lives in this thread's shared-memory column, not in 24 registers:
    // the kernel is latency bound and the registers buy more resident warps
    __shared__ double s_P[12][128];
    const int tx = threadIdx.x;
    for (int c = 0; c < n_chunks; c++) {
        // atoms of chunk c (fast-path order) for which the fast kernel saw a pair below H (any tile)
        unsigned bits = 0u;
        for (int t = 0; t < n_tiles; t++) bits |= flags[((int64_t)t * n_chunks + c) * n_poses + p];
        while (bits != 0u) {
            const int jj = __ffs(bits) - 1;
            bits &= bits - 1u;
            const int j = __ldg(a.forder + c * kFixLJ + jj);
            if (STATS) n_flag++;
            double x, y, z;
            if (src.kind == 1) {
                x = src.xs[p * a.L + j]; y = src.ys[p * a.L + j]; z = src.zs[p * a.L + j];
            } else {
                if (!have_pose) {
                    PoseRT P;
                    load_pose_rt(src, p, P);
#pragma unroll
                    for (int q = 0; q < 9; q++) s_P[q][tx] = P.r[q];
#pragma unroll
                    for (int q = 0; q < 3; q++) s_P[9 + q][tx] = P.t[q];
                    have_pose = true;
                }
                const double ax = __ldg(a.lx + j), ay = __ldg(a.ly + j), az = __ldg(a.lz + j);
                x = __dadd_rn(rot_row(s_P[0][tx], s_P[1][tx], s_P[2][tx], ax, ay, az), s_P[9][tx]);
                y = __dadd_rn(rot_row(s_P[3][tx], s_P[4][tx], s_P[5][tx], ax, ay, az), s_P[10][tx]);
                z = __dadd_rn(rot_row(s_P[6][tx], s_P[7][tx], s_P[8][tx], ax, ay, az), s_P[11][tx]);
            }
            corr += close_contact_corr<VARIANT, STATS>(a, x, y, z, j, n_fix);
        }
    }
    out[p] = e + corr;
    if (STATS) { atomicAdd(a.stats + 2, n_fix); atomicAdd(a.stats + 3, n_flag); }
}

// Small batches (single-pose calls, the reference's closure shape): block = pose, thread = ligand atom.  One thread per
// pose walking its 48 flagged atoms one after the other is ~70 us of dependent loads; here every flagged atom has its own
// thread, the corrections meet in shared memory and thread 0 adds the parts and then the corrections in atom order -- the
// same additions in the same order as hard_fix_kernel (an unflagged atom adds 0.0), hence the same bits.
template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(64)
pose_fix_block_kernel(FixArgs a, PoseSrc src, int64_t n_poses, int n_fast, const double *part, int n_parts,
                      const uint8_t *__restrict__ flags, int n_tiles, int n_chunks, double *out) {   // part may alias out
    extern __shared__ double s_corr[];       // n_fast
    const int64_t p = blockIdx.x;
    unsigned long long n_fix = 0, n_flag = 0;
    for (int k = threadIdx.x; k < n_fast; k += blockDim.x) {
        const int c = k / kFixLJ, jj = k - c * kFixLJ;
        unsigned bits = 0u;
        for (int t = 0; t < n_tiles; t++) bits |= flags[((int64_t)t * n_chunks + c) * n_poses + p];
        double corr = 0.0;
        if ((bits >> jj) & 1u) {
            const int j = __ldg(a.forder + k);
            double x, y, z;
            if (src.kind == 1) {
                x = src.xs[p * a.L + j]; y = src.ys[p * a.L + j]; z = src.zs[p * a.L + j];
            } else {
                PoseRT P;
                load_pose_rt(src, p, P);
                pose_atom_rt(P, __ldg(a.lx + j), __ldg(a.ly + j), __ldg(a.lz + j), x, y, z);
            }
            corr = close_contact_corr<VARIANT, STATS>(a, x, y, z, j, n_fix);
            if (STATS) n_flag++;
        }
        s_corr[k] = corr;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int k = 0; k < n_parts; k++) e += part[(int64_t)k * n_poses + p];
        double corr = 0.0;
        for (int k = 0; k < n_fast; k++) corr += s_corr[k];
        out[p] = e + corr;
    }
    if (STATS && (n_fix | n_flag)) { atomicAdd(a.stats + 2, n_fix); atomicAdd(a.stats + 3, n_flag); }
}

// Item mode: the same correction with thread = ITEM in cell-sorted order (the order direct_items_kernel works in).  The
// lanes of a warp then sit in neighbouring voxels: their candidate lists are the same few cache lines and about equally
// long, which a warp of unrelated conformers (thread = pose) has neither of.  Everything read per item is either
// coalesced (rank -> item, energy and flag as direct_items_kernel stored them, by rank) or ONE 32-byte sector (the
// item's position record); the corrected energy is scattered to item order, one 8-byte store per item -- the only
// pass over e_item before hard_fix_kernel<ITEMS> (called with n_chunks = 0) sums the items of a pose in atom order.
template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(128, kFixBlocksPerSM)
item_fix_kernel(FixArgs a, double ox, double oy, double oz, int n_fast, const uint32_t *__restrict__ perm,
                const float4 *__restrict__ pos, const unsigned long long *__restrict__ n_far, unsigned long long n_items,
                const uint8_t *__restrict__ f_rank, const double *__restrict__ e_rank, double *__restrict__ e_item) {
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_items - *n_far) return;                      // items beyond the lattice are sorted last: energy exactly 0 (memset)
    const uint32_t item = __ldg(perm + r);
    double e = __ldg(e_rank + r);
    if (__ldg(f_rank + r)) {
        const float4 hi = __ldg(pos + 2 * (size_t)item), lo = __ldg(pos + 2 * (size_t)item + 1);
        const int j = __ldg(a.forder + (int)(item % (uint32_t)n_fast));
        const double x = ((double)hi.x + (double)lo.x) + ox, y = ((double)hi.y + (double)lo.y) + oy, z = ((double)hi.z + (double)lo.z) + oz;
        unsigned long long n_fix = 0;
        e += close_contact_corr<VARIANT, STATS>(a, x, y, z, j, n_fix);
        if (STATS) { atomicAdd(a.stats + 2, n_fix); atomicAdd(a.stats + 3, 1ull); }
    }
    e_item[item] = e;
}

// ---- host side -----------------------------------------------------------------------------------
// Library-lifetime device buffers.  Allocated with `new` and never destroyed: no destructor may touch the allocator
// or the CUDA context during static destruction at process exit; mmo_shutdown releases them (direct_drop_caches).
static DevBuf<double4> &g_fixtab = *new DevBuf<double4>();                   // FixArgs::tab
static DevBuf<unsigned long long> &g_stats = *new DevBuf<unsigned long long>(), &g_work = *new DevBuf<unsigned long long>();
static DevBuf<uint8_t> &g_item_scratch = *new DevBuf<uint8_t>();             // item mode scratch arena
constexpr int64_t kPoseFixBlockMax = 4096;       // poses up to which the fp64 pass runs block = pose (latency), thread = pose beyond
constexpr int64_t kItemModeMin = 32768;          // items (poses x ligand atoms) from which item mode pays
constexpr int64_t kItemBatch = (int64_t)64 << 20;  // items per batch: ~65 B of scratch each (4.4 GB)
constexpr int64_t kGlobalFp32MaxPairs = 120000;   // receptor x ligand atoms up to which GLOBAL stays on the fp32 path
static int g_direct_mode = 0;                     // 0 auto, 1 pose kernel always, 2 item kernel for every pose list
void direct_set_mode(int mode) { g_direct_mode = mode; }
static double g_fixtab_for = 0.0;                 // +-hscale the table was built for (sign: SHIFTED / GLOBAL)
void direct_drop_caches() {
    g_fixtab.release(); g_stats.release(); g_work.release(); g_item_scratch.release();
    g_fixtab_for = 0.0;
}

// the fp32 clamp value of a ligand atom of compact element e: the same float product the kernels form (hscale * lparam.w)
static float clamp_H(float hscale, int e) {
    const float xj = (float)std::max(e < kNumElt ? kEltXi[e] : 1.0, 1.0);
    volatile float h = hscale * xj;       // one IEEE float product, no contraction
    return h;
}

static int ensure_fix_tables(float hscale, bool shifted) {
    if (!g_stats.p) MMO_TRY(g_stats.alloc(4));
    const double key = shifted ? (double)hscale : -(double)hscale;
    if (g_fixtab_for != key) {
        std::vector<double4> ht(kEltTab * kEltTab);
        for (int a = 0; a < kEltTab; a++)            // receptor element
            for (int b = 0; b < kEltTab; b++) {      // ligand element
                const bool ok = a < kNumElt && b < kNumElt;
                // d_ij (p6^2 - 2 p6) with p6 = (x_i x_j / r^2)^3  ==  A s^6 - B s^3,  s = 1/r^2
                const double x2 = ok ? kEltXi[a] * kEltXi[b] : NAN, d = ok ? sqrt(kEltDi[a] * kEltDi[b]) : NAN;
                const double x6 = x2 * x2 * x2;
                // what the fast path adds for a clamped pair: the vdW term at r^2 = H, H being that of the LIGAND atom's
                // element, times the shift weight at H (SHIFTED)
                const double H = (double)clamp_H(hscale, b);
                const double t2 = x2 / H, p6 = t2 * t2 * t2;
                const double wH = shifted ? (1.0 - H / 144.0) * (1.0 - H / 144.0) : 1.0;
                ht[a * kEltTab + b] = make_double4(d * (x6 * x6), 2.0 * d * x6, wH * (d * (p6 * p6 - 2.0 * p6)), 0.0);
            }
        MMO_TRY(g_fixtab.upload(ht));
        g_fixtab_for = key;
    }
    return MMO_OK;
}

static int set_fast_smem(size_t smem) {
    static size_t done = 0;
    static int done_epoch = -1;
    if (done_epoch != rt().epoch) { done = 0; done_epoch = rt().epoch; }      // function attributes are per context
    if (smem <= done) return MMO_OK;
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_SHIFTED, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_fp32_kernel<MMO_VARIANT_GLOBAL, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = smem;
    return MMO_OK;
}

// ---- item mode launcher ------------------------------------------------------------------------------
static int set_items_smem(size_t smem) {
    static size_t done = 0;
    static int done_epoch = -1;
    if (done_epoch != rt().epoch) { done = 0; done_epoch = rt().epoch; }
    if (smem <= done) return MMO_OK;
    MMO_CUDA(cudaFuncSetAttribute(direct_items_kernel<MMO_VARIANT_SHIFTED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_items_kernel<MMO_VARIANT_SHIFTED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_items_kernel<MMO_VARIANT_GLOBAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(direct_items_kernel<MMO_VARIANT_GLOBAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = smem;
    return MMO_OK;
}

// poses [p_begin, p_begin + n) of src in item mode; d_out points at the first of them
static int launch_items_batch(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src_all,
                              int64_t p_begin, int64_t n_poses, double *d_out, bool collect_stats, const FastArgs &fa,
                              const FixArgs &xa) {
    Runtime &R = rt();
    PoseSrc src = src_all;      // view of the batch
    if (src.kind == 0) { src.rot9 += 9 * p_begin; src.trans3 += 3 * p_begin; }
    else if (src.kind == 1) { src.xs += p_begin * lig->n; src.ys += p_begin * lig->n; src.zs += p_begin * lig->n; }
    else src.frames += p_begin;
    const int nf = lig->n_fast;
    const size_t n_items = (size_t)n_poses * nf;
    // lattice of cells over everything within 12 A (+ one cell) of the receptor's bounding box
    float lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = (float)(rec->bb_lo[d] - rec->origin[d]) - 12.5f; hi[d] = (float)(rec->bb_hi[d] - rec->origin[d]) + 12.5f; }
    // 0.5 A cells in Z-order (Morton codes of 3 x 10 bits): 64 consecutive items of the sorted list are spatial
    // neighbours at every density -- a dense screen fills single cells (rho <= 0.43 A), a short list spans a compact block
    // of cells -- so there is nothing to tune to the list size.  The cell grows only if the lattice would need more than
    // 1024 cells per axis.  MMO_ITEM_CELL overrides (tuning).
    float cell = kItemCell;
    if (const char *e = getenv("MMO_ITEM_CELL")) { const float v = (float)atof(e); if (v >= 0.25f && v <= 8.0f) cell = v; }
    int nd[3];
    for (;;) {
        bool ok = true;
        for (int d = 0; d < 3; d++) { nd[d] = std::max(1, (int)ceilf((hi[d] - lo[d]) / cell)); ok = ok && nd[d] <= 1023; }
        if (ok) break;
        cell *= 1.25f;
    }
    // the sort only looks at the bits the lattice needs: b per axis with every nd <= 2^b - 1, so that the all-ones code is
    // free for "beyond the lattice" (sorted last) -- 24 bits = three radix passes for a lattice of up to 255 cells per axis
    int axis_bits = 1;
    while (std::max(nd[0], std::max(nd[1], nd[2])) > (1 << axis_bits) - 1) axis_bits++;
    const int end_bit = 3 * axis_bits;
    const uint32_t far_key = (1u << end_bit) - 1u;

    // scratch arena (grow-only, reused by every call: cudaMalloc/cudaFree of ~65 B per item would cost more than the kernels)
    size_t temp_bytes = 0;
    MMO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                             (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n_items, 0, end_bit, R.stream));
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t need = up(n_items * 32) + 2 * up(n_items * 8) + 4 * up(n_items * 4) + up(n_items) + up(temp_bytes);
    if (g_item_scratch.n < need) MMO_TRY(g_item_scratch.alloc(need + need / 8));
    uint8_t *cur = g_item_scratch.p;
    auto carve = [&](size_t b) { uint8_t *r = cur; cur += up(b); return r; };
    float4 *pos = (float4 *)carve(n_items * 32);
    double *e_item = (double *)carve(n_items * 8), *e_rank = (double *)carve(n_items * 8);
    uint32_t *keys = (uint32_t *)carve(n_items * 4), *keys2 = (uint32_t *)carve(n_items * 4);
    uint32_t *vals = (uint32_t *)carve(n_items * 4), *perm = (uint32_t *)carve(n_items * 4);
    uint8_t *f_rank = carve(n_items), *temp = carve(temp_bytes);
    if (!g_work.p) MMO_TRY(g_work.alloc(64));
    MMO_CUDA(cudaMemsetAsync(g_work.p, 0, 64 * sizeof(unsigned long long), R.stream));
    MMO_CUDA(cudaMemsetAsync(e_item, 0, n_items * sizeof(double), R.stream));
    unsigned long long *d_far = g_work.p + 63;
    {
        KernelScope ks(K_ITEM_PREP);
        item_prepare_kernel<<<(unsigned)((n_items + 256 * kPrepIPT - 1) / (256 * kPrepIPT)), 256, 0, R.stream>>>(
            src, n_poses, lig->n, nf, lig->fx.p, lig->fy.p, lig->fz.p, lig->forder.p, lig->fparam.p, rec->origin[0], rec->origin[1],
            rec->origin[2], lo[0], lo[1], lo[2], 1.0f / cell, nd[0], nd[1], nd[2], far_key, pos, keys, vals, d_far);
        MMO_LAUNCH_CHECK();
        MMO_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys2, vals, perm, (int64_t)n_items, 0, end_bit, R.stream));
        count_launch(3);
    }
    ItemArgs ia;
    for (int e = 0; e < kEltTab; e++) { ia.tab_A[e] = fa.tab_A[e]; ia.tab_B[e] = fa.tab_B[e]; }
    ia.xyzq = fa.xyzq; ia.gelt = fa.gelt; ia.blob_box = fa.blob_box; ia.sup_box = rec->sup_box.p;
    ia.n_blobs = rec->n_blobs; ia.n_sup = rec->n_sup; ia.lparam = fa.lparam; ia.n_fast = nf;
    ia.pos = pos; ia.perm = perm; ia.n_far = d_far; ia.n_items = n_items;
    ia.hscale = fa.hscale; ia.e_rank = e_rank; ia.f_rank = f_rank; ia.stats = fa.stats;
    MMO_REQUIRE(rec->n_blobs < 65535, "receptor too large for the direct kernel (%d atoms)", rec->n);
    const int near_cap = (rec->n_blobs + 8 + 7) & ~7;
    const size_t smem = ((size_t)2 * rec->n_sup + (size_t)nf) * sizeof(float4) + 16 * sizeof(float2) +
                        (size_t)(kItemTPB / 32) * NF * LIST_CAP * sizeof(float) + (size_t)(kItemTPB / 32) * near_cap * sizeof(uint16_t) + 16;
    MMO_REQUIRE(smem <= 200 * 1024, "receptor too large for the direct kernel (%d atoms)", rec->n);
    MMO_TRY(set_items_smem(smem));
    const int64_t n_units = ((int64_t)n_items + 63) / 64;
    const unsigned blocks = (unsigned)std::min<int64_t>(2LL * R.sm_count, (n_units + kItemTPB / 32 - 1) / (kItemTPB / 32));
    const bool shifted = variant == MMO_VARIANT_SHIFTED;
    {
        KernelScope ks(K_DIRECT_FP32);
        unsigned long long *w = g_work.p;
        if (shifted && collect_stats) direct_items_kernel<MMO_VARIANT_SHIFTED, true><<<blocks, kItemTPB, smem, R.stream>>>(ia, near_cap, w);
        else if (shifted) direct_items_kernel<MMO_VARIANT_SHIFTED, false><<<blocks, kItemTPB, smem, R.stream>>>(ia, near_cap, w);
        else if (collect_stats) direct_items_kernel<MMO_VARIANT_GLOBAL, true><<<blocks, kItemTPB, smem, R.stream>>>(ia, near_cap, w);
        else direct_items_kernel<MMO_VARIANT_GLOBAL, false><<<blocks, kItemTPB, smem, R.stream>>>(ia, near_cap, w);
        MMO_LAUNCH_CHECK();
    }
    KernelScope ks2(K_HARD_FIX);
    // close contacts per item, in the sorted (spatially coherent) order; then the items of a pose are summed in atom order
    const unsigned iblocks = (unsigned)((n_items + 127) / 128);
    if (shifted && collect_stats) item_fix_kernel<MMO_VARIANT_SHIFTED, true><<<iblocks, 128, 0, R.stream>>>(xa, rec->origin[0], rec->origin[1], rec->origin[2], nf, perm, pos, d_far, n_items, f_rank, e_rank, e_item);
    else if (shifted) item_fix_kernel<MMO_VARIANT_SHIFTED, false><<<iblocks, 128, 0, R.stream>>>(xa, rec->origin[0], rec->origin[1], rec->origin[2], nf, perm, pos, d_far, n_items, f_rank, e_rank, e_item);
    else if (collect_stats) item_fix_kernel<MMO_VARIANT_GLOBAL, true><<<iblocks, 128, 0, R.stream>>>(xa, rec->origin[0], rec->origin[1], rec->origin[2], nf, perm, pos, d_far, n_items, f_rank, e_rank, e_item);
    else item_fix_kernel<MMO_VARIANT_GLOBAL, false><<<iblocks, 128, 0, R.stream>>>(xa, rec->origin[0], rec->origin[1], rec->origin[2], nf, perm, pos, d_far, n_items, f_rank, e_rank, e_item);
    MMO_LAUNCH_CHECK();
    const unsigned fblocks = (unsigned)((n_poses + 127) / 128);
    hard_fix_kernel<MMO_VARIANT_SHIFTED, false, true><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, e_item, nf, nullptr, 1, 0, d_out);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_direct_fp32(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src,
                       int64_t n_poses, double *d_out, bool collect_stats) {
    if (n_poses == 0) return MMO_OK;
    Runtime &R = rt();
    // Without a cut-off every receptor atom contributes a Coulomb term of order 0.1-1 kcal/mol and the fp32
    // rounding noise of the sum grows as sqrt(pairs): the contract is verified up to ~1e5 pairs per pose
    // (worst |err| / tolerance 0.3) and lost near 2e5.  Larger GLOBAL problems take the strict fp64 kernel.
    if (variant == MMO_VARIANT_GLOBAL && (int64_t)rec->n * lig->n > kGlobalFp32MaxPairs)
        return launch_direct_fp64(rec, lig, variant, src, n_poses, d_out);
    // clamp / close-contact threshold on r^2 per ligand atom: H_j = hscale * x_j; floats, so that the fast kernel and the
    // fp64 pass see the same numbers
    const float hscale = (float)(std::max(rec->x_max, 1.0) / kTau);
    MMO_TRY(ensure_fix_tables(hscale, variant == MMO_VARIANT_SHIFTED));
    FastArgs fa;
    fa.n_blobs = rec->n_blobs;
    for (int e = 0; e < kEltTab; e++) vdw_factors(e, &fa.tab_A[e], &fa.tab_B[e]);
    fa.xyzq = rec->xyzq.p; fa.gelt = rec->gelt.p; fa.blob_box = rec->blob_box.p; fa.sup_box = rec->sup_box.p; fa.n_sup = rec->n_sup;
    for (int d = 0; d < 3; d++) fa.origin[d] = rec->origin[d];
    fa.L = lig->n;
    fa.n_fast = lig->n_fast;
    fa.lx = lig->fx.p; fa.ly = lig->fy.p; fa.lz = lig->fz.p;
    fa.forder = lig->forder.p;
    fa.lparam = lig->fparam.p;
    fa.hscale = hscale;
    fa.stats = g_stats.p;
    FixArgs xa;
    xa.pxyzq = rec->xyzq64.p; xa.pxyz32 = rec->xyz32v.p; xa.pelt = rec->elt.p;
    for (int d = 0; d < 3; d++) { xa.vox_lo[d] = rec->vox_lo[d]; xa.vox_dim[d] = rec->vox_dim[d]; }
    xa.vox_inv = 1.0 / rec->vox_edge;
    xa.vox_off = rec->vox_off.p; xa.vox_idx = rec->vox_idx.p;
    xa.L = lig->n;
    xa.lx = lig->x.p; xa.ly = lig->y.p; xa.lz = lig->z.p; xa.lq = lig->q.p; xa.lelt = lig->elt.p;
    xa.forder = lig->forder.p;
    for (int e = 0; e < kEltTab; e++) {
        const double H = (double)clamp_H(hscale, e);
        xa.H[e] = H;
        xa.wrH[e] = (variant == MMO_VARIANT_SHIFTED ? (1.0 - H / 144.0) * (1.0 - H / 144.0) : 1.0) / sqrt(H);
    }
    xa.tab = g_fixtab.p;
    xa.stats = g_stats.p;

    if (collect_stats) MMO_CUDA(cudaMemsetAsync(g_stats.p, 0, 4 * sizeof(unsigned long long), R.stream));
    // Pose lists that are not a scan (arbitrary rotations / explicit conformers) have no coherence a warp of
    // poses could cull with: large ones go through item mode, in batches that bound the scratch memory.
    if (src.kind != 2 && rec->n > 0 && g_direct_mode != 1 && variant == MMO_VARIANT_SHIFTED &&
        ((int64_t)n_poses * lig->n_fast >= kItemModeMin || g_direct_mode == 2)) {
        const int64_t batch = std::max<int64_t>(1, kItemBatch / lig->n_fast);
        for (int64_t p0 = 0; p0 < n_poses; p0 += batch)
            MMO_TRY(launch_items_batch(rec, lig, variant, src, p0, std::min(batch, n_poses - p0), d_out + p0, collect_stats, fa, xa));
        if (collect_stats) {
            unsigned long long h[4];
            MMO_CUDA(cudaMemcpyAsync(h, g_stats.p, sizeof h, cudaMemcpyDeviceToHost, R.stream));
            MMO_CUDA(cudaStreamSynchronize(R.stream));
            R.stat_pairs = (int64_t)h[0]; R.stat_inside = (int64_t)h[1]; R.stat_fp64 = (int64_t)h[2]; R.stat_flagged = (int64_t)h[3];
        }
        return MMO_OK;
    }
    // receptor: one shared-memory tile when it fits (<= 128 groups = 2048 atoms); larger receptors stay in global memory
    // and are read through L1 (untiled kernel): one launch either way
    const bool tiled = rec->n_blobs <= MAX_TILE_GROUPS;
    const int tile_blobs = tiled ? std::max(1, rec->n_blobs) : 0;
    MMO_REQUIRE(rec->n_blobs < 65535, "receptor too large for the direct kernel (%d atoms)", rec->n);
    const int near_cap = ((tiled ? MAX_TILE_GROUPS : rec->n_blobs) + 8 + 7) & ~7;
    const size_t smem = ((size_t)(tile_blobs + 1) * kBlob + (size_t)tile_blobs * 2 + (size_t)lig->n_fast) * sizeof(float4) +
                        16 * sizeof(float2) + (size_t)3 * LJ * PPB * sizeof(float) +
                        (size_t)(TPB / 32) * NF * LIST_CAP * sizeof(float) + (size_t)(tile_blobs + 1) * kBlob + 16 +
                        (size_t)(TPB / 32) * near_cap * sizeof(uint16_t) + 16;
    MMO_REQUIRE(smem <= 227 * 1024, "receptor too large for the direct kernel (%d atoms)", rec->n);
    // work units = (64 poses, chunk split): at least ~20 per resident warp, so that the dynamic hand-out
    // balances the warps to a few percent; persistent grid of one block per SM (fewer for small batches)
    const int n_chunks = lig->n_fast / LJ;
    const int64_t n_groups = (n_poses + 32 * PPT - 1) / (32 * PPT);
    const int64_t resident_warps = (int64_t)kBlocksPerSM * R.sm_count * (TPB / 32);
    const int n_split = (int)std::max<int64_t>(1, std::min<int64_t>(n_chunks, (20 * resident_warps + n_groups - 1) / n_groups));
    // small batches (single-pose calls: the reference's closure shape) also split the receptor, so that one pose is
    // ~100 warps of work and not 12
    const int n_tiles = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(8, std::max(1, rec->n_blobs / 8)), resident_warps / (2 * n_groups * n_split)));
    const int64_t n_units = n_groups * n_split * n_tiles;
    const unsigned blocks = (unsigned)std::min<int64_t>((int64_t)kBlocksPerSM * R.sm_count, (n_units + TPB / 32 - 1) / (TPB / 32));
    const int n_parts = n_tiles * n_split;
    DevBuf<double> part;
    if (n_parts > 1) MMO_TRY(part.alloc((size_t)n_parts * (size_t)n_poses));
    double *d_part = n_parts > 1 ? part.p : d_out;
    DevBuf<uint8_t> flags;
    MMO_TRY(flags.alloc((size_t)n_tiles * (size_t)n_chunks * (size_t)n_poses));
    if (!g_work.p) MMO_TRY(g_work.alloc(64));
    const bool shifted = variant == MMO_VARIANT_SHIFTED;
    if (rec->n > 0) {
        MMO_TRY(set_fast_smem(smem));
        MMO_CUDA(cudaMemsetAsync(g_work.p, 0, 64 * sizeof(unsigned long long), R.stream));
        {
        KernelScope ks(K_DIRECT_FP32);
        {
            const int b0 = 0, nb = rec->n_blobs;
            double *o = d_part;
            uint8_t *f = flags.p;
            unsigned long long *w = g_work.p;
#define MMO_LAUNCH_K1(V, S_, T_)                                                                                                   \
    do {                                                                                                                           \
        if (n_tiles > 1) direct_fp32_kernel<V, S_, T_, true><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, b0, nb, tile_blobs, n_split, n_tiles, near_cap, w, o, f); \
        else direct_fp32_kernel<V, S_, T_, false><<<blocks, TPB, smem, R.stream>>>(fa, src, n_poses, b0, nb, tile_blobs, n_split, 1, near_cap, w, o, f); \
    } while (0)
            if (tiled) {
                if (shifted && collect_stats) MMO_LAUNCH_K1(MMO_VARIANT_SHIFTED, true, true);
                else if (shifted) MMO_LAUNCH_K1(MMO_VARIANT_SHIFTED, false, true);
                else if (collect_stats) MMO_LAUNCH_K1(MMO_VARIANT_GLOBAL, true, true);
                else MMO_LAUNCH_K1(MMO_VARIANT_GLOBAL, false, true);
            } else {
                if (shifted && collect_stats) MMO_LAUNCH_K1(MMO_VARIANT_SHIFTED, true, false);
                else if (shifted) MMO_LAUNCH_K1(MMO_VARIANT_SHIFTED, false, false);
                else if (collect_stats) MMO_LAUNCH_K1(MMO_VARIANT_GLOBAL, true, false);
                else MMO_LAUNCH_K1(MMO_VARIANT_GLOBAL, false, false);
            }
#undef MMO_LAUNCH_K1
            MMO_LAUNCH_CHECK();
        }
        }
        KernelScope ks2(K_HARD_FIX);
        const unsigned fblocks = (unsigned)((n_poses + 127) / 128);
        if (n_poses <= kPoseFixBlockMax) {
            const unsigned pb = (unsigned)n_poses;
            const size_t sm = (size_t)lig->n_fast * sizeof(double);
            MMO_REQUIRE(sm <= 48 * 1024, "ligand too large for the direct kernel (%d atoms)", lig->n);
            if (shifted && collect_stats) pose_fix_block_kernel<MMO_VARIANT_SHIFTED, true><<<pb, 64, sm, R.stream>>>(xa, src, n_poses, lig->n_fast, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
            else if (shifted) pose_fix_block_kernel<MMO_VARIANT_SHIFTED, false><<<pb, 64, sm, R.stream>>>(xa, src, n_poses, lig->n_fast, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
            else if (collect_stats) pose_fix_block_kernel<MMO_VARIANT_GLOBAL, true><<<pb, 64, sm, R.stream>>>(xa, src, n_poses, lig->n_fast, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
            else pose_fix_block_kernel<MMO_VARIANT_GLOBAL, false><<<pb, 64, sm, R.stream>>>(xa, src, n_poses, lig->n_fast, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
        } else
        if (shifted && collect_stats) hard_fix_kernel<MMO_VARIANT_SHIFTED, true, false><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
        else if (shifted) hard_fix_kernel<MMO_VARIANT_SHIFTED, false, false><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
        else if (collect_stats) hard_fix_kernel<MMO_VARIANT_GLOBAL, true, false><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
        else hard_fix_kernel<MMO_VARIANT_GLOBAL, false, false><<<fblocks, 128, 0, R.stream>>>(xa, src, n_poses, d_part, n_parts, flags.p, n_tiles, n_chunks, d_out);
        MMO_LAUNCH_CHECK();
    } else {
        MMO_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n_poses * sizeof(double), R.stream));
    }
    if (collect_stats) {
        unsigned long long h[4];
        MMO_CUDA(cudaMemcpyAsync(h, g_stats.p, sizeof h, cudaMemcpyDeviceToHost, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));
        R.stat_pairs = (int64_t)h[0]; R.stat_inside = (int64_t)h[1]; R.stat_fp64 = (int64_t)h[2]; R.stat_flagged = (int64_t)h[3];
    }
    return MMO_OK;
}

}  // namespace mmo
