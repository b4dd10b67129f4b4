// direct_fp32.cu -- K1: batched UFF Lennard-Jones + Coulomb pair sum, fp32 pair arithmetic.
//
// Replaces the inner loops of Mol.ene_inter_UFF_shifted_brute / _global_brute (src/mol.ml:796-849)
// for many poses per launch.  Layout of the computation (no tensor cores: non-linear pair sum):
//
//   thread  = one pose, block = 256 poses; the receptor (k-d groups of 32 atoms, fp32, relative to the
//             receptor origin) is staged in shared memory, the whole ROI receptor at once when it fits
//   cull    = per (warp, ligand atom): the atom's bounding box over the warp's 32 poses is tested first
//             against the group boxes (one ballot per 32 groups), then against the individual atoms of
//             the near groups (one ballot per group); survivors are compacted into a per-warp list.
//             Shifted variant only: a culled pair has weight exactly 0 in the reference (mol.ml:836)
//   pair    = 20 SASS instructions (1 MUFU.RSQ), see pair_energy(); the list is consumed 8 atoms at a
//             time = 8 independent dependency chains per warp
//   sum     = 8 pair terms in fp32, then one F2F + DADD into a per-thread fp64 accumulator
//
// Accuracy contract (MMO_PREC_FP32): |E - E_ref| <= max(1e-6 |E_ref|, 1e-4 kcal/mol).  fp32 cannot
// deliver that for close contacts (r^-12), so the fast path clamps r^2 at H = x_max_rec*x_max_lig/kTau
// and a second, sparse kernel (hard_fix_kernel) adds  e64(r) - e64(sqrt(H))  in the reference's own
// double arithmetic for the few pairs with r^2 < H, found through the receptor's voxel lists.
#include "common.cuh"
#include "pose.cuh"
#include <math.h>

namespace mmo {

constexpr int LJ = 8;            // ligand atoms per chunk (= one k-d leaf of the ligand)
constexpr int TPB = 256;         // poses per block
constexpr int LIST_CAP = 384;    // per-warp list of near receptor atoms (shared-memory addresses)
constexpr int MAX_TILE_GROUPS = 64;
static_assert(kBlob == 32, "one receptor group per warp-wide test");

struct FastArgs {
    int n_blobs;             // receptor groups of 32 atoms (k-d leaves)
    int n_atoms;             // real receptor atoms (the last group may be padded)
    const float4 *xyzq;
    const float2 *ab;
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

// one receptor atom against one ligand atom; adds w * (EW q_i q_j / r + d_ij (p6^2 - 2 p6)) to acc
template <int VARIANT>
__device__ __forceinline__ float pair_energy(float dx, float dy, float dz, float qi, float Ai, float Bi,
                                             float qj, float Aj, float Bj, float H, float acc) {
    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float r2c = fmaxf(r2, H);                       // close contacts are finished in fp64 elsewhere
    float rinv = rsqrt_fast(r2c);
    float s = rinv * rinv;
    float s3 = s * s * s;
    float v = fmaf(Ai * Aj, s3, -(Bi * Bj));        // (A_i A_j) s^3 - B_i B_j
    float er = (qi * qj) * rinv;                    // qi already carries 332.0637/4
    float e = fmaf(v, s3, er);
    if (VARIANT == MMO_VARIANT_SHIFTED) {
        // shift weight (1 - r^2/144)^2 from the unclamped r^2, saturated to [0,1]: one FFMA.SAT gives
        // exactly 0 beyond the 12 A cut-off (FF.shift_12A, FF.ml:17-20)
        float u = __saturatef(fmaf(r2, -1.0f / 144.0f, 1.0f));
        return fmaf(u * u, e, acc);
    } else {
        return acc + e;
    }
}

// squared distance from point p to the box [lo, hi]
__device__ __forceinline__ float box_dist2(const float4 p, const float *lo, const float *hi) {
    float gx = fmaxf(0.f, fmaxf(lo[0] - p.x, p.x - hi[0]));
    float gy = fmaxf(0.f, fmaxf(lo[1] - p.y, p.y - hi[1]));
    float gz = fmaxf(0.f, fmaxf(lo[2] - p.z, p.z - hi[2]));
    return fmaf(gz, gz, fmaf(gy, gy, gx * gx));
}

// Shared memory (dynamic): receptor tile as an array of 32-byte atoms {x, y, z, 83.0159*q | A, B, 0, 0}
// [tile_atoms + 1], group boxes [2*tile_groups], ligand parameters [n_fast], chunk coordinates [LJ][TPB],
// per-warp near lists (shared-memory byte addresses of the atoms, so that the pair loop needs no index
// arithmetic).  Slot tile_atoms is a dummy atom (far away, no charge, no vdW) that pads a list to 8.
template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(TPB, 2)
direct_fp32_kernel(FastArgs a, PoseSrc src, int64_t n_poses, int tile_groups, double *__restrict__ out) {
    extern __shared__ float4 smem4[];
    const int tile_atoms = tile_groups * kBlob;
    float4 *s_atom = smem4;                                   // 2 * (tile_atoms + 1)
    float4 *s_box = s_atom + 2 * (tile_atoms + 1);            // tile_groups * 2
    float4 *s_lparam = s_box + tile_groups * 2;               // n_fast
    float4 *s_c = s_lparam + a.n_fast;                        // LJ * TPB : {x, y, z, -} of chunk atom jj, pose tid
    unsigned *s_list = (unsigned *)(s_c + LJ * TPB) + (threadIdx.x >> 5) * LIST_CAP;
    const unsigned atom_base = (unsigned)__cvta_generic_to_shared(s_atom);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t p = (int64_t)blockIdx.x * TPB + tid;
    const bool valid = p < n_poses;
    const int64_t pp = valid ? p : n_poses - 1;   // idle lanes shadow the last pose, result discarded
    for (int j = tid; j < a.n_fast; j += TPB) s_lparam[j] = a.lparam[j];

    double acc = 0.0;
    unsigned long long n_eval = 0, n_in = 0;
    const int n_chunks = a.n_fast / LJ;
    const int n_tiles = (a.n_blobs + tile_groups - 1) / tile_groups;

    for (int t = 0; t < n_tiles; t++) {
        // ---- stage a receptor tile (the whole ROI receptor when it fits: one tile, two barriers) ----
        __syncthreads();
        const int b0 = t * tile_groups;
        const int nb = min(tile_groups, a.n_blobs - b0);
        const int n_real = min(nb * kBlob, a.n_atoms - b0 * kBlob);      // real atoms in this tile
        for (int k = tid; k < nb * kBlob; k += TPB) {
            const float2 ab = __ldg(a.ab + (size_t)b0 * kBlob + k);
            s_atom[2 * k] = __ldg(a.xyzq + (size_t)b0 * kBlob + k);
            s_atom[2 * k + 1] = make_float4(ab.x, ab.y, 0.f, 0.f);
        }
        if (tid == 0) {
            s_atom[2 * tile_atoms] = make_float4(1e6f, 1e6f, 1e6f, 0.f);
            s_atom[2 * tile_atoms + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
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
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s_lparam[k].w != 0.f) {
                        double x, y, z;
                        if (src.kind == 1) {
                            const int j = __ldg(a.forder + k);
                            x = src.xs[pp * a.L + j]; y = src.ys[pp * a.L + j]; z = src.zs[pp * a.L + j];
                        } else {
                            pose_atom_rt(P, __ldg(a.lx + k), __ldg(a.ly + k), __ldg(a.lz + k), x, y, z);
                        }
                        v.x = (float)(x - a.origin[0]);
                        v.y = (float)(y - a.origin[1]);
                        v.z = (float)(z - a.origin[2]);
                    }
                    s_c[jj * TPB + tid] = v;
                }
            }
#pragma unroll 1
            for (int jj = 0; jj < LJ; jj++) {
                const float4 lp = s_lparam[c * LJ + jj];
                if (lp.w == 0.f) continue;                              // padding atom (warp-uniform)
                const float4 lc = s_c[jj * TPB + tid];
                // bounding box of this ligand atom over the warp's 32 poses
                float lo[3] = {lc.x, lc.y, lc.z}, hi[3] = {lc.x, lc.y, lc.z};
                if (VARIANT == MMO_VARIANT_SHIFTED) {
#pragma unroll
                    for (int d = 0; d < 3; d++) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
                            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
                        }
                    }
                }
                // ---- level 1: which groups of 32 receptor atoms can be within 12 A of this atom? ----
                unsigned gm0 = 0xffffffffu, gm1 = 0xffffffffu;      // near masks of groups 0-31 / 32-63
                if (VARIANT == MMO_VARIANT_SHIFTED) {
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
                        const unsigned m = __ballot_sync(0xffffffffu, near);
                        if (r == 0) gm0 = m; else gm1 = m;
                    }
                }
                // ---- level 2: per-atom test, 4 groups per step (independent loads and tests), survivors
                //      compacted into the warp's list; the list is consumed 8 atoms at a time ----
                int n = 0;
                for (int g4 = 0; g4 < nb || n > 0; g4 += 4) {
                    if (g4 < nb) {
                        const unsigned m4 = ((g4 < 32 ? gm0 : gm1) >> (g4 & 31)) & 0xfu;
                        if (m4 != 0u) {
                            bool nr[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int atom = (g4 + u) * kBlob + lane;
                                nr[u] = ((m4 >> u) & 1u) && atom < n_real;
                                if (VARIANT == MMO_VARIANT_SHIFTED) {
                                    // (out-of-tile slots are never read: clamp the address, keep the predicate)
                                    const float4 pa = s_atom[2 * min(atom, tile_atoms)];
                                    nr[u] = nr[u] && box_dist2(pa, lo, hi) < 144.0f;
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const unsigned bm = __ballot_sync(0xffffffffu, nr[u]);
                                if (nr[u]) s_list[n + __popc(bm & lt_mask)] = atom_base + (unsigned)((g4 + u) * kBlob + lane) * 32u;
                                n += __popc(bm);
                            }
                        }
                        if (n <= LIST_CAP - 128 && g4 + 4 < nb) continue;       // room for 4 more groups
                    }
                    if (n == 0) continue;
                    // ---- process the list: 8 independent pair chains per step ----
                    if (STATS) n_eval += (unsigned long long)n;
                    const int n8 = (n + 7) & ~7;
                    if (lane < n8 - n) s_list[n + lane] = atom_base + (unsigned)tile_atoms * 32u;     // pad with the dummy atom
                    __syncwarp();
#pragma unroll 1
                    for (int k = 0; k < n8; k += 8) {
                        const uint4 pk0 = *(const uint4 *)(s_list + k), pk1 = *(const uint4 *)(s_list + k + 4);
                        const unsigned ad[8] = {pk0.x, pk0.y, pk0.z, pk0.w, pk1.x, pk1.y, pk1.z, pk1.w};
                        float f = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            float4 ra, rp;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ra.x), "=f"(ra.y), "=f"(ra.z), "=f"(ra.w) : "r"(ad[i]));
                            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+16];" : "=f"(rp.x), "=f"(rp.y) : "r"(ad[i]));
                            float dx = ra.x - lc.x, dy = ra.y - lc.y, dz = ra.z - lc.z;
                            f = pair_energy<VARIANT>(dx, dy, dz, ra.w, rp.x, rp.y, lp.z, lp.x, lp.y, a.H, f);
                            if (STATS) {
                                float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                                if (r2 < 144.0f && ra.x < 1e5f) n_in++;
                            }
                        }
                        acc += (double)f;
                    }
                    n = 0;
                    __syncwarp();
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
// For every pair with r^2 < H the fast path evaluated e(sqrt(H)); this pass adds e(r) - e(sqrt(H)) in
// double.  Same formulas as mol.ml:811-815 / 838-845 (r clamped at 0.01, p6 = (x_ij/r)^6, shift weight),
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
                double d = e - eH;
                if (VARIANT == MMO_VARIANT_SHIFTED) {
                    const double u = 1.0 - r2c * (1.0 / 144.0);        // the fast path weighted e(H) with w(r), not w(H)
                    d *= u * u;
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
    fa.n_atoms = rec->n;
    fa.xyzq = rec->xyzq.p; fa.ab = rec->ab.p; fa.blob_box = rec->blob_box.p;
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
    xa.xx = g_xx.p; xa.dij = g_dij.p; xa.vdwH = g_vdwH.p;
    xa.stats = g_stats.p;

    if (collect_stats) MMO_CUDA(cudaMemsetAsync(g_stats.p, 0, 4 * sizeof(unsigned long long), R.stream));
    const unsigned blocks = (unsigned)((n_poses + TPB - 1) / TPB);
    // receptor tile: everything when it fits (<= 64 groups = 2048 atoms), so that 2 blocks stay resident per SM
    const int tile_blobs = std::max(1, std::min(rec->n_blobs, MAX_TILE_GROUPS));
    const size_t smem = (2 * ((size_t)tile_blobs * kBlob + 1) + (size_t)tile_blobs * 2 + (size_t)lig->n_fast + (size_t)LJ * TPB) * sizeof(float4) +
                        (size_t)(TPB / 32) * LIST_CAP * sizeof(unsigned) + 16;
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
