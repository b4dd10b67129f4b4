// strict_fp64.cu -- the reference's double-precision arithmetic, operation for operation.
// Compiled with -fmad=false: no FMA contraction (OCaml native code never fuses), IEEE sqrt and
// division, and the reference's summation order, so results are bit-identical to the OCaml code
// (and to oracle/mmo_oracle.c).
//
//   direct   : Mol.ene_inter_UFF_shifted_brute / _global_brute   src/mol.ml:796-849
//   comps    : Mol.ene_inter_UFF_shifted_bst_components          src/mol.ml:928-956
//   intra    : Mol.ene_intra_UFFNB_brute                         src/mol.ml:881-903
//   grid     : Mol.ene_inter_UFF_shifted_grid + Lds.pre_calculate_FF_components_grid
//                                                                src/mol.ml:964-989, src/lds.ml:452-469
//   trilin   : G3D.trilin, Mol.ene_inter_UFF_interp              src/G3D.ml:97-157, src/mol.ml:1012-1020
#include "common.cuh"
#include "pose.cuh"
#include "strict_dev.cuh"
#include <math.h>
#include <algorithm>

namespace mmo {

// UFF.ml:32-51: x_ij = sqrt(x_i*x_j), d_ij = sqrt(D_i*D_j), NaN for unsupported elements
__constant__ double c_xij[kEltTab * kEltTab];
__constant__ double c_dij[kEltTab * kEltTab];
static int g_tables_epoch = -1;     // __constant__ memory is per device: re-uploaded after every mmo_init (Runtime::epoch)

static int ensure_tables() {
    if (g_tables_epoch == rt().epoch) return MMO_OK;
    double hx[kEltTab * kEltTab], hd[kEltTab * kEltTab];
    for (int a = 0; a < kEltTab; a++)
        for (int b = 0; b < kEltTab; b++) {
            if (a < kNumElt && b < kNumElt) {
                hx[a * kEltTab + b] = sqrt(kEltXi[a] * kEltXi[b]);   // FF.geo_mean, FF.ml:14-15
                hd[a * kEltTab + b] = sqrt(kEltDi[a] * kEltDi[b]);
            } else {
                hx[a * kEltTab + b] = NAN;
                hd[a * kEltTab + b] = NAN;
            }
        }
    MMO_CUDA(cudaMemcpyToSymbol(c_xij, hx, sizeof hx));
    MMO_CUDA(cudaMemcpyToSymbol(c_dij, hd, sizeof hd));
    g_tables_epoch = rt().epoch;
    return MMO_OK;
}

// coordinates of one pose into shared memory, laid out [coord][atom][thread]
__device__ __forceinline__ void stage_pose(const PoseSrc &src, int64_t p, int L, const double *lx,
                                           const double *ly, const double *lz, double *sm, int nthr, int tid) {
    double *sx = sm, *sy = sm + (size_t)L * nthr, *sz = sm + 2 * (size_t)L * nthr;
    if (src.kind == 1) {
        for (int j = 0; j < L; j++) {
            sx[j * nthr + tid] = src.xs[p * L + j];
            sy[j * nthr + tid] = src.ys[p * L + j];
            sz[j * nthr + tid] = src.zs[p * L + j];
        }
    } else {
        PoseRT P;
        load_pose_rt(src, p, P);
        for (int j = 0; j < L; j++) {
            double x, y, z;
            pose_atom_rt(P, lx[j], ly[j], lz[j], x, y, z);
            sx[j * nthr + tid] = x; sy[j * nthr + tid] = y; sz[j * nthr + tid] = z;
        }
    }
}

// ---- direct, receptor outer / ligand inner (mol.ml:802-818, 828-848) ------------------------------
template <int VARIANT>
__global__ void strict_direct_kernel(int P, const double *__restrict__ px, const double *__restrict__ py,
                                     const double *__restrict__ pz, const double *__restrict__ pq,
                                     const int32_t *__restrict__ pelt,
                                     int L, const double *__restrict__ lx, const double *__restrict__ ly,
                                     const double *__restrict__ lz, const double *__restrict__ lq,
                                     const int32_t *__restrict__ lelt,
                                     PoseSrc src, int64_t n_poses, double *__restrict__ out) {
    extern __shared__ double sm[];
    const int nthr = blockDim.x, tid = threadIdx.x;
    double *sx = sm, *sy = sm + (size_t)L * nthr, *sz = sm + 2 * (size_t)L * nthr;
    double *sq = sm + 3 * (size_t)L * nthr;
    int32_t *se = (int32_t *)(sq + L);
    for (int j = tid; j < L; j += nthr) { sq[j] = lq[j]; se[j] = lelt[j]; }
    int64_t p = (int64_t)blockIdx.x * nthr + tid;
    if (p < n_poses) stage_pose(src, p, L, lx, ly, lz, sm, nthr, tid);
    __syncthreads();
    if (p >= n_poses) return;
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int i = 0; i < P; i++) {
        const double xi = __ldg(px + i), yi = __ldg(py + i), zi = __ldg(pz + i), q_i = __ldg(pq + i);
        const int ei = __ldg(pelt + i) * kEltTab;
        for (int j = 0; j < L; j++) {
            double r2 = d_dist2(xi, yi, zi, sx[j * nthr + tid], sy[j * nthr + tid], sz[j * nthr + tid]);
            if (VARIANT == MMO_VARIANT_SHIFTED) {
                if (r2 < 144.0) {
                    double q_j = sq[j];
                    double r = d_nzd(sqrt(r2));
                    double w = d_shift(r);
                    int t = ei + se[j];
                    const Divisor by_r = make_divisor(r);
                    double p6 = d_pow6(div_by(c_xij[t], by_r));
                    sum_elec = sum_elec + w * div_by(q_i * q_j, by_r);
                    sum_vdW = sum_vdW + w * (c_dij[t] * ((-2.0 * p6) + (p6 * p6)));
                }
            } else {
                double q_j = sq[j];
                double r = d_nzd(sqrt(r2));
                int t = ei + se[j];
                const Divisor by_r = make_divisor(r);
                double p6 = d_pow6(div_by(c_xij[t], by_r));
                sum_elec = sum_elec + div_by(q_i * q_j, by_r);
                sum_vdW = sum_vdW + (c_dij[t] * ((-2.0 * p6) + (p6 * p6)));
            }
        }
    }
    out[p] = (kElecWeight * sum_elec) + sum_vdW;
}

// The same sum for a moderate number of poses (single evaluations of the `ene_inter` closure, the re-scoring stage of the
// fp64 scan): block = pose.  Tile by tile of 64 receptor atoms, thread t counts the ligand atoms within the cut-off of
// receptor atom i0 + t, a block-wide prefix sum gives every thread its place in the reference's (i outer, j inner) order,
// the terms are computed in parallel and stored there, and thread 0 adds them up one by one: the same doubles in the same
// order as the loop above, 0.1 ms per pose instead of the 9.6 ms one thread needs for 1837 x 48 pairs.
constexpr int kDirTile = 64;
template <int VARIANT>
__global__ void __launch_bounds__(kDirTile)
strict_direct_block_kernel(int P, const double *__restrict__ px, const double *__restrict__ py,
                           const double *__restrict__ pz, const double *__restrict__ pq,
                           const int32_t *__restrict__ pelt,
                           int L, const double *__restrict__ lx, const double *__restrict__ ly,
                           const double *__restrict__ lz, const double *__restrict__ lq,
                           const int32_t *__restrict__ lelt,
                           PoseSrc src, int64_t n_poses, double *__restrict__ out) {
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + L, *sz = sm + 2 * L, *sq = sm + 3 * L;
    int32_t *se = (int32_t *)(sq + L);                                   // L ints
    int *s_cnt = (int *)(se + L);                                        // kDirTile + 1
    double2 *terms = (double2 *)(sm + ((4 * L * 2 + L + kDirTile + 2 + 3) / 4) * 2);     // 16-byte aligned, behind the ints
    const int tid = threadIdx.x;
    const int64_t p = blockIdx.x;
    if (src.kind == 1) {
        for (int j = tid; j < L; j += kDirTile) { sx[j] = src.xs[p * L + j]; sy[j] = src.ys[p * L + j]; sz[j] = src.zs[p * L + j]; }
    } else {
        PoseRT Pp;
        load_pose_rt(src, p, Pp);
        for (int j = tid; j < L; j += kDirTile) pose_atom_rt(Pp, lx[j], ly[j], lz[j], sx[j], sy[j], sz[j]);
    }
    for (int j = tid; j < L; j += kDirTile) { sq[j] = lq[j]; se[j] = lelt[j]; }
    __syncthreads();
    double sum_elec = 0.0, sum_vdW = 0.0;            // thread 0 only
    for (int i0 = 0; i0 < P; i0 += kDirTile) {
        const int i = i0 + tid;
        double xi = 0.0, yi = 0.0, zi = 0.0, q_i = 0.0;
        int ei = 0, cnt = 0;
        if (i < P) {
            xi = __ldg(px + i); yi = __ldg(py + i); zi = __ldg(pz + i); q_i = __ldg(pq + i);
            ei = __ldg(pelt + i) * kEltTab;
            if (VARIANT == MMO_VARIANT_SHIFTED) {
                for (int j = 0; j < L; j++) cnt += d_dist2(xi, yi, zi, sx[j], sy[j], sz[j]) < 144.0 ? 1 : 0;
            } else {
                cnt = L;
            }
        }
        // exclusive prefix sum of the 64 counts (two warps)
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
        if ((tid & 31) == 31) s_cnt[tid >> 5] = incl;
        __syncthreads();
        const int base = (tid >> 5) ? s_cnt[0] : 0;
        const int total = s_cnt[0] + s_cnt[1];
        int pos = base + incl - cnt;
        if (i < P) {
            for (int j = 0; j < L; j++) {
                const double r2 = d_dist2(xi, yi, zi, sx[j], sy[j], sz[j]);
                if (VARIANT == MMO_VARIANT_SHIFTED) {
                    if (r2 < 144.0) {
                        const double r = d_nzd(sqrt(r2));
                        const double w = d_shift(r);
                        const int t = ei + se[j];
                        const Divisor by_r = make_divisor(r);
                        const double p6 = d_pow6(div_by(c_xij[t], by_r));
                        terms[pos++] = make_double2(w * div_by(q_i * sq[j], by_r), w * (c_dij[t] * ((-2.0 * p6) + (p6 * p6))));
                    }
                } else {
                    const double r = d_nzd(sqrt(r2));
                    const int t = ei + se[j];
                    const Divisor by_r = make_divisor(r);
                    const double p6 = d_pow6(div_by(c_xij[t], by_r));
                    terms[pos++] = make_double2(div_by(q_i * sq[j], by_r), c_dij[t] * ((-2.0 * p6) + (p6 * p6)));
                }
            }
        }
        __syncthreads();
        if (tid == 0)
            for (int k = 0; k < total; k++) { const double2 u = terms[k]; sum_elec = sum_elec + u.x; sum_vdW = sum_vdW + u.y; }
        __syncthreads();
    }
    if (tid == 0) out[p] = (kElecWeight * sum_elec) + sum_vdW;
}

// ---- components, ligand outer / receptor inner in index order (mol.ml:932-956) --------------------
__global__ void strict_components_kernel(int P, const double *__restrict__ px, const double *__restrict__ py,
                                         const double *__restrict__ pz, const double *__restrict__ pq,
                                         const int32_t *__restrict__ pelt,
                                         int L, const double *__restrict__ lx, const double *__restrict__ ly,
                                         const double *__restrict__ lz, const double *__restrict__ lq,
                                         const int32_t *__restrict__ lelt,
                                         PoseSrc src, int64_t n_poses, double *__restrict__ out_elec,
                                         double *__restrict__ out_vdw) {
    extern __shared__ double sm[];
    const int nthr = blockDim.x, tid = threadIdx.x;
    double *sx = sm, *sy = sm + (size_t)L * nthr, *sz = sm + 2 * (size_t)L * nthr;
    int64_t p = (int64_t)blockIdx.x * nthr + tid;
    if (p < n_poses) stage_pose(src, p, L, lx, ly, lz, sm, nthr, tid);
    __syncthreads();
    if (p >= n_poses) return;
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int j = 0; j < L; j++) {
        const double xj = sx[j * nthr + tid], yj = sy[j * nthr + tid], zj = sz[j * nthr + tid];
        const double q_j = __ldg(lq + j);
        const int ej = __ldg(lelt + j) * kEltTab;
        for (int i = 0; i < P; i++) {
            double d = sqrt(d_dist2(__ldg(px + i), __ldg(py + i), __ldg(pz + i), xj, yj, zj));
            if (d <= 12.0) {   // BST.neighbors q 12.0; w(12) = 0 makes the boundary convention irrelevant
                double r = d_nzd(d);
                double w = d_shift(r);
                int t = ej + __ldg(pelt + i);
                const Divisor by_r = make_divisor(r);
                double p6 = d_pow6(div_by(c_xij[t], by_r));
                sum_elec = sum_elec + w * div_by(__ldg(pq + i) * q_j, by_r);
                sum_vdW = sum_vdW + w * (c_dij[t] * ((-2.0 * p6) + (p6 * p6)));
            }
        }
    }
    out_elec[p] = sum_elec * kElecWeight;
    out_vdw[p] = sum_vdW;
}

// ---- intra-ligand non-bonded (mol.ml:885-903): pairs pre-listed in the reference's (i<j) order ----
__global__ void strict_intra_kernel(int L, int n_pairs, const int32_t *__restrict__ pair_i,
                                    const int32_t *__restrict__ pair_j, const double *__restrict__ lq,
                                    const int32_t *__restrict__ lelt, int64_t n_confs,
                                    const double *__restrict__ xs, const double *__restrict__ ys,
                                    const double *__restrict__ zs, double *__restrict__ out) {
    extern __shared__ double sm[];
    const int nthr = blockDim.x, tid = threadIdx.x;
    double *sx = sm, *sy = sm + (size_t)L * nthr, *sz = sm + 2 * (size_t)L * nthr;
    int64_t p = (int64_t)blockIdx.x * nthr + tid;
    if (p < n_confs)
        for (int j = 0; j < L; j++) {
            sx[j * nthr + tid] = xs[p * L + j];
            sy[j * nthr + tid] = ys[p * L + j];
            sz[j * nthr + tid] = zs[p * L + j];
        }
    __syncthreads();
    if (p >= n_confs) return;
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int k = 0; k < n_pairs; k++) {
        const int i = __ldg(pair_i + k), j = __ldg(pair_j + k);
        double r = d_nzd(sqrt(d_dist2(sx[i * nthr + tid], sy[i * nthr + tid], sz[i * nthr + tid],
                                      sx[j * nthr + tid], sy[j * nthr + tid], sz[j * nthr + tid])));
        int t = __ldg(lelt + i) * kEltTab + __ldg(lelt + j);
        const Divisor by_r = make_divisor(r);
        double p6 = d_pow6(div_by(c_xij[t], by_r));
        sum_elec = sum_elec + div_by(__ldg(lq + i) * __ldg(lq + j), by_r);
        sum_vdW = sum_vdW + c_dij[t] * ((-2.0 * p6) + (p6 * p6));
    }
    out[p] = (kElecWeight * sum_elec) + sum_vdW;
}

// The same energy for a FEW conformers (the literal `ene_intra : Mol.t -> float` closure, one conformer per call): block =
// conformer, the pair terms are computed by all threads in parallel, then one thread adds them up in the reference's
// (i<j) order -- the same doubles as the loop above, 30 us instead of 380 us for one 48-atom conformer.
__global__ void __launch_bounds__(128)
strict_intra_block_kernel(int L, int n_pairs, const int32_t *__restrict__ pair_i, const int32_t *__restrict__ pair_j,
                          const double *__restrict__ lq, const int32_t *__restrict__ lelt, const double *__restrict__ xs,
                          const double *__restrict__ ys, const double *__restrict__ zs, double *__restrict__ out) {
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + L, *sz = sm + 2 * L;
    double2 *terms = (double2 *)(sm + 3 * L + ((3 * L) & 1));
    const int64_t p = blockIdx.x;
    for (int j = threadIdx.x; j < L; j += blockDim.x) { sx[j] = xs[p * L + j]; sy[j] = ys[p * L + j]; sz[j] = zs[p * L + j]; }
    __syncthreads();
    for (int k = threadIdx.x; k < n_pairs; k += blockDim.x) {
        const int i = __ldg(pair_i + k), j = __ldg(pair_j + k);
        const double r = d_nzd(sqrt(d_dist2(sx[i], sy[i], sz[i], sx[j], sy[j], sz[j])));
        const int t = __ldg(lelt + i) * kEltTab + __ldg(lelt + j);
        const Divisor by_r = make_divisor(r);
        const double p6 = d_pow6(div_by(c_xij[t], by_r));
        terms[k] = make_double2(div_by(__ldg(lq + i) * __ldg(lq + j), by_r), c_dij[t] * ((-2.0 * p6) + (p6 * p6)));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum_elec = 0.0, sum_vdW = 0.0;
        for (int k = 0; k < n_pairs; k++) { const double2 u = terms[k]; sum_elec = sum_elec + u.x; sum_vdW = sum_vdW + u.y; }
        out[p] = (kElecWeight * sum_elec) + sum_vdW;
    }
}

// ---- grid build (mol.ml:964-989 per voxel, lds.ml:452-469 clamp + f32 store) ----------------------
// block = brick of 8x4x4 voxels, thread = (voxel, group of kTG types).  The block first filters the
// receptor down to the atoms within 12 A of the brick (ordered compaction: the survivors keep their
// index order, so every voxel still adds its neighbours in the reference's order), staging them in
// shared memory tile by tile; r_ij and w are computed once per atom and reused by the kTG types, as
// in the reference.  Atoms farther than 12 A from a voxel fail the exact `d <= 12` test exactly as
// before, so the culling cannot change a single bit of the maps.
constexpr int kTG = 12;
constexpr int kGridTile = 256;
constexpr int kBrickX = 8, kBrickY = 4, kBrickZ = 4;
__global__ void __launch_bounds__(128, 8)
strict_grid_kernel(int P, const double *__restrict__ px, const double *__restrict__ py,
                   const double *__restrict__ pz, const double *__restrict__ pq,
                   const int32_t *__restrict__ pelt,
                   int dim0, int dim1, int dim2, double q0, double q1, double q2,
                   int nbx, int nby,
                   const uint32_t *__restrict__ mask, int T, const int32_t *__restrict__ telt,
                   const double *__restrict__ tq, const int32_t *__restrict__ tidx, float *__restrict__ maps) {
    __shared__ double sx[kGridTile], sy[kGridTile], sz[kGridTile], sq[kGridTile];
    __shared__ int32_t se[kGridTile];
    __shared__ int s_wcount[4];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t nvox = (size_t)dim0 * dim1 * dim2;
    const int brick = blockIdx.x;
    const int bz = brick / (nbx * nby), by = (brick - bz * nbx * nby) / nbx, bx = brick - (bz * nby + by) * nbx;
    const int i = bx * kBrickX + (tid & 7), j = by * kBrickY + ((tid >> 3) & 3), k = bz * kBrickZ + (tid >> 5);
    const bool in_grid = i < dim0 && j < dim1 && k < dim2;
    const size_t idx = (size_t)i + (size_t)j * dim0 + (size_t)k * dim0 * dim1;     // Grid.idx_of_ijk, grid.ml:98-99
    bool active = in_grid;
    if (active && mask) active = (mask[idx >> 5] >> (idx & 31)) & 1u;
    if (!__syncthreads_or(active)) return;              // brick entirely outside the bitmask (lds.ml:485)
    const double x = (double)i * q0, y = (double)j * q1, z = (double)k * q2;      // grid.xs.(i) etc.
    // the brick's box (all its lattice nodes, clipped to the grid)
    const double blo[3] = {(double)(bx * kBrickX) * q0, (double)(by * kBrickY) * q1, (double)(bz * kBrickZ) * q2};
    const double bhi[3] = {(double)min(bx * kBrickX + kBrickX - 1, dim0 - 1) * q0,
                           (double)min(by * kBrickY + kBrickY - 1, dim1 - 1) * q1,
                           (double)min(bz * kBrickZ + kBrickZ - 1, dim2 - 1) * q2};
    const int t0 = blockIdx.y * kTG;
    const int nt = min(kTG, T - t0);
    double se_acc[kTG], sv_acc[kTG], tqv[kTG];
    int te[kTG];
    unsigned same = 0u;         // bit l: type l has the element of type l - 1 (types arrive sorted by element)
#pragma unroll
    for (int l = 0; l < kTG; l++) {
        se_acc[l] = 0.0; sv_acc[l] = 0.0;
        tqv[l] = (l < nt) ? tq[t0 + l] : 0.0;
        te[l] = (l < nt) ? telt[t0 + l] * kEltTab : 0;
        if (l > 0 && l < nt && te[l] == te[l - 1]) same |= 1u << l;
    }
    int count = 0;                                       // candidates staged in shared memory (block-uniform)
    for (int base = 0; base < P || count > 0; base += 128) {
        // ---- ordered compaction of the next 128 atoms into the tile ----
        if (base < P) {
            const int a = base + tid;
            bool near = false;
            double ax = 0.0, ay = 0.0, az = 0.0;
            if (a < P) {
                ax = px[a]; ay = py[a]; az = pz[a];
                const double gx = fmax(0.0, fmax(blo[0] - ax, ax - bhi[0]));
                const double gy = fmax(0.0, fmax(blo[1] - ay, ay - bhi[1]));
                const double gz = fmax(0.0, fmax(blo[2] - az, az - bhi[2]));
                near = gx * gx + gy * gy + gz * gz <= 144.0 + 1e-6;     // superset of "within 12 A of some voxel"
            }
            const unsigned bm = __ballot_sync(0xffffffffu, near);
            if (lane == 0) s_wcount[wid] = __popc(bm);
            __syncthreads();
            int off = count;
            for (int w = 0; w < wid; w++) off += s_wcount[w];
            const int total = s_wcount[0] + s_wcount[1] + s_wcount[2] + s_wcount[3];
            if (near) {
                const int pos = off + __popc(bm & ((1u << lane) - 1u));
                sx[pos] = ax; sy[pos] = ay; sz[pos] = az; sq[pos] = pq[a]; se[pos] = pelt[a];
            }
            count += total;
            __syncthreads();
            if (count <= kGridTile - 128 && base + 128 < P) continue;      // room for another batch
        }
        // ---- every voxel of the brick visits the staged atoms in order ----
        if (active) {
            for (int c = 0; c < count; c++) {
                double d = sqrt(d_dist2(sx[c], sy[c], sz[c], x, y, z));
                if (d <= 12.0) {
                    const double q_i = sq[c];
                    const double r = d_nzd(d);
                    const double w = d_shift(r);
                    const int ei = se[c];
                    const Divisor by_r = make_divisor(r);       // 2 * kTG quotients by the same r (bit-identical to '/')
                    double vdw = 0.0;           // w * d_ij (p6^2 - 2 p6): depends on the two elements only
#pragma unroll
                    for (int l = 0; l < kTG; l++) {
                        if (l < nt) {
                            if (!((same >> l) & 1u)) {
                                int t = te[l] + ei;       // UFF.vdW_xiDi (get_anum lig 0) prot_anum
                                double p6 = d_pow6(div_by(c_xij[t], by_r));
                                vdw = w * (c_dij[t] * ((-2.0 * p6) + (p6 * p6)));
                            }
                            se_acc[l] = se_acc[l] + w * div_by(q_i * tqv[l], by_r);
                            sv_acc[l] = sv_acc[l] + vdw;
                        }
                    }
                }
            }
        }
        count = 0;
        __syncthreads();
    }
    if (!active) return;
#pragma unroll
    for (int l = 0; l < kTG; l++) {
        if (l < nt) {
            double e = kElecWeight * se_acc[l] + sv_acc[l];
            double v = (kMaxE <= e) ? kMaxE : e;      // OCaml: min max_E e = if max_E <= e then max_E else e
            maps[(size_t)tidx[t0 + l] * nvox + idx] = (float)v;      // the caller's type order
        }
    }
}

__global__ void strict_trilin_kernel(GridGeom g, const float *__restrict__ arr, int64_t n,
                                     const double *__restrict__ xs, const double *__restrict__ ys,
                                     const double *__restrict__ zs, double *__restrict__ out) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = d_trilin(g, arr, xs[p], ys[p], zs[p]);
}

#ifndef MMO_INTERP_UNROLL
#define MMO_INTERP_UNROLL 4
#endif
constexpr int kInterpUnroll = MMO_INTERP_UNROLL;
// Mol.ene_inter_UFF_interp (mol.ml:1012-1020): res := !res +. trilin ... for j = 0 .. l-1
// ZP: read the z-pair copy of the maps (maps = float2 [T][z_dim - 1][y_dim][x_dim], zvox elements per type)
#ifndef MMO_INTERP_MINB
#define MMO_INTERP_MINB 8
#endif
template <bool ZP>
__global__ void __launch_bounds__(128, MMO_INTERP_MINB)
strict_interp_kernel(GridGeom g, const void *__restrict__ maps_, size_t zvox, int L,
                     const double *__restrict__ lx, const double *__restrict__ ly,
                     const double *__restrict__ lz, const int32_t *__restrict__ ltyp,
                     PoseSrc src, int64_t n_poses, double *__restrict__ out) {
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + L, *sz = sm + 2 * L;
    int32_t *st = (int32_t *)(sm + 3 * L);
    for (int j = threadIdx.x; j < L; j += blockDim.x) { sx[j] = lx[j]; sy[j] = ly[j]; sz[j] = lz[j]; st[j] = ltyp[j]; }
    __syncthreads();
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_poses) return;
    double res = 0.0;
    const float *maps = (const float *)maps_;
    const float2 *zmaps = (const float2 *)maps_;
    auto look = [&](int t, double x, double y, double z) {
        return ZP ? d_trilin_zp(g, zmaps + (size_t)t * zvox, x, y, z) : d_trilin(g, maps + (size_t)t * g.nvox, x, y, z);
    };
    // unrolled by kInterpUnroll: the gathers of several atoms are in flight together (the kernel waits on L2), the sum keeps its order
    if (src.kind == 1) {
#pragma unroll kInterpUnroll
        for (int j = 0; j < L; j++)
            res = res + look(st[j], src.xs[p * L + j], src.ys[p * L + j], src.zs[p * L + j]);
    } else {
        PoseRT P;
        load_pose_rt(src, p, P);
#pragma unroll kInterpUnroll
        for (int j = 0; j < L; j++) {
            double x, y, z;
            pose_atom_rt(P, sx[j], sy[j], sz[j], x, y, z);
            res = res + look(st[j], x, y, z);
        }
    }
    out[p] = res;
}

// zpair[t][k][j][i] = {map[t][k][j][i], map[t][k + 1][j][i]}, k < z_dim - 1
__global__ void __launch_bounds__(256)
zpair_build_kernel(const float *__restrict__ maps, size_t nvox, size_t zvox, int T, float2 *__restrict__ zp, size_t xy) {
    const size_t n = zvox * (size_t)T;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t t = e / zvox, v = e - t * zvox;
        zp[e] = make_float2(__ldg(maps + t * nvox + v), __ldg(maps + t * nvox + v + xy));
    }
}

// ---- self-test of div_by against the IEEE division (tests/test_gpu_direct.py) --------------------------------
__global__ void division_selftest_kernel(uint64_t seed, int64_t n, unsigned long long *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto mix = [](uint64_t z) { z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); };
    const uint64_t u = mix(seed + 2 * (uint64_t)i), v = mix(seed + 2 * (uint64_t)i + 1);
    // divisor: the path's range [0.01, 12.5) with random significands; every 16th one a significand of (nearly) all
    // ones or all zeros, the known hard cases of reciprocal-based division
    double b = __longlong_as_double((long long)((u >> 12) | 0x3ff0000000000000ull));          // [1, 2)
    if ((i & 15) == 0) b = __longlong_as_double((long long)(0x3ff0000000000000ull | (0x000fffffffffffffull - (u & 7))));
    if ((i & 15) == 1) b = __longlong_as_double((long long)(0x3ff0000000000000ull | (u & 7)));
    b = ldexp(b, (int)(v % 11) - 7);                                                           // 2^-7 .. 2^3
    double a = __longlong_as_double((long long)((v >> 12) | 0x3ff0000000000000ull));
    a = ldexp(a, (int)((v >> 4) % 31) - 15);
    if (v & 1) a = -a;
    const Divisor d = make_divisor(b);
    const double q = div_by(a, d), want = a / b;
    if (__double_as_longlong(q) != __double_as_longlong(want)) atomicAdd(bad, 1ull);
}

// ---- host launchers ------------------------------------------------------------------------------
static int pick_threads(int L, size_t extra_bytes, size_t *smem) {
    // [3][L][threads] doubles of pose coordinates; keep at most ~100 KB per block
    for (int t : {128, 64, 32}) {
        size_t b = (size_t)3 * L * t * sizeof(double) + extra_bytes;
        if (b <= 100 * 1024 || t == 32) { *smem = b; return t; }
    }
    return 32;
}

int launch_direct_fp64(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src,
                       int64_t n_poses, double *d_out) {
    MMO_TRY(ensure_tables());
    if (n_poses == 0) return MMO_OK;
    size_t smem;
    int L = lig->n;
    {   // up to a few ten thousand poses: block per pose (0.1 ms each, spread over the GPU) beats thread per pose (9.6 ms flat)
        const size_t bsmem = ((size_t)(4 * L * 2 + L + kDirTile + 2 + 3) / 4) * 2 * sizeof(double) + (size_t)kDirTile * L * sizeof(double2);
        if (n_poses < 32768 && bsmem <= 200 * 1024 && rec->n > 0) {
            auto kb = (variant == MMO_VARIANT_SHIFTED) ? strict_direct_block_kernel<MMO_VARIANT_SHIFTED> : strict_direct_block_kernel<MMO_VARIANT_GLOBAL>;
            MMO_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
            KernelScope ks(K_DIRECT_FP64);
            kb<<<(unsigned)n_poses, kDirTile, bsmem, rt().stream>>>(rec->n, rec->x.p, rec->y.p, rec->z.p, rec->q.p, rec->elt.p, L, lig->x.p,
                                                                     lig->y.p, lig->z.p, lig->q.p, lig->elt.p, src, n_poses, d_out);
            MMO_LAUNCH_CHECK();
            return MMO_OK;
        }
    }
    int threads = pick_threads(L, (size_t)L * (sizeof(double) + sizeof(int32_t)) + 16, &smem);
    MMO_REQUIRE(smem <= 220 * 1024, "ligand with %d atoms is too large for the fp64 direct kernel", L);
    int64_t blocks = (n_poses + threads - 1) / threads;
    auto k = (variant == MMO_VARIANT_SHIFTED) ? strict_direct_kernel<MMO_VARIANT_SHIFTED>
                                              : strict_direct_kernel<MMO_VARIANT_GLOBAL>;
    MMO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelScope ks(K_DIRECT_FP64);
    k<<<(unsigned)blocks, threads, smem, rt().stream>>>(rec->n, rec->x.p, rec->y.p, rec->z.p, rec->q.p, rec->elt.p,
                                                        L, lig->x.p, lig->y.p, lig->z.p, lig->q.p, lig->elt.p,
                                                        src, n_poses, d_out);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_components_fp64(const mmo_receptor *rec, const mmo_ligand *lig, const PoseSrc &src,
                           int64_t n_poses, double *d_elec, double *d_vdw) {
    MMO_TRY(ensure_tables());
    if (n_poses == 0) return MMO_OK;
    size_t smem;
    int L = lig->n;
    int threads = pick_threads(L, 0, &smem);
    MMO_REQUIRE(smem <= 220 * 1024, "ligand with %d atoms is too large for the fp64 components kernel", L);
    int64_t blocks = (n_poses + threads - 1) / threads;
    MMO_CUDA(cudaFuncSetAttribute(strict_components_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    strict_components_kernel<<<(unsigned)blocks, threads, smem, rt().stream>>>(
        rec->n, rec->x.p, rec->y.p, rec->z.p, rec->q.p, rec->elt.p, L, lig->x.p, lig->y.p, lig->z.p, lig->q.p,
        lig->elt.p, src, n_poses, d_elec, d_vdw);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_intra_fp64(const mmo_ligand *lig, int64_t n_confs, const double *d_xs, const double *d_ys,
                      const double *d_zs, double *d_out) {
    MMO_TRY(ensure_tables());
    if (n_confs == 0) return MMO_OK;
    size_t smem;
    int L = lig->n;
    {   // a few conformers: block per conformer (terms in parallel, ordered sum by one thread)
        const size_t bsmem = ((size_t)3 * L + 1) * sizeof(double) + (size_t)std::max(lig->n_pairs, 1) * sizeof(double2);
        if (n_confs <= 2LL * rt().sm_count && bsmem <= 200 * 1024) {
            if (bsmem > 48 * 1024) MMO_CUDA(cudaFuncSetAttribute(strict_intra_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
            KernelScope ks(K_INTRA);
            strict_intra_block_kernel<<<(unsigned)n_confs, 128, bsmem, rt().stream>>>(L, lig->n_pairs, lig->pair_i.p, lig->pair_j.p, lig->q.p,
                                                                                   lig->elt.p, d_xs, d_ys, d_zs, d_out);
            MMO_LAUNCH_CHECK();
            return MMO_OK;
        }
    }
    int threads = pick_threads(L, 0, &smem);
    MMO_REQUIRE(smem <= 220 * 1024, "ligand with %d atoms is too large for the intra kernel", L);
    int64_t blocks = (n_confs + threads - 1) / threads;
    MMO_CUDA(cudaFuncSetAttribute(strict_intra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelScope ks(K_INTRA);
    strict_intra_kernel<<<(unsigned)blocks, threads, smem, rt().stream>>>(
        L, lig->n_pairs, lig->pair_i.p, lig->pair_j.p, lig->q.p, lig->elt.p, n_confs, d_xs, d_ys, d_zs, d_out);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_grid_build(const mmo_receptor *rec, const mmo_grid *g, const uint32_t *d_mask_words,
                      const int32_t *d_type_elt, const double *d_type_q, const int32_t *d_type_idx) {
    MMO_TRY(ensure_tables());
    GridGeom G = geom_of(g);
    const int nbx = (g->dims[0] + kBrickX - 1) / kBrickX, nby = (g->dims[1] + kBrickY - 1) / kBrickY,
              nbz = (g->dims[2] + kBrickZ - 1) / kBrickZ;
    dim3 grid((unsigned)(nbx * nby * nbz), (unsigned)((g->T + kTG - 1) / kTG));
    KernelScope ks(K_GRID_BUILD);
    strict_grid_kernel<<<grid, 128, 0, rt().stream>>>(rec->n, rec->x.p, rec->y.p, rec->z.p, rec->q.p, rec->elt.p,
                                                      g->dims[0], g->dims[1], g->dims[2], G.q[0], G.q[1], G.q[2],
                                                      nbx, nby, d_mask_words, g->T, d_type_elt, d_type_q, d_type_idx, g->maps.p);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_trilin(const mmo_grid *g, int type, int64_t n, const double *d_x, const double *d_y,
                  const double *d_z, double *d_out) {
    if (n == 0) return MMO_OK;
    strict_trilin_kernel<<<(unsigned)((n + 255) / 256), 256, 0, rt().stream>>>(
        geom_of(g), g->maps.p + (size_t)type * g->nvox, n, d_x, d_y, d_z, d_out);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

// the z-pair look-up copy of a grid's maps, built on first use (the maps of a grid handle never change after its creation)
int grid_zpairs(const mmo_grid *g, const float2 **out) {
    if (!g->zpair_ready) {
        const size_t xy = (size_t)g->dims[0] * g->dims[1], zvox = xy * (size_t)(g->dims[2] - 1);
        MMO_TRY(g->zpair.alloc(zvox * (size_t)g->T));
        zpair_build_kernel<<<rt().sm_count * 8, 256, 0, rt().stream>>>(g->maps.p, g->nvox, zvox, g->T, g->zpair.p, xy);
        MMO_LAUNCH_CHECK();
        g->zpair_ready = true;
    }
    *out = g->zpair.p;
    return MMO_OK;
}

int launch_interp(const mmo_grid *g, const mmo_ligand *lig, const PoseSrc &src, int64_t n_poses, double *d_out) {
    if (n_poses == 0) return MMO_OK;
    int L = lig->n;
    size_t smem = (size_t)3 * L * sizeof(double) + (size_t)L * sizeof(int32_t) + 16;
    MMO_REQUIRE(smem <= 220 * 1024, "ligand with %d atoms is too large for the interpolation kernel", L);
    // look-ups read the z-pair copy of the maps (half the L2 sectors per look-up); MMO_INTERP_ZPAIR=0 keeps the plain maps
    static const bool want_zp = [] { const char *e = getenv("MMO_INTERP_ZPAIR"); return !(e && e[0] == '0'); }();
    const float2 *zp = nullptr;
    if (want_zp) MMO_TRY(grid_zpairs(g, &zp));
    const size_t zvox = (size_t)g->dims[0] * g->dims[1] * (g->dims[2] - 1);
    if (smem > 48 * 1024) {
        MMO_CUDA(cudaFuncSetAttribute(strict_interp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MMO_CUDA(cudaFuncSetAttribute(strict_interp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    KernelScope ks(K_INTERP);
    if (zp) strict_interp_kernel<true><<<(unsigned)((n_poses + 127) / 128), 128, smem, rt().stream>>>(
            geom_of(g), zp, zvox, L, lig->x.p, lig->y.p, lig->z.p, lig->typ.p, src, n_poses, d_out);
    else strict_interp_kernel<false><<<(unsigned)((n_poses + 127) / 128), 128, smem, rt().stream>>>(
            geom_of(g), g->maps.p, zvox, L, lig->x.p, lig->y.p, lig->z.p, lig->typ.p, src, n_poses, d_out);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

}  // namespace mmo

namespace mmo {
int division_selftest(uint64_t seed, int64_t n, int64_t *mismatches) {
    DevBuf<unsigned long long> bad;
    MMO_TRY(bad.alloc(1));
    MMO_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(unsigned long long), rt().stream));
    division_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, rt().stream>>>(seed, n, bad.p);
    MMO_LAUNCH_CHECK();
    unsigned long long h = 0;
    MMO_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof h, cudaMemcpyDeviceToHost, rt().stream));
    MMO_CUDA(cudaStreamSynchronize(rt().stream));
    *mismatches = (int64_t)h;
    return MMO_OK;
}
}  // namespace mmo
