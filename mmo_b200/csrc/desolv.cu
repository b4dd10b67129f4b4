// desolv.cu -- N4: the Majeux-Scarsi-Caflisch desolvation sums next to the scoring path (rescoring terms).
// Compiled with -fmad=false: IEEE double, no contraction, the reference's evaluation order => bit-identical sums.
//   Lds.protein_desolv        src/lds.ml:204-236   per voxel of the protein's first solvent shell inside the ROI:
//                                                  Const.desolvation * (voxel_vol * sum_j (q_j / d2_j)^2) over the protein
//                                                  atoms within Const.charged_cutoff (BST.neighbors)
//   Lds.desolvation_penalty   src/lds.ml:239-267   per ligand pose: desolvated voxels = protein shell AND ligand shell;
//                                                  prot = sum of the voxels' protein contributions, lig = Const.desolvation *
//                                                  (voxel_vol * sum_voxels sum_j (q_j / d2)^2 for d2 < 144)
//   Lds.first_solvent_shell   src/lds.ml:172-184   (ligand shell: inside r + r_H2O of some atom, inside r of none)
//   Const                     src/const.ml:10-32   r_H2O = 1.4, cut-off 12 A, eps_prot = 4, eps_HOH = 78.5
//   Grid.ijk_of_idx           src/grid.ml:101-105; ROI.is_inside src/ROI.ml:65-66 (strict <)
// Unpinned (un-vendored bst library): the order in which BST.neighbors lists the atoms; restated in atom index
// order, as everywhere else.  Its radius test is taken as dist <= 12.0.
#include "common.cuh"
#include "pose.cuh"
#include <limits.h>
#include <math.h>
#include <algorithm>

namespace mmo {

constexpr int kDesolvTPB = 128;
constexpr int kDesolvTerms = 1024;         // (voxel, atom) terms formed in parallel, then added in order by one thread
constexpr int kDesolvMaxAtoms = 512;       // ligand atoms staged per pose (dynamic shared memory: 40 B each)

__device__ __forceinline__ double dist2_dev(double ux, double uy, double uz, double vx, double vy, double vz) {
    const double dx = ux - vx, dy = uy - vy, dz = uz - vz;      // V3.dist2 u v (src/V3.ml:23-28)
    return dx * dx + dy * dy + dz * dz;
}

// phase 1: qualifying voxels (shell bit set, inside the ROI) appended to a list; their order is irrelevant,
// every voxel's sum is independent.  thread = one 32-voxel word of the mask
__global__ void __launch_bounds__(256)
desolv_select_kernel(const uint32_t *__restrict__ words, size_t nbits, int dim0, int dim1, double q0, double q1, double q2,
                     double cx, double cy, double cz, double r2, uint32_t *__restrict__ list,
                     unsigned long long *__restrict__ count) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= nbits) return;
    uint32_t bits = __ldg(words + w);
    const size_t xy = (size_t)dim0 * dim1;
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        const size_t idx = w * 32 + b;
        if (idx >= nbits) break;
        const int k = (int)(idx / xy), j = (int)((idx - (size_t)k * xy) / dim0), i = (int)(idx - ((size_t)k * xy + (size_t)j * dim0));
        const double x = (double)i * q0, y = (double)j * q1, z = (double)k * q2;
        if (dist2_dev(cx, cy, cz, x, y, z) < r2)                  // ROI.is_inside roi x_p
            list[atomicAdd(count, 1ull)] = (uint32_t)idx;
    }
}

// phase 2: thread = one selected voxel; protein atoms staged through shared memory, summed in index order
__global__ void __launch_bounds__(kDesolvTPB)
desolv_protein_kernel(const uint32_t *__restrict__ list, const unsigned long long *__restrict__ count, int dim0, int dim1,
                      double q0, double q1, double q2, int P, const double *__restrict__ px, const double *__restrict__ py,
                      const double *__restrict__ pz, const double *__restrict__ pq, double voxel_vol, double k_desolv,
                      double *__restrict__ contribs) {
    __shared__ double s_x[kDesolvTPB], s_y[kDesolvTPB], s_z[kDesolvTPB], s_q[kDesolvTPB];
    const unsigned long long n = *count;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((unsigned long long)blockIdx.x * blockDim.x >= n) return;
    const bool live = t < n;
    const uint32_t idx = live ? __ldg(list + t) : 0u;
    const size_t xy = (size_t)dim0 * dim1;
    const int k = (int)(idx / xy), j = (int)((idx - (size_t)k * xy) / dim0), i = (int)(idx - ((size_t)k * xy + (size_t)j * dim0));
    const double x = (double)i * q0, y = (double)j * q1, z = (double)k * q2;
    double res = 0.0;
    for (int a0 = 0; a0 < P; a0 += kDesolvTPB) {
        __syncthreads();
        const int a = a0 + threadIdx.x;
        if (a < P) { s_x[threadIdx.x] = px[a]; s_y[threadIdx.x] = py[a]; s_z[threadIdx.x] = pz[a]; s_q[threadIdx.x] = pq[a]; }
        __syncthreads();
        const int m = min(kDesolvTPB, P - a0);
        if (live)
            for (int c = 0; c < m; c++) {
                // BST.neighbors (Atom.create (-1) x_p) Const.charged_cutoff prot_bst: Atom.dist = sqrt(dist2)
                if (sqrt(dist2_dev(x, y, z, s_x[c], s_y[c], s_z[c])) <= 12.0) {
                    const double d2 = dist2_dev(x, y, z, s_x[c], s_y[c], s_z[c]);     // V3.dist2 x_p x_j
                    const double v = s_q[c] / d2;
                    res = res + v * v;
                }
            }
    }
    if (live) contribs[idx] = k_desolv * (voxel_vol * res);
}

// Lds.desolvation_penalty for many poses: block = pose.  The box of voxels that can belong to the ligand's solvent
// shell is swept in voxel-index order (Bitv.iteri_true), kDesolvTPB voxels per step; every thread decides its voxel
// (ordered compaction of the desolvated ones); their (voxel, atom) terms are formed in parallel and added one by one
// by thread 0 in the reference's order, so both totals are the reference's sums, bit for bit.
__global__ void __launch_bounds__(kDesolvTPB)
desolv_penalty_kernel(const uint32_t *__restrict__ shell, const double *__restrict__ contribs, int dim0, int dim1, int dim2,
                      double step, double q0, double q1, double q2, int L, const double *__restrict__ lx,
                      const double *__restrict__ ly, const double *__restrict__ lz, const double *__restrict__ lq,
                      const double *__restrict__ lr, PoseSrc src, int64_t n_poses, double voxel_vol, double k_desolv,
                      double *__restrict__ out_prot, double *__restrict__ out_lig) {
    extern __shared__ double s_dyn[];
    double *s_x = s_dyn, *s_y = s_x + L, *s_z = s_y + L, *s_q = s_z + L, *s_r = s_q + L;
    __shared__ double s_cv[kDesolvTPB], s_term[kDesolvTerms];
    __shared__ int s_hit[kDesolvTPB];
    __shared__ unsigned s_ballot[kDesolvTPB / 32];
    __shared__ int s_lo[3], s_hi[3];
    const int64_t p = blockIdx.x;
    if (p >= n_poses) return;
    const int tid = threadIdx.x;
    if (tid < 3) { s_lo[tid] = INT_MAX; s_hi[tid] = INT_MIN; }
    PoseRT Pz;
    if (src.kind != 1) load_pose_rt(src, p, Pz);
    __syncthreads();
    for (int j = tid; j < L; j += kDesolvTPB) {
        double x, y, z;
        if (src.kind == 1) { x = src.xs[p * L + j]; y = src.ys[p * L + j]; z = src.zs[p * L + j]; }
        else pose_atom_rt(Pz, __ldg(lx + j), __ldg(ly + j), __ldg(lz + j), x, y, z);
        s_x[j] = x; s_y[j] = y; s_z[j] = z; s_q[j] = __ldg(lq + j); s_r[j] = __ldg(lr + j);
        // atom_bitmask_set's cube for radius r + r_H2O (src/lds.ml:148-170, Grid.coord_of_point src/grid.ml:87-91)
        const int rs = (int)ceil((s_r[j] + 1.4) / step);
        const int ci = (int)((x - 0.0) / step), cj = (int)((y - 0.0) / step), ck = (int)((z - 0.0) / step);
        atomicMin(&s_lo[0], ci - rs); atomicMax(&s_hi[0], ci + rs);
        atomicMin(&s_lo[1], cj - rs); atomicMax(&s_hi[1], cj + rs);
        atomicMin(&s_lo[2], ck - rs); atomicMax(&s_hi[2], ck + rs);
    }
    __syncthreads();
    const int lo0 = max(s_lo[0], 0), lo1 = max(s_lo[1], 0), lo2 = max(s_lo[2], 0);
    const int hi0 = min(s_hi[0], dim0 - 1), hi1 = min(s_hi[1], dim1 - 1), hi2 = min(s_hi[2], dim2 - 1);
    const int b0 = hi0 - lo0 + 1, b1 = hi1 - lo1 + 1, b2 = hi2 - lo2 + 1;
    double prot = 0.0, lig = 0.0;                                  // thread 0's running sums
    if (b0 > 0 && b1 > 0 && b2 > 0) {
        const long total = (long)b0 * b1 * b2;
        const size_t xy = (size_t)dim0 * dim1;
        for (long t0 = 0; t0 < total; t0 += kDesolvTPB) {
            const long t = t0 + tid;
            bool hit = false;
            double cv = 0.0;
            if (t < total) {
                const int ii = (int)(t % b0), jj = (int)((t / b0) % b1), kk = (int)(t / ((long)b0 * b1));
                const int i = lo0 + ii, j = lo1 + jj, k = lo2 + kk;
                const size_t idx = (size_t)i + (size_t)j * dim0 + (size_t)k * xy;
                if ((__ldg(shell + (idx >> 5)) >> (idx & 31)) & 1u) {          // protein shell bit first: cheapest test
                    const double x = (double)i * q0, y = (double)j * q1, z = (double)k * q2;
                    bool in_shell = false, in_vdw = false;
                    for (int a = 0; a < L; a++) {
                        const double d2 = dist2_dev(s_x[a], s_y[a], s_z[a], x, y, z);    // V3.dist2 xyz (make x y z)
                        const double rw = s_r[a] + 1.4;
                        in_shell |= d2 < rw * rw;
                        in_vdw |= d2 < s_r[a] * s_r[a];
                    }
                    if (in_shell && !in_vdw) {
                        hit = true;
                        cv = __ldg(contribs + idx);
                    }
                }
            }
            // ordered compaction of the step's desolvated voxels
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if ((tid & 31) == 0) s_ballot[tid >> 5] = bal;
            __syncthreads();
            int before = 0, n_hit = 0;
#pragma unroll
            for (int w = 0; w < kDesolvTPB / 32; w++) {
                const int c = __popc(s_ballot[w]);
                if (w < (tid >> 5)) before += c;
                n_hit += c;
            }
            if (hit) {
                const int slot = before + __popc(bal & ((1u << (tid & 31)) - 1u));
                s_hit[slot] = tid;
                s_cv[slot] = cv;
            }
            __syncthreads();
            // the reference keeps ONE running sum over (voxel, ligand atom) in that order (lds.ml:252-259): the terms
            // are formed in parallel, kDesolvTerms at a time, and added one by one by thread 0.  A term beyond the hard
            // cut-off is stored as +0.0, which leaves a non-negative running sum unchanged in every bit.
            const int n_terms = n_hit * L;
            for (int e0 = 0; e0 < n_terms; e0 += kDesolvTerms) {
                for (int e = e0 + tid; e < min(n_terms, e0 + kDesolvTerms); e += kDesolvTPB) {
                    const int h = e / L, a = e - h * L;
                    const long tv = t0 + s_hit[h];
                    const int ii = (int)(tv % b0), jj = (int)((tv / b0) % b1), kk = (int)(tv / ((long)b0 * b1));
                    const double x = (double)(lo0 + ii) * q0, y = (double)(lo1 + jj) * q1, z = (double)(lo2 + kk) * q2;
                    const double d2 = dist2_dev(x, y, z, s_x[a], s_y[a], s_z[a]);       // V3.dist2 x_p x_j
                    double term = 0.0;
                    if (d2 < 144.0) {                                                   // Const.charged_cutoff_squared
                        const double v = s_q[a] / d2;
                        term = v * v;
                    }
                    s_term[e - e0] = term;
                }
                __syncthreads();
                if (tid == 0) {
                    const int m = min(kDesolvTerms, n_terms - e0);
                    for (int e = 0; e < m; e++) lig = lig + s_term[e];
                }
                __syncthreads();
            }
            if (tid == 0)
                for (int h = 0; h < n_hit; h++) prot = prot + s_cv[h];
            __syncthreads();
        }
    }
    if (tid == 0) {
        out_prot[p] = prot;
        out_lig[p] = k_desolv * (voxel_vol * lig);
    }
}

static void node_steps(const mmo_mask *m, double q[3]) {
    for (int d = 0; d < 3; d++) {
        const int np = m->dims[d] - 1;
        q[d] = np > 0 ? (m->step * (double)np) / (double)np : 0.0;      // Grid.from_box: frange (src/grid.ml:49-51)
    }
}
static double desolv_constant() {
    const double pi = 4.0 * atan(1.0);                                   // src/math.ml:13
    return (1.0 / 4.0 - 1.0 / 78.5) / (8.0 * pi);                       // src/const.ml:31
}

int launch_desolv_protein(const mmo_receptor *rec, const mmo_mask *shell, const double roi[4], double *d_contribs) {
    Runtime &R = rt();
    double q[3];
    node_steps(shell, q);
    const size_t nwords = (shell->nbits + 31) / 32;
    MMO_REQUIRE(shell->nbits < ((size_t)1 << 32), "mmo_desolv_protein: grid too large (%zu voxels)", shell->nbits);
    DevBuf<uint32_t> list;
    DevBuf<unsigned long long> count;
    // every set bit may qualify; bounded by the voxels of the ROI's bounding cube
    size_t cap = 1;
    for (int d = 0; d < 3; d++) cap *= (size_t)std::min<double>((double)shell->dims[d], 2.0 * roi[3] / shell->step + 3.0);
    MMO_TRY(list.alloc(cap));
    MMO_TRY(count.alloc(1));
    MMO_CUDA(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), R.stream));
    MMO_CUDA(cudaMemsetAsync(d_contribs, 0, shell->nbits * sizeof(double), R.stream));
    desolv_select_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, R.stream>>>(
        shell->words.p, shell->nbits, shell->dims[0], shell->dims[1], q[0], q[1], q[2], roi[0], roi[1], roi[2], roi[3] * roi[3],
        list.p, count.p);
    MMO_LAUNCH_CHECK();
    unsigned long long n = 0;
    MMO_CUDA(cudaMemcpyAsync(&n, count.p, sizeof n, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaStreamSynchronize(R.stream));
    if (n == 0) return MMO_OK;
    const double step = shell->step;
    desolv_protein_kernel<<<(unsigned)((n + kDesolvTPB - 1) / kDesolvTPB), kDesolvTPB, 0, R.stream>>>(
        list.p, count.p, shell->dims[0], shell->dims[1], q[0], q[1], q[2], rec->n, rec->x.p, rec->y.p, rec->z.p, rec->q.p,
        step * step * step, desolv_constant(), d_contribs);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_desolv_penalty(const mmo_mask *shell, const double *d_contribs, const mmo_ligand *lig, const double *d_radii,
                          const PoseSrc &src, int64_t n_poses, double *d_prot, double *d_lig) {
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(lig->n <= kDesolvMaxAtoms, "mmo_desolv_penalty: ligand of %d atoms (limit %d)", lig->n, kDesolvMaxAtoms);
    MMO_REQUIRE(n_poses < ((int64_t)1 << 31), "mmo_desolv_penalty: too many poses in one call");
    double q[3];
    node_steps(shell, q);
    const double step = shell->step;
    desolv_penalty_kernel<<<(unsigned)n_poses, kDesolvTPB, (size_t)lig->n * 5 * sizeof(double), rt().stream>>>(
        shell->words.p, d_contribs, shell->dims[0], shell->dims[1], shell->dims[2], step, q[0], q[1], q[2], lig->n, lig->x.p,
        lig->y.p, lig->z.p, lig->q.p, d_radii, src, n_poses, step * step * step, desolv_constant(), d_prot, d_lig);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

}  // namespace mmo
