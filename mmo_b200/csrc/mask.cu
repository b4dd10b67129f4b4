// mask.cu -- K5: protein vdW occupancy bitmask and the clash prefilter.  Compiled with -fmad=false:
// the comparisons are done on the same IEEE doubles as the reference, so the bits are identical.
//   Lds.atom_bitmask_set / vdW_volume    src/lds.ml:148-196
//   Lds.first_solvent_shell              src/lds.ml:172-184 (set r + r_H2O, then unset r)
//   Lds.bitmask_whole_protein            src/lds.ml:97-145  (vdW_volume with r = 12 A for every atom)
//   Lds.bitmask_ROI_only                 src/lds.ml:269-305
//   Grid.coord_of_point                  src/grid.ml:87-91
//   G3D.vdW_clash_OR / vdW_clash_AND     src/G3D.ml:162-213
//   Mol.protein_ligand_clash             src/mol.ml:1195-1203
#include "common.cuh"
#include "pose.cuh"
#include <math.h>

namespace mmo {

// one block per atom; threads sweep the cube of +-ceil(r/step) voxels around the atom's voxel.
// The reference indexes grid.xs unchecked-by-construction (36 A margin); indices are clipped here.
__global__ void vdw_mask_kernel(int n, const double *__restrict__ px, const double *__restrict__ py,
                                const double *__restrict__ pz, const double *__restrict__ pr,
                                double step, double q0, double q1, double q2,
                                int dim0, int dim1, int dim2, uint32_t *__restrict__ words, int set_bits) {
    const int a = blockIdx.x;
    if (a >= n) return;
    const double x = px[a], y = py[a], z = pz[a], radius = pr[a];
    const int ci = (int)((x - 0.0) / step), cj = (int)((y - 0.0) / step), ck = (int)((z - 0.0) / step);
    const int rs = (int)ceil(radius / step);
    const double r2 = radius * radius;
    const int side = 2 * rs + 1;
    const int total = side * side * side;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        int kk = t % side, jj = (t / side) % side, ii = t / (side * side);
        int i = ci - rs + ii, j = cj - rs + jj, k = ck - rs + kk;
        if (i < 0 || j < 0 || k < 0 || i >= dim0 || j >= dim1 || k >= dim2) continue;
        double gx = (double)i * q0, gy = (double)j * q1, gz = (double)k * q2;
        double dx = x - gx, dy = y - gy, dz = z - gz;      // V3.dist2 xyz (make x y z)
        if (dx * dx + dy * dy + dz * dz < r2) {
            size_t idx = (size_t)i + (size_t)j * dim0 + (size_t)k * dim0 * dim1;
            if (set_bits) atomicOr(words + (idx >> 5), 1u << (idx & 31));
            else atomicAnd(words + (idx >> 5), ~(1u << (idx & 31)));      // atom_bitmask_set ... false
        }
    }
}

// Lds.bitmask_ROI_only (src/lds.ml:269-305): grid points closer than r to c; thread = one 32-voxel word
__global__ void __launch_bounds__(256)
sphere_mask_kernel(double cx, double cy, double cz, double r2, double q0, double q1, double q2,
                   int dim0, int dim1, int dim2, size_t nbits, uint32_t *__restrict__ words) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= nbits) return;
    uint32_t bits = 0u;
    const size_t xy = (size_t)dim0 * dim1;
    for (int b = 0; b < 32; b++) {
        const size_t idx = w * 32 + b;
        if (idx >= nbits) break;
        const int k = (int)(idx / xy), j = (int)((idx - (size_t)k * xy) / dim0), i = (int)(idx - (size_t)k * xy - (size_t)j * dim0);
        const double x = (double)i * q0, y = (double)j * q1, z = (double)k * q2;
        const double dx = cx - x, dy = cy - y, dz = cz - z;            // V3.dist2 c (create x y z)
        if (dx * dx + dy * dy + dz * dz < r2) bits |= 1u << b;
    }
    words[w] = bits;
}

// Bitv.get raises outside the vector (the reference would abort the run); here a voxel outside the mask box reads
// as "not occupied", the same convention d_trilin uses for out-of-grid corners (0.0), in the kernels, on the host
// (scan.cu host_clash_and) and in the oracle alike
__device__ __forceinline__ bool bit_ijk(const uint32_t *__restrict__ w, int dim0, int dim1, int dim2, int i, int j, int k) {
    if ((unsigned)i >= (unsigned)dim0 || (unsigned)j >= (unsigned)dim1 || (unsigned)k >= (unsigned)dim2) return false;
    const size_t idx = (size_t)i + (size_t)j * dim0 + (size_t)k * dim0 * dim1;
    return (__ldg(w + (idx >> 5)) >> (idx & 31)) & 1u;
}

// G3D.vdW_clash_OR (G3D.ml:162-186)
__device__ __forceinline__ bool clash_or(const uint32_t *__restrict__ w, double inv, int dim0, int dim1, int dim2,
                                         double x, double y, double z) {
    const int i0 = (int)(x * inv), j0 = (int)(y * inv), k0 = (int)(z * inv);
    const int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    return bit_ijk(w, dim0, dim1, dim2, i0, j0, k0) || bit_ijk(w, dim0, dim1, dim2, i1, j0, k0) ||
           bit_ijk(w, dim0, dim1, dim2, i1, j1, k0) || bit_ijk(w, dim0, dim1, dim2, i0, j1, k0) ||
           bit_ijk(w, dim0, dim1, dim2, i0, j0, k1) || bit_ijk(w, dim0, dim1, dim2, i1, j0, k1) ||
           bit_ijk(w, dim0, dim1, dim2, i1, j1, k1) || bit_ijk(w, dim0, dim1, dim2, i0, j1, k1);
}

// Mol.protein_ligand_clash for a batch of poses
__global__ void __launch_bounds__(128)
clash_kernel(const uint32_t *__restrict__ words, double inv, int dim0, int dim1, int dim2, int L,
             const double *__restrict__ lx, const double *__restrict__ ly, const double *__restrict__ lz,
             PoseSrc src, int64_t n_poses, uint8_t *__restrict__ flags) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_poses) return;
    bool clash = false;
    if (src.kind == 1) {
        for (int j = 0; j < L && !clash; j++)
            clash = clash_or(words, inv, dim0, dim1, dim2, src.xs[p * L + j], src.ys[p * L + j], src.zs[p * L + j]);
    } else {
        PoseRT P;
        load_pose_rt(src, p, P);
        for (int j = 0; j < L && !clash; j++) {
            double x, y, z;
            pose_atom_rt(P, __ldg(lx + j), __ldg(ly + j), __ldg(lz + j), x, y, z);
            clash = clash_or(words, inv, dim0, dim1, dim2, x, y, z);
        }
    }
    flags[p] = clash ? 1 : 0;
}

// scan prefilter: candidate c of the slab = (active point c / n_rot, rotation c % n_rot);
// survivors are appended (block-ordered) to `frames`
__global__ void __launch_bounds__(256)
scan_prefilter_kernel(const uint32_t *__restrict__ words, double inv, int dim0, int dim1, int dim2, int L,
                      const double *__restrict__ lx, const double *__restrict__ ly, const double *__restrict__ lz,
                      PoseSrc src /* kind 2, frames unused */, const int64_t *__restrict__ points,
                      const int32_t *__restrict__ rot_perm, int64_t n_cand,
                      int64_t *__restrict__ frames, unsigned long long *__restrict__ counter) {
    __shared__ unsigned long long s_base;
    __shared__ int s_warp[8];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    int64_t frame = 0;
    if (c < n_cand) {
        int64_t pt = __ldg(points + c / src.n_rot);
        // rotations are visited in a spatially coherent order (neighbouring lanes = similar rotations);
        // the frame id keeps the reference's numbering, so results do not depend on the visiting order
        int rot_i = __ldg(rot_perm + (int)(c % src.n_rot));
        frame = (int64_t)rot_i + (int64_t)src.n_rot * pt;
        keep = true;
        if (words) {
            PoseRT P;
            load_pose_rt_frame(src, frame, P);
            for (int j = 0; j < L && keep; j++) {
                double x, y, z;
                pose_atom_rt(P, __ldg(lx + j), __ldg(ly + j), __ldg(lz + j), x, y, z);
                keep = !clash_or(words, inv, dim0, dim1, dim2, x, y, z);
            }
        }
    }
    // ordered compaction inside the block, one atomic per block
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { int v = s_warp[w]; s_warp[w] = tot; tot += v; }
        s_base = atomicAdd(counter, (unsigned long long)tot);
    }
    __syncthreads();
    if (keep) frames[s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1))] = frame;
}

// scissors (src/scissors.ml:52-62): keep the protein atoms whose nearest ligand atom is not farther than the cut-off
// (BST.nearest_neighbor -> dist = sqrt(dist2), kept if dist <= cutoff); thread = protein atom, ligand in shared memory
__global__ void __launch_bounds__(256)
carve_kernel(int n_rec, const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
             int n_lig, const double *__restrict__ lx, const double *__restrict__ ly, const double *__restrict__ lz,
             double cutoff, uint8_t *__restrict__ keep) {
    __shared__ double s_x[256], s_y[256], s_z[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double x = i < n_rec ? px[i] : 0.0, y = i < n_rec ? py[i] : 0.0, z = i < n_rec ? pz[i] : 0.0;
    double best = INFINITY;
    for (int j0 = 0; j0 < n_lig; j0 += 256) {
        __syncthreads();
        if (j0 + threadIdx.x < n_lig) { s_x[threadIdx.x] = lx[j0 + threadIdx.x]; s_y[threadIdx.x] = ly[j0 + threadIdx.x]; s_z[threadIdx.x] = lz[j0 + threadIdx.x]; }
        __syncthreads();
        const int m = min(256, n_lig - j0);
        for (int j = 0; j < m; j++) {
            const double dx = x - s_x[j], dy = y - s_y[j], dz = z - s_z[j];
            best = fmin(best, sqrt(dx * dx + dy * dy + dz * dz));
        }
    }
    if (i < n_rec) keep[i] = best <= cutoff ? 1 : 0;
}

int launch_carve(int n_rec, const double *d_px, const double *d_py, const double *d_pz, int n_lig, const double *d_lx,
                 const double *d_ly, const double *d_lz, double cutoff, uint8_t *d_keep) {
    if (n_rec == 0) return MMO_OK;
    carve_kernel<<<(unsigned)((n_rec + 255) / 256), 256, 0, rt().stream>>>(n_rec, d_px, d_py, d_pz, n_lig, d_lx, d_ly, d_lz, cutoff, d_keep);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_sphere_mask(double cx, double cy, double cz, double r, const mmo_mask *m) {
    double q[3];
    for (int d = 0; d < 3; d++) {
        int np = m->dims[d] - 1;
        q[d] = np > 0 ? (m->step * (double)np) / (double)np : 0.0;
    }
    const size_t nwords = (m->nbits + 31) / 32;
    KernelScope ks(K_VDW_MASK);
    sphere_mask_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, rt().stream>>>(cx, cy, cz, r * r, q[0], q[1], q[2], m->dims[0],
                                                                                 m->dims[1], m->dims[2], m->nbits, m->words.p);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_vdw_mask(int n, const double *d_x, const double *d_y, const double *d_z, const double *d_r,
                    const mmo_mask *m, bool set_bits) {
    if (n == 0) return MMO_OK;
    double q[3];
    for (int d = 0; d < 3; d++) {
        int np = m->dims[d] - 1;
        q[d] = np > 0 ? (m->step * (double)np) / (double)np : 0.0;
    }
    KernelScope ks(K_VDW_MASK);
    vdw_mask_kernel<<<n, 128, 0, rt().stream>>>(n, d_x, d_y, d_z, d_r, m->step, q[0], q[1], q[2],
                                                m->dims[0], m->dims[1], m->dims[2], m->words.p, set_bits ? 1 : 0);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_clash(const mmo_mask *m, const mmo_ligand *lig, const PoseSrc &src, int64_t n_poses, uint8_t *d_flags) {
    if (n_poses == 0) return MMO_OK;
    clash_kernel<<<(unsigned)((n_poses + 127) / 128), 128, 0, rt().stream>>>(
        m->words.p, 1.0 / m->step, m->dims[0], m->dims[1], m->dims[2], lig->n, lig->x.p, lig->y.p, lig->z.p,
        src, n_poses, d_flags);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

int launch_scan_prefilter(const mmo_mask *m, const mmo_ligand *lig, const PoseSrc &src, const int64_t *d_points,
                          const int32_t *d_rot_perm, int64_t n_cand, int64_t *d_frames, unsigned long long *d_counter) {
    if (n_cand == 0) return MMO_OK;
    KernelScope ks(K_PREFILTER);
    scan_prefilter_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, rt().stream>>>(
        m ? m->words.p : nullptr, m ? 1.0 / m->step : 0.0, m ? m->dims[0] : 0, m ? m->dims[1] : 0, m ? m->dims[2] : 0,
        lig->n, lig->x.p, lig->y.p, lig->z.p, src, d_points, d_rot_perm, n_cand, d_frames, d_counter);
    MMO_LAUNCH_CHECK();
    return MMO_OK;
}

}  // namespace mmo
