/* mmo_oracle.c -- CPU restatement of the UnixJunkie/MMO scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see mmo_oracle.h).  PARITY UNPINNED: no OCaml
 * toolchain here and no golden vectors in the reference.
 *
 * Every function follows the reference's evaluation order literally; the
 * reference file:line it restates is cited above it (paths relative to
 * /root/reference).  Build: gcc -O2 -ffp-contract=off (see oracle/Makefile).
 */
#include "mmo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* src/FF.ml:5-20 */
static inline double sq(double x) { return x * x; }
static inline double pow3(double x) { return x * x * x; }       /* (x*x)*x */
double orc_pow6(double x) { return pow3(x * x); }
double orc_geo_mean(double x, double y) { return sqrt(x * y); }
double orc_shift_12A(double d) {
    if (d < 12.0) return sq(1.0 - sq(d / 12.0));
    return 0.0;
}
/* src/math.ml:58-62 */
double orc_non_zero_dist(double x) { return (x < 0.01) ? 0.01 : x; }
/* src/UFF.ml:25 with src/const.ml:18 */
double orc_elec_weight(void) { return 332.0637 / 4.0; }
/* src/lds.ml:66-67 with src/const.ml:24 */
double orc_beta(double temp_K) { return 1.0 / (0.0019872041 * temp_K); }

/* src/UFF.ml:10-51 : dense 119x119 table of {x_ij, d_ij}; NaN when unsupported */
static double g_xij[ORC_ANUMS * ORC_ANUMS];
static double g_dij[ORC_ANUMS * ORC_ANUMS];
static int g_uff_ready = 0;
static void uff_init(void) {
    static const int anums[12] = {0, 1, 6, 7, 8, 9, 12, 15, 16, 17, 35, 53};
    static const double xi[12] = {0.0, 2.886, 3.851, 3.660, 3.500, 3.364, 3.021, 4.147, 4.035, 3.947, 4.189, 4.500};
    static const double di[12] = {0.0, 0.044, 0.105, 0.069, 0.060, 0.050, 0.111, 0.305, 0.274, 0.227, 0.251, 0.339};
    if (g_uff_ready) return;
    for (int i = 0; i < ORC_ANUMS * ORC_ANUMS; i++) { g_xij[i] = NAN; g_dij[i] = NAN; }
    for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) {
            int i = anums[a] * ORC_ANUMS + anums[b];
            g_xij[i] = orc_geo_mean(xi[a], xi[b]);
            g_dij[i] = orc_geo_mean(di[a], di[b]);
        }
    g_uff_ready = 1;
}
void orc_vdw_xidi(int a1, int a2, double out[2]) {
    uff_init();
    out[0] = g_xij[a1 * ORC_ANUMS + a2];
    out[1] = g_dij[a1 * ORC_ANUMS + a2];
}
/* src/ptable.ml:41-54 */
double orc_vdw_radius(int anum) {
    switch (anum) {
    case 1: return 1.2; case 6: return 1.7; case 7: return 1.6; case 8: return 1.55;
    case 9: return 1.5; case 12: return 2.2; case 15: return 1.95; case 16: return 1.8;
    case 17: return 1.8; case 35: return 1.9; case 53: return 2.1;
    default: return NAN;
    }
}

/* src/V3.ml:23-28 : u = first argument, v = second */
static inline double dist2(double ux, double uy, double uz, double vx, double vy, double vz) {
    double dx = ux - vx, dy = uy - vy, dz = uz - vz;
    return dx * dx + dy * dy + dz * dz;
}

/* ------------------------------------------------------------------------ */
/* src/mol.ml:796-818 ene_inter_UFF_global_brute */
double orc_ene_inter_global_brute(int P, const double *px, const double *py, const double *pz,
                                  const double *pq, const int32_t *panum,
                                  int L, const double *lx, const double *ly, const double *lz,
                                  const double *lq, const int32_t *lanum) {
    uff_init();
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int i = 0; i < P; i++) {
        double q_i = pq[i];
        int pa = panum[i];
        for (int j = 0; j < L; j++) {
            double q_j = lq[j];
            double r_ij = orc_non_zero_dist(sqrt(dist2(px[i], py[i], pz[i], lx[j], ly[j], lz[j])));
            int t = pa * ORC_ANUMS + lanum[j];
            double p6 = orc_pow6(g_xij[t] / r_ij);
            sum_elec = sum_elec + ((q_i * q_j) / r_ij);
            sum_vdW = sum_vdW + (g_dij[t] * ((-2.0 * p6) + (p6 * p6)));
        }
    }
    return (orc_elec_weight() * sum_elec) + sum_vdW;
}

/* src/mol.ml:822-849 ene_inter_UFF_shifted_brute */
double orc_ene_inter_shifted_brute(int P, const double *px, const double *py, const double *pz,
                                   const double *pq, const int32_t *panum,
                                   int L, const double *lx, const double *ly, const double *lz,
                                   const double *lq, const int32_t *lanum) {
    uff_init();
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int i = 0; i < P; i++) {
        double q_i = pq[i];
        int pa = panum[i];
        for (int j = 0; j < L; j++) {
            double r_ij2 = dist2(px[i], py[i], pz[i], lx[j], ly[j], lz[j]);
            if (r_ij2 < 144.0) {
                double q_j = lq[j];
                double r_ij = orc_non_zero_dist(sqrt(r_ij2));
                double w = orc_shift_12A(r_ij);
                int t = pa * ORC_ANUMS + lanum[j];
                double p6 = orc_pow6(g_xij[t] / r_ij);
                sum_elec = sum_elec + w * ((q_i * q_j) / r_ij);
                sum_vdW = sum_vdW + w * (g_dij[t] * ((-2.0 * p6) + (p6 * p6)));
            }
        }
    }
    return (orc_elec_weight() * sum_elec) + sum_vdW;
}

/* src/mol.ml:928-956 ene_inter_UFF_shifted_bst_components with BST.neighbors replaced by a
 * linear filter "dist <= 12" in receptor index order (the library's order is unpinned, F6);
 * w(12.0) = 0 so the boundary convention cannot change the value. */
void orc_ene_inter_shifted_components(int P, const double *px, const double *py, const double *pz,
                                      const double *pq, const int32_t *panum,
                                      int L, const double *lx, const double *ly, const double *lz,
                                      const double *lq, const int32_t *lanum, double out[2]) {
    uff_init();
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int j = 0; j < L; j++) {
        double q_j = lq[j];
        int la = lanum[j];
        for (int i = 0; i < P; i++) {
            double d = sqrt(dist2(px[i], py[i], pz[i], lx[j], ly[j], lz[j]));
            if (d <= 12.0) {
                double r_ij = orc_non_zero_dist(d);
                double w = orc_shift_12A(r_ij);
                int t = la * ORC_ANUMS + panum[i];
                double p6 = orc_pow6(g_xij[t] / r_ij);
                sum_elec = sum_elec + w * ((pq[i] * q_j) / r_ij);
                sum_vdW = sum_vdW + w * (g_dij[t] * ((-2.0 * p6) + (p6 * p6)));
            }
        }
    }
    out[0] = sum_elec * orc_elec_weight();
    out[1] = sum_vdW;
}

/* src/mol.ml:881-903 ene_intra_UFFNB_brute; interacting = dists >= 3 (mol.ml:203-208) */
double orc_ene_intra_uffnb_brute(int L, const double *lx, const double *ly, const double *lz,
                                 const double *lq, const int32_t *lanum, const int32_t *dists) {
    uff_init();
    double sum_elec = 0.0, sum_vdW = 0.0;
    for (int i = 0; i < L - 1; i++) {
        double q_i = lq[i];
        int a_i = lanum[i];
        for (int j = i + 1; j < L; j++) {
            if (dists[i + j * L] >= 3) {
                double r_ij = orc_non_zero_dist(sqrt(dist2(lx[i], ly[i], lz[i], lx[j], ly[j], lz[j])));
                int t = a_i * ORC_ANUMS + lanum[j];
                double p6 = orc_pow6(g_xij[t] / r_ij);
                sum_elec = sum_elec + (q_i * lq[j]) / r_ij;
                sum_vdW = sum_vdW + g_dij[t] * ((-2.0 * p6) + (p6 * p6));
            }
        }
    }
    return (orc_elec_weight() * sum_elec) + sum_vdW;
}

/* src/mol.ml:964-989 ene_inter_UFF_shifted_grid; neighbours in receptor index order (F6) */
void orc_ene_inter_shifted_grid(int P, const double *px, const double *py, const double *pz,
                                const double *pq, const int32_t *panum,
                                double x, double y, double z,
                                int T, const int32_t *tanum, const double *tq, double *out) {
    uff_init();
    double se[T > 0 ? T : 1], sv[T > 0 ? T : 1];
    for (int l = 0; l < T; l++) { se[l] = 0.0; sv[l] = 0.0; }
    for (int i = 0; i < P; i++) {
        double d = sqrt(dist2(px[i], py[i], pz[i], x, y, z));
        if (d <= 12.0) {
            double q_i = pq[i];
            double r_ij = orc_non_zero_dist(d);
            double w = orc_shift_12A(r_ij);
            for (int l = 0; l < T; l++) {
                int t = tanum[l] * ORC_ANUMS + panum[i];
                double p6 = orc_pow6(g_xij[t] / r_ij);
                se[l] = se[l] + w * ((q_i * tq[l]) / r_ij);
                sv[l] = sv[l] + w * (g_dij[t] * ((-2.0 * p6) + (p6 * p6)));
            }
        }
    }
    for (int l = 0; l < T; l++) out[l] = orc_elec_weight() * se[l] + sv[l];
}

/* ------------------------------------------------------------------------ */
/* src/grid.ml:37-38 */
int orc_grid_num_steps(double dx, double length) { return (int)ceil(length / dx); }
/* src/grid.ml:40-52 */
void orc_grid_from_box(double step, double bx, double by, double bz, int dims[3]) {
    dims[0] = orc_grid_num_steps(step, bx) + 1;
    dims[1] = orc_grid_num_steps(step, by) + 1;
    dims[2] = orc_grid_num_steps(step, bz) + 1;
}
/* src/grid.ml:49-51 : L.frange 0.0 `To (step * n') (n'+1)  -- Batteries' frange is not vendored
 * (unpinned); restated as start + i * (span / (n-1)), exact for dyadic steps (0.5, 0.375, 1, 2). */
double orc_grid_node(double step, int dim, int i) {
    int np = dim - 1;
    if (np <= 0) return 0.0;
    double span = step * (double)np;
    return 0.0 + (double)i * (span / (double)np);
}

/* src/G3D.ml:97-157 trilin */
double orc_trilin(double step, const int dims[3], const float *arr, double px, double py, double pz) {
    double inv = 1.0 / step;                /* grid.ml:41 */
    int x_dim = dims[0], xy_dim = dims[0] * dims[1];
    int i0 = (int)(px * inv), j0 = (int)(py * inv), k0 = (int)(pz * inv);
    int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    /* The reference reads the Bigarray unchecked (G3D.ml:144-151): outside the grid is undefined
     * behaviour there, prevented by the 36 A margin (lds.ml:1832).  Here, and identically in the CUDA
     * kernels, a voxel outside the grid contributes 0.0 (the value of an unmasked voxel). */
    if (i0 < 0 || j0 < 0 || k0 < 0 || i1 >= dims[0] || j1 >= dims[1] || k1 >= dims[2]) return 0.0;
    int j0x = j0 * x_dim, j1x = j1 * x_dim, k0xy = k0 * xy_dim, k1xy = k1 * xy_dim;
    double lx = orc_grid_node(step, dims[0], i0);
    double ly = orc_grid_node(step, dims[1], j0);
    double lz = orc_grid_node(step, dims[2], k0);
    double wlx = (px - lx) * inv, wly = (py - ly) * inv, wlz = (pz - lz) * inv;
    double whx = 1.0 - wlx, why = 1.0 - wly, whz = 1.0 - wlz;
    return ((double)arr[i0 + j0x + k0xy] * (whx * why * whz) +
            (double)arr[i1 + j0x + k0xy] * (wlx * why * whz) +
            (double)arr[i1 + j1x + k0xy] * (wlx * wly * whz) +
            (double)arr[i0 + j1x + k0xy] * (whx * wly * whz) +
            (double)arr[i0 + j0x + k1xy] * (whx * why * wlz) +
            (double)arr[i1 + j0x + k1xy] * (wlx * why * wlz) +
            (double)arr[i1 + j1x + k1xy] * (wlx * wly * wlz) +
            (double)arr[i0 + j1x + k1xy] * (whx * wly * wlz));
}

/* src/mol.ml:1012-1020 ene_inter_UFF_interp */
double orc_ene_inter_interp(double step, const int dims[3], const float *maps,
                            int L, const double *lx, const double *ly, const double *lz,
                            const int32_t *ltyp) {
    size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
    double res = 0.0;
    for (int j = 0; j < L; j++)
        res = res + orc_trilin(step, dims, maps + (size_t)ltyp[j] * nvox, lx[j], ly[j], lz[j]);
    return res;
}

static inline int mask_get(const uint8_t *m, size_t idx) { return (m[idx >> 3] >> (idx & 7)) & 1; }
static inline void mask_put(uint8_t *m, size_t idx, int b) {
    if (b) m[idx >> 3] |= (uint8_t)(1u << (idx & 7));
    else m[idx >> 3] &= (uint8_t)~(1u << (idx & 7));
}

/* src/lds.ml:452-469 pre_calculate_FF_components_grid (+ G3D.ml:47-51,76-80, grid.ml:101-105).
 * Voxels are independent, so the OpenMP split cannot change any value. */
void orc_grid_build(int P, const double *px, const double *py, const double *pz,
                    const double *pq, const int32_t *panum,
                    double step, const int dims[3], const uint8_t *mask,
                    int T, const int32_t *tanum, const double *tq, float *maps) {
    uff_init();
    size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
    int xy = dims[0] * dims[1];
#pragma omp parallel for schedule(dynamic, 256)
    for (long idx = 0; idx < (long)nvox; idx++) {
        if (mask && !mask_get(mask, (size_t)idx)) continue;
        int k = (int)(idx / xy);
        int j = (int)((idx - (long)k * xy) / dims[0]);
        int i = (int)(idx - ((long)k * xy + (long)j * dims[0]));
        double e[T > 0 ? T : 1];
        orc_ene_inter_shifted_grid(P, px, py, pz, pq, panum,
                                   orc_grid_node(step, dims[0], i), orc_grid_node(step, dims[1], j),
                                   orc_grid_node(step, dims[2], k), T, tanum, tq, e);
        for (int l = 0; l < T; l++) {
            double v = (ORC_MAX_E < e[l]) ? ORC_MAX_E : e[l];  /* OCaml min: NaN-propagating compare */
            if (e[l] != e[l]) v = e[l];
            maps[(size_t)l * nvox + (size_t)idx] = (float)v;
        }
    }
}

/* src/lds.ml:269-305 bitmask_ROI_only (caller passes r = R_roi + 24) */
void orc_bitmask_sphere(double step, const int dims[3], double cx, double cy, double cz,
                        double r, uint8_t *mask) {
    double r2 = r * r;
    for (int i = 0; i < dims[0]; i++) {
        double x = orc_grid_node(step, dims[0], i);
        for (int j = 0; j < dims[1]; j++) {
            double y = orc_grid_node(step, dims[1], j);
            for (int k = 0; k < dims[2]; k++) {
                double z = orc_grid_node(step, dims[2], k);
                if (dist2(cx, cy, cz, x, y, z) < r2)
                    mask_put(mask, (size_t)i + (size_t)j * dims[0] + (size_t)k * dims[0] * dims[1], 1);
            }
        }
    }
}

/* src/lds.ml:148-173 atom_bitmask_set (b = true) and 187-196 vdW_volume.
 * The reference indexes grid.xs without clipping (an out-of-range index raises); the 36 A margin
 * makes that unreachable.  Here indices are clipped to the grid. */
void orc_vdw_volume(int P, const double *px, const double *py, const double *pz, const double *pr,
                    double step, const int dims[3], uint8_t *mask) {
    for (int a = 0; a < P; a++) {
        double radius = pr[a];
        int i = (int)((px[a] - 0.0) / step);        /* grid.ml:87-91 coord_of_point */
        int j = (int)((py[a] - 0.0) / step);
        int k = (int)((pz[a] - 0.0) / step);
        int r_steps = (int)ceil(radius / step);
        double r2 = radius * radius;
        for (int ii = i - r_steps; ii <= i + r_steps; ii++) {
            if (ii < 0 || ii >= dims[0]) continue;
            double x = orc_grid_node(step, dims[0], ii);
            for (int jj = j - r_steps; jj <= j + r_steps; jj++) {
                if (jj < 0 || jj >= dims[1]) continue;
                double y = orc_grid_node(step, dims[1], jj);
                for (int kk = k - r_steps; kk <= k + r_steps; kk++) {
                    if (kk < 0 || kk >= dims[2]) continue;
                    double z = orc_grid_node(step, dims[2], kk);
                    if (dist2(px[a], py[a], pz[a], x, y, z) < r2)
                        mask_put(mask, (size_t)ii + (size_t)jj * dims[0] + (size_t)kk * dims[0] * dims[1], 1);
                }
            }
        }
    }
}

/* src/lds.ml:148-170 atom_bitmask_set with an arbitrary boolean */
static void atom_bitmask_put(double ax, double ay, double az, double radius, double step, const int dims[3],
                             uint8_t *mask, int b) {
    int i = (int)((ax - 0.0) / step), j = (int)((ay - 0.0) / step), k = (int)((az - 0.0) / step);
    int r_steps = (int)ceil(radius / step);
    double r2 = radius * radius;
    for (int ii = i - r_steps; ii <= i + r_steps; ii++) {
        if (ii < 0 || ii >= dims[0]) continue;
        double x = orc_grid_node(step, dims[0], ii);
        for (int jj = j - r_steps; jj <= j + r_steps; jj++) {
            if (jj < 0 || jj >= dims[1]) continue;
            double y = orc_grid_node(step, dims[1], jj);
            for (int kk = k - r_steps; kk <= k + r_steps; kk++) {
                if (kk < 0 || kk >= dims[2]) continue;
                double z = orc_grid_node(step, dims[2], kk);
                if (dist2(ax, ay, az, x, y, z) < r2)
                    mask_put(mask, (size_t)ii + (size_t)jj * dims[0] + (size_t)kk * dims[0] * dims[1], b);
            }
        }
    }
}

/* src/lds.ml:172-184 first_solvent_shell: SET inside r_vdW + r_H2O for every atom, then UNSET inside r_vdW
 * (src/const.ml:10 r_H2O = 1.4) */
void orc_first_solvent_shell(int P, const double *px, const double *py, const double *pz, const double *pr,
                             double step, const int dims[3], uint8_t *mask) {
    for (int a = 0; a < P; a++) atom_bitmask_put(px[a], py[a], pz[a], pr[a] + 1.4, step, dims, mask, 1);
    for (int a = 0; a < P; a++) atom_bitmask_put(px[a], py[a], pz[a], pr[a], step, dims, mask, 0);
}

/* ---- N4: desolvation sums (Majeux, Scarsi and Caflisch PROTEINS 2001, eq. 2) ------------------------------- */
/* src/const.ml:31: (1/eps_prot - 1/eps_HOH) / (8 pi), eps_prot = 4.0, eps_HOH = 78.5 (const.ml:17,20), pi = 4 atan 1 */
static double desolvation_constant(void) {
    double pi = 4.0 * atan(1.0);
    return (1.0 / 4.0 - 1.0 / 78.5) / (8.0 * pi);
}

/* src/lds.ml:204-236 protein_desolv: for every set bit of the protein's first solvent shell whose voxel lies inside
 * the ROI (ROI.is_inside, strict <, src/ROI.ml:65-66): res = Const.desolvation * (voxel_vol * sum_j (q_j/d2)^2) over
 * the protein atoms BST.neighbors returns for Const.charged_cutoff = 12 A.  The bst library is not vendored: its
 * neighbour order is unpinned (restated: atom index order) and its radius test is taken as Atom.dist <= tol.
 * res has one double per voxel (A.make n 0.0). */
void orc_protein_desolv(int P, const double *px, const double *py, const double *pz, const double *pq,
                        double step, const int dims[3], const uint8_t *shell, const double roi[4], double *res) {
    double voxel_vol = step * step * step;                         /* src/grid.ml:29-30 */
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    long xy = (long)dims[0] * dims[1];
    for (size_t idx = 0; idx < n; idx++) res[idx] = 0.0;
    for (size_t idx = 0; idx < n; idx++) {                         /* Bitv.iteri_true */
        if (!mask_get(shell, idx)) continue;
        int k = (int)((long)idx / xy);                             /* Grid.ijk_of_idx, src/grid.ml:101-105 */
        int j = (int)(((long)idx - (long)k * xy) / dims[0]);
        int i = (int)((long)idx - ((long)k * xy + (long)j * dims[0]));
        double x = orc_grid_node(step, dims[0], i), y = orc_grid_node(step, dims[1], j), z = orc_grid_node(step, dims[2], k);
        if (!(dist2(roi[0], roi[1], roi[2], x, y, z) < roi[3] * roi[3])) continue;
        for (int a = 0; a < P; a++) {
            if (!(sqrt(dist2(x, y, z, px[a], py[a], pz[a])) <= 12.0)) continue;
            double d2 = dist2(x, y, z, px[a], py[a], pz[a]);       /* V3.dist2 x_p x_j */
            double v = pq[a] / d2;
            res[idx] = res[idx] + (v * v);
        }
        res[idx] = desolvation_constant() * (voxel_vol * res[idx]);
    }
}

/* src/lds.ml:239-267 desolvation_penalty: desolvated = prot_shell AND first_solvent_shell grid lig; for each such voxel
 * in index order: the ligand atoms with d2 < 144 add (q_j/d2)^2, the voxel's protein contribution is added to prot;
 * lig = Const.desolvation * (voxel_vol * sum) */
void orc_desolvation_penalty(double step, const int dims[3], const uint8_t *prot_shell, const double *contribs,
                             int L, const double *lx, const double *ly, const double *lz, const double *lq,
                             const double *lr, double *out_prot, double *out_lig) {
    double voxel_vol = step * step * step;
    size_t n = (size_t)dims[0] * dims[1] * dims[2];
    long xy = (long)dims[0] * dims[1];
    uint8_t *lig_shell = (uint8_t *)calloc((n + 7) / 8, 1);
    orc_first_solvent_shell(L, lx, ly, lz, lr, step, dims, lig_shell);
    double lig_desolv = 0.0, prot_desolv = 0.0;
    for (size_t idx = 0; idx < n; idx++) {
        if (!(mask_get(prot_shell, idx) && mask_get(lig_shell, idx))) continue;     /* Bitv.bw_and */
        int k = (int)((long)idx / xy);
        int j = (int)(((long)idx - (long)k * xy) / dims[0]);
        int i = (int)((long)idx - ((long)k * xy + (long)j * dims[0]));
        double x = orc_grid_node(step, dims[0], i), y = orc_grid_node(step, dims[1], j), z = orc_grid_node(step, dims[2], k);
        for (int a = 0; a < L; a++) {
            double d2 = dist2(x, y, z, lx[a], ly[a], lz[a]);
            if (d2 < 12.0 * 12.0) {                                /* Const.charged_cutoff_squared: hard cut-off */
                double v = lq[a] / d2;
                lig_desolv = lig_desolv + (v * v);
            }
        }
        prot_desolv = prot_desolv + contribs[idx];
    }
    free(lig_shell);
    *out_prot = prot_desolv;
    *out_lig = desolvation_constant() * (voxel_vol * lig_desolv);
}

/* src/lds.ml:97-145 bitmask_whole_protein: grid points whose nearest protein atom is closer than
 * Const.charged_cutoff = 12 A (BST.nearest_neighbor -> V3.dist = sqrt(dist2)); brute force over the atoms */
void orc_bitmask_whole_protein(int P, const double *px, const double *py, const double *pz,
                               double step, const int dims[3], uint8_t *mask) {
    for (int i = 0; i < dims[0]; i++) {
        double x = orc_grid_node(step, dims[0], i);
        for (int j = 0; j < dims[1]; j++) {
            double y = orc_grid_node(step, dims[1], j);
            for (int k = 0; k < dims[2]; k++) {
                double z = orc_grid_node(step, dims[2], k);
                double best = INFINITY;
                for (int a = 0; a < P; a++) {
                    double d = sqrt(dist2(x, y, z, px[a], py[a], pz[a]));
                    if (d < best) best = d;
                }
                if (best < 12.0)
                    mask_put(mask, (size_t)i + (size_t)j * dims[0] + (size_t)k * dims[0] * dims[1], 1);
            }
        }
    }
}

/* src/G3D.ml:162-186 vdW_clash_OR ; 189-213 vdW_clash_AND */
/* Bitv.get raises outside the vector (the reference would abort); checker and library agree that a voxel outside
 * the mask box reads as "not occupied"  */
static int bit_ijk(const int dims[3], const uint8_t *mask, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= dims[0] || j >= dims[1] || k >= dims[2]) return 0;
    return mask_get(mask, (size_t)i + (size_t)j * dims[0] + (size_t)k * dims[0] * dims[1]);
}
static void corner_bits(double step, const int dims[3], const uint8_t *mask,
                        double x, double y, double z, int bits[8]) {
    double inv = 1.0 / step;
    int i0 = (int)(x * inv), j0 = (int)(y * inv), k0 = (int)(z * inv);
    int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    bits[0] = bit_ijk(dims, mask, i0, j0, k0);
    bits[1] = bit_ijk(dims, mask, i1, j0, k0);
    bits[2] = bit_ijk(dims, mask, i1, j1, k0);
    bits[3] = bit_ijk(dims, mask, i0, j1, k0);
    bits[4] = bit_ijk(dims, mask, i0, j0, k1);
    bits[5] = bit_ijk(dims, mask, i1, j0, k1);
    bits[6] = bit_ijk(dims, mask, i1, j1, k1);
    bits[7] = bit_ijk(dims, mask, i0, j1, k1);
}
int orc_vdw_clash_OR(double step, const int dims[3], const uint8_t *mask, double x, double y, double z) {
    int b[8];
    corner_bits(step, dims, mask, x, y, z, b);
    return b[0] || b[1] || b[2] || b[3] || b[4] || b[5] || b[6] || b[7];
}
int orc_vdw_clash_AND(double step, const int dims[3], const uint8_t *mask, double x, double y, double z) {
    int b[8];
    corner_bits(step, dims, mask, x, y, z, b);
    return b[0] && b[1] && b[2] && b[3] && b[4] && b[5] && b[6] && b[7];
}
/* src/mol.ml:1195-1203 protein_ligand_clash */
int orc_protein_ligand_clash(double step, const int dims[3], const uint8_t *mask,
                             int L, const double *lx, const double *ly, const double *lz) {
    for (int i = 0; i < L; i++)
        if (orc_vdw_clash_OR(step, dims, mask, lx[i], ly[i], lz[i])) return 1;
    return 0;
}
/* src/mol.ml:1209-1218 is_ligand_center_vdW_occuppied */
int orc_is_ligand_center_vdW_occupied(int L, const double *lx, const double *ly, const double *lz,
                                      const double *lr, double cx, double cy, double cz) {
    for (int i = 0; i < L; i++)
        if (sqrt(dist2(cx, cy, cz, lx[i], ly[i], lz[i])) < lr[i]) return 1;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* src/math.ml:13-15 */
static double orc_pi(void) { return 4.0 * atan(1.0); }

/* src/SO3.ml:13-30 super_fibonacci; Quat.create w x y z */
void orc_so3_quat(int n_i, int i, double q[4]) {
    double phi = sqrt(2.0);
    double psi = 1.533751168755204288118041;
    double n = (double)n_i;
    double s = (double)i + 0.5;
    double t = s / n;
    double d = (2.0 * orc_pi()) * s;
    double c_r = sqrt(t);
    double c_R = sqrt(1.0 - t);
    double alpha = d / phi;
    double beta = d / psi;
    q[0] = c_r * sin(alpha);
    q[1] = c_r * cos(alpha);
    q[2] = c_R * sin(beta);
    q[3] = c_R * cos(beta);
}
/* src/rot.ml:121-146 of_axis_angle (the live, un-commented body) */
void orc_rot_of_axis_angle(double x, double y, double z, double theta, double r[9]) {
    double c = cos(theta), s = sin(theta);
    double oneMct = 1.0 - c;
    r[0] = c + x * x * oneMct;
    r[1] = x * y * oneMct - z * s;
    r[2] = x * z * oneMct + y * s;
    r[3] = x * y * oneMct + z * s;
    r[4] = c + y * y * oneMct;
    r[5] = y * z * oneMct - x * s;
    r[6] = x * z * oneMct - y * s;
    r[7] = y * z * oneMct + x * s;
    r[8] = c + z * z * oneMct;
}
/* src/SO3.ml:35-39 rotations, via src/quat.ml:32-36 to_axis_angle */
void orc_so3_rotations(int n, double *rot9) {
    for (int i = 0; i < n; i++) {
        double q[4];
        orc_so3_quat(n, i, q);
        double w = q[0], x = q[1], y = q[2], z = q[3];
        double mag = sqrt(x * x + y * y + z * z);
        double theta = 2.0 * atan2(mag, w);
        orc_rot_of_axis_angle(x / mag, y / mag, z / mag, theta, rot9 + 9 * (size_t)i);
    }
}
/* src/rot.ml:22-46 (transposed-sign "direction cosine" convention) */
void orc_rot_rx(double th, double r[9]) {
    double c = cos(th), s = sin(th);
    r[0] = 1.0; r[1] = 0.0; r[2] = 0.0; r[3] = 0.0; r[4] = c; r[5] = s; r[6] = 0.0; r[7] = -s; r[8] = c;
}
void orc_rot_ry(double th, double r[9]) {
    double c = cos(th), s = sin(th);
    r[0] = c; r[1] = 0.0; r[2] = -s; r[3] = 0.0; r[4] = 1.0; r[5] = 0.0; r[6] = s; r[7] = 0.0; r[8] = c;
}
void orc_rot_rz(double th, double r[9]) {
    double c = cos(th), s = sin(th);
    r[0] = c; r[1] = s; r[2] = 0.0; r[3] = -s; r[4] = c; r[5] = 0.0; r[6] = 0.0; r[7] = 0.0; r[8] = 1.0;
}
/* src/rot.ml:52-66 r_xyz */
void orc_rot_r_xyz(double al, double be, double ga, double r[9]) {
    double ac = cos(al), as = sin(al), bc = cos(be), bs = sin(be), gc = cos(ga), gs = sin(ga);
    r[0] = bc * gc;
    r[1] = gc * as * bs - ac * gs;
    r[2] = as * gs + ac * gc * bs;
    r[3] = bc * gs;
    r[4] = ac * gc + as * bs * gs;
    r[5] = ac * bs * gs - gc * as;
    r[6] = -bs;
    r[7] = bc * as;
    r[8] = ac * bc;
}
/* src/rot.ml:71-75 decompose */
void orc_rot_decompose(const double r[9], double abg[3]) {
    double beta = atan2(-r[6], sqrt(r[0] * r[0] + r[3] * r[3]));
    double cb = cos(beta);
    abg[0] = atan2(r[7] / cb, r[8] / cb);
    abg[1] = beta;
    abg[2] = atan2(r[3] / cb, r[0] / cb);
}
/* src/rot.ml:77-94 mult */
void orc_rot_mult(const double a[9], const double b[9], double o[9]) {
    double t[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            t[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
    memcpy(o, t, sizeof t);
}
/* src/rot.ml:97-100 rotate */
void orc_rot_rotate(const double r[9], const double v[3], double o[3]) {
    double x = v[0], y = v[1], z = v[2];
    o[0] = r[0] * x + r[1] * y + r[2] * z;
    o[1] = r[3] * x + r[4] * y + r[5] * z;
    o[2] = r[6] * x + r[7] * y + r[8] * z;
}

/* Batteries A.favg = fsum / n with Kahan-compensated fsum (library not vendored: unpinned) */
double orc_favg(int n, const double *a) {
    double sum = 0.0, c = 0.0;
    for (int i = 0; i < n; i++) {
        double y = a[i] - c;
        double t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
    return sum / (double)n;
}

/* src/mol.ml:603-607 centered_rotate + 593-600 translate_by (rotate_then_translate_copy 669-672) */
void orc_rotate_then_translate(int L, const double *cx, const double *cy, const double *cz,
                               const double rot[9], const double t[3],
                               double *ox, double *oy, double *oz) {
    for (int i = 0; i < L; i++) {
        double v[3] = {cx[i], cy[i], cz[i]}, o[3];
        orc_rot_rotate(rot, v, o);
        ox[i] = o[0] + t[0];
        oy[i] = o[1] + t[1];
        oz[i] = o[2] + t[2];
    }
}
/* src/mol.ml:705-710 center_rotate_translate_copy: center = translate_by (neg mean) (698-702) */
void orc_center_rotate_translate(int L, const double *x, const double *y, const double *z,
                                 const double center[3], const double rot[9], const double t[3],
                                 double *ox, double *oy, double *oz) {
    double nx = -center[0], ny = -center[1], nz = -center[2];
    for (int i = 0; i < L; i++) {
        double v[3] = {x[i] + nx, y[i] + ny, z[i] + nz}, o[3];
        orc_rot_rotate(rot, v, o);
        ox[i] = o[0] + t[0];
        oy[i] = o[1] + t[1];
        oz[i] = o[2] + t[2];
    }
}
/* src/mol.ml:610-631 rotate_bond (Vector3.normalize = v / mag, library unpinned) */
void orc_rotate_bond(double *x, double *y, double *z, int left, int right,
                     int ngroup, const int32_t *group, double alpha) {
    double cx = x[right], cy = y[right], cz = z[right];
    double ax = cx - x[left], ay = cy - y[left], az = cz - z[left];
    double mag = sqrt(ax * ax + ay * ay + az * az);
    double rot[9];
    orc_rot_of_axis_angle(ax / mag, ay / mag, az / mag, alpha, rot);
    for (int g = 0; g < ngroup; g++) {
        int i = group[g];
        double v[3] = {x[i] - cx, y[i] - cy, z[i] - cz}, o[3];
        orc_rot_rotate(rot, v, o);
        x[i] = o[0] + cx;
        y[i] = o[1] + cy;
        z[i] = o[2] + cz;
    }
}
/* src/mol.ml:576-583 radius */
double orc_radius(int L, const double *x, const double *y, const double *z, const double c[3]) {
    double maxi = 0.0;
    for (int i = 0; i < L; i++) {
        double d = 0.01 + sqrt(dist2(c[0], c[1], c[2], x[i], y[i], z[i]));
        if (d > maxi) maxi = d;
    }
    return maxi;
}

/* src/optim.ml:64-80 apply_config: rotate bonds (libm), elongation check, r_xyz rotation, translation.
 * lx/ly/lz = centred ligand (center = origin); returns 1 when Mol.Too_long would be raised */
int orc_apply_config(int L, const double *lx, const double *ly, const double *lz,
                     int n_rbonds, const int32_t *rb_left, const int32_t *rb_right,
                     const int32_t *rg_off, const int32_t *rg_idx,
                     const double *config, int n_config, double *ox, double *oy, double *oz) {
    double *x = (double *)malloc(sizeof(double) * 3 * (size_t)L), *y = x + L, *z = y + L;
    memcpy(x, lx, sizeof(double) * L); memcpy(y, ly, sizeof(double) * L); memcpy(z, lz, sizeof(double) * L);
    double cen[3] = {0.0, 0.0, 0.0};
    for (int b = 0; b + 6 < n_config && b < n_rbonds; b++) {
        orc_rotate_bond(x, y, z, rb_left[b], rb_right[b], rg_off[b + 1] - rg_off[b], rg_idx + rg_off[b], config[6 + b]);
        cen[0] = orc_favg(L, x); cen[1] = orc_favg(L, y); cen[2] = orc_favg(L, z);      /* update_center */
    }
    int too_long = orc_radius(L, x, y, z, cen) > 12.0;
    double rot[9];
    orc_rot_r_xyz(config[3], config[4], config[5], rot);
    orc_rotate_then_translate(L, x, y, z, rot, config, ox, oy, oz);
    free(x);
    return too_long;
}

/* ------------------------------------------------------------------------ */
/* exhaustive rigid scan: src/lds.ml:1040-1114 */
static double scan_score_pose(const orc_scan_args *a, const double *x, const double *y, const double *z) {
    double e;
    if (a->scorer == 0)
        e = orc_ene_inter_shifted_brute(a->P, a->px, a->py, a->pz, a->pq, a->panum,
                                        a->L, x, y, z, a->lq, a->lanum);
    else if (a->scorer == 1)
        e = orc_ene_inter_global_brute(a->P, a->px, a->py, a->pz, a->pq, a->panum,
                                       a->L, x, y, z, a->lq, a->lanum);
    else
        e = orc_ene_inter_interp(a->g_step, a->g_dims, a->maps, a->L, x, y, z, a->ltyp);
    return a->e_intra_const + e;   /* lds.ml:1324-1325 */
}

typedef struct { double s; int64_t f; } topent;
/* keep the k lowest scores; ties resolved towards the smaller frame (= earlier in loop order) */
static int ent_less(topent a, topent b) { return (a.s < b.s) || (a.s == b.s && a.f < b.f); }
static void top_insert(topent *top, int *n, int k, topent e) {
    if (k <= 0) return;
    if (*n == k && !ent_less(e, top[k - 1])) return;
    int pos = (*n < k) ? *n : k - 1;
    while (pos > 0 && ent_less(e, top[pos - 1])) { top[pos] = top[pos - 1]; pos--; }
    top[pos] = e;
    if (*n < k) (*n)++;
}

static void scan_lattice(const orc_scan_args *a, int dims[3], double mins[3]) {
    /* ROI.get_bounds (ROI.ml:76-82) and Grid.from_box over the bounds cube (lds.ml:1065-1069) */
    double r = a->roi_r;
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = a->roi_c[d] - r; hi[d] = a->roi_c[d] + r; mins[d] = lo[d]; }
    orc_grid_from_box(a->trans_step, hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], dims);
}

void orc_scan(const orc_scan_args *a, double *top_scores, int64_t *top_frames, orc_scan_result *res) {
    int dims[3];
    double mins[3];
    scan_lattice(a, dims, mins);
    int L = a->L, n_rot = a->n_rot;
    int x_dim = dims[0], xy_dim = dims[0] * dims[1];
    int64_t nvox = (int64_t)dims[0] * dims[1] * dims[2];
    int64_t p0 = a->first_point, p1 = (a->n_points < 0) ? nvox : a->first_point + a->n_points;
    if (p1 > nvox) p1 = nvox;
    /* lds.ml:1044-1052: the AND prefilter on the lattice point is only armed when the ligand's own
     * centre is inside one of its atoms */
    int center_filter = 0;
    if (a->vdw_mask)
        center_filter = orc_is_ligand_center_vdW_occupied(L, a->lx, a->ly, a->lz, a->lr, 0.0, 0.0, 0.0);
    /* lds.ml:1081-1082 rotated copies of the centred ligand */
    double *rx = (double *)malloc(sizeof(double) * 3 * (size_t)L * (size_t)n_rot);
    double *ry = rx + (size_t)L * n_rot, *rz = ry + (size_t)L * n_rot;
    for (int r = 0; r < n_rot; r++)
        for (int i = 0; i < L; i++) {
            double v[3] = {a->lx[i], a->ly[i], a->lz[i]}, o[3];
            orc_rot_rotate(a->rot9 + 9 * (size_t)r, v, o);
            rx[(size_t)r * L + i] = o[0]; ry[(size_t)r * L + i] = o[1]; rz[(size_t)r * L + i] = o[2];
        }
    int k = a->topk > 0 ? a->topk : 0;
    topent *top = (topent *)malloc(sizeof(topent) * (size_t)(k > 0 ? k : 1));
    int ntop = 0;
    double best = INFINITY;
    int64_t best_frame = -1, n_scored = 0, n_cand = 0;
    double r2 = a->roi_r * a->roi_r;      /* ROI.ml:19-20 out_r2 */
    double *x = (double *)malloc(sizeof(double) * 3 * (size_t)L), *y = x + L, *z = y + L;
    for (int64_t p = p0; p < p1; p++) {   /* idx order == (z, y, x) loop order of lds.ml:1083-1109 */
        int kk = (int)(p / xy_dim);
        int jj = (int)((p - (int64_t)kk * xy_dim) / x_dim);
        int ii = (int)(p - ((int64_t)kk * xy_dim + (int64_t)jj * x_dim));
        double pos[3] = {mins[0] + orc_grid_node(a->trans_step, dims[0], ii),
                         mins[1] + orc_grid_node(a->trans_step, dims[1], jj),
                         mins[2] + orc_grid_node(a->trans_step, dims[2], kk)};
        if (!(dist2(a->roi_c[0], a->roi_c[1], a->roi_c[2], pos[0], pos[1], pos[2]) < r2)) continue;
        if (center_filter && orc_vdw_clash_AND(a->m_step, a->m_dims, a->vdw_mask, pos[0], pos[1], pos[2]))
            continue;
        for (int r = 0; r < n_rot; r++) {
            n_cand++;
            /* translate_copy_to (mol.ml:687-696): x + (pos - center), center = origin */
            for (int i = 0; i < L; i++) {
                x[i] = rx[(size_t)r * L + i] + (pos[0] - 0.0);
                y[i] = ry[(size_t)r * L + i] + (pos[1] - 0.0);
                z[i] = rz[(size_t)r * L + i] + (pos[2] - 0.0);
            }
            if (a->vdw_mask && orc_protein_ligand_clash(a->m_step, a->m_dims, a->vdw_mask, L, x, y, z))
                continue;
            double s = scan_score_pose(a, x, y, z);
            n_scored++;
            int64_t frame = (int64_t)r + (int64_t)n_rot * p;
            topent e = {s, frame};
            top_insert(top, &ntop, k, e);
            if (s < best) { best = s; best_frame = frame; }
        }
    }
    for (int i = 0; i < ntop; i++) { top_scores[i] = top[i].s; top_frames[i] = top[i].f; }
    res->n_scored = n_scored; res->n_candidates = n_cand;
    res->best_score = best; res->best_frame = best_frame; res->n_top = ntop;
    res->lattice_dims[0] = dims[0]; res->lattice_dims[1] = dims[1]; res->lattice_dims[2] = dims[2];
    free(x); free(top); free(rx);
}

void orc_scan_score_frames(const orc_scan_args *a, int n, const int64_t *frames, double *out) {
    int dims[3];
    double mins[3];
    scan_lattice(a, dims, mins);
    int L = a->L;
    int x_dim = dims[0], xy_dim = dims[0] * dims[1];
    double *x = (double *)malloc(sizeof(double) * 3 * (size_t)L), *y = x + L, *z = y + L;
    for (int f = 0; f < n; f++) {
        int64_t p = frames[f] / a->n_rot;
        int r = (int)(frames[f] - p * a->n_rot);
        int kk = (int)(p / xy_dim);
        int jj = (int)((p - (int64_t)kk * xy_dim) / x_dim);
        int ii = (int)(p - ((int64_t)kk * xy_dim + (int64_t)jj * x_dim));
        double pos[3] = {mins[0] + orc_grid_node(a->trans_step, dims[0], ii),
                         mins[1] + orc_grid_node(a->trans_step, dims[1], jj),
                         mins[2] + orc_grid_node(a->trans_step, dims[2], kk)};
        for (int i = 0; i < L; i++) {
            double v[3] = {a->lx[i], a->ly[i], a->lz[i]}, o[3];
            orc_rot_rotate(a->rot9 + 9 * (size_t)r, v, o);
            x[i] = o[0] + (pos[0] - 0.0); y[i] = o[1] + (pos[1] - 0.0); z[i] = o[2] + (pos[2] - 0.0);
        }
        out[f] = scan_score_pose(a, x, y, z);
    }
    free(x);
}

/* ------------------------------------------------------------------------ */
/* CPU arm for bench.py: many poses of one rigid centred ligand, shifted brute scorer,
 * one contiguous block of poses per OpenMP thread (mirrors Parany's fork-per-core model). */
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_score_poses_shifted_mt(int P, const double *px, const double *py, const double *pz,
                                const double *pq, const int32_t *panum,
                                int L, const double *cx, const double *cy, const double *cz,
                                const double *lq, const int32_t *lanum,
                                int64_t n_poses, const double *rot9, const double *trans3,
                                double *out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *x = (double *)malloc(sizeof(double) * 3 * (size_t)L), *y = x + L, *z = y + L;
#pragma omp for schedule(static)
        for (int64_t p = 0; p < n_poses; p++) {
            orc_rotate_then_translate(L, cx, cy, cz, rot9 + 9 * p, trans3 + 3 * p, x, y, z);
            out[p] = orc_ene_inter_shifted_brute(P, px, py, pz, pq, panum, L, x, y, z, lq, lanum);
        }
        free(x);
    }
}
