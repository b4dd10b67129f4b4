/* mmo_oracle.h -- CPU restatement of the UnixJunkie/MMO scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mmo_b200/ may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and the cpu_baseline / reference
 * legs of bench.py use it, and only as the checker / the CPU arm.
 *
 * PARITY UNPINNED: the reference is OCaml, there is no OCaml toolchain in this
 * image, and the reference ships no golden energies or known-answer tests for
 * this path.  The oracle is therefore pinned only by
 *  (1) a literal re-statement of the reference's evaluation order
 *      (each function cites the file:line it follows),
 *  (2) the hand-derived known answers of SURVEY.md Appendix B,
 *  (3) an independent mpmath evaluation of the same formulas in tests/.
 *
 * All arithmetic is IEEE double, compiled with -ffp-contract=off because OCaml
 * native code never contracts a*b+c into an FMA.
 */
#ifndef MMO_ORACLE_H
#define MMO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ANUMS 119             /* UFF.ml:31 */
#define ORC_CUTOFF 12.0           /* const.ml:10 */
#define ORC_MAX_E 100000.0        /* params.ml:26 */

/* ---- FF.ml / math.ml scalars ---- */
double orc_pow6(double x);                     /* FF.ml:11-12 */
double orc_shift_12A(double d);                /* FF.ml:17-20 */
double orc_geo_mean(double x, double y);       /* FF.ml:14-15 */
double orc_non_zero_dist(double x);            /* math.ml:58-62 */
double orc_elec_weight(void);                  /* UFF.ml:25 */
double orc_beta(double temp_K);                /* lds.ml:66-67 */
/* UFF.ml:32-51; returns x_ij in out[0], d_ij in out[1] (NaN if unsupported) */
void orc_vdw_xidi(int anum1, int anum2, double out[2]);
double orc_vdw_radius(int anum);               /* ptable.ml:41-54 */

/* ---- direct energies, mol.ml:796-903, 928-989 ---- */
double orc_ene_inter_global_brute(int P, const double *px, const double *py, const double *pz,
                                  const double *pq, const int32_t *panum,
                                  int L, const double *lx, const double *ly, const double *lz,
                                  const double *lq, const int32_t *lanum);
double orc_ene_inter_shifted_brute(int P, const double *px, const double *py, const double *pz,
                                   const double *pq, const int32_t *panum,
                                   int L, const double *lx, const double *ly, const double *lz,
                                   const double *lq, const int32_t *lanum);
/* the Bst variant with neighbours visited in receptor index order (library order is unpinned);
 * out[0] = EW*sum_elec, out[1] = sum_vdW */
void orc_ene_inter_shifted_components(int P, const double *px, const double *py, const double *pz,
                                      const double *pq, const int32_t *panum,
                                      int L, const double *lx, const double *ly, const double *lz,
                                      const double *lq, const int32_t *lanum, double out[2]);
/* dists: N*N topological distances, element (i,j) at i + j*N (mol.ml:151-152) */
double orc_ene_intra_uffnb_brute(int L, const double *lx, const double *ly, const double *lz,
                                 const double *lq, const int32_t *lanum, const int32_t *dists);
void orc_ene_inter_shifted_grid(int P, const double *px, const double *py, const double *pz,
                                const double *pq, const int32_t *panum,
                                double x, double y, double z,
                                int T, const int32_t *tanum, const double *tq, double *out);

/* ---- grid.ml / G3D.ml ---- */
int orc_grid_num_steps(double dx, double length);       /* grid.ml:37-38 */
/* fills dims[3]; node coordinate i on any axis is i*((step*n')/n') (grid.ml:49-51) */
void orc_grid_from_box(double step, double bx, double by, double bz, int dims[3]);
double orc_grid_node(double step, int dim, int i);
double orc_trilin(double step, const int dims[3], const float *arr,
                  double px, double py, double pz);      /* G3D.ml:97-157 */
double orc_ene_inter_interp(double step, const int dims[3], const float *maps /* T maps, type-major */,
                            int L, const double *lx, const double *ly, const double *lz,
                            const int32_t *ltyp);        /* mol.ml:1012-1020 */
/* lds.ml:452-469: maps must be zero-filled (G3D.create); mask may be NULL (all voxels);
 * mask bit idx is (mask[idx>>3] >> (idx&7)) & 1 */
void orc_grid_build(int P, const double *px, const double *py, const double *pz,
                    const double *pq, const int32_t *panum,
                    double step, const int dims[3], const uint8_t *mask,
                    int T, const int32_t *tanum, const double *tq, float *maps);
/* lds.ml:269-305 sphere mask of radius r around c */
void orc_bitmask_sphere(double step, const int dims[3], double cx, double cy, double cz,
                        double r, uint8_t *mask);

/* ---- vdW occupancy mask and clash tests, lds.ml:148-196, G3D.ml:162-213, mol.ml:1195-1218 ---- */
void orc_vdw_volume(int P, const double *px, const double *py, const double *pz, const double *pr,
                    double step, const int dims[3], uint8_t *mask);
/* lds.ml:172-184 first solvent shell; lds.ml:97-145 voxels within 12 A of the protein */
void orc_first_solvent_shell(int P, const double *px, const double *py, const double *pz, const double *pr,
                             double step, const int dims[3], uint8_t *mask);
void orc_bitmask_whole_protein(int P, const double *px, const double *py, const double *pz,
                               double step, const int dims[3], uint8_t *mask);
/* N4: lds.ml:204-236 protein_desolv (res: one double per voxel), lds.ml:239-267 desolvation_penalty */
void orc_protein_desolv(int P, const double *px, const double *py, const double *pz, const double *pq,
                        double step, const int dims[3], const uint8_t *shell, const double roi[4], double *res);
void orc_desolvation_penalty(double step, const int dims[3], const uint8_t *prot_shell, const double *contribs,
                             int L, const double *lx, const double *ly, const double *lz, const double *lq,
                             const double *lr, double *out_prot, double *out_lig);
int orc_vdw_clash_OR(double step, const int dims[3], const uint8_t *mask, double x, double y, double z);
int orc_vdw_clash_AND(double step, const int dims[3], const uint8_t *mask, double x, double y, double z);
int orc_protein_ligand_clash(double step, const int dims[3], const uint8_t *mask,
                             int L, const double *lx, const double *ly, const double *lz);
int orc_is_ligand_center_vdW_occupied(int L, const double *lx, const double *ly, const double *lz,
                                      const double *lr, double cx, double cy, double cz);

/* ---- rotations: rot.ml, quat.ml, SO3.ml ---- */
void orc_so3_quat(int n, int i, double q[4]);            /* SO3.ml:18-30, (w,x,y,z) */
void orc_so3_rotations(int n, double *rot9);             /* SO3.ml:35-39 */
void orc_rot_of_axis_angle(double x, double y, double z, double theta, double r[9]); /* rot.ml:121-146 */
void orc_rot_rx(double theta, double r[9]);              /* rot.ml:22-28 */
void orc_rot_ry(double theta, double r[9]);              /* rot.ml:31-37 */
void orc_rot_rz(double theta, double r[9]);              /* rot.ml:40-46 */
void orc_rot_r_xyz(double a, double b, double g, double r[9]);  /* rot.ml:52-66 */
void orc_rot_decompose(const double r[9], double abg[3]);      /* rot.ml:71-75 */
void orc_rot_mult(const double r1[9], const double r2[9], double out[9]);  /* rot.ml:77-94 */
void orc_rot_rotate(const double r[9], const double v[3], double out[3]);  /* rot.ml:97-100 */

/* ---- pose builders, mol.ml:593-710 ---- */
/* mean of an array as Batteries' A.favg (unpinned: Kahan-compensated sum / n) */
double orc_favg(int n, const double *a);
/* centered_rotate then translate_by: out = R*(x) + t, input already centred (mol.ml:669-672) */
void orc_rotate_then_translate(int L, const double *cx, const double *cy, const double *cz,
                               const double rot[9], const double t[3],
                               double *ox, double *oy, double *oz);
/* center_rotate_translate_copy (mol.ml:705-710): subtract 'center', rotate, translate */
void orc_center_rotate_translate(int L, const double *x, const double *y, const double *z,
                                 const double center[3], const double rot[9], const double t[3],
                                 double *ox, double *oy, double *oz);
/* rotate_bond (mol.ml:610-631), in place; group lists the rgroup atoms (axis tip excluded) */
void orc_rotate_bond(double *x, double *y, double *z, int left, int right,
                     int ngroup, const int32_t *group, double alpha);
double orc_radius(int L, const double *x, const double *y, const double *z, const double center[3]); /* mol.ml:576-583 */

int orc_apply_config(int L, const double *lx, const double *ly, const double *lz,
                     int n_rbonds, const int32_t *rb_left, const int32_t *rb_right,
                     const int32_t *rg_off, const int32_t *rg_idx,
                     const double *config, int n_config, double *ox, double *oy, double *oz);   /* optim.ml:64-80 */

/* ---- exhaustive rigid scan, lds.ml:1040-1114 ---- */
typedef struct {
    int P; const double *px, *py, *pz, *pq; const int32_t *panum;           /* receptor */
    int L; const double *lx, *ly, *lz, *lq, *lr; const int32_t *lanum, *ltyp; /* centred ligand */
    /* scorer: 0 = shifted brute (BrL), 1 = global brute (BrG), 2 = interpolated */
    int scorer;
    double g_step; int g_dims[3]; const float *maps;                        /* for scorer 2 */
    const uint8_t *vdw_mask; double m_step; int m_dims[3];                   /* prefilter; NULL = off */
    double roi_c[3], roi_r;
    double trans_step;
    int n_rot; const double *rot9;
    double e_intra_const;
    int topk;
    int64_t first_point, n_points;  /* lattice-point sub-range (idx = i + j*xd + k*xd*yd); n_points<0 = all */
} orc_scan_args;
typedef struct {
    int64_t n_scored;        /* poses that reached the scorer */
    int64_t n_candidates;    /* in-ROI lattice points * rotations */
    double best_score; int64_t best_frame;   /* frame = rot_i + n_rot*(i + j*x_dim + k*xy_dim) */
    int n_top;               /* filled entries in top_scores/top_frames (ascending score) */
    int lattice_dims[3];
} orc_scan_result;
void orc_scan(const orc_scan_args *a, double *top_scores, int64_t *top_frames, orc_scan_result *res);
/* scores of an explicit list of frames (for spot checks at full size) */
void orc_scan_score_frames(const orc_scan_args *a, int n, const int64_t *frames, double *out);

/* ---- Monte-Carlo chain, lds.ml:741-1000 (restated with a counter-based RNG, see .c) ---- */
typedef struct {
    int L; const double *lx, *ly, *lz, *lq; const int32_t *lanum, *ltyp, *dists; /* centred ligand */
    int n_rbonds; const int32_t *rb_left, *rb_right, *rg_off, *rg_idx;
    /* inter scorer: 0 shifted brute, 2 interpolated */
    int scorer;
    int P; const double *px, *py, *pz, *pq; const int32_t *panum;
    double g_step; int g_dims[3]; const float *maps;
    double roi_c[3], roi_r;
    int tweak_rbonds, hard_roi, no_flip, intra_nb;
    double beta;
    int n_steps;
    uint64_t seed;
    double rot0[9], pos0[3];
} orc_mc_args;
typedef struct {
    double best_E, prev_E; double best_rot[9], best_pos[3];
    int64_t n_accept_rigid, n_reject_rigid, n_accept_conf, n_reject_conf, n_ooroi, n_ezero;
    int too_long;     /* Mol.Too_long raised: run aborted at that frame */
    int frames_done;
    double max_rot, max_trans;
} orc_mc_result;
/* trace (optional, may be NULL): per frame [curr_E, E_inter, E_intra, accepted] */
void orc_mc_run(const orc_mc_args *a, orc_mc_result *res, double *best_xyz /* 3L */, double *trace);
/* the RNG both the oracle chain and the CUDA chain use (NOT OCaml's Random.State) */
double orc_rng_uniform(uint64_t seed, uint64_t counter);

/* ---- multi-threaded CPU arm for bench.py (one pose block per OpenMP thread) ---- */
int orc_num_threads(void);
void orc_score_poses_shifted_mt(int P, const double *px, const double *py, const double *pz,
                                const double *pq, const int32_t *panum,
                                int L, const double *cx, const double *cy, const double *cz,
                                const double *lq, const int32_t *lanum,
                                int64_t n_poses, const double *rot9, const double *trans3,
                                double *out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
