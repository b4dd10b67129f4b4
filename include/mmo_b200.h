/* mmo_b200.h -- C ABI of libmmo_b200.so: the B200 (sm_100a) implementation of MMO's
 * ligand-receptor scoring hot path.
 *
 * The reference (UnixJunkie/MMO) has no FFI on this path: the "operator API" is a set of OCaml
 * functions in src/mol.ml, src/G3D.ml, src/grid.ml, src/SO3.ml and the drivers in src/lds.ml.
 * Each entry point below names the reference function it replaces (file:line relative to
 * /root/reference).  INTEGRATION.md shows the OCaml `external` + C stub a maintainer adds.
 *
 * Conventions
 *  - every function returns 0 (MMO_OK) on success, a negative MMO_E* code otherwise; the message
 *    is available from mmo_last_error() (thread local).  No C++ exception crosses the boundary.
 *  - the caller owns all host buffers; the library owns device memory behind opaque handles.
 *  - OCaml `float array` is a flat unboxed double[] and can be passed as is; OCaml `int array`
 *    must be converted to int32_t by the stub.
 *  - one device per process (mmo_init), created AFTER any fork() (CUDA contexts do not survive
 *    fork; run lds with --par 1/1).  Handles are not thread-safe.
 *  - there is NO CPU fallback: without a usable CUDA device every compute call fails.
 */
#ifndef MMO_B200_H
#define MMO_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MMO_OK 0
#define MMO_EINVAL (-1)   /* bad argument */
#define MMO_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, out of memory) */
#define MMO_ESTATE (-3)   /* library not initialised / handle destroyed */
#define MMO_ENCCL (-4)    /* NCCL not loadable or a collective failed */
#define MMO_ENOMEM (-5)   /* host allocation failed */

/* -ff BrG | BrL/Bst  (src/lds.ml:570-584) */
#define MMO_VARIANT_GLOBAL 0   /* Mol.ene_inter_UFF_global_brute,  src/mol.ml:796-818 */
#define MMO_VARIANT_SHIFTED 1  /* Mol.ene_inter_UFF_shifted_brute, src/mol.ml:822-849 (== _shifted_bst 928-960) */

/* MMO_PREC_FP32: fp32 pair arithmetic, fp64 accumulation, close-contact pairs re-evaluated in fp64;
 *                per-pose energy within max(1e-6*|E|, 1e-4 kcal/mol) of the reference arithmetic.
 * MMO_PREC_FP64: every operation in IEEE double, no FMA contraction, summation order of
 *                src/mol.ml:828-848 (receptor outer, ligand inner): bit-identical to the reference
 *                arithmetic, hence identical argmin index and top-k order. */
#define MMO_PREC_FP32 0
#define MMO_PREC_FP64 1

typedef struct mmo_receptor mmo_receptor;
typedef struct mmo_ligand mmo_ligand;
typedef struct mmo_grid mmo_grid;
typedef struct mmo_mask mmo_mask;

/* ---------------------------------------------------------------- runtime ---------------- */
int mmo_init(int device);          /* select the CUDA device, create the library stream */
int mmo_shutdown(void);
const char *mmo_last_error(void);
int mmo_device_count(int *n);
const char *mmo_build_info(void);  /* "sm_100a, CUDA x.y, built <date>" */
/* number of kernels this library has launched since mmo_init (bench.py's gpu_launches) */
int64_t mmo_launch_count(void);
/* per-kernel device time (CUDA events around every launch on the library stream); ids below */
#define MMO_K_DIRECT_FP32 0
#define MMO_K_HARD_FIX 1
#define MMO_K_DIRECT_FP64 2
#define MMO_K_INTRA 3
#define MMO_K_GRID_BUILD 4
#define MMO_K_INTERP 5
#define MMO_K_PREFILTER 6
#define MMO_K_REDUCE 7
#define MMO_K_VDW_MASK 8
#define MMO_K_MC 9
int mmo_kernel_timing(int on);    /* switches collection on/off and zeroes the counters */
int mmo_kernel_time_get(int kernel_id, double *ms, int64_t *launches);
/* device-side timing on the library stream (CUDA events) and an L2 flush, for bench.py */
int mmo_timer_start(void);
int mmo_timer_stop(float *ms);
int mmo_l2_flush(void);
int mmo_sync(void);
/* FP32 FMA-chain and HBM copy micro-benchmarks: the measured roofline denominators */
int mmo_measure_fp32_peak(double *tflops);
int mmo_measure_fp64_peak(double *tflops);
int mmo_measure_hbm_copy(double *gbs);
/* the gather roof of the interpolated lookup (K4): 8-corner reads of pseudo-random cells of T type-major f32 maps of
 * the given dims (L2 resident when they fit), no arithmetic; lookups/s (x 32 B = gathered bytes/s) */
int mmo_measure_l2_gather(const int32_t dims[3], int32_t T, double *lookups_per_s);
/* the same for the layout the lookup kernel actually reads: the z-pair copy of the maps ({v[k], v[k+1]} as one 8-byte
 * element), four 8-byte reads in two rows of x per lookup */
int mmo_measure_l2_gather_zpair(const int32_t dims[3], int32_t T, double *lookups_per_s);
/* raw device buffers, so that a caller can keep pose batches resident in HBM */
int mmo_dev_alloc(size_t bytes, void **dptr);
int mmo_dev_free(void *dptr);
int mmo_h2d(void *dptr, const void *host, size_t bytes);
int mmo_d2h(void *host, const void *dptr, size_t bytes);
/* page-locked host memory for the end-to-end path */
int mmo_host_alloc(size_t bytes, void **hptr);
int mmo_host_free(void *hptr);

/* ---------------------------------------------------------------- molecules -------------- */
/* Receptor atoms (Mol.t of the protein, src/mol.ml:17-35: xs, ys, zs, q_a, elt_a).
 * Unsupported elements (not in src/UFF.ml:10-22) give NaN energies, like the reference. */
int mmo_receptor_create(int32_t n, const double *xs, const double *ys, const double *zs,
                        const double *q, const int32_t *anum, mmo_receptor **out);
int mmo_receptor_destroy(mmo_receptor *rec);

/* Ligand (Mol.t of a ligand).  xs/ys/zs are the template conformer; for every *_poses entry point
 * it must already be centred at the origin (lds.ml:44-52 preprocess_ligand).
 *   r      vdW radii (mol.ml r_a), may be NULL when the centre-occupancy prefilter is not used
 *   typ    FF atom types (mol.ml:456-462), may be NULL unless interpolated scoring is used
 *   dists  N*N topological distances, (i,j) at i + j*N (mol.ml:151-152), may be NULL unless
 *          intra-ligand energies are used; interacting pairs are those with dists >= 3
 *   rotatable bonds (mol.ml:27-31): axis left[b] -> right[b], group atoms (tip excluded) in CSR form */
int mmo_ligand_create(int32_t n, const double *xs, const double *ys, const double *zs,
                      const double *q, const double *r, const int32_t *anum, const int32_t *typ,
                      const int32_t *dists,
                      int32_t n_rbonds, const int32_t *rb_left, const int32_t *rb_right,
                      const int32_t *rg_off, const int32_t *rg_idx, mmo_ligand **out);
int mmo_ligand_destroy(mmo_ligand *lig);

/* ---------------------------------------------------------------- direct pair path -------- */
/* Mol.ene_inter_UFF_{global,shifted}_brute prot lig  (src/mol.ml:796-849), batched over n_poses
 * ligand copies given by explicit coordinates: pose p, atom j at xs[p*L + j].
 * n_poses = 1 is the literal `ene_inter : Mol.t -> float` closure of src/lds.ml:1952-1955. */
int mmo_score_coords(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                     int64_t n_poses, const double *xs, const double *ys, const double *zs,
                     double *out_E);
/* same energies for poses given as (rotation, translation) of the centred template:
 * Mol.rotate_then_translate_copy lig rot pos (src/mol.ml:669-672) then ene_inter.
 * rot9: n_poses row-major 3x3 (Rot.t a..i, src/rot.ml:5-7); trans3: n_poses x 3. */
int mmo_score_poses(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                    int64_t n_poses, const double *rot9, const double *trans3, double *out_E);
/* device-resident variant (all three pointers are device memory from mmo_dev_alloc) */
int mmo_score_poses_dev(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                        int64_t n_poses, const double *d_rot9, const double *d_trans3, double *d_out_E);
/* the same for explicit coordinates already resident on the device (pose-major, stride = ligand atoms) */
int mmo_score_coords_dev(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                         int64_t n_poses, const double *d_xs, const double *d_ys, const double *d_zs, double *d_out_E);
/* Mol.ene_inter_UFF_shifted_bst_components (src/mol.ml:928-956): EW*sum_elec and sum_vdW apart
 * (fp64, ligand-outer order with receptor atoms in index order) */
int mmo_score_coords_components(const mmo_receptor *rec, const mmo_ligand *lig, int64_t n_poses,
                                const double *xs, const double *ys, const double *zs,
                                double *out_elec, double *out_vdw);
/* Mol.ene_intra_UFFNB_brute (src/mol.ml:881-903) for n_confs conformers of the ligand;
 * fp64 in the reference's (i<j) order: bit-identical */
int mmo_intra_nb(const mmo_ligand *lig, int64_t n_confs, const double *xs, const double *ys,
                 const double *zs, double *out_E);
/* statistics of the last FP32 direct launch: pairs whose distance was evaluated, pairs inside the
 * 12 A cut-off, close-contact pairs re-evaluated in fp64 */
int mmo_last_pair_stats(int64_t *pairs_evaluated, int64_t *pairs_inside, int64_t *pairs_fp64);
/* The strict (bit-identical) kernels divide many numerators by one r through a shared correctly rounded reciprocal and
 * two fused corrections; this runs that routine against the IEEE division on n pseudo-random (a, b) of the path's
 * ranges, hard significands included, and reports how many quotients differ in any bit (must be 0). */
int mmo_selftest_division(uint64_t seed, int64_t n, int64_t *mismatches);
/* Which FP32 kernel scores a pose list (mmo_score_poses / mmo_score_coords; scans always take the pose kernel):
 * 0 = automatic (item kernel from 32 k pose-atoms on: incoherent lists are sorted by atom position first),
 * 1 = pose kernel (warp = 64 consecutive poses), 2 = item kernel (shifted variant only: MMO_VARIANT_GLOBAL has
 * nothing to cull and always takes the pose kernel, or the fp64 kernel beyond 1.2e5 receptor x ligand atom pairs,
 * where fp32 rounding noise would break the contract).  Same accuracy contract either way. */
int mmo_direct_set_mode(int mode);
/* (pose, ligand atom) combinations the FP32 kernel flagged for the fp64 close-contact pass in the last launch */
int mmo_last_fix_stats(int64_t *atoms_flagged);
/* pair statistics cost two extra instructions per pair: off by default, switch on for accounting runs */
int mmo_set_collect_stats(int on);

/* ---------------------------------------------------------------- energy grids ------------ */
/* Grid.from_box (src/grid.ml:40-52): dims = ceil(len/step)+1, lowest corner at the origin */
int mmo_grid_from_box(double step, double bx, double by, double bz, int32_t dims[3]);
/* Lds.pre_calculate_FF_components_grid (src/lds.ml:452-469) with Mol.ene_inter_UFF_shifted_grid
 * (src/mol.ml:964-989): T maps of x_dim*y_dim*z_dim float32, index i + j*x_dim + k*x_dim*y_dim,
 * value = f32(min(1e5, E)) on masked voxels and 0 elsewhere.  mask_bits: bit idx at
 * (mask[idx>>3] >> (idx&7)) & 1, NULL = every voxel.  out_maps (host, type-major) and out_grid
 * (device-resident handle) may each be NULL. */
int mmo_grid_build(const mmo_receptor *rec, double step, const int32_t dims[3],
                   const uint8_t *mask_bits, int32_t T, const int32_t *type_anum,
                   const double *type_q, float *out_maps, mmo_grid **out_grid);
/* maps read back from the reference's .ba1 cache (G3D.of_ba1_file, src/G3D.ml:55-64) */
int mmo_grid_upload(double step, const int32_t dims[3], int32_t T, const float *maps, mmo_grid **out);
int mmo_grid_download(const mmo_grid *grid, float *maps);
int mmo_grid_destroy(mmo_grid *grid);
/* G3D.to_ba1_file / of_ba1_file (src/G3D.ml:14-64): raw little-endian f32 + `.dims` side-car */
int mmo_grid_write_ba1(const mmo_grid *grid, int32_t type, const char *path);
int mmo_grid_read_ba1(const char *const *paths, int32_t T, mmo_grid **out);
/* G3D.trilin grid arr p (src/G3D.ml:97-157) for n points against map `type` */
int mmo_trilin(const mmo_grid *grid, int32_t type, int64_t n, const double *xs, const double *ys,
               const double *zs, double *out);
/* Mol.ene_inter_UFF_interp grid ff_comps lig (src/mol.ml:1012-1020); fp64 on f32 data in the
 * reference's order: bit-identical.  The ligand needs `typ`. */
int mmo_score_interp_coords(const mmo_grid *grid, const mmo_ligand *lig, int64_t n_poses,
                            const double *xs, const double *ys, const double *zs, double *out_E);
int mmo_score_interp_poses(const mmo_grid *grid, const mmo_ligand *lig, int64_t n_poses,
                           const double *rot9, const double *trans3, double *out_E);
int mmo_score_interp_poses_dev(const mmo_grid *grid, const mmo_ligand *lig, int64_t n_poses,
                               const double *d_rot9, const double *d_trans3, double *d_out_E);

/* ---------------------------------------------------------------- vdW occupancy mask ------ */
/* Lds.vdW_volume grid prot (src/lds.ml:148-196): bit set where dist^2 < r^2 to some atom */
int mmo_vdw_mask_build(int32_t n, const double *xs, const double *ys, const double *zs,
                       const double *radii, double step, const int32_t dims[3], uint8_t *out_bits,
                       mmo_mask **out_mask);
/* N3: the other bitmasks of the simulation grid, built on the device, same bit layout.
 *   Lds.first_solvent_shell   (src/lds.ml:172-184): inside r_vdW + 1.4 A of some atom and inside r_vdW of none
 *   Lds.bitmask_whole_protein (src/lds.ml:97-145) : nearest protein atom closer than 12 A
 *   Lds.bitmask_ROI_only      (src/lds.ml:269-305): closer than roi_r + 24 A to the ROI centre (the grid-build mask) */
int mmo_mask_first_solvent_shell(int32_t n, const double *xs, const double *ys, const double *zs, const double *radii,
                                 double step, const int32_t dims[3], uint8_t *out_bits, mmo_mask **out_mask);
int mmo_mask_whole_protein(int32_t n, const double *xs, const double *ys, const double *zs,
                           double step, const int32_t dims[3], uint8_t *out_bits, mmo_mask **out_mask);
int mmo_mask_roi_only(const double roi_c[3], double roi_r, double step, const int32_t dims[3], uint8_t *out_bits,
                      mmo_mask **out_mask);
int mmo_mask_upload(double step, const int32_t dims[3], const uint8_t *bits, mmo_mask **out);
int mmo_mask_download(const mmo_mask *mask, uint8_t *out_bits);
/* N1: the `<rec>.bitmask` cache file of lds (Utls.bitmask_to_file / bitmask_from_file, src/utls.ml:12-20, used at
 * src/lds.ml:1543-1553): one text line of '0'/'1', Bitv.M.to_string.  bitv is not vendored; per its documentation M
 * puts the most significant bit first (character 0 = bit n-1): msb_first = 1.  UNPINNED -- msb_first = 0 writes /
 * reads the other order (Bitv.L) should a real lds file turn out to use it. */
int mmo_mask_write_bitmask(const mmo_mask *mask, const char *path, int msb_first);
int mmo_mask_read_bitmask(const char *path, double step, const int32_t dims[3], int msb_first, mmo_mask **out);
/* N1: the `.ba1.zst` variant of the map cache (Utls.zstd_compress_file / zstd_uncompress_file, src/utls.ml:22-45):
 * like the reference, the external `zstd` program does the work (`zstd --rm -qf F`, `zstd -dqfk F.zst`); fails with
 * MMO_EINVAL when it is not installed.  mmo_grid_read_ba1 accepts paths ending in .zst (lds.ml:540-553: uncompressed
 * beside the file, read, transient copy removed; the .dims side-car is never compressed). */
int mmo_zstd_compress_file(const char *path);
int mmo_zstd_uncompress_file(const char *path_zst);
int mmo_mask_destroy(mmo_mask *mask);
/* Mol.protein_ligand_clash (src/mol.ml:1195-1203): out_flags[p] = 1 when any atom of pose p has any
 * of its 8 surrounding voxels set (G3D.vdW_clash_OR, src/G3D.ml:162-186) */
int mmo_clash_poses(const mmo_mask *mask, const mmo_ligand *lig, int64_t n_poses,
                    const double *rot9, const double *trans3, uint8_t *out_flags);

/* ---------------------------------------------------------------- N3: binding-site carve --- */
/* scissors (src/scissors.ml:48-62): out_keep[i] = 1 for the protein atoms whose nearest ligand atom is within the
 * cut-off (BST.nearest_neighbor: dist = sqrt(dist2), kept if dist <= cutoff; default 5 A around the ligand) */
int mmo_carve_near_ligand(int32_t n_rec, const double *px, const double *py, const double *pz, int32_t n_lig,
                          const double *lx, const double *ly, const double *lz, double cutoff, uint8_t *out_keep,
                          int32_t *n_kept);

/* ---------------------------------------------------------------- N4: desolvation sums ---- */
/* Majeux, Scarsi and Caflisch (PROTEINS 2001) eq. (2), the rescoring terms next to the pair path.
 * Lds.protein_desolv roi grid prot_bst prot_solvent_shell prot (src/lds.ml:204-236): for every voxel of the protein's
 * first solvent shell (mmo_mask_first_solvent_shell of the receptor) inside the ROI, Const.desolvation * (voxel_vol *
 * sum over protein atoms within 12 A of (q_j / d2)^2); 0.0 elsewhere.  out_contribs (host, one double per voxel, index
 * i + j*x_dim + k*xy_dim) and out (device-resident handle; it keeps its own copy of prot_shell) may each be NULL.
 * Neighbour order inside BST.neighbors is unpinned (library not vendored): atoms are added in index order. */
typedef struct mmo_desolv mmo_desolv;
int mmo_desolv_protein(const mmo_receptor *rec, const mmo_mask *prot_shell, const double roi_c[3], double roi_r,
                       double *out_contribs, mmo_desolv **out);
/* Lds.desolvation_penalty grid prot_desolv_contribs prot_solvent_shell _prot lig (src/lds.ml:239-267) for n_poses
 * poses of the ligand (created with radii): the ligand's own first solvent shell (src/lds.ml:172-184) ANDed with the
 * protein's; out_prot[p] = sum of the desolvated voxels' protein contributions, out_lig[p] = Const.desolvation *
 * (voxel_vol * sum over those voxels and the ligand atoms with d2 < 144 of (q_j / d2)^2), both summed in voxel index
 * then atom order (Bitv.iteri_true): bit-identical to the reference's doubles */
int mmo_desolv_penalty_coords(const mmo_desolv *d, const mmo_ligand *lig, int64_t n_poses, const double *xs,
                              const double *ys, const double *zs, double *out_prot, double *out_lig);
int mmo_desolv_penalty_poses(const mmo_desolv *d, const mmo_ligand *lig, int64_t n_poses, const double *rot9,
                             const double *trans3, double *out_prot, double *out_lig);
int mmo_desolv_destroy(mmo_desolv *d);

/* ---------------------------------------------------------------- rotations (host, libm) -- */
/* SO3.rotations n (src/SO3.ml:18-39 + src/quat.ml:32-36 + src/rot.ml:136-146): n row-major 3x3 */
int mmo_so3_rotations(int32_t n, double *rot9);
int mmo_rot_r_xyz(double alpha, double beta, double gamma, double rot9[9]);  /* src/rot.ml:52-66 */
int mmo_rot_decompose(const double rot9[9], double abg[3]);                  /* src/rot.ml:71-75 */

/* ---------------------------------------------------------------- pose builders (host) ----- */
/* Optim.apply_config centered_lig conf (src/optim.ml:64-80) = the place_ligand tool
 * (src/place_ligand.ml:11-59): config = x y z alpha beta gamma [rbond angles]; the ligand handle holds the
 * centred conformer.  too_long = 1 when Mol.check_elongation_exn would raise (coordinates still returned). */
int mmo_apply_config(const mmo_ligand *lig, const double *config, int32_t n_config, double *out_xs,
                     double *out_ys, double *out_zs, int32_t *too_long);
/* lig_rot_sample (src/lig_rot_sample.ml:23-45): Mol.center_rotate_translate_copy mol rot center for n
 * rotations; out arrays hold n x L coordinates */
int mmo_rotated_copies(const mmo_ligand *lig, const double center[3], int32_t n, const double *rot9,
                       double *out_xs, double *out_ys, double *out_zs);

/* ---------------------------------------------------------------- exhaustive rigid scan ---- */
/* Lds.exhaustive_rigid_ligand_docking (src/lds.ml:1040-1114).
 * Lattice = Grid.from_box trans_step over ROI.get_bounds; loop order z, y, x, rotation;
 * frame = rot_i + n_rot*(i + j*x_dim + k*xy_dim); score = e_intra_const + ene_inter. */
typedef struct {
    const mmo_receptor *rec;      /* direct scorer (NULL when grid is given) */
    const mmo_grid *grid;         /* interpolated scorer (NULL when rec is given) */
    const mmo_ligand *lig;        /* centred ligand */
    const mmo_mask *vdw_mask;     /* protein vdW volume prefilter; NULL = no prefilter */
    int32_t variant, prec;
    double roi_c[3], roi_r;       /* ROI.sphere in simulation-box coordinates */
    double trans_step;
    int32_t n_rot;
    const double *rot9;           /* host, n_rot x 9 */
    double e_intra_const;
    int32_t topk;
    int64_t first_point, n_points;  /* lattice-point sub-range for sharding; n_points < 0 = all */
} mmo_scan_params;
typedef struct {
    int64_t n_candidates;   /* in-ROI lattice points x rotations */
    int64_t n_scored;       /* poses that passed the prefilter and were scored */
    double best_score;
    int64_t best_frame;
    double best_pos[3];
    int32_t best_rot_i;
    int32_t n_top;
    int32_t lattice_dims[3];
    int64_t pairs_evaluated, pairs_inside;   /* direct scorer only */
    float device_ms;        /* GPU time of the scan, CUDA events on the library stream */
} mmo_scan_result;
int mmo_scan(const mmo_scan_params *p, double *top_scores, int64_t *top_frames, mmo_scan_result *res);
/* the same scan as a resident job: rotations, lattice and buffers stay in HBM between calls, so a
 * caller (bench.py, a multi-GPU driver) can run it slab by slab over the active lattice points
 * (in-ROI points that passed the centre prefilter, in loop order) and read the result at the end */
typedef struct mmo_scan_job mmo_scan_job;
int mmo_scan_create(const mmo_scan_params *p, int collect_stats, mmo_scan_job **out);
int mmo_scan_num_points(const mmo_scan_job *job, int64_t *n_active_points);
int mmo_scan_run(mmo_scan_job *job, int64_t first_active, int64_t n_active);
int mmo_scan_result_get(const mmo_scan_job *job, double *top_scores, int64_t *top_frames, mmo_scan_result *res);
int mmo_scan_destroy(mmo_scan_job *job);
/* The rotation set of a run stays resident on the device between mmo_scan calls together with its visiting order (a
 * k-d sort on the host; lds builds the set once per run, src/lds.ml:1748-1752).  How a call recognises the set it is
 * handed: mode 1 (default) copies the caller's rotations to the device and compares them there with the resident copy
 * (7.2 MB for 1e5 rotations: 0.2 ms from pinned memory; every byte of the call's input moves); mode 0 compares them on
 * the host (one memcmp, 0.4 ms for the same set) and skips the upload -- for hosts with a slow link.  Either way the
 * visiting order is rebuilt only when the bytes differ.  In mode 1 the one-shot mmo_scan does not wait for the verdict:
 * it scans on the resident set at once, uploads and compares the caller's bytes on a second stream behind the first
 * slab's kernels (so a pageable source costs the same as a pinned one), and scans again only if the set turns out to be
 * another one; mmo_scan_rot_rescans counts those second scans since the library was loaded. */
int mmo_scan_set_rot_cache(int mode);
int mmo_scan_rot_rescans(int64_t *count);
/* k smallest of n device-resident energies (a conformer screen's top-k, src/lds.ml:1055-1064 semantics: ascending,
 * NaN last, ties to the smaller id); id of entry p = id_base + p.  out arrays hold k entries. */
int mmo_topk_select_dev(const double *d_E, int64_t n, int32_t k, int64_t id_base, double *out_scores,
                        int64_t *out_ids, int32_t *out_n);
/* K-way merge of per-GPU top-k lists: (score, frame) ascending, ties to the smaller frame */
int mmo_topk_merge(int32_t n_lists, int32_t k, const double *scores, const int64_t *frames,
                   const int32_t *counts, double *out_scores, int64_t *out_frames, int32_t *out_n);

/* ---------------------------------------------------------------- multi-GPU top-k ---------- */
/* One process per GPU.  Poses / lattice points / chains are sharded with no data-path collective;
 * the only exchange is one ncclAllGather of the per-GPU top-k lists (k x 16 B per rank) followed
 * by the merge above.  NCCL (libnccl.so.2) is dlopen'ed on first use.  Rank 0 creates the id with
 * mmo_nccl_unique_id and hands it to the other ranks out of band (a file, an env variable, any key-value store). */
int mmo_nccl_unique_id(uint8_t id[128]);
int mmo_nccl_init(int32_t rank, int32_t nranks, const uint8_t id[128]);
int mmo_nccl_finalize(void);
int mmo_topk_allgather_merge(int32_t k, int32_t n_local, const double *scores, const int64_t *frames,
                             double *out_scores, int64_t *out_frames, int32_t *out_n);

/* ---------------------------------------------------------------- Monte-Carlo chains ------- */
/* Lds.simulate_lig frame loop (src/lds.ml:882-995) for n_chains independent chains in one launch:
 * alternating rigid-body (Move.rand_rot / rand_trans, src/move.ml:20-54) and conformer moves
 * (Mol.tweak_rbond / flip_rbond / rotate_bond, src/mol.ml:610-647), interpolated E_inter
 * (src/mol.ml:1012-1020; give `grid`, rec = NULL) or direct shifted E_inter (--no-interp, src/mol.ml:822-849;
 * give `rec`, grid = NULL; fp64 pair terms, lane-strided summation: ~1e-13 relative to the reference order),
 * E_intra = Mol.ene_intra_UFFNB_brute when intra_nb, Metropolis at
 * beta = 1/(kB*T) (lds.ml:66-75, 931-934), acceptance windows (src/SW.ml) and adaptive step sizes
 * (lds.ml:586-621).  The reference's behaviours D1-D6/D14 of SURVEY Appendix D are mirrored
 * (in particular: nothing is accepted or rejected unless hard_roi is set).  The random stream and
 * sin/cos/exp are those of include/mmo_detmath.h, not OCaml's: chain c consumes
 * mmo_rng_uniform(seeds[c], 0, 1, 2, ...). */
typedef struct {
    double roi_c[3], roi_r;      /* ROI.sphere in simulation-box coordinates */
    double temperature_K;        /* -T, default 293.15 */
    int32_t n_steps;             /* -steps */
    int32_t tweak_rbonds;        /* 0 = --rigid-ligand */
    int32_t hard_roi;            /* --hard-ROI */
    int32_t no_flip;             /* --no-flip */
    int32_t intra_nb;            /* --intra-NB (0 = --no-E-intra) */
} mmo_mc_params;
typedef struct {
    double best_E, prev_E;
    double best_rot[9], best_pos[3];
    double max_rot, max_trans;   /* adapted step sizes at the end of the run */
    int64_t n_accept_rigid, n_reject_rigid, n_accept_conf, n_reject_conf, n_ooroi, n_ezero;
    int32_t too_long;            /* Mol.Too_long was raised: the run stopped at frames_done */
    int32_t frames_done;
} mmo_mc_result;
/* start_rot9 / start_pos3: per chain (lds.ml:2034-2047); best_xyz: n_chains x 3 x L (x.. y.. z..),
 * may be NULL; trace_chain0: n_steps x {curr_E, E_inter, E_intra, accepted(-1 = no test)}, may be NULL; the frames
 * from results[0].frames_done on (a run cut short by Mol.Too_long) hold NaN */
int mmo_mc_run(const mmo_receptor *rec, const mmo_grid *grid, const mmo_ligand *lig, const mmo_mc_params *p,
               int64_t n_chains, const uint64_t *seeds, const double *start_rot9,
               const double *start_pos3, mmo_mc_result *results, double *best_xyz,
               double *trace_chain0);

/* Lds.place_ligand_in_ROI (src/lds.ml:308-345): n_starts random start poses (rotation, position) of the centred
 * ligand inside the ROI sphere, rejecting poses in which a ligand heavy atom comes closer than 0.8 (r_i + r_j) to a
 * protein heavy atom (Mol.heavy_atom_clash, src/mol.ml:1170-1190; px.. = ALL receptor atoms, simulation-box
 * coordinates).  Host code; random numbers mmo_rng_uniform(seed, 0, 1, ...) (not OCaml's Random.State), libm sin/cos.
 * clash_check = 0 skips the rejection: in the reference Ptable.vdW_max is A.max of an array holding nan
 * (src/ptable.ml:32,84), so BST.neighbors is called with a nan radius and, if it then returns nothing (bst and
 * batteries are not vendored: unpinned), no pose is ever rejected -- pass 0 to reproduce that reading, which is also
 * the only one under which the reference can start on its own data/ example (blind rigid placement into the buried
 * 3A2J pocket never passes the test as written).  Fails after 100 000 draws, as the reference exits. */
int mmo_place_ligand_in_roi(int32_t n_rec, const double *px, const double *py, const double *pz, const int32_t *panum,
                            const mmo_ligand *lig, const double roi_c[3], double roi_r, uint64_t seed, int32_t n_starts,
                            int32_t clash_check, double *out_rot9, double *out_pos3, int32_t *n_trials);

/* ---------------------------------------------------------------- N2: ligand / receptor files on the host ----
 * mol2pqrs (src/mol2pqrs.ml:10-58, src/mol_graph.ml:45-200, src/mol2.ml:139-320) and the .pqrs reader
 * (src/pqrs.ml:19-87, src/mol.ml:368-440) inside the library: no subprocess, no GPU needed except for
 * mmo_molfile_ligand.  A molfile holds every molecule of the file in file order; molecules the reference skips
 * (disconnected atom, malformed block) are counted in n_skipped. */
typedef struct mmo_molfile mmo_molfile;
int mmo_molfile_read_mol2(const char *path, mmo_molfile **out);
/* Mol2.read_one_from_file (src/mol2.ml:350-356) as `scissors` reads its two inputs: the first molecule only, atoms and
 * bonds as written (lone pairs dropped), no graph analysis -- a protein is not one connected molecule */
int mmo_molfile_read_mol2_atoms(const char *path, mmo_molfile **out);
int mmo_molfile_read_pqrs(const char *path, int is_receptor, mmo_molfile **out);
int mmo_molfile_count(const mmo_molfile *f, int32_t *n_mols, int32_t *n_skipped);
int mmo_molfile_shape(const mmo_molfile *f, int32_t k, int32_t *n_atoms, int32_t *n_rbonds, int32_t *rg_total,
                      char *name, int32_t name_cap);
/* any output pointer may be NULL; dists is n*n with element (i, j) at i + j*n (src/mol.ml:151-152);
 * rg_off has n_rbonds + 1 entries, rg_idx the movable atoms of every bond without the axis tip (src/pqrs.ml:80-87) */
int mmo_molfile_get(const mmo_molfile *f, int32_t k, double *xs, double *ys, double *zs, double *q, double *r,
                    int32_t *anum, int32_t *typ, int32_t *dists, int32_t *rb_left, int32_t *rb_right,
                    int32_t *rg_off, int32_t *rg_idx);
/* FF atom types of the file: (anum, exact charge) -> id in first-seen order (src/mol.ml:280-293, 456-469) */
int mmo_molfile_types(const mmo_molfile *f, int32_t *n_types, int32_t *type_anum, double *type_q);
/* lds --less-charges (src/lds.ml:1887-1894, src/mol.ml:256-260, src/utls.ml:127-132): every partial charge of the
 * file rounded to two decimals (away from zero), then the FF types re-assigned: fewer (anum, q) types, fewer maps */
int mmo_molfile_reduce_charges(mmo_molfile *f);
/* the text mol2pqrs writes (values through %g) */
int mmo_molfile_write_pqrs(const mmo_molfile *f, const char *path);
/* molecule k as a device-resident ligand handle; centered != 0 applies Mol.translate_to lig V3.origin (lds.ml:44-52) */
int mmo_molfile_ligand(const mmo_molfile *f, int32_t k, int centered, mmo_ligand **out);
/* Mol2.output_one of Mol.update_mol2 mol2 m (src/mol2.ml:326-343, src/mol.ml:544-552): molecule k (read from a
 * mol2 file) written n_copies times, copy c with the coordinates xs/ys/zs[c * n_atoms ..] (NULL = the file's own);
 * append != 0 adds to an existing file */
int mmo_molfile_write_mol2(const mmo_molfile *f, int32_t k, int32_t n_copies, const double *xs, const double *ys,
                           const double *zs, const char *path, int append);
/* host-only bodies of the two pose-feed tools (no GPU, no ligand handle needed):
 * lig_rot_sample (src/lig_rot_sample.ml:23-45): Mol.center_rotate_translate_copy mol rot (Mol.get_center mol) for n
 * rotations, out arrays n x n_atoms; place_ligand (src/place_ligand.ml:36-59): Mol.center, then Optim.apply_config
 * (src/optim.ml:64-80) with config = x y z alpha beta gamma [one angle per rotatable bond] */
int mmo_molfile_rotated_copies(const mmo_molfile *f, int32_t k, int32_t n, const double *rot9, double *out_xs,
                               double *out_ys, double *out_zs);
int mmo_molfile_apply_config(const mmo_molfile *f, int32_t k, const double *config, int32_t n_config, double *out_xs,
                             double *out_ys, double *out_zs, int32_t *too_long);
int mmo_molfile_destroy(mmo_molfile *f);

#ifdef __cplusplus
}
#endif
#endif
