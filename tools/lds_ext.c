/* lds_ext.c -- `lds --ext` (exhaustive rigid ligand docking) as a plain C program on the C ABI of libmmo_b200.so.
 *
 * Host side of the drop-in written in compiled code, with no Python and no PyTorch in the process: what the OCaml
 * `lds` does around Lds.exhaustive_rigid_ligand_docking (src/lds.ml:1299-1352, 1685-2092), restated in C where
 * OCaml cannot be compiled (this image has no OCaml toolchain; the OCaml externs are in mmo_b200/ocaml/).
 *
 *   lds_ext -lig L.{mol2,pqrs} -rec R.pqrs -roi ROI.bild --ext <dx>,<n_rot> [-top k] [--no-prefilter] [--fp64] [-dev i]
 *
 *   preprocess_protein   src/lds.ml:20-41   receptor into a positive-octant box with a 36 A margin, ROI follows
 *   ligand               src/lds.ml:44-52   centred on the origin; E_intra constant for a rigid ligand (lds.ml:706-712)
 *   prefilter            src/lds.ml:1985    vdW_volume of the whole receptor on the 0.5 A simulation grid
 *   scorer               -ff BrL --no-interp: Mol.ene_inter_UFF_shifted_brute on the ROI receptor atoms
 *                        (atoms beyond R_roi + R_lig + 12 A have weight exactly 0, FF.ml:17-20)
 * Output: one line per kept pose "rank score frame", then the best pose (lds.ml:1110-1114).
 * Build: make -C mmo_b200/csrc lds_ext */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/mmo_b200.h"

#define CK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "lds_ext: %s failed (%d): %s\n", #call, rc__, mmo_last_error()); return 1; } } while (0)

static double favg(const double *a, int n) {        /* Batteries A.favg as restated in the oracle: Kahan sum / n */
    double s = 0.0, c = 0.0;
    for (int i = 0; i < n; i++) { double y = a[i] - c, t = s + y; c = (t - s) - y; s = t; }
    return s / n;
}

static int ends_with(const char *s, const char *suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}

int main(int argc, char **argv) {
    const char *lig_fn = NULL, *rec_fn = NULL, *roi_fn = NULL;
    double dx = 1.0;
    int n_rot = 1000, topk = 10, prefilter = 1, prec = MMO_PREC_FP32, dev = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-lig") && i + 1 < argc) lig_fn = argv[++i];
        else if (!strcmp(argv[i], "-rec") && i + 1 < argc) rec_fn = argv[++i];
        else if (!strcmp(argv[i], "-roi") && i + 1 < argc) roi_fn = argv[++i];
        else if (!strcmp(argv[i], "--ext") && i + 1 < argc) { if (sscanf(argv[++i], "%lf,%d", &dx, &n_rot) != 2) { fprintf(stderr, "--ext wants <dx>,<n_rot>\n"); return 2; } }
        else if (!strcmp(argv[i], "-top") && i + 1 < argc) topk = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-dev") && i + 1 < argc) dev = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--no-prefilter")) prefilter = 0;
        else if (!strcmp(argv[i], "--fp64")) prec = MMO_PREC_FP64;
        else { fprintf(stderr, "usage: %s -lig L.{mol2,pqrs} -rec R.pqrs -roi ROI.bild --ext <dx>,<n_rot> [-top k] [--no-prefilter] [--fp64] [-dev i]\n", argv[0]); return 2; }
    }
    if (!lig_fn || !rec_fn || !roi_fn) { fprintf(stderr, "lds_ext: -lig, -rec and -roi are mandatory\n"); return 2; }
    CK(mmo_init(dev));

    /* ---- inputs ---- */
    mmo_molfile *rf = NULL, *lf = NULL;
    CK(mmo_molfile_read_pqrs(rec_fn, 1, &rf));
    if (ends_with(lig_fn, ".mol2")) CK(mmo_molfile_read_mol2(lig_fn, &lf));
    else CK(mmo_molfile_read_pqrs(lig_fn, 0, &lf));
    int32_t n_lig_mols = 0, P = 0, L = 0;
    CK(mmo_molfile_count(lf, &n_lig_mols, NULL));
    if (n_lig_mols < 1) { fprintf(stderr, "lds_ext: no usable ligand in %s\n", lig_fn); return 1; }
    CK(mmo_molfile_shape(rf, 0, &P, NULL, NULL, NULL, 0));
    CK(mmo_molfile_shape(lf, 0, &L, NULL, NULL, NULL, 0));
    double roi[4];
    {
        FILE *f = fopen(roi_fn, "r");
        if (!f) { fprintf(stderr, "lds_ext: cannot open %s\n", roi_fn); return 1; }
        char line[512];
        int found = 0;
        while (fgets(line, sizeof line, f))                 /* ROI.from_bild, src/ROI.ml:22-32 */
            if (!strncmp(line, ".sphere ", 8) && sscanf(line + 8, "%lf %lf %lf %lf", roi, roi + 1, roi + 2, roi + 3) == 4) found++;
        fclose(f);
        if (found != 1) { fprintf(stderr, "lds_ext: %s must hold exactly one .sphere line\n", roi_fn); return 1; }
    }
    double *px = malloc(sizeof(double) * P * 5), *py = px + P, *pz = py + P, *pq = pz + P, *pr = pq + P;
    int32_t *pa = malloc(sizeof(int32_t) * P);
    CK(mmo_molfile_get(rf, 0, px, py, pz, pq, pr, pa, NULL, NULL, NULL, NULL, NULL, NULL));

    /* ---- preprocess_protein (lds.ml:20-41): box = bounding box of the vdW spheres + 2 x 36 A, centred protein ---- */
    double rmax = 0.0, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < P; i++) if (pr[i] > rmax) rmax = pr[i];
    for (int i = 0; i < P; i++) {
        const double c[3] = {px[i], py[i], pz[i]};
        for (int d = 0; d < 3; d++) { if (c[d] - rmax < lo[d]) lo[d] = c[d] - rmax; if (c[d] + rmax > hi[d]) hi[d] = c[d] + rmax; }
    }
    const double margin = 12.0 * 3.0;
    double sim[3], delta[3];
    const double old_c[3] = {favg(px, P), favg(py, P), favg(pz, P)};
    for (int d = 0; d < 3; d++) { sim[d] = (hi[d] - lo[d]) + 2.0 * margin; delta[d] = sim[d] * 0.5 - old_c[d]; }
    for (int i = 0; i < P; i++) { px[i] += delta[0]; py[i] += delta[1]; pz[i] += delta[2]; }
    for (int d = 0; d < 3; d++) roi[d] += delta[d];

    /* ---- ligand: centred handle, rigid-ligand E_intra (lds.ml:706-712, 1318-1319) ---- */
    mmo_ligand *lig = NULL;
    CK(mmo_molfile_ligand(lf, 0, 1, &lig));
    double *lx = malloc(sizeof(double) * L * 3), *ly = lx + L, *lz = ly + L;
    CK(mmo_molfile_get(lf, 0, lx, ly, lz, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    const double lc[3] = {favg(lx, L), favg(ly, L), favg(lz, L)};
    double lig_r = 0.0;
    for (int j = 0; j < L; j++) {
        lx[j] += 0.0 - lc[0]; ly[j] += 0.0 - lc[1]; lz[j] += 0.0 - lc[2];
        const double d = sqrt(lx[j] * lx[j] + ly[j] * ly[j] + lz[j] * lz[j]);
        if (d > lig_r) lig_r = d;
    }
    lig_r += 0.01;                                           /* Mol.radius, src/mol.ml:576-583 */
    double e_intra = 0.0;
    CK(mmo_intra_nb(lig, 1, lx, ly, lz, &e_intra));

    /* ---- ROI receptor for the direct scorer ---- */
    const double reach = roi[3] + lig_r + 12.0;
    int n_roi = 0;
    double *qx = malloc(sizeof(double) * P * 4), *qy = qx + P, *qz = qy + P, *qq = qz + P;
    int32_t *qa = malloc(sizeof(int32_t) * P);
    for (int i = 0; i < P; i++) {
        const double ddx = px[i] - roi[0], ddy = py[i] - roi[1], ddz = pz[i] - roi[2];
        if (ddx * ddx + ddy * ddy + ddz * ddz < reach * reach) { qx[n_roi] = px[i]; qy[n_roi] = py[i]; qz[n_roi] = pz[i]; qq[n_roi] = pq[i]; qa[n_roi] = pa[i]; n_roi++; }
    }
    mmo_receptor *rec = NULL;
    CK(mmo_receptor_create(n_roi, qx, qy, qz, qq, qa, &rec));

    /* ---- vdW prefilter on the 0.5 A simulation grid (params.ml:22, lds.ml:1985-1986) ---- */
    mmo_mask *mask = NULL;
    if (prefilter) {
        int32_t dims[3];
        CK(mmo_grid_from_box(0.5, sim[0], sim[1], sim[2], dims));
        CK(mmo_vdw_mask_build(P, px, py, pz, pr, 0.5, dims, NULL, &mask));
    }

    /* ---- the scan ---- */
    double *rot = malloc(sizeof(double) * 9 * (size_t)n_rot);
    CK(mmo_so3_rotations(n_rot, rot));
    mmo_scan_params sp;
    memset(&sp, 0, sizeof sp);
    sp.rec = rec; sp.grid = NULL; sp.lig = lig; sp.vdw_mask = mask;
    sp.variant = MMO_VARIANT_SHIFTED; sp.prec = prec;
    for (int d = 0; d < 3; d++) sp.roi_c[d] = roi[d];
    sp.roi_r = roi[3]; sp.trans_step = dx; sp.n_rot = n_rot; sp.rot9 = rot;
    sp.e_intra_const = e_intra; sp.topk = topk; sp.first_point = 0; sp.n_points = -1;
    double *ts = malloc(sizeof(double) * (topk > 0 ? topk : 1));
    int64_t *tf = malloc(sizeof(int64_t) * (topk > 0 ? topk : 1));
    mmo_scan_result res;
    CK(mmo_scan(&sp, ts, tf, &res));

    printf("# receptor atoms %d (ROI %d), ligand atoms %d, lattice %dx%dx%d, candidates %lld, scored %lld, device %.2f ms\n",
           (int)P, n_roi, (int)L, res.lattice_dims[0], res.lattice_dims[1], res.lattice_dims[2],
           (long long)res.n_candidates, (long long)res.n_scored, res.device_ms);
    for (int k = 0; k < res.n_top; k++) printf("%d\t%.17g\t%lld\n", k, ts[k], (long long)tf[k]);
    printf("best\t%.17g\t%lld\trot %d\tpos %.17g %.17g %.17g\n", res.best_score, (long long)res.best_frame, res.best_rot_i,
           res.best_pos[0], res.best_pos[1], res.best_pos[2]);
    if (mask) mmo_mask_destroy(mask);
    mmo_receptor_destroy(rec); mmo_ligand_destroy(lig); mmo_molfile_destroy(rf); mmo_molfile_destroy(lf);
    mmo_shutdown();
    return 0;
}
