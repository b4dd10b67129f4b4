/* lds_mc.c -- `lds` in its default mode (protein-ligand Monte Carlo, Lds.simulate_lig) as a plain C program on the C ABI
 * of libmmo_b200.so: the reference's flag spellings (src/lds.ml:1691-1734), every (ligand, start) pair one GPU chain.
 *
 *   lds_mc -lig L.{mol2,pqrs} -rec R.pqrs -roi ROI.bild -steps <int>[k|M] {--intra-NB | --no-E-intra}
 *          [-s seed] [-starts n] [-T kelvin] [--hard-ROI] [--no-flip] [--rigid-ligand] [--no-interp [-ff BrL|Bst]] [--less-charges] [-dev i]
 *
 *   preprocess_protein    src/lds.ml:20-41     receptor into a positive-octant box with a 36 A margin, ROI follows
 *   steps_of_string       src/lds.ml:558-570   10k, 2M
 *   E_intra flags         src/lds.ml:1759-1771 exactly one of them; the rdkit / torchani ones are outside the scope
 *   default scorer        src/lds.ml:1952-1979 interpolated maps on the 0.5 A grid behind the ROI-only bitmask
 *                         (bitmask_ROI_only, pre_calculate_FF_components_grid) = mmo_mask_roi_only + mmo_grid_build;
 *                         --no-interp: the direct shifted pair sum (-ff BrL / Bst; BrG has no cut-off: not in the kernel)
 *   start poses           src/lds.ml:308-345   place_ligand_in_ROI = mmo_place_ligand_in_roi.  The heavy-atom clash rejection
 *                         is OFF by default: in the reference Ptable.vdW_max is nan (A.max over an array holding nan,
 *                         src/ptable.ml:32,84), BST.neighbors gets a nan radius and nothing is ever rejected -- which is
 *                         also the only way `lds` can start on its own example (a 48-atom ligand cannot be dropped
 *                         clash-free into the buried 3A2J pocket within 100 000 rigid draws).  --clash-check (not an
 *                         lds flag) turns the rejection the source text intends on.
 *   chains                src/lds.ml:882-995   simulate_lig frame loop = mmo_mc_run, one RNG per (ligand, start)
 *                         (lds.ml:1997-2000)
 * Output: one line per chain (ligand, start, best E, counters), then the reference's own rate line
 * "%d frames in %.2f (s) @ %.2f (Hz)" (lds.ml:2086-2087).  File outputs (*_E.txt, .bild, docking_scores.tsv) stay with
 * the OCaml driver.  Random numbers are those of include/mmo_detmath.h, not OCaml's Random.State.
 * Build: make -C mmo_b200/csrc lds_mc */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/mmo_b200.h"

#define CK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "lds_mc: %s failed (%d): %s\n", #call, rc__, mmo_last_error()); return 1; } } while (0)

static double favg(const double *a, int n) {        /* Batteries A.favg as restated in the oracle: Kahan sum / n */
    double s = 0.0, c = 0.0;
    for (int i = 0; i < n; i++) { double y = a[i] - c, t = s + y; c = (t - s) - y; s = t; }
    return s / n;
}
static int ends_with(const char *s, const char *suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}
static long steps_of_string(const char *s) {        /* src/lds.ml:558-570 */
    char *end = NULL;
    long v = strtol(s, &end, 10);
    if (end == s) return -1;
    if (*end == 0) return v;
    if (end[1] != 0) return -1;
    if (*end == 'k') return 1000L * v;
    if (*end == 'M') return 1000000L * v;
    fprintf(stderr, "Lds.steps_of_string: unsupported suffix: %c\n", *end);
    return -1;
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
    const char *lig_fn = NULL, *rec_fn = NULL, *roi_fn = NULL, *ff = "Bst";
    long n_steps = -1;
    long long seed = 0;
    int have_seed = 0, starts = 1, hard_roi = 0, no_flip = 0, tweak_rbonds = 1, intra_nb = 0, no_e_intra = 0, no_interp = 0, dev = 0, clash_check = 0, less_charges = 0;
    double temp = 293.15;                                       /* Const.room_temp_K */
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-lig") && i + 1 < argc) lig_fn = argv[++i];
        else if (!strcmp(argv[i], "-rec") && i + 1 < argc) rec_fn = argv[++i];
        else if (!strcmp(argv[i], "-roi") && i + 1 < argc) roi_fn = argv[++i];
        else if (!strcmp(argv[i], "-steps") && i + 1 < argc) n_steps = steps_of_string(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) { seed = atoll(argv[++i]); have_seed = 1; }
        else if (!strcmp(argv[i], "-starts") && i + 1 < argc) starts = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-T") && i + 1 < argc) temp = atof(argv[++i]);
        else if (!strcmp(argv[i], "-ff") && i + 1 < argc) ff = argv[++i];
        else if (!strcmp(argv[i], "-dev") && i + 1 < argc) dev = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--hard-ROI")) hard_roi = 1;
        else if (!strcmp(argv[i], "--no-flip")) no_flip = 1;
        else if (!strcmp(argv[i], "--rigid-ligand")) tweak_rbonds = 0;
        else if (!strcmp(argv[i], "--intra-NB")) intra_nb = 1;
        else if (!strcmp(argv[i], "--no-E-intra")) no_e_intra = 1;
        else if (!strcmp(argv[i], "--no-interp")) no_interp = 1;
        else if (!strcmp(argv[i], "--clash-check")) clash_check = 1;
        else if (!strcmp(argv[i], "--less-charges")) less_charges = 1;
        else if (!strcmp(argv[i], "--intra-QM") || !strcmp(argv[i], "--intra-UFF") || !strcmp(argv[i], "--intra-MMFF") ||
                 !strcmp(argv[i], "--no-vdW-clash")) {
            fprintf(stderr, "lds_mc: %s needs rdkit / torchani: outside this build (use --intra-NB or --no-E-intra)\n", argv[i]);
            return 2;
        } else {
            fprintf(stderr, "usage:\n  %s\n  -lig <FILE.mol2>: ligand\n  -rec <FILE.pqrs>: receptor protein\n  [-s <int>]: random seed\n"
                            "  [-steps <int>[k|M]]: maximum number of frames\n  [-starts <int>]: how many ligand starting positions\n"
                            "  [-ff]: FF eval. style: {BrL|Bst*} (default=*)\n  [-T <float>]: temperature in Kelvin (default=%g)\n"
                            "  [--no-interp]: turn OFF FF interpolation (slow)\n  [--rigid-ligand]: freeze ligand's rotatable bonds\n"
                            "  [--no-E-intra]: ignore lig_E_intra\n  [--intra-NB]: lig_E_intra is UFF Non-Bonded interactions only\n"
                            "  [--no-flip]: disable rbond flip move in MC simulations\n  [--hard-ROI]: enforce ROI during simulation (default=false)\n"
                            "  -roi <FILE.bild>: ROI sphere in original PDB coordinates\n  [-dev <int>]: CUDA device\n", argv[0], 293.15);
            return 1;
        }
    }
    if (!lig_fn || !rec_fn || !roi_fn || n_steps < 0) { fprintf(stderr, "lds_mc: -lig, -rec, -roi and -steps are mandatory\n"); return 2; }
    if (intra_nb + no_e_intra != 1) { fprintf(stderr, "Lds.main: which lig_E_intra FF to use?\n"); return 2; }     /* lds.ml:1771 */
    if (strcmp(ff, "Bst") && strcmp(ff, "BrL")) { fprintf(stderr, "lds_mc: -ff %s: only the shifted variants (BrL, Bst) run in the chain kernel\n", ff); return 2; }
    if (starts < 1 || n_steps > 2000000000L) { fprintf(stderr, "lds_mc: bad -starts / -steps\n"); return 2; }
    if (!have_seed) seed = (long long)time(NULL);              /* RNG.entropy_160b: any fresh seed */
    CK(mmo_init(dev));

    /* ---- inputs ---- */
    mmo_molfile *rf = NULL, *lf = NULL;
    CK(mmo_molfile_read_pqrs(rec_fn, 1, &rf));
    if (ends_with(lig_fn, ".mol2")) CK(mmo_molfile_read_mol2(lig_fn, &lf));
    else CK(mmo_molfile_read_pqrs(lig_fn, 0, &lf));
    if (less_charges) CK(mmo_molfile_reduce_charges(lf));     /* lds.ml:1887-1894, before the FF types are used */
    int32_t n_ligs = 0, P = 0;
    CK(mmo_molfile_count(lf, &n_ligs, NULL));
    if (n_ligs < 1) { fprintf(stderr, "lds_mc: no usable ligand in %s\n", lig_fn); return 1; }
    CK(mmo_molfile_shape(rf, 0, &P, NULL, NULL, NULL, 0));
    double roi[4];
    {
        FILE *f = fopen(roi_fn, "r");
        if (!f) { fprintf(stderr, "lds_mc: cannot open %s\n", roi_fn); return 1; }
        char line[512];
        int found = 0;
        while (fgets(line, sizeof line, f))                 /* ROI.from_bild, src/ROI.ml:22-32 */
            if (!strncmp(line, ".sphere ", 8) && sscanf(line + 8, "%lf %lf %lf %lf", roi, roi + 1, roi + 2, roi + 3) == 4) found++;
        fclose(f);
        if (found != 1) { fprintf(stderr, "lds_mc: %s must hold exactly one .sphere line\n", roi_fn); return 1; }
    }
    double *px = malloc(sizeof(double) * P * 5), *py = px + P, *pz = py + P, *pq = pz + P, *pr = pq + P;
    int32_t *pa = malloc(sizeof(int32_t) * P);
    CK(mmo_molfile_get(rf, 0, px, py, pz, pq, pr, pa, NULL, NULL, NULL, NULL, NULL, NULL));

    /* ---- preprocess_protein (lds.ml:20-41) ---- */
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

    /* ---- scorer: interpolated maps (default) or the direct pair sum (--no-interp) ---- */
    mmo_receptor *rec = NULL;
    mmo_grid *grid = NULL;
    CK(mmo_receptor_create(P, px, py, pz, pq, pa, &rec));
    const double t_grid0 = now_s();
    if (!no_interp) {
        int32_t dims[3], T = 0;
        CK(mmo_grid_from_box(0.5, sim[0], sim[1], sim[2], dims));                       /* Params.grid_step, params.ml:22 */
        const size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
        uint8_t *bits = malloc((nvox + 7) / 8 + 8);
        CK(mmo_mask_roi_only(roi, roi[3], 0.5, dims, bits, NULL));                      /* ROI_only roi, lds.ml:269-305 */
        CK(mmo_molfile_types(lf, &T, NULL, NULL));
        int32_t *ta = malloc(sizeof(int32_t) * T);
        double *tq = malloc(sizeof(double) * T);
        CK(mmo_molfile_types(lf, &T, ta, tq));
        CK(mmo_grid_build(rec, 0.5, dims, bits, T, ta, tq, NULL, &grid));
        CK(mmo_sync());
        fprintf(stderr, "lds_mc: %d FF maps of %dx%dx%d voxels in %.2f s\n", (int)T, dims[0], dims[1], dims[2], now_s() - t_grid0);
        free(bits); free(ta); free(tq);
    }

    /* ---- chains: every (ligand, start) pair ---- */
    mmo_mc_params mp;
    memset(&mp, 0, sizeof mp);
    for (int d = 0; d < 3; d++) mp.roi_c[d] = roi[d];
    mp.roi_r = roi[3]; mp.temperature_K = temp; mp.n_steps = (int32_t)n_steps;
    mp.tweak_rbonds = tweak_rbonds; mp.hard_roi = hard_roi; mp.no_flip = no_flip; mp.intra_nb = intra_nb;
    double *rot9 = malloc(sizeof(double) * 12 * (size_t)starts), *pos3 = rot9 + 9 * (size_t)starts;
    uint64_t *seeds = malloc(sizeof(uint64_t) * (size_t)starts);
    mmo_mc_result *res = malloc(sizeof(mmo_mc_result) * (size_t)starts);
    const double t0 = now_s();
    long long frames = 0;
    printf("#ligand\tstart\tbest_E\tframes\tacc_rigid\trej_rigid\tacc_conf\trej_conf\tout_of_ROI\ttoo_long\n");
    for (int32_t k = 0; k < n_ligs; k++) {
        mmo_ligand *lig = NULL;
        char name[256];
        int32_t n_rb = 0;
        CK(mmo_molfile_shape(lf, k, NULL, &n_rb, NULL, name, (int32_t)sizeof name));
        CK(mmo_molfile_ligand(lf, k, 1, &lig));                                          /* preprocess_ligand: centred */
        int32_t trials = 0;
        CK(mmo_place_ligand_in_roi(P, px, py, pz, pa, lig, roi, roi[3], (uint64_t)seed + (uint64_t)k * 1000003ull, starts, clash_check,
                                   rot9, pos3, &trials));
        fprintf(stderr, "Lds.place_ligands_in_ROI: %d w/ %d trials (%s: %d rbonds)\n", starts, (int)trials, name, (int)n_rb);
        for (int s = 0; s < starts; s++) seeds[s] = (uint64_t)seed + 1ull + (uint64_t)k * (uint64_t)starts + (uint64_t)s;
        CK(mmo_mc_run(no_interp ? rec : NULL, no_interp ? NULL : grid, lig, &mp, starts, seeds, rot9, pos3, res, NULL, NULL));
        for (int s = 0; s < starts; s++) {
            printf("%s\t%d\t%.17g\t%d\t%lld\t%lld\t%lld\t%lld\t%lld\t%d\n", name, s, res[s].best_E, (int)res[s].frames_done,
                   (long long)res[s].n_accept_rigid, (long long)res[s].n_reject_rigid, (long long)res[s].n_accept_conf,
                   (long long)res[s].n_reject_conf, (long long)res[s].n_ooroi, (int)res[s].too_long);
            frames += res[s].frames_done;
        }
        CK(mmo_ligand_destroy(lig));
    }
    const double el = now_s() - t0;
    printf("%ld frames in %.2f (s) @ %.2f (Hz)\n", n_steps, el, (double)frames / el);  /* lds.ml:2086-2087, over all chains */
    if (grid) mmo_grid_destroy(grid);
    mmo_receptor_destroy(rec); mmo_molfile_destroy(rf); mmo_molfile_destroy(lf);
    mmo_shutdown();
    return 0;
}
