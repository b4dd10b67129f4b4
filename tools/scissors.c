/* scissors.c -- the reference's `scissors -l ligand.mol2 -p protein.mol2 [-d cutoff] -o site.pqrs [-v]`
 * (src/scissors.ml:24-66) as a plain C program on the C ABI of libmmo_b200.so: same flags, same usage text, same output
 * (one "%g %g %g %g %g %s" line per kept protein atom, Mol2.pqrs_line_of_atom, src/mol2.ml:67-69; no header line).
 *
 *   inputs   Mol2.read_one_from_file on both files              = mmo_molfile_read_mol2_atoms
 *   carve    nearest ligand atom within the cut-off (default 5 A) = mmo_carve_near_ligand (one CUDA launch)
 * Build: make -C mmo_b200/csrc scissors */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/mmo_b200.h"

#define CK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "scissors: %s failed (%d): %s\n", #call, rc__, mmo_last_error()); return 1; } } while (0)

static const char *symbol(int anum) {          /* src/ptable.ml:86-100 */
    switch (anum) {
    case 1: return "H"; case 6: return "C"; case 7: return "N"; case 8: return "O"; case 9: return "F"; case 12: return "Mg";
    case 15: return "P"; case 16: return "S"; case 17: return "Cl"; case 35: return "Br"; case 53: return "I";
    default: return "X";
    }
}

int main(int argc, char **argv) {
    const double default_cutoff = 5.0;          /* (Angstrom) around ligand's heavy atoms */
    const char *lig_fn = NULL, *prot_fn = NULL, *out_fn = NULL;
    double cutoff = default_cutoff;
    int dev = 0;
    if (argc == 1) {
        fprintf(stderr, "usage:\n  %s\n  -l <ligand.mol2>: xtal ligand input file\n  -p <protein.mol2>: receptor protein input file\n"
                        "  [-d <float>]: distance cutoff (default=%.2f)\n  -o <output.pqrs>: ligand-defined binding site output file\n"
                        "  [-v]: verbose/debug mode\n", argv[0], default_cutoff);
        return 1;
    }
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-l") && i + 1 < argc) lig_fn = argv[++i];
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) prot_fn = argv[++i];
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out_fn = argv[++i];
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) cutoff = atof(argv[++i]);
        else if (!strcmp(argv[i], "-dev") && i + 1 < argc) dev = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-v")) {}
        else { fprintf(stderr, "scissors: unknown option %s\n", argv[i]); return 2; }       /* CLI.finalize */
    }
    if (!lig_fn || !prot_fn || !out_fn) { fprintf(stderr, "scissors: -l, -p and -o are mandatory\n"); return 2; }
    mmo_molfile *lf = NULL, *pf = NULL;
    CK(mmo_molfile_read_mol2_atoms(lig_fn, &lf));
    CK(mmo_molfile_read_mol2_atoms(prot_fn, &pf));
    int32_t nl = 0, np = 0, k = 0;
    CK(mmo_molfile_count(lf, &k, NULL));
    if (k < 1) { fprintf(stderr, "Mol2.read_one_from_file: could not read %s\n", lig_fn); return 1; }
    CK(mmo_molfile_count(pf, &k, NULL));
    if (k < 1) { fprintf(stderr, "Mol2.read_one_from_file: could not read %s\n", prot_fn); return 1; }
    CK(mmo_molfile_shape(lf, 0, &nl, NULL, NULL, NULL, 0));
    CK(mmo_molfile_shape(pf, 0, &np, NULL, NULL, NULL, 0));
    double *l = malloc(sizeof(double) * 3 * (size_t)nl), *p = malloc(sizeof(double) * 5 * (size_t)np);
    int32_t *pa = malloc(sizeof(int32_t) * (size_t)np);
    uint8_t *keep = malloc((size_t)np);
    if (!l || !p || !pa || !keep) return 1;
    CK(mmo_molfile_get(lf, 0, l, l + nl, l + 2 * nl, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    CK(mmo_molfile_get(pf, 0, p, p + np, p + 2 * np, p + 3 * np, p + 4 * np, pa, NULL, NULL, NULL, NULL, NULL, NULL));
    CK(mmo_init(dev));
    int32_t n_kept = 0;
    CK(mmo_carve_near_ligand(np, p, p + np, p + 2 * np, nl, l, l + nl, l + 2 * nl, cutoff, keep, &n_kept));
    FILE *o = fopen(out_fn, "w");
    if (!o) { fprintf(stderr, "scissors: cannot create %s\n", out_fn); return 1; }
    for (int i = 0; i < np; i++)
        if (keep[i]) fprintf(o, "%g %g %g %g %g %s\n", p[i], p[np + i], p[2 * np + i], p[3 * np + i], p[4 * np + i], symbol(pa[i]));
    fclose(o);
    fprintf(stderr, "scissors: %d of %d protein atoms within %g A of the ligand\n", (int)n_kept, (int)np, cutoff);
    mmo_molfile_destroy(lf); mmo_molfile_destroy(pf);
    mmo_shutdown();
    return 0;
}
