/* place_ligand.c -- the reference's `place_ligand input.mol2 output.mol2 x y z a b g [rbonds]`
 * (src/place_ligand.ml:9-62) as a plain C program on the C ABI of libmmo_b200.so: same argv, same usage text,
 * same fatal cases, host only (no GPU: nothing here calls mmo_init).
 *
 *   centred ligand   Mol.center x (place_ligand.ml:36-41)                 \
 *   transform        Optim.apply_config centered_lig cfg (optim.ml:64-80)  } mmo_molfile_apply_config
 *   output           Mol.update_mol2 + Mol2.write_one_to_file               = mmo_molfile_write_mol2
 * (a, b, g) are the Cartesian angles of Rot.r_xyz (rot.ml:52-66); without rbond values only the rigid-body move is
 * applied; a wrong number of rbond values is fatal (place_ligand.ml:49-55).
 * Build: make -C mmo_b200/csrc place_ligand */
#include <stdio.h>
#include <stdlib.h>
#include "../include/mmo_b200.h"

#define CK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "place_ligand: %s failed (%d): %s\n", #call, rc__, mmo_last_error()); return 1; } } while (0)

static int to_double(const char *s, double *v) {       /* float_of_string raises Failure */
    char *end = NULL;
    *v = strtod(s, &end);
    return end != s && *end == 0;
}

int main(int argc, char **argv) {
    if (argc == 1) {
        fprintf(stderr, "usage:\n%s input.mol2 output.mol2 x y z a b g [rbonds]\n", argv[0]);
        return 1;
    }
    if (argc < 9) { fprintf(stderr, "place_ligand: x y z a b g are mandatory\n"); return 2; }   /* Invalid_argument "index out of bounds" */
    const int n_config = argc - 3;                       /* 6 rigid-body DOFs + the rbond angles given */
    const int rbonds_given = argc - 9;
    double *config = (double *)malloc(sizeof(double) * (size_t)n_config);
    if (!config) return 1;
    for (int i = 0; i < n_config; i++)
        if (!to_double(argv[3 + i], &config[i])) { fprintf(stderr, "place_ligand: not a number: %s\n", argv[3 + i]); return 2; }
    fprintf(stderr, "rbonds given: %d\n", rbonds_given);
    mmo_molfile *f = NULL;
    CK(mmo_molfile_read_mol2(argv[1], &f));
    int32_t n_mols = 0, n_atoms = 0, n_rbonds = 0;
    CK(mmo_molfile_count(f, &n_mols, NULL));
    if (n_mols < 1) { fprintf(stderr, "place_ligand: cannot parse %s\n", argv[1]); return 2; }
    /* Mol2.read_one: molecules after the first are ignored; mol2pqrs would list them all and the reference then
     * fails with "several ligands": keep that */
    if (n_mols > 1) { fprintf(stderr, "place_ligand: several ligands in %s\n", argv[1]); return 2; }
    CK(mmo_molfile_shape(f, 0, &n_atoms, &n_rbonds, NULL, NULL, 0));
    if (rbonds_given > 0 && rbonds_given != n_rbonds) {
        fprintf(stderr, "place_ligand: %d rbond values but mol has %d\n", rbonds_given, n_rbonds);
        return 1;
    }
    double *xs = (double *)malloc(sizeof(double) * 3 * (size_t)n_atoms);
    if (!xs) return 1;
    double *ys = xs + n_atoms, *zs = ys + n_atoms;
    int32_t too_long = 0;
    CK(mmo_molfile_apply_config(f, 0, config, n_config, xs, ys, zs, &too_long));
    if (too_long) { fprintf(stderr, "place_ligand: Mol.Too_long\n"); return 2; }    /* uncaught exception in the reference */
    CK(mmo_molfile_write_mol2(f, 0, 1, xs, ys, zs, argv[2], 0));
    free(xs);
    free(config);
    CK(mmo_molfile_destroy(f));
    return 0;
}
