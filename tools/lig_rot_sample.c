/* lig_rot_sample.c -- the reference's `lig_rot_sample num_samples input.mol2 output.mol2` (src/lig_rot_sample.ml:11-45)
 * as a plain C program on the C ABI of libmmo_b200.so: same argv, same usage text and exit code, host only (no GPU:
 * nothing here calls mmo_init).
 *
 *   rotations  SO3.sample n -> Quat.to_axis_angle -> Rot.of_axis_angle   = mmo_so3_rotations (SO3.ml:13-39)
 *   copies     Mol.center_rotate_translate_copy mol rot orig_center       = mmo_molfile_rotated_copies (mol.ml:705-710)
 *   output     Mol.update_mol2 + Mol2.output_one, one block per rotation  = mmo_molfile_write_mol2 (mol2.ml:326-343)
 * The mol2pqrs subprocess of the reference (Utls.mol2pqrs, lig_rot_sample.ml:31) is the library's own reader.
 * Build: make -C mmo_b200/csrc lig_rot_sample */
#include <stdio.h>
#include <stdlib.h>
#include "../include/mmo_b200.h"

#define CK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "lig_rot_sample: %s failed (%d): %s\n", #call, rc__, mmo_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "usage:\n%s num_samples input.mol2 output.mol2\n", argv[0]);
        return 1;
    }
    char *end = NULL;
    const long n = strtol(argv[1], &end, 10);
    if (end == argv[1] || *end != 0 || n < 0 || n > 100000000L) {      /* int_of_string raises Failure */
        fprintf(stderr, "lig_rot_sample: bad num_samples: %s\n", argv[1]);
        return 2;
    }
    mmo_molfile *f = NULL;
    CK(mmo_molfile_read_mol2(argv[2], &f));
    int32_t n_mols = 0, n_atoms = 0;
    CK(mmo_molfile_count(f, &n_mols, NULL));
    if (n_mols < 1) { fprintf(stderr, "lig_rot: no ligand in %s\n", argv[2]); return 2; }
    if (n_mols > 1) { fprintf(stderr, "lig_rot: several ligands in %s\n", argv[2]); return 2; }
    CK(mmo_molfile_shape(f, 0, &n_atoms, NULL, NULL, NULL, 0));
    /* in slices, so that a million rotations of a large ligand do not need gigabytes of host memory */
    const long slice = 4096;
    double *rot9 = (double *)malloc(sizeof(double) * 9 * (size_t)(n > 0 ? n : 1));
    double *xs = (double *)malloc(sizeof(double) * 3 * (size_t)slice * (size_t)n_atoms);
    if (!rot9 || !xs) { fprintf(stderr, "lig_rot_sample: out of memory\n"); return 1; }
    double *ys = xs + (size_t)slice * n_atoms, *zs = ys + (size_t)slice * n_atoms;
    CK(mmo_so3_rotations((int32_t)n, rot9));
    CK(mmo_molfile_write_mol2(f, 0, 0, NULL, NULL, NULL, argv[3], 0));          /* create / truncate */
    for (long r0 = 0; r0 < n; r0 += slice) {
        const int32_t m = (int32_t)(n - r0 < slice ? n - r0 : slice);
        CK(mmo_molfile_rotated_copies(f, 0, m, rot9 + 9 * (size_t)r0, xs, ys, zs));
        CK(mmo_molfile_write_mol2(f, 0, m, xs, ys, zs, argv[3], 1));
    }
    free(xs);
    free(rot9);
    CK(mmo_molfile_destroy(f));
    return 0;
}
