// host_math.cu -- host-side (libm, IEEE double) mirrors of the reference's pose-feed modules.
// OCaml's sin/cos/atan2/sqrt are the C library's, so these produce the same bits as the reference
// on the same platform.  Compiled with -ffp-contract=off (OCaml never fuses a*b+c).
//   SO3.ml:13-39, quat.ml:32-36, rot.ml:52-75,121-146, grid.ml:37-52
#include "common.cuh"
#include "../../include/mmo_detmath.h"
#include <math.h>
#include <string.h>
#include <algorithm>

namespace mmo {

static double pi_() { return 4.0 * atan(1.0); }   // math.ml:13

// rot.ml:136-146 (the live body of of_axis_angle)
static void rot_of_axis_angle(double x, double y, double z, double theta, double *r) {
    double c = cos(theta), s = sin(theta);
    double omc = 1.0 - c;
    r[0] = c + x * x * omc;
    r[1] = x * y * omc - z * s;
    r[2] = x * z * omc + y * s;
    r[3] = x * y * omc + z * s;
    r[4] = c + y * y * omc;
    r[5] = y * z * omc - x * s;
    r[6] = x * z * omc - y * s;
    r[7] = y * z * omc + x * s;
    r[8] = c + z * z * omc;
}

// SO3.ml:18-39: super-Fibonacci quaternion (w,x,y,z) -> Quat.to_axis_angle -> Rot.of_axis_angle
void so3_rotations(int n, double *rot9) {
    const double phi = sqrt(2.0);
    const double psi = 1.533751168755204288118041;
    const double nf = (double)n;
    const double two_pi = 2.0 * pi_();
    for (int i = 0; i < n; i++) {
        double s = (double)i + 0.5;
        double t = s / nf;
        double d = two_pi * s;
        double c_r = sqrt(t);
        double c_R = sqrt(1.0 - t);
        double alpha = d / phi;
        double beta = d / psi;
        double w = c_r * sin(alpha), x = c_r * cos(alpha), y = c_R * sin(beta), z = c_R * cos(beta);
        double mag = sqrt(x * x + y * y + z * z);
        double theta = 2.0 * atan2(mag, w);
        rot_of_axis_angle(x / mag, y / mag, z / mag, theta, rot9 + 9 * (size_t)i);
    }
}

// rot.ml:52-66
void rot_r_xyz(double al, double be, double ga, double r[9]) {
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

// rot.ml:71-75
void rot_decompose(const double r[9], double abg[3]) {
    double beta = atan2(-r[6], sqrt(r[0] * r[0] + r[3] * r[3]));
    double cb = cos(beta);
    abg[0] = atan2(r[7] / cb, r[8] / cb);
    abg[1] = beta;
    abg[2] = atan2(r[3] / cb, r[0] / cb);
}

// grid.ml:37-38
int grid_num_steps(double dx, double length) { return (int)ceil(length / dx); }

// grid.ml:49-51: xs = frange 0 `To (step*n') (n'+1); node i = i * ((step*n')/n')
double grid_node(double step, int dim, int i) {
    int np = dim - 1;
    if (np <= 0) return 0.0;
    double span = step * (double)np;
    return (double)i * (span / (double)np);
}

// rot.ml:97-100
static void rot_apply(const double *r, double x, double y, double z, double *o) {
    o[0] = r[0] * x + r[1] * y + r[2] * z;
    o[1] = r[3] * x + r[4] * y + r[5] * z;
    o[2] = r[6] * x + r[7] * y + r[8] * z;
}

// Batteries A.favg restated as in the oracle (Kahan-compensated sum / n; library not vendored)
static double favg(const std::vector<double> &a) {
    double sum = 0.0, c = 0.0;
    for (double v : a) { double y = v - c; double t = sum + y; c = (t - sum) - y; sum = t; }
    return sum / (double)a.size();
}

}  // namespace mmo

namespace mmo {

double favg_host(const double *a, int n) {
    double sum = 0.0, c = 0.0;
    for (int i = 0; i < n; i++) { double y = a[i] - c; double t = sum + y; c = (t - sum) - y; sum = t; }
    return sum / (double)n;
}

// Optim.apply_config centered_lig conf (src/optim.ml:64-80), the body of the place_ligand tool
// (src/place_ligand.ml:55-59): rotate every rotatable bond by conf[6+b] (Mol.rotate_bond, src/mol.ml:610-631),
// check the elongation (Mol.check_elongation_exn lig 12.0, src/mol.ml:576-591), then
// Mol.rotate_then_translate_copy lig (Rot.r_xyz a b g) (x, y, z) (src/mol.ml:669-672).  Host code, libm.
int apply_config_host(const HostLig &lig, const double *config, int32_t n_config, double *out_xs, double *out_ys,
                      double *out_zs, int32_t *too_long) {
    MMO_REQUIRE(n_config == 6 || n_config == 6 + lig.n_rbonds, "mmo_apply_config: %d values for a ligand with %d rotatable bonds",
                n_config, lig.n_rbonds);
    const int L = lig.n;
    std::vector<double> x(lig.x, lig.x + L), y(lig.y, lig.y + L), z(lig.z, lig.z + L);
    double cen[3] = {0.0, 0.0, 0.0};       // centered_lig.center
    for (int b = 0; b + 6 < n_config; b++) {
        const int left = lig.rb_left[b], right = lig.rb_right[b];
        const double cx = x[right], cy = y[right], cz = z[right];
        const double ax = cx - x[left], ay = cy - y[left], az = cz - z[left];
        const double mag = sqrt(ax * ax + ay * ay + az * az);
        double rot[9];
        rot_of_axis_angle(ax / mag, ay / mag, az / mag, config[6 + b], rot);
        for (int g = lig.rg_off[b]; g < lig.rg_off[b + 1]; g++) {
            const int i = lig.rg_idx[g];
            double o[3];
            rot_apply(rot, x[i] - cx, y[i] - cy, z[i] - cz, o);
            x[i] = o[0] + cx; y[i] = o[1] + cy; z[i] = o[2] + cz;
        }
        cen[0] = favg(x); cen[1] = favg(y); cen[2] = favg(z);     // update_center
    }
    double maxi = 0.0;
    for (int i = 0; i < L; i++) {
        double dx = cen[0] - x[i], dy = cen[1] - y[i], dz = cen[2] - z[i];
        maxi = std::max(maxi, 0.01 + sqrt(dx * dx + dy * dy + dz * dz));
    }
    if (too_long) *too_long = maxi > 12.0 ? 1 : 0;                                 // Mol.Too_long
    double rot[9];
    rot_r_xyz(config[3], config[4], config[5], rot);
    for (int i = 0; i < L; i++) {
        double o[3];
        rot_apply(rot, x[i], y[i], z[i], o);
        out_xs[i] = o[0] + config[0]; out_ys[i] = o[1] + config[1]; out_zs[i] = o[2] + config[2];
    }
    return MMO_OK;
}

void rotated_copies_host(const HostLig &lig, const double center[3], int32_t n, const double *rot9,
                         double *out_xs, double *out_ys, double *out_zs) {
    const int L = lig.n;
    const double nx = -center[0], ny = -center[1], nz = -center[2];      // Mol.center: translate_by (V3.neg mean)
    for (int r = 0; r < n; r++)
        for (int i = 0; i < L; i++) {
            double o[3];
            rot_apply(rot9 + 9 * (size_t)r, lig.x[i] + nx, lig.y[i] + ny, lig.z[i] + nz, o);
            out_xs[(size_t)r * L + i] = o[0] + center[0];
            out_ys[(size_t)r * L + i] = o[1] + center[1];
            out_zs[(size_t)r * L + i] = o[2] + center[2];
        }
}

}  // namespace mmo

extern "C" {

int mmo_apply_config(const mmo_ligand *lig, const double *config, int32_t n_config, double *out_xs, double *out_ys,
                     double *out_zs, int32_t *too_long) try {
    MMO_REQUIRE(lig && config && out_xs && out_ys && out_zs, "mmo_apply_config: null argument");
    const mmo::HostLig h = {lig->n, lig->hx.data(), lig->hy.data(), lig->hz.data(), lig->n_rbonds, lig->rb_left.data(),
                            lig->rb_right.data(), lig->rg_off.data(), lig->rg_idx.data()};
    return mmo::apply_config_host(h, config, n_config, out_xs, out_ys, out_zs, too_long);
} MMO_CATCH_ALL

// lig_rot_sample (src/lig_rot_sample.ml:23-45): n SO3-rotated copies of the ligand about its own centre:
// Mol.center_rotate_translate_copy mol rot orig_center (src/mol.ml:705-710).  `center` = Mol.get_center mol.
int mmo_rotated_copies(const mmo_ligand *lig, const double center[3], int32_t n, const double *rot9,
                       double *out_xs, double *out_ys, double *out_zs) try {
    MMO_REQUIRE(lig && center && n >= 0 && (n == 0 || (rot9 && out_xs && out_ys && out_zs)), "mmo_rotated_copies: bad arguments");
    const mmo::HostLig h = {lig->n, lig->hx.data(), lig->hy.data(), lig->hz.data(), 0, nullptr, nullptr, nullptr, nullptr};
    mmo::rotated_copies_host(h, center, n, rot9, out_xs, out_ys, out_zs);
    return MMO_OK;
} MMO_CATCH_ALL

// Lds.place_ligand_in_ROI rng prot_bst prot roi centered_ligs n (src/lds.ml:308-345): uniform points in the ROI's
// englobing cube (Bbox.rand_point_inside, src/bbox.ml:50-54) until one lies inside the sphere (ROI.is_inside, strict <),
// then Move.rand_rot_full (src/move.ml:20-39: three rand_rot of +-pi about a random axis, R' = R_axis * R); the pose is
// kept unless Mol.heavy_atom_clash (src/mol.ml:1170-1190: two heavy atoms closer than 0.8 (r_i + r_j),
// src/ptable.ml:61-80).  100 000 draws over all starts are fatal, as in the reference.  Random numbers:
// mmo_rng_uniform(seed, 0, 1, 2, ...) (include/mmo_detmath.h) instead of OCaml's Random.State, consumed in
// the reference's order (constructor arguments right to left: z, y, x; then theta, axis three times).  libm sin/cos.
int mmo_place_ligand_in_roi(int32_t n_rec, const double *px, const double *py, const double *pz, const int32_t *panum,
                            const mmo_ligand *lig, const double roi_c[3], double roi_r, uint64_t seed, int32_t n_starts,
                            int32_t clash_check, double *out_rot9, double *out_pos3, int32_t *n_trials) try {
    MMO_REQUIRE(lig && roi_c && roi_r > 0.0 && n_starts >= 0 && n_rec >= 0, "mmo_place_ligand_in_roi: bad arguments");
    MMO_REQUIRE(n_starts == 0 || (out_rot9 && out_pos3), "mmo_place_ligand_in_roi: null output");
    MMO_REQUIRE(n_rec == 0 || (px && py && pz && panum), "mmo_place_ligand_in_roi: null receptor arrays");
    const int L = lig->n;
    double low[3], dims[3];
    for (int d = 0; d < 3; d++) { low[d] = roi_c[d] - roi_r; dims[d] = (roi_c[d] + roi_r) - low[d]; }   // Bbox.create_2v
    const double r2 = roi_r * roi_r;
    uint64_t ctr = 0;
    int trials = 0;
    for (int s = 0; s < n_starts; s++) {
        for (;;) {
            const double z = low[2] + mmo_rng_uniform(seed, ctr++) * dims[2];
            const double y = low[1] + mmo_rng_uniform(seed, ctr++) * dims[1];
            const double x = low[0] + mmo_rng_uniform(seed, ctr++) * dims[0];
            trials++;
            if (trials == 100000) {
                mmo::set_error("Lds.place_ligands_in_ROI: start conformer cannot be placed in ROI after 100k trials");
                return MMO_EINVAL;
            }
            const double ddx = roi_c[0] - x, ddy = roi_c[1] - y, ddz = roi_c[2] - z;
            if (!(ddx * ddx + ddy * ddy + ddz * ddz < r2)) continue;
            double rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            for (int k = 0; k < 3; k++) {
                const double pi = 4.0 * atan(1.0);
                const double theta = mmo_rng_uniform(seed, ctr++) * (2.0 * pi) - pi;      // Move.rand_angle
                int axis = (int)(mmo_rng_uniform(seed, ctr++) * 3.0);                    // Random.State.int rng 3
                if (axis > 2) axis = 2;
                const double c = cos(theta), sn = sin(theta);
                double rb[9];
                if (axis == 0) { const double t[9] = {1, 0, 0, 0, c, sn, 0, -sn, c}; memcpy(rb, t, sizeof t); }       // Rot.rx
                else if (axis == 1) { const double t[9] = {c, 0, -sn, 0, 1, 0, sn, 0, c}; memcpy(rb, t, sizeof t); }  // Rot.ry
                else { const double t[9] = {c, sn, 0, -sn, c, 0, 0, 0, 1}; memcpy(rb, t, sizeof t); }                 // Rot.rz
                double o[9];
                for (int a = 0; a < 3; a++)                                              // Rot.mult rotate_by rot
                    for (int b = 0; b < 3; b++)
                        o[3 * a + b] = rb[3 * a] * rot[b] + rb[3 * a + 1] * rot[3 + b] + rb[3 * a + 2] * rot[6 + b];
                memcpy(rot, o, sizeof o);
            }
            bool clash = false;
            if (clash_check) {
                for (int j = 0; j < L && !clash; j++) {
                    if (lig->hanum[j] <= 1) continue;
                    double q[3];
                    mmo::rot_apply(rot, lig->hx[j], lig->hy[j], lig->hz[j], q);            // rotate_then_translate_copy
                    q[0] += x; q[1] += y; q[2] += z;
                    const double rj = mmo::vdw_radius(lig->hanum[j]);
                    for (int i = 0; i < n_rec; i++) {
                        if (panum[i] <= 1) continue;
                        const double lim = 0.8 * (rj + mmo::vdw_radius(panum[i]));         // Ptable.vdW_clash_params2
                        const double ex = q[0] - px[i], ey = q[1] - py[i], ez = q[2] - pz[i];
                        if (ex * ex + ey * ey + ez * ez < lim * lim) { clash = true; break; }
                    }
                }
            }
            if (clash) continue;
            memcpy(out_rot9 + 9 * (size_t)s, rot, sizeof rot);
            out_pos3[3 * (size_t)s] = x; out_pos3[3 * (size_t)s + 1] = y; out_pos3[3 * (size_t)s + 2] = z;
            break;
        }
    }
    if (n_trials) *n_trials = trials;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_so3_rotations(int32_t n, double *rot9) try {
    MMO_REQUIRE(n >= 0 && (n == 0 || rot9 != nullptr), "mmo_so3_rotations: bad arguments");
    mmo::so3_rotations(n, rot9);
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_rot_r_xyz(double a, double b, double g, double rot9[9]) try {
    MMO_REQUIRE(rot9 != nullptr, "mmo_rot_r_xyz: null pointer");
    mmo::rot_r_xyz(a, b, g, rot9);
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_rot_decompose(const double rot9[9], double abg[3]) try {
    MMO_REQUIRE(rot9 != nullptr && abg != nullptr, "mmo_rot_decompose: null pointer");
    mmo::rot_decompose(rot9, abg);
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_grid_from_box(double step, double bx, double by, double bz, int32_t dims[3]) try {
    MMO_REQUIRE(dims != nullptr && step > 0.0, "mmo_grid_from_box: bad arguments");
    dims[0] = mmo::grid_num_steps(step, bx) + 1;
    dims[1] = mmo::grid_num_steps(step, by) + 1;
    dims[2] = mmo::grid_num_steps(step, bz) + 1;
    return MMO_OK;
} MMO_CATCH_ALL

}  // extern "C"
