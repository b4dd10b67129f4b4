// host_math.cu -- host-side (libm, IEEE double) mirrors of the reference's pose-feed modules.
// OCaml's sin/cos/atan2/sqrt are the C library's, so these produce the same bits as the reference
// on the same platform.  Compiled with -ffp-contract=off (OCaml never fuses a*b+c).
//   SO3.ml:13-39, quat.ml:32-36, rot.ml:52-75,121-146, grid.ml:37-52
#include "common.cuh"
#include <math.h>

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

}  // namespace mmo

extern "C" {

int mmo_so3_rotations(int32_t n, double *rot9) {
    MMO_REQUIRE(n >= 0 && (n == 0 || rot9 != nullptr), "mmo_so3_rotations: bad arguments");
    mmo::so3_rotations(n, rot9);
    return MMO_OK;
}
int mmo_rot_r_xyz(double a, double b, double g, double rot9[9]) {
    MMO_REQUIRE(rot9 != nullptr, "mmo_rot_r_xyz: null pointer");
    mmo::rot_r_xyz(a, b, g, rot9);
    return MMO_OK;
}
int mmo_rot_decompose(const double rot9[9], double abg[3]) {
    MMO_REQUIRE(rot9 != nullptr && abg != nullptr, "mmo_rot_decompose: null pointer");
    mmo::rot_decompose(rot9, abg);
    return MMO_OK;
}
int mmo_grid_from_box(double step, double bx, double by, double bz, int32_t dims[3]) {
    MMO_REQUIRE(dims != nullptr && step > 0.0, "mmo_grid_from_box: bad arguments");
    dims[0] = mmo::grid_num_steps(step, bx) + 1;
    dims[1] = mmo::grid_num_steps(step, by) + 1;
    dims[2] = mmo::grid_num_steps(step, bz) + 1;
    return MMO_OK;
}

}  // extern "C"
