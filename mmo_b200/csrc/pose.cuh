// pose.cuh -- ligand pose coordinates in the reference's own arithmetic (IEEE double, no FMA).
// Explicit __dmul_rn/__dadd_rn so the result does not depend on the file's -fmad setting.
//   Rot.rotate            src/rot.ml:97-100      x' = (a*x + b*y) + c*z
//   Mol.centered_rotate   src/mol.ml:603-607
//   Mol.translate_by      src/mol.ml:593-600     x' + t.x
//   scan: translate_copy_to rot_lig pos (src/lds.ml:1094, src/mol.ml:687-696): x' + (pos.x - 0.0)
#pragma once
#include "common.cuh"

namespace mmo {

struct PoseRT {
    double r[9];
    double t[3];
};

__device__ __forceinline__ double rot_row(double a, double b, double c, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)), __dmul_rn(c, z));
}

// scan frame -> (rotation, lattice node): frame = rot_i + n_rot*(i + j*x_dim + k*xy_dim) (src/lds.ml:1100)
__device__ __forceinline__ void load_pose_rt_frame(const PoseSrc &s, int64_t frame, PoseRT &o) {
    // 32-bit division whenever the frame id allows it (C2: 9702 lattice nodes x 1e5 rotations < 2^32): the 64-bit
    // software division costs ~100 instructions, and the fp32 kernel decodes every pose once per ligand chunk
    int64_t pt;
    int rot_i;
    if ((unsigned long long)frame < 0x100000000ull) {
        const unsigned f = (unsigned)frame, q = f / (unsigned)s.n_rot;
        pt = q;
        rot_i = (int)(f - q * (unsigned)s.n_rot);
    } else {
        pt = frame / s.n_rot;
        rot_i = (int)(frame - pt * s.n_rot);
    }
    int xy = s.lat_dims[0] * s.lat_dims[1];
    int k = (int)(pt / xy);
    int j = (int)((pt - (int64_t)k * xy) / s.lat_dims[0]);
    int i = (int)(pt - ((int64_t)k * xy + (int64_t)j * s.lat_dims[0]));
    const double *r = s.rot9 + 9 * (int64_t)rot_i;
#pragma unroll
    for (int q = 0; q < 9; q++) o.r[q] = __ldg(r + q);
    // lattice node: x_min + i * ((step*n')/n')   (src/lds.ml:1071-1073, src/grid.ml:49-51)
    o.t[0] = __dadd_rn(s.lat_min[0], __dmul_rn((double)i, s.lat_q[0]));
    o.t[1] = __dadd_rn(s.lat_min[1], __dmul_rn((double)j, s.lat_q[1]));
    o.t[2] = __dadd_rn(s.lat_min[2], __dmul_rn((double)k, s.lat_q[2]));
}

// rotation + translation of pose p (kinds 0 and 2)
__device__ __forceinline__ void load_pose_rt(const PoseSrc &s, int64_t p, PoseRT &o) {
    if (s.kind == 0) {
        const double *r = s.rot9 + 9 * p;
#pragma unroll
        for (int k = 0; k < 9; k++) o.r[k] = __ldg(r + k);
        o.t[0] = __ldg(s.trans3 + 3 * p);
        o.t[1] = __ldg(s.trans3 + 3 * p + 1);
        o.t[2] = __ldg(s.trans3 + 3 * p + 2);
    } else {
        load_pose_rt_frame(s, __ldg(s.frames + p), o);
    }
}

__device__ __forceinline__ void pose_atom_rt(const PoseRT &P, double ax, double ay, double az,
                                             double &x, double &y, double &z) {
    x = __dadd_rn(rot_row(P.r[0], P.r[1], P.r[2], ax, ay, az), P.t[0]);
    y = __dadd_rn(rot_row(P.r[3], P.r[4], P.r[5], ax, ay, az), P.t[1]);
    z = __dadd_rn(rot_row(P.r[6], P.r[7], P.r[8], ax, ay, az), P.t[2]);
}

}  // namespace mmo
