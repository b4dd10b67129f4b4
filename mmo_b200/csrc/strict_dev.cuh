// strict_dev.cuh -- device functions of the reference's double arithmetic, shared by the strict
// translation units (all compiled with -fmad=false).
#pragma once
#include "common.cuh"

namespace mmo {

// ---- scalars: FF.ml:5-20, math.ml:58-62 ------------------------------------------------------
__device__ __forceinline__ double d_sq(double x) { return x * x; }
__device__ __forceinline__ double d_pow6(double x) { double y = x * x; return (y * y) * y; }
__device__ __forceinline__ double d_nzd(double x) { return (x < 0.01) ? 0.01 : x; }
// V3.dist2 u v (V3.ml:23-28)
__device__ __forceinline__ double d_dist2(double ux, double uy, double uz, double vx, double vy, double vz) {
    double dx = ux - vx, dy = uy - vy, dz = uz - vz;
    return dx * dx + dy * dy + dz * dz;
}

// ---- a / b for many a and one b, bit-identical to the IEEE division ------------------------------
// rc = RN(1/b) once (one real division), then per quotient: q0 = RN(a rc) (within 1.5 ulp of a/b), r0 = a - q0 b
// (exact in an fma), q1 = RN(q0 + r0 rc) (a faithful rounding of a/b), r1 = a - q1 b (exact), q2 = RN(q1 + r1 rc).
// With y the correctly rounded reciprocal and q1 faithful, the last fused step rounds a/b correctly (Markstein 1990;
// Cornea, Harrison, Tang 2002) unless b's significand is all ones: that case takes the plain division.  No
// overflow, underflow or subnormals in this path's ranges (b in [0.01, 12], |a| < 1e6).
struct Divisor { double b, rc; bool plain; };
__device__ __forceinline__ Divisor make_divisor(double b) {
    Divisor d;
    d.b = b;
    d.rc = __drcp_rn(b);          // correctly rounded reciprocal
    d.plain = ((__double2hiint(b) & 0x000fffff) == 0x000fffff) && (__double2loint(b) == (int)0xffffffff);
    return d;
}
__device__ __forceinline__ double div_by(double a, const Divisor &d) {
    if (d.plain) return a / d.b;
    double q = a * d.rc;
    double r = fma(-q, d.b, a);
    q = fma(r, d.rc, q);
    r = fma(-q, d.b, a);
    return fma(r, d.rc, q);
}

// FF.shift_12A (FF.ml:17-20): (1 - (d/12)^2)^2 below the cut-off; d / 12.0 through the shared-reciprocal division
__device__ __forceinline__ double d_shift(double d) {
    Divisor by12;
    by12.b = 12.0; by12.rc = 1.0 / 12.0; by12.plain = false;      // RN(1/12), folded at compile time
    return (d < 12.0) ? d_sq(1.0 - d_sq(div_by(d, by12))) : 0.0;
}

// ---- trilinear interpolation (G3D.ml:97-157) ---------------------------------------------------
struct GridGeom {
    double inv;          // grid.one_div_step
    double q[3];         // node i at i*q[d]
    int x_dim, xy_dim;
    int dims[3];
    size_t nvox;
};
__device__ __forceinline__ double d_trilin(const GridGeom &g, const float *__restrict__ arr,
                                           double px, double py, double pz) {
    const int i0 = (int)(px * g.inv), j0 = (int)(py * g.inv), k0 = (int)(pz * g.inv);
    const int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    // the reference reads unchecked (undefined outside the grid); here an outside voxel contributes 0.0
    if (i0 < 0 || j0 < 0 || k0 < 0 || i1 >= g.dims[0] || j1 >= g.dims[1] || k1 >= g.dims[2]) return 0.0;
    const int j0x = j0 * g.x_dim, j1x = j1 * g.x_dim, k0xy = k0 * g.xy_dim, k1xy = k1 * g.xy_dim;
    const double lx = (double)i0 * g.q[0], ly = (double)j0 * g.q[1], lz = (double)k0 * g.q[2];
    const double wlx = (px - lx) * g.inv, wly = (py - ly) * g.inv, wlz = (pz - lz) * g.inv;
    const double whx = 1.0 - wlx, why = 1.0 - wly, whz = 1.0 - wlz;
    return ((double)__ldg(arr + (i0 + j0x + k0xy)) * (whx * why * whz) +
            (double)__ldg(arr + (i1 + j0x + k0xy)) * (wlx * why * whz) +
            (double)__ldg(arr + (i1 + j1x + k0xy)) * (wlx * wly * whz) +
            (double)__ldg(arr + (i0 + j1x + k0xy)) * (whx * wly * whz) +
            (double)__ldg(arr + (i0 + j0x + k1xy)) * (whx * why * wlz) +
            (double)__ldg(arr + (i1 + j0x + k1xy)) * (wlx * why * wlz) +
            (double)__ldg(arr + (i1 + j1x + k1xy)) * (wlx * wly * wlz) +
            (double)__ldg(arr + (i0 + j1x + k1xy)) * (whx * wly * wlz));
}

// the same interpolation from the z-pair copy of the map (mmo_grid::zpair): the same eight floats, products and order of
// additions as d_trilin, hence the same double
__device__ __forceinline__ double d_trilin_zp(const GridGeom &g, const float2 *__restrict__ zp,
                                              double px, double py, double pz) {
    const int i0 = (int)(px * g.inv), j0 = (int)(py * g.inv), k0 = (int)(pz * g.inv);
    const int i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
    if (i0 < 0 || j0 < 0 || k0 < 0 || i1 >= g.dims[0] || j1 >= g.dims[1] || k1 >= g.dims[2]) return 0.0;
    const int j0x = j0 * g.x_dim, j1x = j1 * g.x_dim, k0xy = k0 * g.xy_dim;
    const double lx = (double)i0 * g.q[0], ly = (double)j0 * g.q[1], lz = (double)k0 * g.q[2];
    const double wlx = (px - lx) * g.inv, wly = (py - ly) * g.inv, wlz = (pz - lz) * g.inv;
    const double whx = 1.0 - wlx, why = 1.0 - wly, whz = 1.0 - wlz;
    const float2 a = __ldg(zp + (i0 + j0x + k0xy)), b = __ldg(zp + (i1 + j0x + k0xy));
    const float2 c = __ldg(zp + (i1 + j1x + k0xy)), d = __ldg(zp + (i0 + j1x + k0xy));
    return ((double)a.x * (whx * why * whz) +
            (double)b.x * (wlx * why * whz) +
            (double)c.x * (wlx * wly * whz) +
            (double)d.x * (whx * wly * whz) +
            (double)a.y * (whx * why * wlz) +
            (double)b.y * (wlx * why * wlz) +
            (double)c.y * (wlx * wly * wlz) +
            (double)d.y * (whx * wly * wlz));
}

static inline GridGeom geom_of(const mmo_grid *g) {
    GridGeom G;
    G.inv = 1.0 / g->step;                  // grid.ml:41
    for (int d = 0; d < 3; d++) {
        int np = g->dims[d] - 1;
        G.q[d] = np > 0 ? (g->step * (double)np) / (double)np : 0.0;    // grid.ml:49-51
    }
    G.x_dim = g->dims[0];
    for (int d = 0; d < 3; d++) G.dims[d] = g->dims[d];
    G.xy_dim = g->dims[0] * g->dims[1];
    G.nvox = g->nvox;
    return G;
}


}  // namespace mmo
