/* mmo_detmath.h -- platform-independent double-precision sin/cos/exp and a counter-based RNG.
 *
 * The Monte-Carlo driver of the reference (src/lds.ml:741-1000, src/move.ml) calls libm's sin, cos
 * and exp and OCaml's Random.State.  Neither can be matched bit for bit on a GPU: CUDA's libm
 * differs from glibc in the last ulp, and OCaml's generator changed between 4.14 and 5.x
 * (src/RNG.ml:9-16).  The GPU chains and the CPU oracle therefore both use the
 * functions below -- plain IEEE +,-,*,/ in a fixed order, no FMA (compile with -fmad=false /
 * -ffp-contract=off) -- so that a chain is bit-reproducible between the two, which is what lets
 * the parity tests compare whole trajectories instead of distributions.
 *
 * Accuracy: < 2 ulp on the ranges the driver uses (|angle| <= 2*pi, exp argument <= 0).
 * Polynomial coefficients: the classic fdlibm kernels (k_sin.c, k_cos.c, e_exp.c constants).
 */
#ifndef MMO_DETMATH_H
#define MMO_DETMATH_H
#include <stdint.h>

#ifdef __CUDACC__
#define MMO_HD __host__ __device__ static inline
#else
#define MMO_HD static inline
#endif

MMO_HD double mmo_det_floor(double x) {
    /* |x| < 2^31 on every call site */
    double t = (double)(long long)x;
    return (t > x) ? t - 1.0 : t;
}

/* sin and cos of x, |x| <= ~1e5 */
MMO_HD void mmo_det_sincos(double x, double *s_out, double *c_out) {
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    const double pio2_1t = 6.07710050650619224932e-11; /* pi/2 - pio2_1 */
    double kf = mmo_det_floor(x * two_over_pi + 0.5);
    long long k = (long long)kf;
    double r = (x - kf * pio2_1) - kf * pio2_1t;
    double z = r * r;
    /* sin kernel on [-pi/4, pi/4] */
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    double sr = r + (r * z) * (S1 + z * ps);
    /* cos kernel */
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double cr = (1.0 - 0.5 * z) + z * pc;
    int q = (int)(k & 3);
    double s, c;
    if (q == 0) { s = sr; c = cr; }
    else if (q == 1) { s = cr; c = -sr; }
    else if (q == 2) { s = -sr; c = -cr; }
    else { s = -cr; c = sr; }
    *s_out = s;
    *c_out = c;
}

/* exp(x) for x <= 0 (Metropolis factor); returns 0 below -700 */
MMO_HD double mmo_det_exp(double x) {
    if (x > 0.0) x = 0.0;
    if (x < -700.0) return 0.0;
    const double inv_ln2 = 1.44269504088896338700e+00;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    double kf = mmo_det_floor(x * inv_ln2 + 0.5);
    double r = (x - kf * ln2_hi) - kf * ln2_lo;      /* |r| <= 0.35 */
    /* exp(r) = 1 + r + r^2/2 + ... degree 13 Horner */
    double p = 1.0 / 6227020800.0;
    p = 1.0 / 479001600.0 + r * p;
    p = 1.0 / 39916800.0 + r * p;
    p = 1.0 / 3628800.0 + r * p;
    p = 1.0 / 362880.0 + r * p;
    p = 1.0 / 40320.0 + r * p;
    p = 1.0 / 5040.0 + r * p;
    p = 1.0 / 720.0 + r * p;
    p = 1.0 / 120.0 + r * p;
    p = 1.0 / 24.0 + r * p;
    p = 1.0 / 6.0 + r * p;
    p = 0.5 + r * p;
    p = 1.0 + r * p;
    p = 1.0 + r * p;
    /* scale by 2^k, k in [-1010, 0]: build the power of two from its exponent bits */
    long long k = (long long)kf;
    union { uint64_t u; double d; } sc;
    sc.u = (uint64_t)(k + 1023) << 52;
    return p * sc.d;
}

/* counter-based generator: splitmix64 finaliser of (seed, counter) -> uniform double in [0, 1) */
MMO_HD uint64_t mmo_rng_u64(uint64_t seed, uint64_t counter) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (counter + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return z;
}
MMO_HD double mmo_rng_uniform(uint64_t seed, uint64_t counter) {
    return (double)(mmo_rng_u64(seed, counter) >> 11) * (1.0 / 9007199254740992.0);
}

#endif
