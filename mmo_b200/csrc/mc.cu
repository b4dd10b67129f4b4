// mc.cu -- K6: many independent protein-ligand Monte-Carlo chains per launch.
//
// Replaces the frame loop of Lds.simulate_lig (src/lds.ml:882-995) with the moves of src/move.ml and
// src/mol.ml:593-710 (rand_rot, rand_trans, tweak_rbond / flip_rbond, rotate_bond,
// center_rotate_translate_copy), the interpolated scorer (src/mol.ml:1012-1020, src/G3D.ml:97-157),
// the intra-ligand non-bonded energy (src/mol.ml:881-903), the Metropolis test (lds.ml:931-934), the
// acceptance windows (src/SW.ml) and the adaptive step sizes (lds.ml:586-621).
//
// One warp per chain.  Scalars (rotation, position, energies, RNG counter) are kept redundantly in
// every lane -- all lanes execute the same IEEE operations, so they stay identical -- while atoms,
// pair terms and trilinear look-ups are spread over the lanes.  Sums are then accumulated in the
// reference's order from a shared-memory staging row read as broadcasts, which keeps every energy
// bit-identical to the sequential loops of the reference (and of oracle/mmo_oracle_mc.c).
// sin/cos/exp and the random stream come from include/mmo_detmath.h on both sides (see there).
// Compiled with -fmad=false.  Reference quirks D1-D6, D14 of SURVEY Appendix D are mirrored.
#include "common.cuh"
#include "strict_dev.cuh"
#include "../../include/mmo_detmath.h"
#include <math.h>
#include <string.h>

namespace mmo {

constexpr int kBlockSize = 100;      // params.ml:29
constexpr int kWarpsPerBlock = 4;

struct Sw { unsigned bits[4]; int head, n, accepts, rejects; };   // SW.ml, window of 100 events

__device__ __forceinline__ void sw_reset(Sw &s) { s.bits[0] = s.bits[1] = s.bits[2] = s.bits[3] = 0u; s.head = 0; s.n = 0; s.accepts = 0; s.rejects = 0; }
__device__ __forceinline__ void sw_process(Sw &s, bool evt) {      // SW.ml:20-34
    int slot;
    if (s.n == kBlockSize) {
        slot = s.head;
        bool old = (s.bits[slot >> 5] >> (slot & 31)) & 1u;
        if (old) s.accepts--; else s.rejects--;
        s.head = (s.head + 1) % kBlockSize;
    } else {
        slot = (s.head + s.n) % kBlockSize;
        s.n++;
    }
    unsigned m = 1u << (slot & 31);
    if (evt) { s.bits[slot >> 5] |= m; s.accepts++; } else { s.bits[slot >> 5] &= ~m; s.rejects++; }
}
__device__ __forceinline__ double sw_ratio(const Sw &s) { return (double)s.accepts / (double)(s.accepts + s.rejects); }

struct McArgs {
    // ligand (centred template)
    int L;
    const double *lx, *ly, *lz, *lq;
    const int32_t *lelt, *ltyp;
    int n_pairs;
    const int32_t *pair_i, *pair_j;
    int n_rbonds;
    const int32_t *rb_left, *rb_right, *rg_off, *rg_idx;
    // interpolated scorer (maps != nullptr) ...
    GridGeom g;
    const float *maps;
    // ... or direct shifted scorer over the receptor atoms
    int P;
    const double4 *pxyzq;
    const int32_t *pelt;
    // UFF tables (kEltTab^2)
    const double *xij, *dij;
    double roi_c[3], roi_r;
    int tweak_rbonds, hard_roi, no_flip, intra_nb;
    double beta;
    int n_steps;
    int64_t n_chains;
    const uint64_t *seeds;
    const double *rot0, *pos0;          // per chain
    double p_max_rot, p_max_trans, p_max_rbond_rot, pi;   // params.ml:11-20 computed on the host (libm atan)
    // outputs, per chain
    double *best_E, *prev_E, *best_rot, *best_pos, *best_xyz, *step_sizes;   // step_sizes: max_rot, max_trans
    long long *counters;                // 8 per chain: acc_rigid, rej_rigid, acc_conf, rej_conf, ooroi, ezero, too_long, frames
    double *trace;                      // optional: chain 0 only, 4 doubles per frame
};

// rot.ml:22-46
__device__ __forceinline__ void det_rot_axis(int axis, double th, double *r) {
    double s, c;
    mmo_det_sincos(th, &s, &c);
    if (axis == 0) { r[0] = 1.0; r[1] = 0.0; r[2] = 0.0; r[3] = 0.0; r[4] = c; r[5] = s; r[6] = 0.0; r[7] = -s; r[8] = c; }
    else if (axis == 1) { r[0] = c; r[1] = 0.0; r[2] = -s; r[3] = 0.0; r[4] = 1.0; r[5] = 0.0; r[6] = s; r[7] = 0.0; r[8] = c; }
    else { r[0] = c; r[1] = s; r[2] = 0.0; r[3] = -s; r[4] = c; r[5] = 0.0; r[6] = 0.0; r[7] = 0.0; r[8] = 1.0; }
}
// rot.ml:77-94
__device__ __forceinline__ void rot_mult(const double *a, const double *b, double *o) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            o[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
}
// rot.ml:97-100
__device__ __forceinline__ void rot_apply(const double *r, double x, double y, double z, double &ox, double &oy, double &oz) {
    ox = r[0] * x + r[1] * y + r[2] * z;
    oy = r[3] * x + r[4] * y + r[5] * z;
    oz = r[6] * x + r[7] * y + r[8] * z;
}
// Batteries A.favg restated as in the oracle: Kahan-compensated sum / n (same order in every lane);
// the three coordinates at once: three independent Kahan chains in one loop (the same operations per chain, interleaved)
__device__ __forceinline__ void favg3_smem(const double *ax, const double *ay, const double *az, int n, double *out) {
    double s0 = 0.0, c0 = 0.0, s1 = 0.0, c1 = 0.0, s2 = 0.0, c2 = 0.0;
    for (int i = 0; i < n; i++) {
        const double y0 = ax[i] - c0, y1 = ay[i] - c1, y2 = az[i] - c2;
        const double t0 = s0 + y0, t1 = s1 + y1, t2 = s2 + y2;
        c0 = (t0 - s0) - y0; c1 = (t1 - s1) - y1; c2 = (t2 - s2) - y2;
        s0 = t0; s1 = t1; s2 = t2;
    }
    out[0] = s0 / (double)n; out[1] = s1 / (double)n; out[2] = s2 / (double)n;
}

// Mol.ene_intra_UFFNB_brute (mol.ml:881-903): the pair terms are computed 32 at a time, one per lane,
// parked in shared memory and then added up in the reference's (i<j) order by every lane alike
// (broadcast reads: one LDS.128 + two DADD per term), which keeps the sum bit-identical.
constexpr int kTermDoubles = 128;      // per warp: two buffers of 32 double2 term slots
__device__ __forceinline__ double2 intra_term(const McArgs &a, const double *x, const double *y, const double *z, int k) {
    double2 t = make_double2(0.0, 0.0);
    if (k < a.n_pairs) {
        const int i = __ldg(a.pair_i + k), j = __ldg(a.pair_j + k);
        const double r = d_nzd(sqrt(d_dist2(x[i], y[i], z[i], x[j], y[j], z[j])));
        const int tt = __ldg(a.lelt + i) * kEltTab + __ldg(a.lelt + j);
        const Divisor by_r = make_divisor(r);
        const double p6 = d_pow6(div_by(__ldg(a.xij + tt), by_r));
        t.x = div_by(__ldg(a.lq + i) * __ldg(a.lq + j), by_r);
        t.y = __ldg(a.dij + tt) * ((-2.0 * p6) + (p6 * p6));
    }
    return t;
}
// Software pipelined: the terms of round r + 1 are computed (sqrt and division chains) between the store of round r
// and its in-order summation, two independent dependency chains the scheduler can interleave; two term buffers.
__device__ double intra_energy(const McArgs &a, const double *x, const double *y, const double *z, int lane,
                               double2 *terms) {
    double se = 0.0, sv = 0.0;
    double2 t = intra_term(a, x, y, z, lane);
    int buf = 0;
    for (int base = 0; base < a.n_pairs; base += 32) {
        double2 *tb = terms + 32 * buf;
        __syncwarp();
        tb[lane] = t;
        __syncwarp();
        t = intra_term(a, x, y, z, base + 32 + lane);                 // next round (zeros beyond the last pair)
        const int lim = min(32, a.n_pairs - base);
        if (lim == 32) {
#pragma unroll
            for (int l = 0; l < 32; l++) {
                const double2 u = tb[l];
                se = se + u.x;
                sv = sv + u.y;
            }
        } else {
            for (int l = 0; l < lim; l++) {
                const double2 u = tb[l];
                se = se + u.x;
                sv = sv + u.y;
            }
        }
        buf ^= 1;
    }
    return (kElecWeight * se) + sv;
}

// Mol.ene_inter_UFF_interp (mol.ml:1012-1020): one trilinear look-up per lane, summed in atom order
__device__ double interp_energy(const McArgs &a, const double *x, const double *y, const double *z, int lane,
                                double2 *terms) {
    double res = 0.0;
    for (int base = 0; base < a.L; base += 32) {
        const int j = base + lane;
        double t = 0.0;
        if (j < a.L) t = d_trilin(a.g, a.maps + (size_t)__ldg(a.ltyp + j) * a.g.nvox, x[j], y[j], z[j]);
        __syncwarp();
        terms[lane].x = t;
        __syncwarp();
        const int lim = min(32, a.L - base);
        for (int l = 0; l < lim; l++) res = res + terms[l].x;
    }
    return res;
}

// Mol.ene_inter_UFF_shifted_brute (mol.ml:822-849) for one chain: receptor atoms are dealt to the lanes,
// every pair term is computed with the reference's own operations (sqrt, divisions, no FMA); the 32
// per-lane partial sums are then added in lane order.  Only the summation order differs from the
// reference (receptor-outer/ligand-inner over ALL atoms), i.e. the value agrees to ~1e-13 relative.
__device__ double direct_energy(const McArgs &a, const double *x, const double *y, const double *z, int lane,
                                double2 *terms) {
    // bounding box of the ligand: receptor atoms farther than 12 A from it have weight 0 (mol.ml:836)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int j = lane; j < a.L; j += 32) {
        lo[0] = fmin(lo[0], x[j]); hi[0] = fmax(hi[0], x[j]);
        lo[1] = fmin(lo[1], y[j]); hi[1] = fmax(hi[1], y[j]);
        lo[2] = fmin(lo[2], z[j]); hi[2] = fmax(hi[2], z[j]);
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    double se = 0.0, sv = 0.0;
    for (int i = lane; i < a.P; i += 32) {
        const double2 p01 = __ldg((const double2 *)(a.pxyzq + i));
        const double2 p23 = __ldg((const double2 *)(a.pxyzq + i) + 1);
        const double gx = fmax(0.0, fmax(lo[0] - p01.x, p01.x - hi[0]));
        const double gy = fmax(0.0, fmax(lo[1] - p01.y, p01.y - hi[1]));
        const double gz = fmax(0.0, fmax(lo[2] - p23.x, p23.x - hi[2]));
        if (gx * gx + gy * gy + gz * gz >= 144.0) continue;
        const double q_i = p23.y;
        const int ei = __ldg(a.pelt + i) * kEltTab;
        for (int j = 0; j < a.L; j++) {
            const double r2 = d_dist2(p01.x, p01.y, p23.x, x[j], y[j], z[j]);
            if (r2 < 144.0) {
                const double r = d_nzd(sqrt(r2));
                const double w = d_shift(r);
                const int t = ei + __ldg(a.lelt + j);
                const Divisor by_r = make_divisor(r);
                const double p6 = d_pow6(div_by(__ldg(a.xij + t), by_r));
                se = se + w * div_by(q_i * __ldg(a.lq + j), by_r);
                sv = sv + w * (__ldg(a.dij + t) * ((-2.0 * p6) + (p6 * p6)));
            }
        }
    }
    __syncwarp();
    terms[lane] = make_double2(se, sv);
    __syncwarp();
    double te = 0.0, tv = 0.0;
    for (int l = 0; l < 32; l++) { te = te + terms[l].x; tv = tv + terms[l].y; }
    return (kElecWeight * te) + tv;
}

__device__ __forceinline__ double inter_energy(const McArgs &a, const double *x, const double *y, const double *z,
                                               int lane, double2 *terms) {
    return a.maps ? interp_energy(a, x, y, z, lane, terms) : direct_energy(a, x, y, z, lane, terms);
}

// Two builds of the same kernel: MINB = 1 keeps every chain's state in registers (248: lowest latency per frame,
// at most 8 chains per SM) for launches that do not fill the GPU anyway; MINB = 7 caps the registers at 72 (state
// spills to local memory, slower per chain) so that 28 chains are resident per SM when there are thousands.
template <int MINB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
mc_chains_kernel(McArgs a) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t chain = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
    if (chain >= a.n_chains) return;                    // whole warp leaves together
    const int L = a.L, nrb = a.n_rbonds;
    const int per_warp = ((9 * L + 2 * max(nrb, 1) + 1) & ~1) + kTermDoubles;     // even, + 2 x 32 double2 term slots
    double *cx = smem + (size_t)wib * per_warp, *cy = cx + L, *cz = cy + L;      // conf
    double *px = cz + L, *py = px + L, *pz = py + L;                              // conf' (proposed)
    double *lx = pz + L, *ly = lx + L, *lz = ly + L;                              // lig'
    double *dr = lz + L, *drp = dr + max(nrb, 1);                                 // per-bond step sizes of conf / conf'
    double2 *terms = (double2 *)(cx + per_warp - kTermDoubles);                   // 16-byte aligned: per_warp is even
    Sw *sw_bond = (Sw *)(smem + (size_t)kWarpsPerBlock * per_warp) + (size_t)wib * max(nrb, 1);

    const bool flexible = a.tweak_rbonds && nrb > 0;
    const long long rbf = a.no_flip ? 0x7fffffffffffffffLL : (long long)kBlockSize;
    const double target_low = 0.5 - 0.05, target_high = 0.5 + 0.05;    // lds.ml:651-652
    const uint64_t seed = a.seeds[chain];
    uint64_t ctr = 0;

    for (int j = lane; j < L; j += 32) { cx[j] = a.lx[j]; cy[j] = a.ly[j]; cz[j] = a.lz[j]; }
    for (int b = lane; b < nrb; b += 32) { dr[b] = a.p_max_rbond_rot; sw_reset(sw_bond[b]); }
    double ccen[3] = {0.0, 0.0, 0.0}, pcen[3] = {0.0, 0.0, 0.0};        // conf.center, conf'.center
    Sw sw_rigid;
    sw_reset(sw_rigid);
    double max_rot = a.p_max_rot, max_trans = a.p_max_trans;
    double rot[9], pos[3], rot0[9], pos0[3];
#pragma unroll
    for (int k = 0; k < 9; k++) rot0[k] = rot[k] = a.rot0[chain * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; k++) pos0[k] = pos[k] = a.pos0[chain * 3 + k];
    double best_rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, best_pos[3] = {0, 0, 0};
    double *bxyz = a.best_xyz + chain * 3 * (int64_t)L;
    __syncwarp();
    // start_conf = rotate_then_translate_copy centered_lig rot0 pos0 (lds.ml:758)
    for (int j = lane; j < L; j += 32) {
        double x, y, z;
        rot_apply(rot, cx[j], cy[j], cz[j], x, y, z);
        lx[j] = x + pos[0]; ly[j] = y + pos[1]; lz[j] = z + pos[2];
        bxyz[j] = lx[j]; bxyz[L + j] = ly[j]; bxyz[2 * L + j] = lz[j];
    }
    __syncwarp();
    double const_intra = 0.0;
    if (a.intra_nb && !flexible) const_intra = intra_energy(a, cx, cy, cz, lane, terms);
    double prev_E_intra = !a.intra_nb ? 0.0 : (flexible ? intra_energy(a, lx, ly, lz, lane, terms) : const_intra);
    double prev_E_inter = inter_energy(a, lx, ly, lz, lane, terms);
    double prev_E = prev_E_inter + prev_E_intra;
    double best_E = prev_E;
    long long rigid_step = 0, conf_step = 0;
    long long n_acc_r = 0, n_rej_r = 0, n_acc_c = 0, n_rej_c = 0, n_ooroi = 0, n_ezero = 0, too_long = 0;
    int frame = 0;
    for (; frame < a.n_steps; frame++) {
        const bool rigid = (frame & 1) == 0;
        int just_rotated = -1;
        int which = 0;                                  // conf' is: 0 = conf, 1 = proposed copy, 2 = centred template
        if (!rigid) {
            if (flexible) {
                for (int j = lane; j < L; j += 32) { px[j] = cx[j]; py[j] = cy[j]; pz[j] = cz[j]; }
                for (int b = lane; b < nrb; b += 32) drp[b] = dr[b];
                __syncwarp();
                int bond;
                double alpha;
                {
                    double u = mmo_rng_uniform(seed, ctr++);
                    bond = (int)(u * (double)nrb);
                    if (bond >= nrb) bond = nrb - 1;
                    double u2 = mmo_rng_uniform(seed, ctr++);
                    if (conf_step > 0 && conf_step % rbf == 0) alpha = (2.0 * a.pi) * u2 - a.pi;     // Mol.flip_rbond
                    else { double d = drp[bond]; alpha = (2.0 * d) * u2 - d; }                       // Mol.tweak_rbond
                }
                // Mol.rotate_bond (mol.ml:610-631)
                const int left = __ldg(a.rb_left + bond), right = __ldg(a.rb_right + bond);
                const double ox = px[right], oy = py[right], oz = pz[right];
                const double ax = ox - px[left], ay = oy - py[left], az = oz - pz[left];
                const double mag = sqrt(ax * ax + ay * ay + az * az);
                double br[9];
                {
                    const double ux = ax / mag, uy = ay / mag, uz = az / mag;
                    double s, c;
                    mmo_det_sincos(alpha, &s, &c);
                    const double omc = 1.0 - c;                  // rot.ml:136-146
                    br[0] = c + ux * ux * omc; br[1] = ux * uy * omc - uz * s; br[2] = ux * uz * omc + uy * s;
                    br[3] = ux * uy * omc + uz * s; br[4] = c + uy * uy * omc; br[5] = uy * uz * omc - ux * s;
                    br[6] = ux * uz * omc - uy * s; br[7] = uy * uz * omc + ux * s; br[8] = c + uz * uz * omc;
                }
                __syncwarp();
                const int g0 = __ldg(a.rg_off + bond), g1 = __ldg(a.rg_off + bond + 1);
                for (int g = g0 + lane; g < g1; g += 32) {
                    const int i = __ldg(a.rg_idx + g);
                    double x, y, z;
                    rot_apply(br, px[i] - ox, py[i] - oy, pz[i] - oz, x, y, z);
                    px[i] = x + ox; py[i] = y + oy; pz[i] = z + oz;
                }
                __syncwarp();
                favg3_smem(px, py, pz, L, pcen);                                                       // update_center
                just_rotated = bond;
                // Mol.check_elongation_exn lig 12.0 (mol.ml:576-591)
                double maxi = 0.0;
                for (int j = lane; j < L; j += 32) {
                    double d = 0.01 + sqrt(d_dist2(pcen[0], pcen[1], pcen[2], px[j], py[j], pz[j]));
                    maxi = fmax(maxi, d);
                }
                for (int o = 16; o > 0; o >>= 1) maxi = fmax(maxi, __shfl_xor_sync(0xffffffffu, maxi, o));
                if (maxi > 12.0) { too_long = 1; break; }          // Mol.Too_long ends this run (lds.ml:996-997)
                which = 1;
            } else {
                which = 2;
            }
        }
        double rotp[9], posp[3];
#pragma unroll
        for (int k = 0; k < 9; k++) rotp[k] = rot[k];
#pragma unroll
        for (int k = 0; k < 3; k++) posp[k] = pos[k];
        if (rigid) {
            // right-to-left evaluation: rand_trans (z, y, x) before rand_rot (theta, axis)
            const double dz = 2.0 * mmo_rng_uniform(seed, ctr++) - 1.0;
            const double dy = 2.0 * mmo_rng_uniform(seed, ctr++) - 1.0;
            const double dx = 2.0 * mmo_rng_uniform(seed, ctr++) - 1.0;
            posp[0] = pos[0] + dx * max_trans; posp[1] = pos[1] + dy * max_trans; posp[2] = pos[2] + dz * max_trans;
            const double theta = (2.0 * max_rot) * mmo_rng_uniform(seed, ctr++) - max_rot;
            int axis = (int)(mmo_rng_uniform(seed, ctr++) * 3.0);
            if (axis > 2) axis = 2;
            double rb[9];
            det_rot_axis(axis, theta, rb);
            rot_mult(rb, rot, rotp);                               // move.ml:31
        }
        // lig' = center_rotate_translate_copy conf' rot' pos' (mol.ml:705-710)
        {
            const double *sx = which == 1 ? px : (which == 2 ? a.lx : cx);
            const double *sy = which == 1 ? py : (which == 2 ? a.ly : cy);
            const double *sz = which == 1 ? pz : (which == 2 ? a.lz : cz);
            const double *cen = which == 1 ? pcen : ccen;
            const double nx = which == 2 ? -0.0 : -cen[0], ny = which == 2 ? -0.0 : -cen[1], nz = which == 2 ? -0.0 : -cen[2];
            for (int j = lane; j < L; j += 32) {
                double x, y, z;
                rot_apply(rotp, sx[j] + nx, sy[j] + ny, sz[j] + nz, x, y, z);
                lx[j] = x + posp[0]; ly[j] = y + posp[1]; lz[j] = z + posp[2];
            }
        }
        __syncwarp();
        if (!rigid && a.intra_nb) prev_E_intra = flexible ? intra_energy(a, lx, ly, lz, lane, terms) : const_intra;   // D2
        prev_E_inter = inter_energy(a, lx, ly, lz, lane, terms);
        const double curr_E = prev_E_inter + prev_E_intra;
        int accepted = -1;
        const double ddx = a.roi_c[0] - (0.0 + posp[0]), ddy = a.roi_c[1] - (0.0 + posp[1]), ddz = a.roi_c[2] - (0.0 + posp[2]);
        const double dist_roi = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        bool do_reset = false;
        if (a.hard_roi) {                                    // D1: everything below hangs off --hard-ROI
            if (dist_roi > a.roi_r) { do_reset = true; n_ooroi++; }
            else if (prev_E_inter == 0.0) { do_reset = true; n_ezero++; }          // D6
            else {
                bool acc = curr_E <= prev_E;
                if (!acc) acc = mmo_rng_uniform(seed, ctr++) < mmo_det_exp((-(curr_E - prev_E)) * a.beta);
                accepted = acc ? 1 : 0;
                if (rigid) { sw_process(sw_rigid, acc); if (acc) n_acc_r++; else n_rej_r++; }
                else {
                    if (just_rotated > -1 && lane == 0) sw_process(sw_bond[just_rotated], acc);
                    if (acc) n_acc_c++; else n_rej_c++;
                }
                __syncwarp();
                if (acc) {
#pragma unroll
                    for (int k = 0; k < 9; k++) rot[k] = rotp[k];
#pragma unroll
                    for (int k = 0; k < 3; k++) pos[k] = posp[k];
                    prev_E = curr_E;
                    if (which == 1) {                              // conf := conf'
                        for (int j = lane; j < L; j += 32) { cx[j] = px[j]; cy[j] = py[j]; cz[j] = pz[j]; }
                        for (int b = lane; b < nrb; b += 32) dr[b] = drp[b];
                        ccen[0] = pcen[0]; ccen[1] = pcen[1]; ccen[2] = pcen[2];
                    } else if (which == 2) {
                        for (int j = lane; j < L; j += 32) { cx[j] = a.lx[j]; cy[j] = a.ly[j]; cz[j] = a.lz[j]; }
                        for (int b = lane; b < nrb; b += 32) dr[b] = a.p_max_rbond_rot;
                        ccen[0] = ccen[1] = ccen[2] = 0.0;
                    }
                }
                if (curr_E < best_E) {                             // D3
                    best_E = curr_E;
#pragma unroll
                    for (int k = 0; k < 9; k++) best_rot[k] = rotp[k];
#pragma unroll
                    for (int k = 0; k < 3; k++) best_pos[k] = posp[k];
                    for (int j = lane; j < L; j += 32) { bxyz[j] = lx[j]; bxyz[L + j] = ly[j]; bxyz[2 * L + j] = lz[j]; }
                }
                if (rigid && rigid_step > 0 && rigid_step % kBlockSize == 0) {     // lds.ml:586-600
                    const double ar = sw_ratio(sw_rigid);
                    if (ar <= target_low) { max_trans = 0.95 * max_trans; max_rot = 0.95 * max_rot; }
                    else if (ar >= target_high) {
                        max_trans = 1.05 * max_trans;
                        const double m = 1.05 * max_rot;
                        max_rot = (a.pi <= m) ? a.pi : m;
                    }
                }
                if (flexible && !rigid && conf_step > 0 && conf_step % ((long long)kBlockSize * nrb) == 0) {
                    __syncwarp();
                    double *tgt = acc ? dr : drp;                  // D4
                    for (int b = lane; b < nrb; b += 32) {
                        const double ar = sw_ratio(sw_bond[b]);
                        if (ar <= target_low) tgt[b] = 0.95 * tgt[b];
                        else if (ar >= target_high) { const double m = 1.05 * tgt[b]; tgt[b] = (a.pi <= m) ? a.pi : m; }
                    }
                }
                __syncwarp();
            }
        }
        if (do_reset) {                                      // reset_run_params (lds.ml:632-648)
            max_rot = a.p_max_rot; max_trans = a.p_max_trans;
#pragma unroll
            for (int k = 0; k < 9; k++) { rot[k] = rot0[k]; best_rot[k] = (k % 4 == 0) ? 1.0 : 0.0; }
#pragma unroll
            for (int k = 0; k < 3; k++) { pos[k] = pos0[k]; best_pos[k] = 0.0; }
            prev_E = INFINITY; best_E = INFINITY;
            for (int j = lane; j < L; j += 32) {
                double x, y, z;
                rot_apply(rot, a.lx[j], a.ly[j], a.lz[j], x, y, z);
                bxyz[j] = x + pos[0]; bxyz[L + j] = y + pos[1]; bxyz[2 * L + j] = z + pos[2];
            }
            sw_reset(sw_rigid);
        }
        if (a.trace && chain == 0 && lane == 0) {
            a.trace[4 * (size_t)frame] = curr_E; a.trace[4 * (size_t)frame + 1] = prev_E_inter;
            a.trace[4 * (size_t)frame + 2] = prev_E_intra; a.trace[4 * (size_t)frame + 3] = (double)accepted;
        }
        if (rigid) rigid_step++; else conf_step++;
        __syncwarp();
    }
    if (lane == 0) {
        a.best_E[chain] = best_E;
        a.prev_E[chain] = prev_E;
        for (int k = 0; k < 9; k++) a.best_rot[chain * 9 + k] = best_rot[k];
        for (int k = 0; k < 3; k++) a.best_pos[chain * 3 + k] = best_pos[k];
        a.step_sizes[chain * 2] = max_rot; a.step_sizes[chain * 2 + 1] = max_trans;
        long long *c = a.counters + chain * 8;
        c[0] = n_acc_r; c[1] = n_rej_r; c[2] = n_acc_c; c[3] = n_rej_c; c[4] = n_ooroi; c[5] = n_ezero; c[6] = too_long; c[7] = frame;
    }
}

// library-lifetime tables: never destroyed at process exit, released by mmo_shutdown (mc_drop_caches)
static DevBuf<double> &g_mc_xij = *new DevBuf<double>(), &g_mc_dij = *new DevBuf<double>();
void mc_drop_caches() { g_mc_xij.release(); g_mc_dij.release(); }

static int ensure_mc_tables() {
    if (g_mc_xij.p) return MMO_OK;
    std::vector<double> hx(kEltTab * kEltTab), hd(kEltTab * kEltTab);
    for (int a = 0; a < kEltTab; a++)
        for (int b = 0; b < kEltTab; b++) {
            bool ok = a < kNumElt && b < kNumElt;
            hx[a * kEltTab + b] = ok ? sqrt(kEltXi[a] * kEltXi[b]) : NAN;    // FF.geo_mean
            hd[a * kEltTab + b] = ok ? sqrt(kEltDi[a] * kEltDi[b]) : NAN;
        }
    MMO_TRY(g_mc_xij.upload(hx));
    MMO_TRY(g_mc_dij.upload(hd));
    return MMO_OK;
}

}  // namespace mmo

using namespace mmo;

extern "C" int mmo_mc_run(const mmo_receptor *rec, const mmo_grid *grid, const mmo_ligand *lig, const mmo_mc_params *p,
                          int64_t n_chains, const uint64_t *seeds, const double *start_rot9,
                          const double *start_pos3, mmo_mc_result *results, double *best_xyz,
                          double *trace_chain0) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(lig && p, "mmo_mc_run: null handle");
    MMO_REQUIRE((rec != nullptr) != (grid != nullptr), "mmo_mc_run: give exactly one of rec (direct, --no-interp) or grid (interpolated)");
    MMO_REQUIRE(!grid || lig->has_typ, "mmo_mc_run: the ligand needs FF atom types (interpolated scorer)");
    MMO_REQUIRE(!p->intra_nb || lig->has_dists, "mmo_mc_run: --intra-NB needs topological distances");
    MMO_REQUIRE(n_chains >= 0 && p->n_steps >= 0, "mmo_mc_run: negative size");
    if (n_chains == 0) return MMO_OK;
    MMO_REQUIRE(seeds && start_rot9 && start_pos3 && results, "mmo_mc_run: null buffer");
    for (int j = 0; grid && j < lig->n; j++)
        MMO_REQUIRE(lig->htyp[j] >= 0 && lig->htyp[j] < grid->T, "mmo_mc_run: atom %d has type %d, grid holds %d maps", j, lig->htyp[j], grid->T);
    MMO_TRY(ensure_mc_tables());
    Runtime &R = rt();
    const int L = lig->n;
    DevBuf<uint64_t> d_seeds;
    DevBuf<double> d_rot, d_pos, d_bestE, d_prevE, d_brot, d_bpos, d_bxyz, d_steps, d_trace;
    DevBuf<long long> d_cnt;
    MMO_TRY(d_seeds.upload(seeds, (size_t)n_chains));
    MMO_TRY(d_rot.upload(start_rot9, (size_t)n_chains * 9));
    MMO_TRY(d_pos.upload(start_pos3, (size_t)n_chains * 3));
    MMO_TRY(d_bestE.alloc((size_t)n_chains)); MMO_TRY(d_prevE.alloc((size_t)n_chains));
    MMO_TRY(d_brot.alloc((size_t)n_chains * 9)); MMO_TRY(d_bpos.alloc((size_t)n_chains * 3));
    MMO_TRY(d_bxyz.alloc((size_t)n_chains * 3 * L)); MMO_TRY(d_steps.alloc((size_t)n_chains * 2));
    MMO_TRY(d_cnt.alloc((size_t)n_chains * 8));
    if (trace_chain0) {
        // frames after a Mol.Too_long break are never written by the kernel: NaN, not stale pool memory
        // (frames_done of chain 0 is the valid length)
        MMO_TRY(d_trace.alloc((size_t)std::max(1, p->n_steps) * 4));
        MMO_CUDA(cudaMemsetAsync(d_trace.p, 0xff, (size_t)std::max(1, p->n_steps) * 4 * sizeof(double), R.stream));
    }
    McArgs a;
    a.L = L; a.lx = lig->x.p; a.ly = lig->y.p; a.lz = lig->z.p; a.lq = lig->q.p; a.lelt = lig->elt.p; a.ltyp = lig->typ.p;
    a.n_pairs = lig->n_pairs; a.pair_i = lig->pair_i.p; a.pair_j = lig->pair_j.p;
    a.n_rbonds = lig->n_rbonds; a.rb_left = lig->d_rb_left.p; a.rb_right = lig->d_rb_right.p;
    a.rg_off = lig->d_rg_off.p; a.rg_idx = lig->d_rg_idx.p;
    if (grid) { a.g = geom_of(grid); a.maps = grid->maps.p; } else { memset(&a.g, 0, sizeof a.g); a.maps = nullptr; }
    a.P = rec ? rec->n : 0; a.pxyzq = rec ? rec->xyzq64.p : nullptr; a.pelt = rec ? rec->elt.p : nullptr;
    a.xij = g_mc_xij.p; a.dij = g_mc_dij.p;
    for (int d = 0; d < 3; d++) a.roi_c[d] = p->roi_c[d];
    a.roi_r = p->roi_r;
    a.tweak_rbonds = p->tweak_rbonds; a.hard_roi = p->hard_roi; a.no_flip = p->no_flip; a.intra_nb = p->intra_nb;
    a.beta = 1.0 / (0.0019872041 * p->temperature_K);           // lds.ml:66-67, const.ml:24
    a.n_steps = p->n_steps; a.n_chains = n_chains;
    a.seeds = d_seeds.p; a.rot0 = d_rot.p; a.pos0 = d_pos.p;
    a.pi = 4.0 * atan(1.0);                                      // math.ml:13
    a.p_max_rot = 15.0 * (a.pi / 180.0);                         // params.ml:11
    a.p_max_trans = 0.15;                                        // params.ml:14
    a.p_max_rbond_rot = 5.0 * (a.pi / 180.0);                    // params.ml:17
    a.best_E = d_bestE.p; a.prev_E = d_prevE.p; a.best_rot = d_brot.p; a.best_pos = d_bpos.p; a.best_xyz = d_bxyz.p;
    a.step_sizes = d_steps.p; a.counters = d_cnt.p; a.trace = trace_chain0 ? d_trace.p : nullptr;
    const int nrb1 = std::max(lig->n_rbonds, 1);
    const size_t smem = (size_t)kWarpsPerBlock * (((9 * L + 2 * nrb1 + 1) & ~1) + kTermDoubles) * sizeof(double) + (size_t)kWarpsPerBlock * nrb1 * sizeof(Sw);
    MMO_REQUIRE(smem <= 200 * 1024, "mmo_mc_run: ligand too large (%d atoms, %d rotatable bonds)", L, lig->n_rbonds);
    MMO_CUDA(cudaFuncSetAttribute(mc_chains_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMO_CUDA(cudaFuncSetAttribute(mc_chains_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((n_chains + kWarpsPerBlock - 1) / kWarpsPerBlock);
    {
        KernelScope ks(K_MC);
        // more chains than the register-resident build can keep on the GPU (2 blocks per SM)?
        if ((int64_t)blocks > 2LL * R.sm_count) mc_chains_kernel<7><<<blocks, kWarpsPerBlock * 32, smem, R.stream>>>(a);
        else mc_chains_kernel<1><<<blocks, kWarpsPerBlock * 32, smem, R.stream>>>(a);
    }
    MMO_LAUNCH_CHECK();
    std::vector<double> hE(n_chains), hP(n_chains), hR((size_t)n_chains * 9), hT((size_t)n_chains * 3), hS((size_t)n_chains * 2);
    std::vector<long long> hC((size_t)n_chains * 8);
    MMO_CUDA(cudaMemcpyAsync(hE.data(), d_bestE.p, hE.size() * 8, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaMemcpyAsync(hP.data(), d_prevE.p, hP.size() * 8, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaMemcpyAsync(hR.data(), d_brot.p, hR.size() * 8, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaMemcpyAsync(hT.data(), d_bpos.p, hT.size() * 8, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaMemcpyAsync(hS.data(), d_steps.p, hS.size() * 8, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaMemcpyAsync(hC.data(), d_cnt.p, hC.size() * 8, cudaMemcpyDeviceToHost, R.stream));
    if (best_xyz) MMO_CUDA(cudaMemcpyAsync(best_xyz, d_bxyz.p, (size_t)n_chains * 3 * L * 8, cudaMemcpyDeviceToHost, R.stream));
    if (trace_chain0) MMO_CUDA(cudaMemcpyAsync(trace_chain0, d_trace.p, (size_t)p->n_steps * 4 * 8, cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaStreamSynchronize(R.stream));
    for (int64_t c = 0; c < n_chains; c++) {
        mmo_mc_result &r = results[c];
        r.best_E = hE[c]; r.prev_E = hP[c];
        for (int k = 0; k < 9; k++) r.best_rot[k] = hR[c * 9 + k];
        for (int k = 0; k < 3; k++) r.best_pos[k] = hT[c * 3 + k];
        r.max_rot = hS[c * 2]; r.max_trans = hS[c * 2 + 1];
        r.n_accept_rigid = hC[c * 8]; r.n_reject_rigid = hC[c * 8 + 1]; r.n_accept_conf = hC[c * 8 + 2];
        r.n_reject_conf = hC[c * 8 + 3]; r.n_ooroi = hC[c * 8 + 4]; r.n_ezero = hC[c * 8 + 5];
        r.too_long = (int32_t)hC[c * 8 + 6]; r.frames_done = (int32_t)hC[c * 8 + 7];
    }
    return MMO_OK;
} MMO_CATCH_ALL
