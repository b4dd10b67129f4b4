// mc.cu -- K6: many independent protein-ligand Monte-Carlo chains per launch.
//
// Replaces the frame loop of Lds.simulate_lig (src/lds.ml:882-995) with the moves of src/move.ml and
// src/mol.ml:593-710 (rand_rot, rand_trans, tweak_rbond / flip_rbond, rotate_bond,
// center_rotate_translate_copy), the interpolated scorer (src/mol.ml:1012-1020, src/G3D.ml:97-157),
// the intra-ligand non-bonded energy (src/mol.ml:881-903), the Metropolis test (lds.ml:931-934), the
// acceptance windows (src/SW.ml) and the adaptive step sizes (lds.ml:586-621).
//
// One block per chain (see "block per chain" below): warp 0 runs the frame loop, its scalars replicated in its
// lanes, the other warps produce the energy terms in parallel, and warp 0 adds them up one by one in the
// reference's order, which keeps every energy bit-identical to the sequential loops of the reference (and of
// oracle/mmo_oracle_mc.c).
// sin/cos/exp and the random stream come from include/mmo_detmath.h on both sides (see there).
// Compiled with -fmad=false.  Reference quirks D1-D6, D14 of SURVEY Appendix D are mirrored.
#include "common.cuh"
#include "strict_dev.cuh"
#include "../../include/mmo_detmath.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace mmo {

constexpr int kBlockSize = 100;      // params.ml:29

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
    const double4 *pair_tab;            // {x_ij, d_ij, q_i q_j, bits(i | j << 16)}: one coalesced load per term
    int n_rbonds;
    const int32_t *rb_left, *rb_right, *rg_off, *rg_idx;
    // interpolated scorer (maps != nullptr) ...
    GridGeom g;
    const float *maps;
    const float2 *zp;                   // the z-pair copy of the maps (mmo_grid::zpair) the look-ups read, zvox elements per type
    size_t zvox;
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
    long long *prof;                    // optional (MMO_MC_PROFILE): chain 0's clock cycles per phase of the frame loop
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

// one term of Mol.ene_intra_UFFNB_brute (mol.ml:881-903): {q_i q_j / r, d_ij (p6^2 - 2 p6)} of interacting pair k
__device__ __forceinline__ double4 ld_tab(const double4 *p) {
    const double2 u = __ldg((const double2 *)p), v = __ldg((const double2 *)p + 1);
    return make_double4(u.x, u.y, v.x, v.y);
}
__device__ __forceinline__ double2 intra_term(const double4 e, const double *x, const double *y, const double *z) {
    const unsigned ij = (unsigned)__double2loint(e.w);
    const int i = (int)(ij & 0xffffu), j = (int)(ij >> 16);
    const double r = d_nzd(sqrt(d_dist2(x[i], y[i], z[i], x[j], y[j], z[j])));
    const Divisor by_r = make_divisor(r);
    const double p6 = d_pow6(div_by(e.x, by_r));                  // UFF.vdW_xiDi: x_ij = sqrt(x_i x_j)
    double2 t;
    t.x = div_by(e.z, by_r);                                      // (q_i *. q_j) /. r
    t.y = e.y * ((-2.0 * p6) + (p6 * p6));
    return t;
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

// ---- block per chain ------------------------------------------------------------------------------------
// One block of NT threads runs one chain.  Warp 0 owns the chain: every scalar of the frame loop (rotation, position,
// energies, RNG counter, windows) lives in its registers, replicated in its 32 lanes, and it alone adds up the energy
// terms, one by one in the reference's order -- which is what keeps every energy bit-identical to the sequential
// loops of the reference and sets the latency floor of a frame (991 dependent double additions for the fixture).
// The other warps ("helpers") only produce terms: after warp 0 has published the trial coordinates (one
// __syncthreads per frame), they compute the intra-ligand pair terms chunk by chunk, all chunks' terms in parallel
// across the block, and signal every finished chunk on its own named barrier (bar.arrive); warp 0 sums chunk c
// (bar.sync on its barrier) while the helpers are already on chunk c + 1.  The trilinear look-ups of E_inter go the
// same way (one per thread of the first warps, barrier 15).  Commands to the helpers are double buffered by sequence
// number, so that warp 0 can post the next one while a helper still reads the last.
constexpr int kMaxChunks = 13;                     // named barriers 1..13 carry the intra chunks, 14 the look-ups, 15 the hand-over
constexpr int kBarLookups = 14, kBarHandOver = 15;  // (barrier 0 stays __syncthreads' own: every thread, one place in the code)
struct McCmd { int src, do_intra, do_inter, quit; };

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct McShared {                                  // pointers into the block's dynamic shared memory
    double *cx, *cy, *cz, *px, *py, *pz, *lx, *ly, *lz, *dr, *drp, *iterms, *keep;
    double2 *terms, *scratch;
    Sw *sw_bond;
    volatile McCmd *cmd;
    int *ltyp, *rb;                                // FF types [L]; rotatable bonds: left[nrb] right[nrb] rg_off[nrb + 1] rg_idx[...]
};

// helper side of one command: the intra chunks (table entry of the next term prefetched while the current one is
// computed), then the look-up terms (threads of the first n_lw helper warps)
template <int NT>
__device__ __forceinline__ void mc_produce(const McArgs &a, const McShared &S, const McCmd c, int tid, int n_lw, int R, int CH) {
    const int h = tid - 32, H = NT - 32;
    const double *x = c.src ? S.cx : S.lx, *y = c.src ? S.cy : S.ly, *z = c.src ? S.cz : S.lz;
    const long long tp0 = clock64();
    if (c.do_intra) {
        // invariant: at the start of chunk r, e = the table entry of this thread's first term of the chunk (if it has one)
        double4 e = make_double4(0.0, 0.0, 0.0, 0.0);
        if (R > 0 && h < min(a.n_pairs, CH)) e = ld_tab(a.pair_tab + h);
        for (int r = 0; r < R; r++) {
            const int k1 = min(a.n_pairs, (r + 1) * CH);
            const int kf = (r + 1) * CH + h;                       // this thread's first term of the next chunk
            const bool has_next = (r + 1 < R) && kf < min(a.n_pairs, (r + 2) * CH);
            int k = r * CH + h;
            if (k >= k1 && has_next) e = ld_tab(a.pair_tab + kf);
            for (; k < k1; k += H) {
                const int kn = k + H;
                double4 en = e;
                if (kn < k1) en = ld_tab(a.pair_tab + kn);
                else if (has_next) en = ld_tab(a.pair_tab + kf);
                S.terms[k] = intra_term(e, x, y, z);
                e = en;
            }
            for (; k < (r + 1) * CH; k += H) S.terms[k] = make_double2(0.0, 0.0);      // padding of the last chunk(s)
            __threadfence_block();
            __syncwarp();                 // bar.arrive is warp-aligned: the lanes' trip counts above differ
            bar_arrive(1 + r, NT);
        }
        if (a.prof && blockIdx.x == 0 && h == 0) a.prof[7] += clock64() - tp0;
    }
    if (c.do_inter && h < n_lw * 32) {
        for (int j = h; j < a.L; j += n_lw * 32)
            S.iterms[j] = d_trilin_zp(a.g, a.zp + (size_t)S.ltyp[j] * a.zvox, x[j], y[j], z[j]);
        __threadfence_block();
        __syncwarp();
        bar_arrive(kBarLookups, n_lw * 32 + 32);
    }
}

// warp 0, before the hand-over barrier: publish what the helpers have to produce
__device__ __forceinline__ void mc_post(const McShared &S, int lane, int &seq, int src, bool do_intra, bool interp, bool quit) {
    if (lane == 0) {
        volatile McCmd *dst = S.cmd + (seq & 1);
        dst->src = src; dst->do_intra = do_intra ? 1 : 0; dst->do_inter = interp ? 1 : 0; dst->quit = quit ? 1 : 0;
    }
    seq++;
    __syncwarp();                                 // lane 0 back with the others: bar.sync is warp-aligned
}

// warp 0, after the hand-over: add everything up in the reference's order as the helpers deliver it
template <int NT, bool PROF>
__device__ __forceinline__ void mc_sum(const McArgs &a, const McShared &S, int lane, int src, bool do_intra, bool do_inter,
                                       int n_lw, int R, int CH, double &E_intra, double &E_inter, long long *pc) {
    const bool interp = do_inter && a.maps != nullptr;
    long long t0 = PROF ? clock64() : 0;
    const double *x = src ? S.cx : S.lx, *y = src ? S.cy : S.ly, *z = src ? S.cz : S.lz;
    if (do_inter && !interp) E_inter = direct_energy(a, x, y, z, lane, S.scratch);
    if (do_intra) {
        double se = 0.0, sv = 0.0;                 // Mol.ene_intra_UFFNB_brute (mol.ml:881-903), pairs in (i<j) order
        for (int r = 0; r < R; r++) {
            const long long tb = PROF ? clock64() : 0;
            bar_sync(1 + r, NT);
            if (PROF) pc[6] += clock64() - tb;
            // Rotating buffer of 8 terms: the load of term k + 8 is issued when term k is consumed, ~66 cycles (8 dependent
            // DADDs) ahead of its own use, so the chain never waits on shared memory.  Every chunk holds CH terms (a
            // multiple of 8): the helpers fill the slots beyond the last pair with +0.0, which leaves the (never -0.0)
            // running sums unchanged.
            const double2 *tk = S.terms + r * CH;
            double2 u[8];
#pragma unroll
            for (int q = 0; q < 8; q++) u[q] = tk[q];
            for (int k = 8; k < CH; k += 8) {
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    se = se + u[q].x;
                    sv = sv + u[q].y;
                    u[q] = tk[k + q];
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++) { se = se + u[q].x; sv = sv + u[q].y; }
        }
        E_intra = (kElecWeight * se) + sv;
        if (PROF) { const long long t = clock64(); pc[4] += t - t0; t0 = t; }
    }
    if (interp) {
        bar_sync(kBarLookups, n_lw * 32 + 32);
        double res = 0.0;                          // Mol.ene_inter_UFF_interp (mol.ml:1012-1020): res := !res +. trilin ...
        int j = 0;
        for (; j + 8 <= a.L; j += 8) {
            double u[8];
#pragma unroll
            for (int q = 0; q < 8; q++) u[q] = S.iterms[j + q];
#pragma unroll
            for (int q = 0; q < 8; q++) res = res + u[q];
        }
        for (; j < a.L; j++) res = res + S.iterms[j];
        E_inter = res;
    }
    if (PROF) pc[3] += clock64() - t0;
}

template <int NT, int MINB, bool PROF>
__global__ void __launch_bounds__(NT, MINB)
mc_chain_kernel(McArgs a) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t chain = blockIdx.x;
    const int L = a.L, nrb = a.n_rbonds, nrb1 = max(nrb, 1);
    // chunks of the intra terms (one named barrier each)
    const int H = NT - 32;
    const int R = a.n_pairs > 0 ? min(kMaxChunks, (a.n_pairs + H - 1) / H) : 0;
    const int CH = R > 0 ? (((a.n_pairs + R - 1) / R + 7) & ~7) : 8;      // a multiple of 8 (warp 0's rotating buffer)
    McShared S;
    S.cx = smem; S.cy = S.cx + L; S.cz = S.cy + L;                                  // conf
    S.px = S.cz + L; S.py = S.px + L; S.pz = S.py + L;                              // conf' (proposed)
    S.lx = S.pz + L; S.ly = S.lx + L; S.lz = S.ly + L;                              // lig'
    S.dr = S.lz + L; S.drp = S.dr + nrb1;                                           // per-bond step sizes of conf / conf'
    S.iterms = S.drp + nrb1;                                                        // L look-up terms
    S.keep = S.iterms + L;                                                          // rot0[9] pos0[3] best_rot[9] best_pos[3]
    double *end = S.keep + 24;
    end += ((size_t)(end - smem) & 1);                                              // 16-byte alignment
    S.scratch = (double2 *)end;                                                     // 32 slots (direct scorer)
    S.terms = S.scratch + 32;                                                       // n_pairs intra terms
    S.sw_bond = (Sw *)(S.terms + max(R * CH, 1));
    S.cmd = (volatile McCmd *)(S.sw_bond + nrb1);
    S.ltyp = (int *)(S.cmd + 2);
    S.rb = S.ltyp + L;
    const int n_lw = min(NT / 32 - 1, (L + 31) / 32);      // helper warps that share the look-ups
    // constant tables of the chain in shared memory (global loads would sit on warp 0's critical path)
    const int rg_total = nrb > 0 ? __ldg(a.rg_off + nrb) : 0;
    for (int j = tid; j < L; j += NT) S.ltyp[j] = a.maps ? __ldg(a.ltyp + j) : 0;
    for (int b = tid; b < nrb; b += NT) { S.rb[b] = __ldg(a.rb_left + b); S.rb[nrb + b] = __ldg(a.rb_right + b); }
    for (int b = tid; b <= nrb && nrb > 0; b += NT) S.rb[2 * nrb + b] = __ldg(a.rg_off + b);
    for (int g = tid; g < rg_total; g += NT) S.rb[3 * nrb + 1 + g] = __ldg(a.rg_idx + g);
    __syncthreads();
    const int *const s_left = S.rb, *const s_right = S.rb + nrb, *const s_rgoff = S.rb + 2 * nrb, *const s_rgidx = S.rb + 3 * nrb + 1;

    // ---- every warp runs the loop below and meets the others at ONE hand-over barrier per evaluation; warp 0 (w0) owns the
    //      chain, the other warps only produce terms.  The state declared here lives in warp 0's registers (the helpers'
    //      copies are dead).  phase 0: constant E_intra of the rigid ligand, 1: start energies, 2: the frames.
    const bool w0 = wid == 0;
    int seq = 0;
    double *const cx = S.cx, *const cy = S.cy, *const cz = S.cz, *const px = S.px, *const py = S.py, *const pz = S.pz;
    double *const lx = S.lx, *const ly = S.ly, *const lz = S.lz, *const dr = S.dr, *const drp = S.drp;
    double *const rot0 = S.keep, *const pos0 = S.keep + 9, *const best_rot = S.keep + 12, *const best_pos = S.keep + 21;
    Sw *const sw_bond = S.sw_bond;
    const bool flexible = a.tweak_rbonds && nrb > 0;
    // frame counters fit 32 bits (n_steps is an int32): 32-bit remainders instead of the 64-bit software division
    const int rbf = a.no_flip ? 0x7fffffff : kBlockSize;
    const double target_low = 0.5 - 0.05, target_high = 0.5 + 0.05;    // lds.ml:651-652
    const uint64_t seed = a.seeds[chain];
    uint64_t ctr = 0;

    if (w0) {
        for (int j = lane; j < L; j += 32) { cx[j] = a.lx[j]; cy[j] = a.ly[j]; cz[j] = a.lz[j]; }
        for (int b = lane; b < nrb; b += 32) { dr[b] = a.p_max_rbond_rot; sw_reset(sw_bond[b]); }
    }
    double ccen[3] = {0.0, 0.0, 0.0}, pcen[3] = {0.0, 0.0, 0.0};        // conf.center, conf'.center
    Sw sw_rigid;
    sw_reset(sw_rigid);
    double max_rot = a.p_max_rot, max_trans = a.p_max_trans;
    double rot[9], pos[3];
#pragma unroll
    for (int k = 0; k < 9; k++) rot[k] = a.rot0[chain * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = a.pos0[chain * 3 + k];
    double *bxyz = a.best_xyz + chain * 3 * (int64_t)L;
    if (w0) {
        if (lane < 9) { rot0[lane] = a.rot0[chain * 9 + lane]; best_rot[lane] = (lane % 4 == 0) ? 1.0 : 0.0; }
        if (lane < 3) { pos0[lane] = a.pos0[chain * 3 + lane]; best_pos[lane] = 0.0; }
        __syncwarp();
        // start_conf = rotate_then_translate_copy centered_lig rot0 pos0 (lds.ml:758)
        for (int j = lane; j < L; j += 32) {
            double x, y, z;
            rot_apply(rot, cx[j], cy[j], cz[j], x, y, z);
            lx[j] = x + pos[0]; ly[j] = y + pos[1]; lz[j] = z + pos[2];
            bxyz[j] = lx[j]; bxyz[L + j] = ly[j]; bxyz[2 * L + j] = lz[j];
        }
    }
    double const_intra = 0.0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // cycles: 0 conformer move, 1 rigid move + lig', 2 hand-over, 3 E_inter, 4 E_intra sum, 5 accept/bookkeeping
    double prev_E_intra = 0.0, prev_E_inter = 0.0, prev_E = 0.0, best_E = 0.0;
    int rigid_step = 0, conf_step = 0;
    int n_acc_r = 0, n_rej_r = 0, n_acc_c = 0, n_rej_c = 0, n_ooroi = 0, n_ezero = 0, too_long = 0;
    int frame = 0;
    int phase = (a.intra_nb && !flexible) ? 0 : 1;
    // per-frame state of warp 0 that lives across the hand-over
    bool rigid = true;
    int just_rotated = -1, which = 0;                   // conf' is: 0 = conf, 1 = proposed copy, 2 = centred template
    double rotp[9], posp[3];
    long long tf0 = 0;
    int cmd_src = 0;
    bool cmd_intra = false, cmd_inter = false;
    for (;;) {
      bool quit = false;
      if (w0) {
        // ---- warp 0 before the hand-over: what is evaluated next, on which coordinates ----
        if (phase == 0) { cmd_src = 1; cmd_intra = true; cmd_inter = false; }
        else if (phase == 1) { cmd_src = 0; cmd_intra = a.intra_nb && flexible; cmd_inter = true; }
        else if (frame >= a.n_steps) quit = true;
        else {
        rigid = (frame & 1) == 0;
        just_rotated = -1;
        which = 0;
        if (PROF) tf0 = clock64();
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
                const int left = s_left[bond], right = s_right[bond];
                const double ox = px[right], oy = py[right], oz = pz[right];
                const double ax = ox - px[left], ay = oy - py[left], az = oz - pz[left];
                const double mag = sqrt(ax * ax + ay * ay + az * az);
                double br[9];
                {
                    const Divisor by_mag = make_divisor(mag);      // three IEEE quotients by the same |v| (V3.normalize), bit-identical
                    const double ux = div_by(ax, by_mag), uy = div_by(ay, by_mag), uz = div_by(az, by_mag);
                    double s, c;
                    mmo_det_sincos(alpha, &s, &c);
                    const double omc = 1.0 - c;                  // rot.ml:136-146
                    br[0] = c + ux * ux * omc; br[1] = ux * uy * omc - uz * s; br[2] = ux * uz * omc + uy * s;
                    br[3] = ux * uy * omc + uz * s; br[4] = c + uy * uy * omc; br[5] = uy * uz * omc - ux * s;
                    br[6] = ux * uz * omc - uy * s; br[7] = uy * uz * omc + ux * s; br[8] = c + uz * uz * omc;
                }
                __syncwarp();
                const int g0 = s_rgoff[bond], g1 = s_rgoff[bond + 1];
                for (int g = g0 + lane; g < g1; g += 32) {
                    const int i = s_rgidx[g];
                    double x, y, z;
                    rot_apply(br, px[i] - ox, py[i] - oy, pz[i] - oz, x, y, z);
                    px[i] = x + ox; py[i] = y + oy; pz[i] = z + oz;
                }
                __syncwarp();
                favg3_smem(px, py, pz, L, pcen);                                                       // update_center
                just_rotated = bond;
                // Mol.check_elongation_exn lig 12.0 (mol.ml:576-591): maxi = max_j (0.01 +. dist center xyz_j).  sqrt and +. are
                // monotonic, so the maximum is taken over the squared distances and one square root gives the same double
                double maxd2 = 0.0;
                for (int j = lane; j < L; j += 32) maxd2 = fmax(maxd2, d_dist2(pcen[0], pcen[1], pcen[2], px[j], py[j], pz[j]));
                for (int o = 16; o > 0; o >>= 1) maxd2 = fmax(maxd2, __shfl_xor_sync(0xffffffffu, maxd2, o));
                const double maxi = 0.01 + sqrt(maxd2);
                if (maxi > 12.0) { too_long = 1; quit = true; }    // Mol.Too_long ends this run (lds.ml:996-997)
                which = 1;
            } else {
                which = 2;
            }
        }
        if (PROF) { const long long t = clock64(); pc[0] += t - tf0; tf0 = t; }
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
        // D2: a conformer frame overwrites prev_E_intra with the trial's value, accepted or not
        if (!rigid && a.intra_nb && !flexible) prev_E_intra = const_intra;
        if (PROF) { const long long t = clock64(); pc[1] += t - tf0; }
        cmd_src = 0; cmd_intra = !rigid && a.intra_nb && flexible; cmd_inter = true;
        }
        if (PROF) tf0 = clock64();
        mc_post(S, lane, seq, cmd_src, cmd_intra, cmd_inter && a.maps != nullptr, quit);
      }
      // ---- the hand-over: ONE barrier instruction for every warp of the block (coordinates and command visible) ----
      bar_sync(kBarHandOver, NT);
      if (!w0) {
        const volatile McCmd *src = S.cmd + (seq & 1);
        McCmd c;
        c.src = src->src; c.do_intra = src->do_intra; c.do_inter = src->do_inter; c.quit = src->quit;
        seq++;
        if (c.quit) break;
        mc_produce<NT>(a, S, c, tid, n_lw, R, CH);
        continue;
      }
      if (quit) break;
      if (PROF) pc[2] += clock64() - tf0;
      // ---- warp 0 after the hand-over: the sums, then the bookkeeping of the phase ----
      if (phase == 0) {
        double dummy = 0.0;
        mc_sum<NT, PROF>(a, S, lane, 1, true, false, n_lw, R, CH, const_intra, dummy, pc);
        prev_E_intra = const_intra;
        phase = 1;
        continue;
      }
      mc_sum<NT, PROF>(a, S, lane, 0, cmd_intra, true, n_lw, R, CH, prev_E_intra, prev_E_inter, pc);
      if (phase == 1) {
        prev_E = prev_E_inter + prev_E_intra;
        best_E = prev_E;
#pragma unroll
        for (int k = 0; k < 8; k++) pc[k] = 0;
        phase = 2;
        continue;
      }
      {
        if (PROF) tf0 = clock64();
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
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < 9; k++) best_rot[k] = rotp[k];
#pragma unroll
                        for (int k = 0; k < 3; k++) best_pos[k] = posp[k];
                    }
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
                if (flexible && !rigid && conf_step > 0 && conf_step % (kBlockSize * nrb) == 0) {
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
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 9; k++) rot[k] = rot0[k];
#pragma unroll
            for (int k = 0; k < 3; k++) pos[k] = pos0[k];
            __syncwarp();
            if (lane < 9) best_rot[lane] = (lane % 4 == 0) ? 1.0 : 0.0;
            if (lane < 3) best_pos[lane] = 0.0;
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
        if (PROF) pc[5] += clock64() - tf0;
        frame++;
      }
    }
    if (!w0) return;
    if (PROF && a.prof && chain == 0 && lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; k++) a.prof[k] = pc[k];
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

// ---- warp per chain -------------------------------------------------------------------------------------
// The round-1 kernel, kept for launches with thousands of chains per GPU: one warp does everything a block does above
// (terms 32 at a time, then the same in-order sum), 20 us per frame instead of 7, but 28 chains are resident per SM
// instead of 4, which wins once the chains outnumber what the block kernel can keep on the GPU.  Same arithmetic, same
// results.
constexpr int kWarpsPerBlock = 4;
constexpr int kTermDoubles = 128;      // per warp: two buffers of 32 double2 term slots
__device__ __forceinline__ double2 w_intra_term(const McArgs &a, const double *x, const double *y, const double *z, int k) {
    if (k >= a.n_pairs) return make_double2(0.0, 0.0);
    return intra_term(ld_tab(a.pair_tab + k), x, y, z);
}
// Software pipelined: the terms of round r + 1 are computed (sqrt and division chains) between the store of round r
// and its in-order summation, two independent dependency chains the scheduler can interleave; two term buffers.
__device__ double w_intra_energy(const McArgs &a, const double *x, const double *y, const double *z, int lane,
                               double2 *terms) {
    double se = 0.0, sv = 0.0;
    double2 t = w_intra_term(a, x, y, z, lane);
    int buf = 0;
    for (int base = 0; base < a.n_pairs; base += 32) {
        double2 *tb = terms + 32 * buf;
        __syncwarp();
        tb[lane] = t;
        __syncwarp();
        t = w_intra_term(a, x, y, z, base + 32 + lane);                 // next round (zeros beyond the last pair)
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
__device__ double w_interp_energy(const McArgs &a, const double *x, const double *y, const double *z, int lane,
                                double2 *terms) {
    double res = 0.0;
    for (int base = 0; base < a.L; base += 32) {
        const int j = base + lane;
        double t = 0.0;
        if (j < a.L) t = d_trilin_zp(a.g, a.zp + (size_t)__ldg(a.ltyp + j) * a.zvox, x[j], y[j], z[j]);
        __syncwarp();
        terms[lane].x = t;
        __syncwarp();
        const int lim = min(32, a.L - base);
        for (int l = 0; l < lim; l++) res = res + terms[l].x;
    }
    return res;
}

__device__ __forceinline__ double w_inter_energy(const McArgs &a, const double *x, const double *y, const double *z,
                                               int lane, double2 *terms) {
    return a.maps ? w_interp_energy(a, x, y, z, lane, terms) : direct_energy(a, x, y, z, lane, terms);
}

// Two builds of the same kernel: MINB = 1 keeps every chain's state in registers (248: lowest latency per frame,
// at most 8 chains per SM) for launches that do not fill the GPU anyway; MINB = 7 caps the registers at 72 (state
// spills to local memory, slower per chain) so that 28 chains are resident per SM when there are thousands.
template <int MINB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
mc_warp_kernel(McArgs a) {
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
    if (a.intra_nb && !flexible) const_intra = w_intra_energy(a, cx, cy, cz, lane, terms);
    double prev_E_intra = !a.intra_nb ? 0.0 : (flexible ? w_intra_energy(a, lx, ly, lz, lane, terms) : const_intra);
    double prev_E_inter = w_inter_energy(a, lx, ly, lz, lane, terms);
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
        if (!rigid && a.intra_nb) prev_E_intra = flexible ? w_intra_energy(a, lx, ly, lz, lane, terms) : const_intra;   // D2
        prev_E_inter = w_inter_energy(a, lx, ly, lz, lane, terms);
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
    if (lig->n_pairs > 0 && !lig->mc_pair_tab.p) {
        // per interacting pair: UFF.vdW_xiDi of the two elements (x_ij, d_ij), q_i *. q_j, and the two atom indices
        MMO_REQUIRE(L <= 65535, "mmo_mc_run: ligand too large (%d atoms)", L);
        std::vector<double4> tab((size_t)lig->n_pairs);
        for (int k = 0; k < lig->n_pairs; k++) {
            const int i = lig->h_pair_i[k], j = lig->h_pair_j[k];
            const int ei = elt_index(lig->hanum[i]), ej = elt_index(lig->hanum[j]);
            const bool ok = ei < kNumElt && ej < kNumElt;
            tab[k].x = ok ? sqrt(kEltXi[ei] * kEltXi[ej]) : NAN;       // FF.geo_mean (same doubles as the kEltTab^2 tables)
            tab[k].y = ok ? sqrt(kEltDi[ei] * kEltDi[ej]) : NAN;
            tab[k].z = lig->hq[i] * lig->hq[j];
            const long long bits = (long long)((unsigned)i | ((unsigned)j << 16));
            memcpy(&tab[k].w, &bits, sizeof bits);
        }
        MMO_TRY(const_cast<mmo_ligand *>(lig)->mc_pair_tab.upload(tab));
    }
    a.pair_tab = lig->mc_pair_tab.p;
    a.n_rbonds = lig->n_rbonds; a.rb_left = lig->d_rb_left.p; a.rb_right = lig->d_rb_right.p;
    a.rg_off = lig->d_rg_off.p; a.rg_idx = lig->d_rg_idx.p;
    a.zp = nullptr; a.zvox = 0;
    if (grid) {
        a.g = geom_of(grid); a.maps = grid->maps.p;
        MMO_TRY(grid_zpairs(grid, &a.zp));
        a.zvox = (size_t)grid->dims[0] * grid->dims[1] * (grid->dims[2] - 1);
    } else { memset(&a.g, 0, sizeof a.g); a.maps = nullptr; }
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
    DevBuf<long long> d_prof;
    const bool want_prof = getenv("MMO_MC_PROFILE") != nullptr;
    if (want_prof) { MMO_TRY(d_prof.alloc(8)); MMO_CUDA(cudaMemsetAsync(d_prof.p, 0, 64, R.stream)); }
    a.prof = want_prof ? d_prof.p : nullptr;
    const int nrb1 = std::max(lig->n_rbonds, 1);
    // conf, conf', lig' (9 L), step sizes (2 nrb), look-up terms (L), kept scalars (24), alignment (1), direct-scorer
    // scratch (32 double2), intra terms (n_pairs double2), per-bond windows, two command slots
    const size_t smem = ((size_t)10 * L + 2 * nrb1 + 24 + 1) * sizeof(double) + ((size_t)32 + lig->n_pairs + 13 * 9 + 16) * sizeof(double2) +
                        (size_t)nrb1 * sizeof(Sw) + 2 * sizeof(McCmd) +
                        ((size_t)L + 3 * (size_t)lig->n_rbonds + 1 + lig->rg_idx.size()) * sizeof(int) + 16;
    MMO_REQUIRE(smem <= 200 * 1024, "mmo_mc_run: ligand too large (%d atoms, %d rotatable bonds, %d interacting pairs)", L, lig->n_rbonds, lig->n_pairs);
    // One block of 128 threads per chain while the block kernel can keep (nearly) all chains on the GPU at once (4 blocks per
    // SM): lowest latency per frame.  Beyond ~12 chains per SM the warp-per-chain kernel (28 chains resident per SM) has
    // the higher throughput.  MMO_MC_THREADS = 32 | 64 | 128 | 256 overrides (32 = warp per chain).
    int nt = (n_chains > 12LL * R.sm_count) ? 32 : 128;
    if (const char *e = getenv("MMO_MC_THREADS")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) nt = v; }
    {
        KernelScope ks(K_MC);
        if (nt == 32) {
            const size_t wsmem = (size_t)kWarpsPerBlock * (((9 * L + 2 * nrb1 + 1) & ~1) + kTermDoubles) * sizeof(double) + (size_t)kWarpsPerBlock * nrb1 * sizeof(Sw);
            MMO_REQUIRE(wsmem <= 200 * 1024, "mmo_mc_run: ligand too large (%d atoms, %d rotatable bonds)", L, lig->n_rbonds);
            MMO_CUDA(cudaFuncSetAttribute(mc_warp_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            const unsigned blocks = (unsigned)((n_chains + kWarpsPerBlock - 1) / kWarpsPerBlock);
            mc_warp_kernel<7><<<blocks, kWarpsPerBlock * 32, wsmem, R.stream>>>(a);
        } else if (nt == 64) {
            MMO_CUDA(cudaFuncSetAttribute(mc_chain_kernel<64, 10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mc_chain_kernel<64, 10, false><<<(unsigned)n_chains, 64, smem, R.stream>>>(a);
        } else if (nt == 128 && want_prof) {       // MMO_MC_PROFILE: the build with cycle counters around every phase
            MMO_CUDA(cudaFuncSetAttribute(mc_chain_kernel<128, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mc_chain_kernel<128, 4, true><<<(unsigned)n_chains, 128, smem, R.stream>>>(a);
        } else if (nt == 128) {
            MMO_CUDA(cudaFuncSetAttribute(mc_chain_kernel<128, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mc_chain_kernel<128, 4, false><<<(unsigned)n_chains, 128, smem, R.stream>>>(a);
        } else {
            MMO_CUDA(cudaFuncSetAttribute(mc_chain_kernel<256, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mc_chain_kernel<256, 2, false><<<(unsigned)n_chains, 256, smem, R.stream>>>(a);
        }
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
    if (want_prof) {
        long long h[8];
        MMO_CUDA(cudaMemcpy(h, d_prof.p, sizeof h, cudaMemcpyDeviceToHost));
        const double f = std::max(1, p->n_steps);
        fprintf(stderr, "[mmo_mc_run] chain 0, cycles per frame (%d threads per chain, %lld chains): conformer move %.0f, rigid move + lig' %.0f, "
                        "hand-over %.0f, E_inter %.0f, E_intra sum %.0f (of which waiting at chunk barriers %.0f; helper thread 0 produces for %.0f), accept %.0f\n",
                nt, (long long)n_chains, h[0] / f, h[1] / f, h[2] / f, h[3] / f, h[4] / f, h[6] / f, h[7] / f, h[5] / f);
    }
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
