/* mmo_oracle_mc.c -- CPU restatement of the protein-ligand Monte-Carlo frame loop.
 *
 * TEST INFRASTRUCTURE ONLY (see mmo_oracle.h).  PARITY UNPINNED.
 *
 * Follows Lds.simulate_lig (src/lds.ml:741-1000) statement by statement, including the behaviours
 * listed in SURVEY Appendix D:
 *   D1  dangling else: the Metropolis block only runs under --hard-ROI (lds.ml:908-990)
 *   D2  prev_E_intra is overwritten by the proposed conformer's value, never restored (900-905)
 *   D3  best_* is updated from the trial regardless of acceptance (959-969)
 *   D4  per-bond acceptance windows are shared between molecule copies, per-bond step sizes are
 *       deep-copied: step-size updates made on a rejected conformer are lost (976-988, mol.ml:57-62)
 *   D5  Mol.Too_long ends the run of that ligand (794, 996-997)
 *   D6  E_inter = 0.0 exactly resets the run (920-927)
 *   D14 draws inside one frame happen in OCaml's right-to-left order: translation (z, y, x) before
 *       rotation (theta, axis) (893-894, move.ml:49-52)
 * Two things are deliberately NOT the reference's: the random stream (OCaml's Random.State is
 * version dependent, RNG.ml:9-16) and libm's sin/cos/exp; both are replaced by include/mmo_detmath.h
 * so that the CUDA chains can be compared with this code bit for bit.
 */
#include "mmo_oracle.h"
#include "../include/mmo_detmath.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_SIZE 100   /* params.ml:29 */

/* SW.ml: sliding window of the last `size` accept/reject events */
typedef struct { int ev[BLOCK_SIZE]; int head, n, accepts, rejects; } sw_t;
static void sw_reset(sw_t *s) { s->head = 0; s->n = 0; s->accepts = 0; s->rejects = 0; }
static void sw_process(sw_t *s, int evt) {            /* SW.ml:20-34 */
    if (s->n == BLOCK_SIZE) {                          /* queue longer than size after the push: pop oldest */
        int old = s->ev[s->head];
        if (old) s->accepts--; else s->rejects--;
        s->ev[s->head] = evt;
        s->head = (s->head + 1) % BLOCK_SIZE;
    } else {
        s->ev[(s->head + s->n) % BLOCK_SIZE] = evt;
        s->n++;
    }
    if (evt) s->accepts++; else s->rejects++;
}
static double sw_ratio(const sw_t *s) { return (double)s->accepts / (double)(s->accepts + s->rejects); } /* SW.ml:36-37 */

typedef struct { uint64_t seed, ctr; } rng_t;
static double rng_float(rng_t *r, double scale) { return scale * mmo_rng_uniform(r->seed, r->ctr++); } /* Random.State.float */
static int rng_int(rng_t *r, int n) { int v = (int)(mmo_rng_uniform(r->seed, r->ctr++) * (double)n); return v >= n ? n - 1 : v; }

double orc_rng_uniform(uint64_t seed, uint64_t counter) { return mmo_rng_uniform(seed, counter); }

/* rot.ml:22-46 with the deterministic sin/cos */
static void det_rot_axis(int axis, double th, double r[9]) {
    double s, c;
    mmo_det_sincos(th, &s, &c);
    if (axis == 0) { r[0] = 1.0; r[1] = 0.0; r[2] = 0.0; r[3] = 0.0; r[4] = c; r[5] = s; r[6] = 0.0; r[7] = -s; r[8] = c; }
    else if (axis == 1) { r[0] = c; r[1] = 0.0; r[2] = -s; r[3] = 0.0; r[4] = 1.0; r[5] = 0.0; r[6] = s; r[7] = 0.0; r[8] = c; }
    else { r[0] = c; r[1] = s; r[2] = 0.0; r[3] = -s; r[4] = c; r[5] = 0.0; r[6] = 0.0; r[7] = 0.0; r[8] = 1.0; }
}
/* rot.ml:136-146 */
static void det_rot_of_axis_angle(double x, double y, double z, double theta, double r[9]) {
    double s, c;
    mmo_det_sincos(theta, &s, &c);
    double omc = 1.0 - c;
    r[0] = c + x * x * omc; r[1] = x * y * omc - z * s; r[2] = x * z * omc + y * s;
    r[3] = x * y * omc + z * s; r[4] = c + y * y * omc; r[5] = y * z * omc - x * s;
    r[6] = x * z * omc - y * s; r[7] = y * z * omc + x * s; r[8] = c + z * z * omc;
}

typedef struct {
    double *x, *y, *z;     /* L */
    double center[3];
    double *dr;            /* n_rbonds, deep-copied with the molecule (mol.ml:57-62) */
} conf_t;

static conf_t conf_alloc(int L, int nrb) {
    conf_t c;
    c.x = (double *)malloc(sizeof(double) * (size_t)(3 * L + (nrb > 0 ? nrb : 1)));
    c.y = c.x + L; c.z = c.y + L; c.dr = c.z + L;
    c.center[0] = c.center[1] = c.center[2] = 0.0;
    return c;
}
static void conf_copy(conf_t *d, const conf_t *s, int L, int nrb) {
    memcpy(d->x, s->x, sizeof(double) * (size_t)(3 * L + (nrb > 0 ? nrb : 1)));
    memcpy(d->center, s->center, sizeof d->center);
}

/* mol.ml:610-631 rotate_bond on a conformer (deterministic sin/cos), then update_center (353-356) */
static void conf_rotate_bond(conf_t *m, int L, int left, int right, int ng, const int32_t *grp, double alpha) {
    double cx = m->x[right], cy = m->y[right], cz = m->z[right];
    double ax = cx - m->x[left], ay = cy - m->y[left], az = cz - m->z[left];
    double mag = sqrt(ax * ax + ay * ay + az * az);
    double rot[9];
    det_rot_of_axis_angle(ax / mag, ay / mag, az / mag, alpha, rot);
    for (int g = 0; g < ng; g++) {
        int i = grp[g];
        double v[3] = {m->x[i] - cx, m->y[i] - cy, m->z[i] - cz}, o[3];
        orc_rot_rotate(rot, v, o);
        m->x[i] = o[0] + cx; m->y[i] = o[1] + cy; m->z[i] = o[2] + cz;
    }
    m->center[0] = orc_favg(L, m->x); m->center[1] = orc_favg(L, m->y); m->center[2] = orc_favg(L, m->z);
}

static double inter_energy(const orc_mc_args *a, const double *x, const double *y, const double *z) {
    if (a->scorer == 2) return orc_ene_inter_interp(a->g_step, a->g_dims, a->maps, a->L, x, y, z, a->ltyp);
    return orc_ene_inter_shifted_brute(a->P, a->px, a->py, a->pz, a->pq, a->panum, a->L, x, y, z, a->lq, a->lanum);
}

void orc_mc_run(const orc_mc_args *a, orc_mc_result *res, double *best_xyz, double *trace) {
    const int L = a->L, nrb = a->n_rbonds;
    const int flexible = a->tweak_rbonds && nrb > 0;
    const double pi = 4.0 * atan(1.0);                          /* math.ml:13 */
    const double p_max_rot = 15.0 * (pi / 180.0);               /* params.ml:11 */
    const double p_max_trans = 0.15;                            /* params.ml:14 */
    const double p_max_rbond_rot = 5.0 * (pi / 180.0);          /* params.ml:17 */
    const double p_max_rbond_flip = pi;                         /* params.ml:20 */
    const long rbf = a->no_flip ? 0x7fffffffffffffffL : BLOCK_SIZE;   /* params.ml:33, lds.ml:1826 */
    const double target_low = 0.5 - 0.05, target_high = 0.5 + 0.05;   /* lds.ml:651-652 */
    rng_t rng = {a->seed, 0};

    conf_t centered = conf_alloc(L, nrb), conf = conf_alloc(L, nrb), confp = conf_alloc(L, nrb);
    memcpy(centered.x, a->lx, sizeof(double) * L); memcpy(centered.y, a->ly, sizeof(double) * L);
    memcpy(centered.z, a->lz, sizeof(double) * L);
    for (int i = 0; i < nrb; i++) centered.dr[i] = p_max_rbond_rot;      /* mol.ml:397 */
    conf_copy(&conf, &centered, L, nrb);                                 /* lds.ml:763 */
    sw_t sw_rigid;
    sw_reset(&sw_rigid);
    sw_t *sw_bond = (sw_t *)malloc(sizeof(sw_t) * (size_t)(nrb > 0 ? nrb : 1));
    for (int i = 0; i < nrb; i++) sw_reset(&sw_bond[i]);

    double max_rot = p_max_rot, max_trans = p_max_trans;
    double rot[9], pos[3], best_rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, best_pos[3] = {0, 0, 0};
    memcpy(rot, a->rot0, sizeof rot); memcpy(pos, a->pos0, sizeof pos);
    double *lx = (double *)malloc(sizeof(double) * 6 * (size_t)L), *ly = lx + L, *lz = ly + L;
    double *bx = lz + L, *by = bx + L, *bz = by + L;
    /* start_conf = rotate_then_translate_copy centered_lig rot0 pos0 (lds.ml:758) */
    orc_rotate_then_translate(L, centered.x, centered.y, centered.z, rot, pos, lx, ly, lz);
    memcpy(bx, lx, sizeof(double) * 3 * (size_t)L);
    /* ene_intra_builder (lds.ml:702-712) */
    double const_intra = 0.0;
    if (a->intra_nb && !flexible)
        const_intra = orc_ene_intra_uffnb_brute(L, centered.x, centered.y, centered.z, a->lq, a->lanum, a->dists);
    double prev_E_intra = !a->intra_nb ? 0.0 : (flexible ? orc_ene_intra_uffnb_brute(L, lx, ly, lz, a->lq, a->lanum, a->dists) : const_intra);
    double prev_E_inter = inter_energy(a, lx, ly, lz);
    double prev_E = prev_E_inter + prev_E_intra;
    double best_E = prev_E;
    long rigid_step = 0, conf_step = 0;
    memset(res, 0, sizeof *res);
    int frame = 0;
    for (; frame < a->n_steps; frame++) {
        const int rigid = (frame % 2) == 0;
        int just_rotated = -1;
        const conf_t *cp = &conf;                       /* conf' */
        if (!rigid) {                                   /* lds.ml:781-798 rotate_bond closure */
            if (flexible) {
                conf_copy(&confp, &conf, L, nrb);
                int bond;
                double alpha;
                if (conf_step > 0 && conf_step % rbf == 0) {        /* Mol.flip_rbond */
                    bond = rng_int(&rng, nrb);
                    alpha = rng_float(&rng, 2.0 * p_max_rbond_flip) - p_max_rbond_flip;
                } else {                                            /* Mol.tweak_rbond */
                    bond = rng_int(&rng, nrb);
                    double dr = confp.dr[bond];
                    alpha = rng_float(&rng, 2.0 * dr) - dr;
                }
                conf_rotate_bond(&confp, L, a->rb_left[bond], a->rb_right[bond],
                                 a->rg_off[bond + 1] - a->rg_off[bond], a->rg_idx + a->rg_off[bond], alpha);
                just_rotated = bond;
                if (orc_radius(L, confp.x, confp.y, confp.z, confp.center) > 12.0) {   /* Mol.Too_long */
                    res->too_long = 1;
                    break;
                }
                cp = &confp;
            } else {
                cp = &centered;                         /* (fun _ _ _ -> (-1, centered_lig)) */
            }
        }
        double rotp[9], posp[3];
        memcpy(rotp, rot, sizeof rot); memcpy(posp, pos, sizeof pos);
        if (rigid) {
            /* OCaml evaluates the tuple right to left: rand_trans first, its V3.make arguments z, y, x */
            double dz = rng_float(&rng, 2.0) - 1.0;
            double dy = rng_float(&rng, 2.0) - 1.0;
            double dx = rng_float(&rng, 2.0) - 1.0;
            posp[0] = pos[0] + dx * max_trans; posp[1] = pos[1] + dy * max_trans; posp[2] = pos[2] + dz * max_trans;
            double theta = rng_float(&rng, 2.0 * max_rot) - max_rot;     /* move.ml:14-15 */
            int axis = rng_int(&rng, 3);
            double rb[9];
            det_rot_axis(axis, theta, rb);
            orc_rot_mult(rb, rot, rotp);                                  /* move.ml:31 */
        }
        /* lig' = center_rotate_translate_copy conf' rot' pos' (mol.ml:705-710) */
        orc_center_rotate_translate(L, cp->x, cp->y, cp->z, cp->center, rotp, posp, lx, ly, lz);
        if (!rigid && a->intra_nb)
            prev_E_intra = flexible ? orc_ene_intra_uffnb_brute(L, lx, ly, lz, a->lq, a->lanum, a->dists) : const_intra;
        prev_E_inter = inter_energy(a, lx, ly, lz);
        double curr_E = prev_E_inter + prev_E_intra;
        int accepted = -1;
        /* lig'.center = origin + pos' (mol.ml:698-710) */
        double ddx = a->roi_c[0] - (0.0 + posp[0]), ddy = a->roi_c[1] - (0.0 + posp[1]), ddz = a->roi_c[2] - (0.0 + posp[2]);
        double dist_roi = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        int do_reset = 0;
        if (a->hard_roi) {
            if (dist_roi > a->roi_r) {
                do_reset = 1;
                res->n_ooroi++;
            } else if (prev_E_inter == 0.0) {
                do_reset = 1;
                res->n_ezero++;
            } else {
                int acc = (curr_E <= prev_E);
                if (!acc) acc = rng_float(&rng, 1.0) < mmo_det_exp((-(curr_E - prev_E)) * a->beta);
                accepted = acc;
                if (rigid) { sw_process(&sw_rigid, acc); if (acc) res->n_accept_rigid++; else res->n_reject_rigid++; }
                else if (just_rotated > -1) { sw_process(&sw_bond[just_rotated], acc); }
                if (!rigid) { if (acc) res->n_accept_conf++; else res->n_reject_conf++; }
                if (acc) {
                    memcpy(rot, rotp, sizeof rot); memcpy(pos, posp, sizeof pos);
                    prev_E = curr_E;
                    if (cp == &confp) conf_copy(&conf, &confp, L, nrb);       /* conf := conf' */
                    else if (cp == &centered) conf_copy(&conf, &centered, L, nrb);
                }
                if (curr_E < best_E) {
                    best_E = curr_E;
                    memcpy(best_rot, rotp, sizeof rotp); memcpy(best_pos, posp, sizeof posp);
                    memcpy(bx, lx, sizeof(double) * 3 * (size_t)L);
                }
                if (rigid && rigid_step > 0 && rigid_step % BLOCK_SIZE == 0) {   /* lds.ml:586-600 */
                    double ar = sw_ratio(&sw_rigid);
                    if (ar <= target_low) { max_trans = 0.95 * max_trans; max_rot = 0.95 * max_rot; }
                    else if (ar >= target_high) {
                        max_trans = 1.05 * max_trans;
                        double m = 1.05 * max_rot;
                        max_rot = (pi <= m) ? pi : m;          /* min Math.pi (...) */
                    }
                }
                if (flexible && !rigid && conf_step > 0 && conf_step % ((long)BLOCK_SIZE * nrb) == 0) {
                    /* update_rbond_rot_params on conf': kept only if conf' became conf (D4) */
                    conf_t *tgt = acc ? &conf : &confp;
                    for (int i = 0; i < nrb; i++) {
                        double ar = sw_ratio(&sw_bond[i]);
                        if (ar <= target_low) tgt->dr[i] = 0.95 * tgt->dr[i];
                        else if (ar >= target_high) { double m = 1.05 * tgt->dr[i]; tgt->dr[i] = (pi <= m) ? pi : m; }
                    }
                }
            }
        }
        if (do_reset) {                                  /* reset_run_params, lds.ml:632-648 */
            max_rot = p_max_rot; max_trans = p_max_trans;
            memcpy(rot, a->rot0, sizeof rot); memcpy(pos, a->pos0, sizeof pos);
            double id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            memcpy(best_rot, id, sizeof id);
            best_pos[0] = best_pos[1] = best_pos[2] = 0.0;
            prev_E = INFINITY; best_E = INFINITY;
            orc_rotate_then_translate(L, centered.x, centered.y, centered.z, rot, pos, bx, by, bz);
            sw_reset(&sw_rigid);
        }
        if (trace) { trace[4 * frame] = curr_E; trace[4 * frame + 1] = prev_E_inter; trace[4 * frame + 2] = prev_E_intra; trace[4 * frame + 3] = (double)accepted; }
        if (rigid) rigid_step++; else conf_step++;
    }
    res->frames_done = frame;
    res->best_E = best_E; res->prev_E = prev_E;
    memcpy(res->best_rot, best_rot, sizeof best_rot); memcpy(res->best_pos, best_pos, sizeof best_pos);
    res->max_rot = max_rot; res->max_trans = max_trans;
    if (best_xyz) memcpy(best_xyz, bx, sizeof(double) * 3 * (size_t)L);
    free(lx); free(sw_bond); free(centered.x); free(conf.x); free(confp.x);
}
