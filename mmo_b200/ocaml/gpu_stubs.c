/* gpu_stubs.c -- C stubs between OCaml (gpu.ml) and libmmo_b200.so (include/mmo_b200.h).
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no caml/ headers); see INTEGRATION.md.
 * Every stub: unbox arguments, one library call, box the result; a non-zero status raises Failure. */
#include <stdlib.h>
#include <string.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/alloc.h>
#include <caml/fail.h>
#include <caml/custom.h>
#include <caml/bigarray.h>
#include "mmo_b200.h"

static void check(int rc) { if (rc != MMO_OK) caml_failwith(mmo_last_error()); }
static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) caml_raise_out_of_memory(); return p; }

/* ---- handles as custom blocks with finalisers ------------------------------------------------ */
#define HANDLE(name, ctype, destroy)                                                          \
  static void fin_##name(value v) { ctype *p = *(ctype **)Data_custom_val(v); if (p) destroy(p); } \
  static struct custom_operations ops_##name = { "mmo_b200." #name, fin_##name, custom_compare_default, \
    custom_hash_default, custom_serialize_default, custom_deserialize_default, custom_compare_ext_default, \
    custom_fixed_length_default };                                                            \
  static value box_##name(ctype *p) { value v = caml_alloc_custom(&ops_##name, sizeof(ctype *), 0, 1); \
    *(ctype **)Data_custom_val(v) = p; return v; }
HANDLE(rec, mmo_receptor, mmo_receptor_destroy)
HANDLE(lig, mmo_ligand, mmo_ligand_destroy)
HANDLE(grid, mmo_grid, mmo_grid_destroy)
HANDLE(mask, mmo_mask, mmo_mask_destroy)
HANDLE(desolv, mmo_desolv, mmo_desolv_destroy)
#define Rec_val(v) (*(mmo_receptor **)Data_custom_val(v))
#define Lig_val(v) (*(mmo_ligand **)Data_custom_val(v))
#define Grid_val(v) (*(mmo_grid **)Data_custom_val(v))
#define Mask_val(v) (*(mmo_mask **)Data_custom_val(v))
#define Desolv_val(v) (*(mmo_desolv **)Data_custom_val(v))

/* OCaml float array = flat unboxed doubles */
#define DARR(v) ((const double *)(v))
#define DLEN(v) (Wosize_val(v) / Double_wosize)

/* OCaml int array (tagged) -> int32_t[] */
static int32_t *ints_of(value a) {
  mlsize_t n = Wosize_val(a);
  int32_t *r = (int32_t *)xmalloc(sizeof(int32_t) * (n ? n : 1));
  for (mlsize_t i = 0; i < n; i++) r[i] = (int32_t)Long_val(Field(a, i));
  return r;
}

CAMLprim value mmo_ml_init(value dev) { check(mmo_init(Int_val(dev))); return Val_unit; }
CAMLprim value mmo_ml_shutdown(value u) { (void)u; check(mmo_shutdown()); return Val_unit; }

CAMLprim value mmo_ml_receptor_create(value xs, value ys, value zs, value q, value elt) {
  CAMLparam5(xs, ys, zs, q, elt);
  int32_t *an = ints_of(elt);
  mmo_receptor *h = NULL;
  int rc = mmo_receptor_create((int32_t)DLEN(xs), DARR(xs), DARR(ys), DARR(zs), DARR(q), an, &h);
  free(an);
  check(rc);
  CAMLreturn(box_rec(h));
}

CAMLprim value mmo_ml_ligand_create(value xs, value ys, value zs, value q, value r, value elt, value typ,
                                    value dists, value lefts, value rights, value rgroups) {
  CAMLparam5(xs, ys, zs, q, r); CAMLxparam5(elt, typ, dists, lefts, rights); CAMLxparam1(rgroups);
  int32_t n = (int32_t)DLEN(xs), nrb = (int32_t)Wosize_val(lefts);
  int32_t *an = ints_of(elt), *ty = ints_of(typ), *di = ints_of(dists), *le = ints_of(lefts), *ri = ints_of(rights);
  int32_t *off = (int32_t *)xmalloc(sizeof(int32_t) * (nrb + 1));
  off[0] = 0;
  for (int b = 0; b < nrb; b++) off[b + 1] = off[b] + (int32_t)Wosize_val(Field(rgroups, b));
  int32_t *idx = (int32_t *)xmalloc(sizeof(int32_t) * (off[nrb] ? off[nrb] : 1));
  for (int b = 0; b < nrb; b++)
    for (mlsize_t k = 0; k < Wosize_val(Field(rgroups, b)); k++) idx[off[b] + k] = (int32_t)Long_val(Field(Field(rgroups, b), k));
  mmo_ligand *h = NULL;
  int rc = mmo_ligand_create(n, DARR(xs), DARR(ys), DARR(zs), DARR(q), DARR(r), an, ty,
                             Wosize_val(dists) ? di : NULL, nrb, le, ri, off, idx, &h);
  free(an); free(ty); free(di); free(le); free(ri); free(off); free(idx);
  check(rc);
  CAMLreturn(box_lig(h));
}
CAMLprim value mmo_ml_ligand_create_bc(value *a, int n) {
  (void)n; return mmo_ml_ligand_create(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
}

/* ene_inter : Mol.t -> float for one pose (lds.ml:1952-1955) */
CAMLprim value mmo_ml_score_coords(value rec, value lig, value variant, value prec, value xs, value ys, value zs) {
  CAMLparam5(rec, lig, variant, prec, xs); CAMLxparam2(ys, zs);
  double e;
  check(mmo_score_coords(Rec_val(rec), Lig_val(lig), Int_val(variant), Int_val(prec), 1, DARR(xs), DARR(ys), DARR(zs), &e));
  CAMLreturn(caml_copy_double(e));
}
CAMLprim value mmo_ml_score_coords_bc(value *a, int n) { (void)n; return mmo_ml_score_coords(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); }

CAMLprim value mmo_ml_score_components(value rec, value lig, value xs, value ys, value zs) {
  CAMLparam5(rec, lig, xs, ys, zs);
  CAMLlocal3(res, be, bv);
  double e, v;
  check(mmo_score_coords_components(Rec_val(rec), Lig_val(lig), 1, DARR(xs), DARR(ys), DARR(zs), &e, &v));
  be = caml_copy_double(e);        /* allocate first, store afterwards: an allocation may move res */
  bv = caml_copy_double(v);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, be);
  Store_field(res, 1, bv);
  CAMLreturn(res);
}

CAMLprim value mmo_ml_intra_nb(value lig, value xs, value ys, value zs) {
  CAMLparam4(lig, xs, ys, zs);
  double e;
  check(mmo_intra_nb(Lig_val(lig), 1, DARR(xs), DARR(ys), DARR(zs), &e));
  CAMLreturn(caml_copy_double(e));
}

CAMLprim value mmo_ml_grid_build(value rec, value step, value xd, value yd, value zd, value mask_opt, value types, value maps) {
  CAMLparam5(rec, step, xd, yd, zd); CAMLxparam3(mask_opt, types, maps);
  int32_t dims[3] = {Int_val(xd), Int_val(yd), Int_val(zd)};
  size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
  int32_t T = (int32_t)Wosize_val(types);
  int32_t *ta = (int32_t *)xmalloc(sizeof(int32_t) * T);
  double *tq = (double *)xmalloc(sizeof(double) * T);
  for (int t = 0; t < T; t++) { ta[t] = (int32_t)Long_val(Field(Field(types, t), 0)); tq[t] = Double_val(Field(Field(types, t), 1)); }
  uint8_t *bits = NULL;
  if (Is_block(mask_opt)) {                                   /* Some (bool array) */
    value m = Field(mask_opt, 0);
    bits = (uint8_t *)xmalloc((nvox + 7) / 8 + 8); memset(bits, 0, (nvox + 7) / 8 + 8);
    for (size_t i = 0; i < nvox; i++) if (Bool_val(Field(m, i))) bits[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
  float *host = (float *)xmalloc(sizeof(float) * nvox * T);
  mmo_grid *h = NULL;
  int rc = mmo_grid_build(Rec_val(rec), Double_val(step), dims, bits, T, ta, tq, host, &h);
  if (rc == MMO_OK)
    for (int t = 0; t < T; t++) memcpy(Caml_ba_data_val(Field(maps, t)), host + (size_t)t * nvox, sizeof(float) * nvox);
  free(ta); free(tq); free(bits); free(host);
  check(rc);
  CAMLreturn(box_grid(h));
}
CAMLprim value mmo_ml_grid_build_bc(value *a, int n) { (void)n; return mmo_ml_grid_build(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); }

CAMLprim value mmo_ml_grid_upload(value step, value xd, value yd, value zd, value maps) {
  CAMLparam5(step, xd, yd, zd, maps);
  int32_t dims[3] = {Int_val(xd), Int_val(yd), Int_val(zd)};
  size_t nvox = (size_t)dims[0] * dims[1] * dims[2];
  int32_t T = (int32_t)Wosize_val(maps);
  float *host = (float *)xmalloc(sizeof(float) * nvox * T);
  for (int t = 0; t < T; t++) memcpy(host + (size_t)t * nvox, Caml_ba_data_val(Field(maps, t)), sizeof(float) * nvox);
  mmo_grid *h = NULL;
  int rc = mmo_grid_upload(Double_val(step), dims, T, host, &h);
  free(host);
  check(rc);
  CAMLreturn(box_grid(h));
}

CAMLprim value mmo_ml_score_interp(value grid, value lig, value xs, value ys, value zs) {
  CAMLparam5(grid, lig, xs, ys, zs);
  double e;
  check(mmo_score_interp_coords(Grid_val(grid), Lig_val(lig), 1, DARR(xs), DARR(ys), DARR(zs), &e));
  CAMLreturn(caml_copy_double(e));
}

CAMLprim value mmo_ml_trilin(value grid, value t, value x, value y, value z) {
  CAMLparam5(grid, t, x, y, z);
  double px = Double_val(x), py = Double_val(y), pz = Double_val(z), e;
  check(mmo_trilin(Grid_val(grid), Int_val(t), 1, &px, &py, &pz, &e));
  CAMLreturn(caml_copy_double(e));
}

CAMLprim value mmo_ml_vdw_mask_build(value xs, value ys, value zs, value radii, value step, value xd, value yd, value zd) {
  CAMLparam5(xs, ys, zs, radii, step); CAMLxparam3(xd, yd, zd);
  int32_t dims[3] = {Int_val(xd), Int_val(yd), Int_val(zd)};
  mmo_mask *h = NULL;
  check(mmo_vdw_mask_build((int32_t)DLEN(xs), DARR(xs), DARR(ys), DARR(zs), DARR(radii), Double_val(step), dims, NULL, &h));
  CAMLreturn(box_mask(h));
}
CAMLprim value mmo_ml_vdw_mask_build_bc(value *a, int n) { (void)n; return mmo_ml_vdw_mask_build(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); }

/* Lds.exhaustive_rigid_ligand_docking (lds.ml:1040-1114) */
CAMLprim value mmo_ml_scan(value rec_opt, value grid_opt, value lig, value mask_opt, value variant, value prec,
                           value roi, value trans_step, value rots, value e_intra, value topk) {
  CAMLparam5(rec_opt, grid_opt, lig, mask_opt, variant); CAMLxparam5(prec, roi, trans_step, rots, e_intra); CAMLxparam1(topk);
  CAMLlocal3(res, tops, bbest);
  mmo_scan_params p;
  memset(&p, 0, sizeof p);
  p.rec = Is_block(rec_opt) ? Rec_val(Field(rec_opt, 0)) : NULL;
  p.grid = Is_block(grid_opt) ? Grid_val(Field(grid_opt, 0)) : NULL;
  p.lig = Lig_val(lig);
  p.vdw_mask = Is_block(mask_opt) ? Mask_val(Field(mask_opt, 0)) : NULL;
  p.variant = Int_val(variant); p.prec = Int_val(prec);
  for (int d = 0; d < 3; d++) p.roi_c[d] = Double_flat_field(roi, d);
  p.roi_r = Double_flat_field(roi, 3);
  p.trans_step = Double_val(trans_step);
  p.n_rot = (int32_t)(DLEN(rots) / 9);
  p.rot9 = DARR(rots);
  p.e_intra_const = Double_val(e_intra);
  p.topk = Int_val(topk);
  p.first_point = 0; p.n_points = -1;
  int k = p.topk > 0 ? p.topk : 1;
  double *ts = (double *)xmalloc(sizeof(double) * k);
  int64_t *tf = (int64_t *)xmalloc(sizeof(int64_t) * k);
  mmo_scan_result r;
  int rc = mmo_scan(&p, ts, tf, &r);
  if (rc != MMO_OK) { free(ts); free(tf); check(rc); }
  tops = caml_alloc_float_array(r.n_top);
  for (int i = 0; i < r.n_top; i++) Store_double_flat_field(tops, i, ts[i]);
  free(ts); free(tf);
  bbest = caml_copy_double(r.best_score);
  res = caml_alloc_tuple(3);
  Store_field(res, 0, tops);
  Store_field(res, 1, bbest);
  Store_field(res, 2, Val_long(r.best_frame));
  CAMLreturn(res);
}
CAMLprim value mmo_ml_scan_bc(value *a, int n) { (void)n; return mmo_ml_scan(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]); }

/* ---- N4: desolvation sums (lds.ml:204-267) ---------------------------------------------------- */
CAMLprim value mmo_ml_first_solvent_shell(value xs, value ys, value zs, value radii, value step, value xd, value yd, value zd) {
  CAMLparam5(xs, ys, zs, radii, step); CAMLxparam3(xd, yd, zd);
  int32_t dims[3] = {Int_val(xd), Int_val(yd), Int_val(zd)};
  mmo_mask *h = NULL;
  check(mmo_mask_first_solvent_shell((int32_t)DLEN(xs), DARR(xs), DARR(ys), DARR(zs), DARR(radii), Double_val(step), dims, NULL, &h));
  CAMLreturn(box_mask(h));
}
CAMLprim value mmo_ml_first_solvent_shell_bc(value *a, int n) { (void)n; return mmo_ml_first_solvent_shell(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); }

/* the library copies the shell mask into the desolv handle: the OCaml mask value may be collected independently */
CAMLprim value mmo_ml_desolv_protein(value rec, value shell, value roi) {
  CAMLparam3(rec, shell, roi);
  double c[3] = {Double_flat_field(roi, 0), Double_flat_field(roi, 1), Double_flat_field(roi, 2)};
  mmo_desolv *h = NULL;
  check(mmo_desolv_protein(Rec_val(rec), Mask_val(shell), c, Double_flat_field(roi, 3), NULL, &h));
  CAMLreturn(box_desolv(h));
}

CAMLprim value mmo_ml_desolv_penalty(value d, value lig, value xs, value ys, value zs) {
  CAMLparam5(d, lig, xs, ys, zs);
  CAMLlocal3(res, bp, bl);
  double prot, ligp;
  check(mmo_desolv_penalty_coords(Desolv_val(d), Lig_val(lig), 1, DARR(xs), DARR(ys), DARR(zs), &prot, &ligp));
  bp = caml_copy_double(prot);
  bl = caml_copy_double(ligp);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, bp);
  Store_field(res, 1, bl);
  CAMLreturn(res);
}

/* Lds.place_ligand_in_ROI (lds.ml:308-345) */
CAMLprim value mmo_ml_place_ligand_in_roi(value xs, value ys, value zs, value anums, value lig, value roi, value seed,
                                          value starts, value clash) {
  CAMLparam5(xs, ys, zs, anums, lig); CAMLxparam4(roi, seed, starts, clash);
  CAMLlocal3(res, rots, poss);
  int32_t n = Int_val(starts), trials = 0;
  int32_t *an = ints_of(anums);
  double c[3] = {Double_flat_field(roi, 0), Double_flat_field(roi, 1), Double_flat_field(roi, 2)};
  double *r9 = (double *)xmalloc(sizeof(double) * 12 * (size_t)(n > 0 ? n : 1)), *p3 = r9 + 9 * (size_t)(n > 0 ? n : 1);
  int rc = mmo_place_ligand_in_roi((int32_t)DLEN(xs), DARR(xs), DARR(ys), DARR(zs), an, Lig_val(lig), c, Double_flat_field(roi, 3),
                                   (uint64_t)Long_val(seed), n, Bool_val(clash), r9, p3, &trials);
  free(an);
  if (rc != MMO_OK) { free(r9); check(rc); }
  rots = caml_alloc_float_array(9 * (mlsize_t)n);
  poss = caml_alloc_float_array(3 * (mlsize_t)n);
  for (int i = 0; i < 9 * n; i++) Store_double_flat_field(rots, i, r9[i]);
  for (int i = 0; i < 3 * n; i++) Store_double_flat_field(poss, i, p3[i]);
  free(r9);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, rots);
  Store_field(res, 1, poss);
  CAMLreturn(res);
}
CAMLprim value mmo_ml_place_ligand_in_roi_bc(value *a, int n) { (void)n; return mmo_ml_place_ligand_in_roi(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); }

/* Lds.simulate_lig frame loop (lds.ml:741-1000, frames 882-995) for many (ligand, start) chains in one launch.
 * grid option (interpolated E_inter) or receptor option (--no-interp), centred ligand, (roi x y z r), temperature,
 * steps, flags (tweak_rbonds, hard_roi, no_flip, intra_nb), seeds (int array), start rotations (9 per chain, flat),
 * start positions (3 per chain, flat)  ->  (best_E array, best_rot flat, best_pos flat, best_xyz flat (3 L per chain),
 * frames_done array) */
CAMLprim value mmo_ml_mc_run(value rec_opt, value grid_opt, value lig, value roi, value temp, value steps, value flags,
                             value seeds, value rots, value poss) {
  CAMLparam5(rec_opt, grid_opt, lig, roi, temp); CAMLxparam5(steps, flags, seeds, rots, poss);
  CAMLlocal5(res, bE, bR, bP, bX);
  CAMLlocal1(fr);
  mmo_mc_params p;
  memset(&p, 0, sizeof p);
  for (int d = 0; d < 3; d++) p.roi_c[d] = Double_flat_field(roi, d);
  p.roi_r = Double_flat_field(roi, 3);
  p.temperature_K = Double_val(temp);
  p.n_steps = Int_val(steps);
  p.tweak_rbonds = Bool_val(Field(flags, 0)); p.hard_roi = Bool_val(Field(flags, 1));
  p.no_flip = Bool_val(Field(flags, 2)); p.intra_nb = Bool_val(Field(flags, 3));
  int64_t n = (int64_t)Wosize_val(seeds);
  uint64_t *sd = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)n);
  for (int64_t c = 0; c < n; c++) sd[c] = (uint64_t)Long_val(Field(seeds, c));
  mmo_mc_result *r = (mmo_mc_result *)xmalloc(sizeof(mmo_mc_result) * (size_t)n);
  /* best_xyz holds 3 L doubles per chain; L = Mol.num_atoms centered_lig travels in the flags tuple */
  const int32_t L = (int32_t)Long_val(Field(flags, 4));
  if ((int64_t)DLEN(rots) != 9 * n || (int64_t)DLEN(poss) != 3 * n) { free(sd); free(r); caml_failwith("mc_run: rotations / positions do not match the seeds"); }
  double *xyz = (double *)xmalloc(sizeof(double) * 3 * (size_t)L * (size_t)n);
  int rc = mmo_mc_run(Is_block(rec_opt) ? Rec_val(Field(rec_opt, 0)) : NULL, Is_block(grid_opt) ? Grid_val(Field(grid_opt, 0)) : NULL,
                      Lig_val(lig), &p, n, sd, DARR(rots), DARR(poss), r, xyz, NULL);
  free(sd);
  if (rc != MMO_OK) { free(r); free(xyz); check(rc); }
  bE = caml_alloc_float_array((mlsize_t)n);
  bR = caml_alloc_float_array(9 * (mlsize_t)n);
  bP = caml_alloc_float_array(3 * (mlsize_t)n);
  bX = caml_alloc_float_array(3 * (mlsize_t)L * (mlsize_t)n);
  fr = caml_alloc_tuple((mlsize_t)n);
  for (int64_t c = 0; c < n; c++) {
    Store_double_flat_field(bE, c, r[c].best_E);
    for (int k = 0; k < 9; k++) Store_double_flat_field(bR, 9 * c + k, r[c].best_rot[k]);
    for (int k = 0; k < 3; k++) Store_double_flat_field(bP, 3 * c + k, r[c].best_pos[k]);
    Store_field(fr, c, Val_long(r[c].too_long ? -r[c].frames_done : r[c].frames_done));   /* negative: Mol.Too_long stopped the run */
  }
  for (mlsize_t i = 0; i < 3 * (mlsize_t)L * (mlsize_t)n; i++) Store_double_flat_field(bX, i, xyz[i]);
  free(r); free(xyz);
  res = caml_alloc_tuple(5);
  Store_field(res, 0, bE); Store_field(res, 1, bR); Store_field(res, 2, bP); Store_field(res, 3, bX); Store_field(res, 4, fr);
  CAMLreturn(res);
}
CAMLprim value mmo_ml_mc_run_bc(value *a, int n) { (void)n; return mmo_ml_mc_run(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); }
