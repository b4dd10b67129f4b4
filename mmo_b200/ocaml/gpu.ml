(* gpu.ml -- OCaml side of the libmmo_b200 binding (see INTEGRATION.md).
   NOT COMPILED IN THIS REPOSITORY'S IMAGE (no OCaml toolchain); kept mechanical on purpose.
   Each external maps 1:1 to a stub in gpu_stubs.c which maps 1:1 to include/mmo_b200.h. *)

type receptor
type ligand
type grid
type mask

(* MMO_VARIANT_* / MMO_PREC_* of include/mmo_b200.h *)
let variant_global = 0
let variant_shifted = 1
let prec_fp32 = 0
let prec_fp64 = 1

external init : int -> unit = "mmo_ml_init"
external shutdown : unit -> unit = "mmo_ml_shutdown"

external receptor_create :
  float array -> float array -> float array -> float array -> int array -> receptor
  = "mmo_ml_receptor_create"

(* xs ys zs q_a r_a elt_a t_a dists_a rb_left rb_right rgroups *)
external ligand_create_raw :
  float array -> float array -> float array -> float array -> float array ->
  int array -> int array -> int array -> int array -> int array -> int array array -> ligand
  = "mmo_ml_ligand_create_bc" "mmo_ml_ligand_create"

external score_coords :
  receptor -> ligand -> int -> int -> float array -> float array -> float array -> float
  = "mmo_ml_score_coords_bc" "mmo_ml_score_coords"

external score_components :
  receptor -> ligand -> float array -> float array -> float array -> float * float
  = "mmo_ml_score_components"

external intra_nb : ligand -> float array -> float array -> float array -> float
  = "mmo_ml_intra_nb"

(* step x_dim y_dim z_dim mask(as bool array option) types -> maps (filled in place) *)
external grid_build :
  receptor -> float -> int -> int -> int -> bool array option -> (int * float) array ->
  (float, Bigarray.float32_elt, Bigarray.c_layout) Bigarray.Array1.t array -> grid
  = "mmo_ml_grid_build_bc" "mmo_ml_grid_build"

external grid_upload :
  float -> int -> int -> int ->
  (float, Bigarray.float32_elt, Bigarray.c_layout) Bigarray.Array1.t array -> grid
  = "mmo_ml_grid_upload"

external score_interp : grid -> ligand -> float array -> float array -> float array -> float
  = "mmo_ml_score_interp"

external trilin : grid -> int -> float -> float -> float -> float = "mmo_ml_trilin"

external vdw_mask_build :
  float array -> float array -> float array -> float array -> float -> int -> int -> int -> mask
  = "mmo_ml_vdw_mask_build_bc" "mmo_ml_vdw_mask_build"

(* rec option, grid option, lig, mask option, variant, prec, (roi x y z r), trans_step, rotations
   (float array of 9*n: Rot.t records are flat), const_ene_intra, topk
   -> (top scores ascending, best score, best frame) *)
external scan :
  receptor option -> grid option -> ligand -> mask option -> int -> int ->
  float array -> float -> float array -> float -> int -> float array * float * int
  = "mmo_ml_scan_bc" "mmo_ml_scan"

(* N4: Lds.protein_desolv roi grid prot_bst prot_solvent_shell prot (lds.ml:204-236).
   rec, protein first-solvent-shell mask, (roi x y z r) -> device-resident contributions *)
type desolv
external first_solvent_shell :
  float array -> float array -> float array -> float array -> float -> int -> int -> int -> mask
  = "mmo_ml_first_solvent_shell_bc" "mmo_ml_first_solvent_shell"
external desolv_protein : receptor -> mask -> float array -> desolv = "mmo_ml_desolv_protein"
(* Lds.desolvation_penalty (lds.ml:239-267) of one ligand conformer: (prot, lig) *)
external desolv_penalty : desolv -> ligand -> float array -> float array -> float array -> float * float
  = "mmo_ml_desolv_penalty"

(* Lds.place_ligand_in_ROI (lds.ml:308-345): receptor xs ys zs anums, centred ligand, (roi x y z r), seed, starts,
   clash_check -> flat rotations (9 per start), flat positions (3 per start) *)
external place_ligand_in_roi :
  float array -> float array -> float array -> int array -> ligand -> float array -> int -> int -> bool ->
  float array * float array
  = "mmo_ml_place_ligand_in_roi_bc" "mmo_ml_place_ligand_in_roi"

(* Lds.simulate_lig frame loop (lds.ml:741-1000) for many chains in one launch -- chains = (ligand, start) pairs
   (lds.ml:1997-2000, 2034-2050).  rec option (--no-interp) / grid option (interpolated), centred ligand,
   (roi x y z r), temperature (K), steps, (tweak_rbonds, hard_roi, no_flip, intra_nb, num_atoms), seeds,
   start rotations (9 floats per chain), start positions (3 per chain)
   -> best_E per chain, best rotations (flat), best positions (flat), best coordinates (3 L per chain: xs ys zs),
      frames done per chain (negative when Mol.Too_long ended the run) *)
external mc_run :
  receptor option -> grid option -> ligand -> float array -> float -> int ->
  (bool * bool * bool * bool * int) -> int array -> float array -> float array ->
  float array * float array * float array * float array * int array
  = "mmo_ml_mc_run_bc" "mmo_ml_mc_run"

(* ---- drop-in closures ------------------------------------------------------------------ *)

let ligand_create (m: Mol.t): ligand =
  let lefts = Array.map (fun b -> b.Bond.left) m.Mol.rbonds_a in
  let rights = Array.map (fun b -> b.Bond.right) m.Mol.rbonds_a in
  ligand_create_raw m.Mol.xs m.Mol.ys m.Mol.zs m.Mol.q_a m.Mol.r_a m.Mol.elt_a m.Mol.t_a
    m.Mol.dists_a lefts rights m.Mol.rgroups_a

(* Mol.ene_inter_UFF_shifted_brute prot (mol.ml:822-849) / _shifted_bst (958-960) *)
let ene_inter_shifted ?(prec = prec_fp32) rec_h lig_h (lig': Mol.t): float =
  score_coords rec_h lig_h variant_shifted prec lig'.Mol.xs lig'.Mol.ys lig'.Mol.zs

(* Mol.ene_inter_UFF_global_brute prot (mol.ml:796-818) *)
let ene_inter_global ?(prec = prec_fp32) rec_h lig_h (lig': Mol.t): float =
  score_coords rec_h lig_h variant_global prec lig'.Mol.xs lig'.Mol.ys lig'.Mol.zs

(* Mol.ene_inter_UFF_interp grid ff_comps (mol.ml:1012-1020) *)
let ene_inter_interp grid_h lig_h (lig': Mol.t): float =
  score_interp grid_h lig_h lig'.Mol.xs lig'.Mol.ys lig'.Mol.zs

(* Mol.ene_intra_UFFNB_brute (mol.ml:881-903) *)
let ene_intra lig_h (lig': Mol.t): float =
  intra_nb lig_h lig'.Mol.xs lig'.Mol.ys lig'.Mol.zs

(* Lds.desolvation_penalty grid prot_desolv_contribs prot_solvent_shell prot lig (lds.ml:239-267) *)
let desolvation_penalty desolv_h lig_h (lig': Mol.t): float * float =
  desolv_penalty desolv_h lig_h lig'.Mol.xs lig'.Mol.ys lig'.Mol.zs

(* Lds.simulate_lig for every start of one ligand at once: what main's array_pariteri over starts (lds.ml:2050) does
   with one forked process per start.  Returns, per start, (best_E, best_rot as Rot.t-ordered floats, best_pos). *)
let simulate_lig_starts ?rec_h ?grid_h lig_h (centered_lig: Mol.t) ~roi ~temperature ~nsteps
    ~tweak_rbonds ~hard_roi ~no_flip ~intra_nb ~(seeds: int array) ~(rots: float array) ~(poss: float array) =
  mc_run rec_h grid_h lig_h roi temperature nsteps
    (tweak_rbonds, hard_roi, no_flip, intra_nb, Mol.num_atoms centered_lig) seeds rots poss
