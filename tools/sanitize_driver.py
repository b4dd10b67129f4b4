#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (tools/sanitize.sh): sizes chosen so that memcheck /
racecheck / synccheck / initcheck finish in seconds each.  Results are compared with nothing here (the parity tests do
that); the point is the sanitizer's verdict on the final kernels of the round."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmo_b200  # noqa: E402
from mmo_b200 import pqrs, workloads  # noqa: E402

mmo_b200.init(0)
which = sys.argv[1:] or ["direct", "items", "strict", "grid", "scan", "mc", "masks", "desolv"]
c2 = workloads.load_c2("ligdecs")
rec_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
rec = mmo_b200.Receptor.from_mol(rec_m)
lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
R, t = workloads.random_poses_in_sphere(600, c2["roi"][:3], 6.0, seed=3)
if "direct" in which:        # pose kernel + hard_fix (close contacts), both variants
    mmo_b200.lib().mmo_direct_set_mode(1)
    mmo_b200.Mol.score_poses(rec, lig, R, t)                     # receptor slices + fp64 pass with block = pose
    mmo_b200.Mol.score_poses(rec, lig, R[:100], t[:100], variant=mmo_b200.VARIANT_GLOBAL)
    mmo_b200.Mol.score_poses(rec, lig, R[:1], t[:1])             # the single-pose call (pinned staging upload)
    R5, t5 = workloads.random_poses_in_sphere(4200, c2["roi"][:3], 6.0, seed=6)
    mmo_b200.Mol.score_poses(rec, lig, R5, t5)                   # one slice + fp64 pass with thread = pose
    mmo_b200.lib().mmo_direct_set_mode(0)
if "items" in which:         # item kernel: prepare, sort, direct_items_kernel, item_fix, per-pose sum; 10 000-atom receptor
    rec5_m = workloads.synthetic_receptor(3000, "sphere", 22.0, seed=5, origin=(40.0, 40.0, 40.0))
    rec5 = mmo_b200.Receptor.from_mol(rec5_m)
    lig5_m = workloads.c5_ligand()
    lig5 = mmo_b200.Ligand.from_mol(lig5_m, centered=False)
    X, Y, Z = workloads.c5_conformers(lig5_m, 700, (40.0, 40.0, 40.0), radius=8.0)
    mmo_b200.lib().mmo_direct_set_mode(2)
    e = mmo_b200.Mol.ene_inter_UFF_shifted_brute(rec5, lig5, X, Y, Z)
    mmo_b200.lib().mmo_direct_set_mode(0)
if "strict" in which:
    mmo_b200.Mol.score_poses(rec, lig, R[:64], t[:64], prec=mmo_b200.PREC_FP64)                  # block per pose
    mmo_b200.Mol.ene_intra_UFFNB_brute(lig, np.tile(lig.xs, (8, 1)), np.tile(lig.ys, (8, 1)), np.tile(lig.zs, (8, 1)))   # block per conformer
    rec_s = mmo_b200.Receptor.from_mol(workloads.carve(c2["rec"], c2["roi"][:3], 6.0))
    Rb, tb = workloads.random_poses_in_sphere(33000, c2["roi"][:3], 6.0, seed=4)
    mmo_b200.Mol.score_poses(rec_s, lig, Rb, tb, prec=mmo_b200.PREC_FP64)                        # thread per pose
    mmo_b200.Mol.ene_intra_UFFNB_brute(lig, np.tile(lig.xs, (400, 1)), np.tile(lig.ys, (400, 1)), np.tile(lig.zs, (400, 1)))   # thread per conformer
ta, tq = pqrs.assign_ff_types([c2["lig"]])
cc = np.array(c2["roi"][:3])
if "grid" in which or "mc" in which or "scan" in which:
    gd = mmo_b200.Grid.from_box(1.0, *(cc + 23.0))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_legs
    gmask = bench_legs.sphere_mask_bits(1.0, gd, cc, 21.0)
    grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 1.0, gd, ta, tq, mask_bits=gmask, want_host=False)
    mmo_b200.Mol.interp_poses(grid, lig, R[:200], t[:200])
if "scan" in which:
    rot = mmo_b200.SO3.rotations(40)
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 2.0)
    mmo_b200.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, rec=rec)
    mmo_b200.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, rec=rec, prec=mmo_b200.PREC_FP64)
    mmo_b200.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, grid=grid)
if "mc" in which:
    seeds = np.arange(6, dtype=np.uint64) + 7
    Rm, tm = workloads.random_poses_in_sphere(6, c2["roi"][:3], 3.0, seed=41)
    for nt in ("128", "64", "256", "32"):
        os.environ["MMO_MC_THREADS"] = nt
        mmo_b200.Lds.simulate_lig(grid, lig, c2["roi"], 60, seeds, Rm, tm, want_xyz=True, want_trace=True)
    mmo_b200.Lds.simulate_lig(None, lig, c2["roi"], 20, seeds[:2], Rm[:2], tm[:2], rec=rec)
    del os.environ["MMO_MC_THREADS"]
if "masks" in which or "desolv" in which:
    m = c2["rec"]
    sd = mmo_b200.Grid.from_box(1.0, *c2["sim_dims"])
    vdw = mmo_b200.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, 1.0, sd)
    shell = mmo_b200.Lds.first_solvent_shell(m.xs, m.ys, m.zs, m.r, 1.0, sd)
    mmo_b200.Lds.bitmask_ROI_only(c2["roi"], 1.0, sd)
    mmo_b200.Mol.protein_ligand_clash(vdw, lig, R[:100], t[:100] - 40.0)        # partly outside the mask box
if "desolv" in which:
    rec_all = mmo_b200.Receptor.from_mol(m)
    dh, _ = mmo_b200.Lds.protein_desolv(c2["roi"], rec_all, shell, want_host=False)
    mmo_b200.Lds.desolvation_penalty(dh, lig, rot9=R[:8], trans3=t[:8])
print("sanitize_driver done:", " ".join(which), "launches", mmo_b200.launch_count())
