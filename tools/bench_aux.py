#!/usr/bin/env python
"""Secondary measurements (not the bench.py headline): configs C3 (grid build + interpolated scoring),
C4 (many MC chains) and C5-shaped direct scoring of arbitrary conformers.  Prints one JSON object.
Run on a B200:  python tools/bench_aux.py [--quick]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mmo_b200  # noqa: E402
from mmo_b200 import pqrs, workloads  # noqa: E402

K = {"direct_fp32": 0, "hard_fix": 1, "direct_fp64": 2, "intra": 3, "grid_build": 4, "interp": 5, "mc": 9}


def ktime(L, kid):
    ms, n = C.c_double(), C.c_int64()
    L.mmo_kernel_time_get(kid, C.byref(ms), C.byref(n))
    return ms.value, n.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    mmo_b200.init(0)
    L = mmo_b200.lib()
    out = {}
    hbm = C.c_double(); L.mmo_measure_hbm_copy(C.byref(hbm))
    fp64 = C.c_double(); L.mmo_measure_fp64_peak(C.byref(fp64))
    fp32 = C.c_double(); L.mmo_measure_fp32_peak(C.byref(fp32))
    out["peaks"] = {"hbm_copy_gbs": hbm.value, "fp64_fma_tflops": fp64.value, "fp32_fma_tflops": fp32.value}

    # ---- C3: 5000-atom synthetic receptor, 30 A box at 0.375 A (81^3 voxels), 22 ligand types --------
    rec_m = workloads.synthetic_receptor(5000, "cube", 60.0, seed=workloads.SEED)
    lig_m = pqrs.read_ligands_pqrs(os.path.join(workloads.GOLDEN, "ligdecs.pqrs"))[0]
    ta, tq = pqrs.assign_ff_types([lig_m])
    # Grid.from_box puts the lowest corner at the origin: shift the receptor so that the 30 A box
    # centred in it starts at the origin
    rec_m.xs -= 15.0; rec_m.ys -= 15.0; rec_m.zs -= 15.0
    rec = mmo_b200.Receptor.from_mol(rec_m)
    dims = mmo_b200.Grid.from_box(0.375, 30.0, 30.0, 30.0)
    L.mmo_kernel_timing(1)
    reps = 2 if args.quick else 5
    for _ in range(reps):
        g, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta, tq, want_host=False)
    ms, n = ktime(L, K["grid_build"])
    nvox = dims[0] * dims[1] * dims[2]
    out["c3_grid_build"] = {"dims": dims, "types": len(ta), "receptor_atoms": rec_m.n, "ms_per_build": ms / reps,
                            "launches_per_build": n / reps, "voxel_types_per_s": nvox * len(ta) / (ms / reps * 1e-3),
                            "bytes_written": nvox * len(ta) * 4, "write_gbs": nvox * len(ta) * 4 / (ms / reps * 1e-3) / 1e9,
                            "atom_voxel_pairs": nvox * rec_m.n}
    # interpolated scoring of 1e6 rigid poses whose atoms stay inside the grid
    lig = mmo_b200.Ligand.from_mol(lig_m, centered=True)
    n_poses = 200_000 if args.quick else 1_000_000
    rng = np.random.default_rng(workloads.SEED)
    R = workloads.random_rotations(n_poses, rng)
    rl = workloads.lig_radius((lig.xs, lig.ys, lig.zs))
    t = rng.uniform(rl + 0.4, 30.0 - rl - 0.4, (n_poses, 3))
    d_rot = C.c_void_p(); d_t = C.c_void_p(); d_e = C.c_void_p()
    L.mmo_dev_alloc(C.c_size_t(R.nbytes), C.byref(d_rot)); L.mmo_dev_alloc(C.c_size_t(t.nbytes), C.byref(d_t))
    L.mmo_dev_alloc(C.c_size_t(n_poses * 8), C.byref(d_e))
    L.mmo_h2d(d_rot, R.ctypes.data_as(C.c_void_p), C.c_size_t(R.nbytes)); L.mmo_h2d(d_t, t.ctypes.data_as(C.c_void_p), C.c_size_t(t.nbytes))
    L.mmo_kernel_timing(1)
    for _ in range(reps + 1):
        L.mmo_l2_flush()
        assert L.mmo_score_interp_poses_dev(g.h, lig.h, C.c_int64(n_poses), d_rot, d_t, d_e) == 0
    L.mmo_sync()
    ms, n = ktime(L, K["interp"])
    lookups = n_poses * lig.n
    out["c3_interp"] = {"poses": n_poses, "ms_per_launch": ms / n, "poses_per_s": n_poses / (ms / n * 1e-3),
                        "atom_lookups_per_s": lookups / (ms / n * 1e-3),
                        "algorithmic_gbs": lookups * 48 / (ms / n * 1e-3) / 1e9, "maps_mb": nvox * len(ta) * 4 / 1e6}
    # direct fp32 scoring of the same poses against the 5000-atom receptor (incoherent poses: C5 shape)
    n_dir = 50_000 if args.quick else 200_000
    L.mmo_kernel_timing(1)
    for _ in range(3):
        assert L.mmo_score_poses_dev(rec.h, lig.h, 1, 0, C.c_int64(n_dir), d_rot, d_t, d_e) == 0
    L.mmo_sync()
    ms, n = ktime(L, K["direct_fp32"])
    fms, _ = ktime(L, K["hard_fix"])
    out["c5_shape_direct"] = {"poses": n_dir, "receptor_atoms": rec_m.n, "ligand_atoms": lig.n,
                              "ms_per_launch": ms / n, "fix_ms_per_launch": fms / n,
                              "poses_per_s": n_dir / ((ms + fms) / n * 1e-3),
                              "nominal_pairs_per_s": n_dir * rec_m.n * lig.n / ((ms + fms) / n * 1e-3)}
    # ---- C5: conformer screen, 70-atom ligand x 10 000-atom receptor sphere, explicit coordinates, top-100 ----
    rec5_m = workloads.synthetic_receptor(10000, "sphere", 34.0, seed=workloads.SEED + 1, origin=(60.0, 60.0, 60.0))
    lig5_m = workloads.c5_ligand()
    rec5 = mmo_b200.Receptor.from_mol(rec5_m)
    lig5 = mmo_b200.Ligand.from_mol(lig5_m, centered=False)
    n5 = 50_000 if args.quick else 250_000
    X5, Y5, Z5 = workloads.c5_conformers(lig5_m, n5, (60.0, 60.0, 60.0))
    dxyz = [C.c_void_p() for _ in range(3)]
    d_e5 = C.c_void_p()
    for dp, arr in zip(dxyz, (X5, Y5, Z5)):
        L.mmo_dev_alloc(C.c_size_t(arr.nbytes), C.byref(dp)); L.mmo_h2d(dp, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes))
    L.mmo_dev_alloc(C.c_size_t(n5 * 8), C.byref(d_e5))
    L.mmo_set_collect_stats(1)
    assert L.mmo_score_coords_dev(rec5.h, lig5.h, 1, 0, C.c_int64(n5), dxyz[0], dxyz[1], dxyz[2], d_e5) == 0
    pe, pi_, pf = C.c_int64(), C.c_int64(), C.c_int64()
    L.mmo_last_pair_stats(C.byref(pe), C.byref(pi_), C.byref(pf))
    L.mmo_set_collect_stats(0)
    c5 = {}
    for mode, name in ((2, "item_kernel"), (1, "pose_kernel")):
        L.mmo_direct_set_mode(mode)
        nn = n5 if mode == 2 else n5 // 10          # the pose kernel evaluates ~20x the pairs: time a tenth
        L.mmo_kernel_timing(1)
        tot = 0.0
        ms1 = C.c_float()
        for it in range(3):
            L.mmo_l2_flush(); L.mmo_sync(); L.mmo_timer_start()
            assert L.mmo_score_coords_dev(rec5.h, lig5.h, 1, 0, C.c_int64(nn), dxyz[0], dxyz[1], dxyz[2], d_e5) == 0
            L.mmo_timer_stop(C.byref(ms1))
            if it > 0:
                tot += ms1.value
        ms, n = ktime(L, K["direct_fp32"]); fms, _ = ktime(L, K["hard_fix"]); pms, _ = ktime(L, 10)
        c5[name] = {"conformers": nn, "ms_per_call": tot / 2, "poses_per_s": nn / (tot / 2 * 1e-3),
                    "pair_kernel_ms": ms / 3, "fix_ms": fms / 3, "prepare_and_sort_ms": pms / 3}
    L.mmo_direct_set_mode(0)
    e5 = np.empty(n5)
    t0 = time.perf_counter()
    L.mmo_d2h(e5.ctypes.data_as(C.c_void_p), d_e5, C.c_size_t(n5 * 8))
    top = np.argpartition(e5, 100)[:100]
    c5["topk_100_host_ms"] = 1e3 * (time.perf_counter() - t0)
    c5.update({"receptor_atoms": rec5_m.n, "ligand_atoms": lig5_m.n, "pairs_evaluated_per_pose": pe.value / n5,
               "pairs_inside_cutoff_per_pose": pi_.value / n5, "fp64_fix_pairs_per_pose": pf.value / n5,
               "nominal_pairs_per_s": c5["item_kernel"]["poses_per_s"] * rec5_m.n * lig5_m.n,
               "algorithmic_tflops": c5["item_kernel"]["poses_per_s"] * (27.0 * pi_.value + 8.0 * (pe.value - pi_.value)) / n5 / 1e12,
               "best_E": float(e5[top].min())})
    out["c5_conformer_screen"] = c5
    L.mmo_kernel_timing(1)
    n64 = 5_000 if args.quick else 20_000
    assert L.mmo_score_poses_dev(rec.h, lig.h, 1, 1, C.c_int64(n64), d_rot, d_t, d_e) == 0
    L.mmo_sync()
    ms, n = ktime(L, K["direct_fp64"])
    out["direct_fp64_strict"] = {"poses": n64, "ms": ms, "poses_per_s": n64 / (ms * 1e-3),
                                 "nominal_pairs_per_s": n64 * rec_m.n * lig.n / (ms * 1e-3)}

    # ---- C4: MC chains, interpolated E_inter + intra NB, --hard-ROI -----------------------------------
    c2 = workloads.load_c2("ligdecs")
    rec2_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
    rec2 = mmo_b200.Receptor.from_mol(rec2_m)
    import oracle
    cc = np.array(c2["roi"][:3])
    gd = mmo_b200.Grid.from_box(0.5, *(cc + 23.0))
    gmask = oracle.bitmask_sphere(0.5, gd, cc, 21.0)
    ta2, tq2 = pqrs.assign_ff_types([c2["lig"]])
    L.mmo_kernel_timing(1)
    g2, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec2, 0.5, gd, ta2, tq2, mask_bits=gmask, want_host=False)
    ms, _ = ktime(L, K["grid_build"])
    out["c4_grid_build_ms"] = ms
    lig2 = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    n_chains, n_steps = (512, 1000) if args.quick else (4096, 10000)
    seeds = np.arange(n_chains, dtype=np.uint64) + workloads.SEED
    Rm, tm = workloads.random_poses_in_sphere(n_chains, c2["roi"][:3], 3.0, seed=41)
    L.mmo_kernel_timing(1)
    t0 = time.perf_counter()
    res, _, _ = mmo_b200.Lds.simulate_lig(g2, lig2, c2["roi"], n_steps, seeds, Rm, tm)
    wall = time.perf_counter() - t0
    ms, _ = ktime(L, K["mc"])
    done = sum(r["frames_done"] for r in res)
    out["c4_mc"] = {"chains": n_chains, "steps": n_steps, "kernel_ms": ms, "wall_s": wall,
                    "chain_steps_per_s": done / (ms * 1e-3), "median_best_E": float(np.median([r["best_E"] for r in res])),
                    "too_long": int(sum(r["too_long"] for r in res))}
    # CPU oracle chain for the side-by-side
    t0 = time.perf_counter()
    cx, cy, cz = c2["centered"]
    gm = g2.download()
    oracle.mc_run(c2["lig"], cx, cy, cz, c2["roi"], n_steps, int(seeds[0]), Rm[0], tm[0], maps=gm, g_step=0.5, g_dims=gd)
    out["c4_mc"]["cpu_oracle_chain_steps_per_s_1core"] = n_steps / (time.perf_counter() - t0)
    # ---- N4: desolvation sums on the reference's own grid (0.5 A over the simulation box) ---------------
    rm = c2["rec"]
    sdims = mmo_b200.Grid.from_box(0.5, *c2["sim_dims"])
    shell = mmo_b200.Lds.first_solvent_shell(rm.xs, rm.ys, rm.zs, rm.r, 0.5, sdims)
    rec_all = mmo_b200.Receptor.from_mol(rm)
    L.mmo_sync()
    t0 = time.perf_counter()
    dh, _ = mmo_b200.Lds.protein_desolv(c2["roi"], rec_all, shell, want_host=False)
    t_prot = time.perf_counter() - t0
    n_pen = 2000 if args.quick else 20000
    Rp, tp = workloads.random_poses_in_sphere(n_pen, c2["roi"][:3], 6.0, seed=43)
    mmo_b200.Lds.desolvation_penalty(dh, lig2, rot9=Rp[:64], trans3=tp[:64])
    t0 = time.perf_counter()
    dp, dl = mmo_b200.Lds.desolvation_penalty(dh, lig2, rot9=Rp, trans3=tp)
    t_pen = time.perf_counter() - t0
    # the oracle on a few of the same poses, one host core
    t0 = time.perf_counter()
    contribs = oracle.protein_desolv(rm, 0.5, sdims, shell.bits, c2["roi"])
    t_prot_cpu = time.perf_counter() - t0
    X, Y, Z = oracle.pose_coords(lig2.xs, lig2.ys, lig2.zs, Rp[:3], tp[:3])
    t0 = time.perf_counter()
    for q in range(3):
        wp, wl = oracle.desolvation_penalty(0.5, sdims, shell.bits, contribs, X[q], Y[q], Z[q], c2["lig"].q, c2["lig"].r)
        assert wp == dp[q] and wl == dl[q]
    t_pen_cpu = (time.perf_counter() - t0) / 3
    out["n4_desolvation"] = {"grid_dims": list(sdims), "receptor_atoms": rm.n, "protein_desolv_wall_ms": t_prot * 1e3,
                             "protein_desolv_cpu_oracle_ms": t_prot_cpu * 1e3, "penalty_poses": n_pen,
                             "penalty_wall_ms": t_pen * 1e3, "penalty_poses_per_s": n_pen / t_pen,
                             "penalty_cpu_oracle_poses_per_s_1core": 1.0 / t_pen_cpu,
                             "median_prot": float(np.median(dp)), "median_lig": float(np.median(dl))}
    L.mmo_kernel_timing(0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
