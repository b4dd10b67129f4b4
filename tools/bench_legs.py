#!/usr/bin/env python
"""The secondary legs of bench.py (its `aux` object): every BASELINE.json config beside the C2 headline, measured
on the same box in the same run, each with its own roofline denominator measured live.

  c3_grid_build   configs[2]: 22 UFF maps of 81^3 voxels (0.375 A over 30 A) from a 5000-atom receptor   (K3, FP64 ALU + GB/s written)
  c3_lookup       configs[2]: interpolated scoring of 1e6 poses of ligdecs.mol2                          (K4, L2 gather roof)
  c4_mc           configs[3]: 512 MC chains per GPU x 10 000 frames (N = 8: the 4096 chains of C4)       (K6)
  c5_screen       configs[4]: 1e6 conformers (70 atoms) x 10 000-atom receptor sharded over the N GPUs,
                  top-100 conformer ids merged by the NCCL all-gather                                    (K1 item mode, FP32 ALU)
  c2_fp64_scan    configs[1] in MMO_PREC_FP64 (two-stage scan: argmin and top-k order of strict scoring)

Times are CUDA events on the library stream (max over ranks), units are summed over ranks.  Stand-alone:
`python tools/bench_legs.py [--quick]` prints the same object for one GPU."""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

K_DIRECT_FP32, K_HARD_FIX, K_DIRECT_FP64, K_INTRA, K_GRID_BUILD, K_INTERP, K_PREFILTER, K_REDUCE, K_VDW_MASK, K_MC, K_ITEM_PREP = range(11)
_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_int64)


class Clocks:
    """nvidia-smi clocks / throttle reasons sampled while a leg runs (the recipe's clocks line)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, enabled=True):
        self.samples, self.stop_evt, self.enabled = [], threading.Event(), enabled
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        dev = os.environ.get("LOCAL_RANK", "0")
        while not self.stop_evt.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", dev, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                return
            self.stop_evt.wait(0.1)

    def __enter__(self):
        if self.enabled:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop_evt.set()
        if self.enabled:
            self.th.join(timeout=10)

    def summary(self):
        s = self.samples
        if not s:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi sample"]}
        sm = sorted(float(x[0]) for x in s)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(x[2 + i].lower().startswith("active") for x in s)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(s[0][1]), "reasons": reasons, "samples": len(sm)}


class Ctx:
    """library + (optional) process group of one bench run"""

    def __init__(self, L, rank=0, world=1, dist=None, torch=None):
        self.L, self.rank, self.world, self.dist, self.torch = L, rank, world, dist, torch

    def ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.mmo_last_error().decode())

    def ktime(self, kid):
        ms, n = C.c_double(), C.c_int64()
        self.ck(self.L.mmo_kernel_time_get(kid, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def barrier(self):
        self.ck(self.L.mmo_sync())
        if self.dist is not None:
            self.dist.barrier()

    def reduce(self, vals, op="max"):
        """max / sum of a list of floats over the ranks"""
        if self.dist is None:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def timed(self, fn):
        """device time of fn() in ms: CUDA events on the library stream (fn's launches go there), sync on both sides"""
        ms = C.c_float()
        self.ck(self.L.mmo_sync())
        self.ck(self.L.mmo_timer_start())
        fn()
        self.ck(self.L.mmo_timer_stop(C.byref(ms)))
        return ms.value


def peaks(ctx, dims=(81, 81, 81), T=22):
    L = ctx.L
    fp32, fp64, hbm, gat = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    ctx.ck(L.mmo_measure_fp32_peak(C.byref(fp32)))
    ctx.ck(L.mmo_measure_fp64_peak(C.byref(fp64)))
    ctx.ck(L.mmo_measure_hbm_copy(C.byref(hbm)))
    d = (C.c_int32 * 3)(*dims)
    ctx.ck(L.mmo_measure_l2_gather(d, C.c_int32(T), C.byref(gat)))
    gzp = C.c_double()
    ctx.ck(L.mmo_measure_l2_gather_zpair(d, C.c_int32(T), C.byref(gzp)))
    return {"fp32_fma_tflops": fp32.value, "fp64_fma_tflops": fp64.value, "hbm_copy_gbs": hbm.value,
            "l2_gather_lookups_per_s": gat.value, "l2_gather_gbs": gat.value * 32 / 1e9,
            "l2_gather_zpair_lookups_per_s": gzp.value, "l2_gather_zpair_gbs": gzp.value * 32 / 1e9,
            "how": "FMA chains (8 per thread, full occupancy), 1 GiB float4 copy, 8-corner gathers of random cells of 22 "
                   "L2-resident 81^3 f32 maps (32 B per lookup) in the plain layout (8 x 4 B in four rows) and in the z-pair "
                   "layout the lookup kernel reads (4 x 8 B in two rows); all measured in this run on this GPU"}


def sphere_mask_bits(step, dims, c, r):
    """bit idx (LSB first) set where the grid node is closer than r to c: the shape of Lds.bitmask_ROI_only (lds.ml:269-305)"""
    x = np.arange(dims[0]) * step; y = np.arange(dims[1]) * step; z = np.arange(dims[2]) * step
    d2 = ((c[2] - z) ** 2)[:, None, None] + ((c[1] - y) ** 2)[None, :, None] + ((c[0] - x) ** 2)[None, None, :]
    return np.packbits((d2 < r * r).reshape(-1), bitorder="little")


# ---------------------------------------------------------------------------------------------------------------
def leg_c3(ctx, pk, quick=False):
    import mmo_b200
    from mmo_b200 import pqrs, workloads
    from scipy.spatial import cKDTree
    L = ctx.L
    rec_m = workloads.synthetic_receptor(5000, "cube", 60.0, seed=workloads.SEED)
    rec_m.xs -= 15.0; rec_m.ys -= 15.0; rec_m.zs -= 15.0        # Grid.from_box: lowest corner at the origin
    lig_m = pqrs.read_ligands_pqrs(os.path.join(workloads.GOLDEN, "ligdecs.pqrs"))[0]
    ta, tq = pqrs.assign_ff_types([lig_m])
    T = len(ta)
    rec = mmo_b200.Receptor.from_mol(rec_m)
    dims = mmo_b200.Grid.from_box(0.375, 30.0, 30.0, 30.0)
    nvox = dims[0] * dims[1] * dims[2]
    # (voxel, receptor atom) pairs inside the 12 A cut-off: the terms Mol.ene_inter_UFF_shifted_grid evaluates
    g = [np.arange(d) * 0.375 for d in dims]
    V = np.stack(np.meshgrid(g[0], g[1], g[2], indexing="ij"), -1).reshape(-1, 3)
    pairs_in = int(cKDTree(np.stack([rec_m.xs, rec_m.ys, rec_m.zs], 1)).query_ball_point(V, 12.0, return_length=True, workers=-1).sum())
    reps = 2 if quick else 5
    grid = None
    with Clocks(ctx.rank == 0) as ck:
        grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta, tq, want_host=False)   # warm-up
        ctx.ck(L.mmo_kernel_timing(1))
        ctx.barrier()
        for _ in range(reps):
            ctx.ck(L.mmo_l2_flush())
            grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta, tq, want_host=False)
        ms, n = ctx.ktime(K_GRID_BUILD)
    ms_build = ctx.reduce([ms / max(1, n)])[0]
    flops = pairs_in * (13.0 + 14.0 * T)          # 13 shared + 14 per type, in-cut-off terms only
    tf = flops / (ms_build * 1e-3) / 1e12
    build = {"workload": "C3 grid build: 5000-atom synthetic receptor, 81^3 voxels at 0.375 A, 22 ligand FF types, no mask",
             "ms": ms_build, "launches": int(n), "voxel_types_per_s": ctx.world * nvox * T / (ms_build * 1e-3),
             "scaling": "replicas (every GPU builds the whole 46.8 MB map set it will look up)",
             "atom_voxel_pairs_inside_cutoff": pairs_in, "bytes_written": nvox * T * 4,
             "write_gbs": nvox * T * 4 / (ms_build * 1e-3) / 1e9,
             "roofline": {"bound": "fp64", "kernel": "strict_grid_kernel", "achieved": tf, "peak": pk["fp64_fma_tflops"],
                          "unit": "TFLOP/s", "frac": tf / pk["fp64_fma_tflops"],
                          "flops_per_launch": flops, "flop_model": "(13 + 14 T) per (voxel, atom) term inside 12 A (SURVEY 8d), terms outside not credited",
                          "hbm_write_frac": nvox * T * 4 / (ms_build * 1e-3) / 1e9 / pk["hbm_copy_gbs"]},
             "clocks": ck.summary()}

    # ---- K4: 1e6 rigid poses whose atoms stay inside the grid ----
    lig = mmo_b200.Ligand.from_mol(lig_m, centered=True)
    n_poses = 200_000 if quick else 1_000_000
    rng = np.random.default_rng(workloads.SEED + ctx.rank)
    R = workloads.random_rotations(n_poses, rng)
    rl = workloads.lig_radius((lig.xs, lig.ys, lig.zs))
    t = rng.uniform(rl + 0.4, 30.0 - rl - 0.4, (n_poses, 3))
    d_rot, d_t, d_e = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.ck(L.mmo_dev_alloc(C.c_size_t(R.nbytes), C.byref(d_rot))); ctx.ck(L.mmo_dev_alloc(C.c_size_t(t.nbytes), C.byref(d_t)))
    ctx.ck(L.mmo_dev_alloc(C.c_size_t(n_poses * 8), C.byref(d_e)))
    ctx.ck(L.mmo_h2d(d_rot, R.ctypes.data_as(C.c_void_p), C.c_size_t(R.nbytes)))
    ctx.ck(L.mmo_h2d(d_t, t.ctypes.data_as(C.c_void_p), C.c_size_t(t.nbytes)))
    out = {}
    with Clocks(ctx.rank == 0) as ck:
        for name, flush in (("l2_warm", False), ("l2_flushed", True)):
            ctx.ck(L.mmo_score_interp_poses_dev(grid.h, lig.h, C.c_int64(n_poses), d_rot, d_t, d_e))      # warm-up
            ctx.ck(L.mmo_kernel_timing(1))
            ctx.barrier()
            for _ in range(reps):
                if flush:
                    ctx.ck(L.mmo_l2_flush())
                ctx.ck(L.mmo_score_interp_poses_dev(grid.h, lig.h, C.c_int64(n_poses), d_rot, d_t, d_e))
            ms, n = ctx.ktime(K_INTERP)
            out[name] = ctx.reduce([ms / max(1, n)])[0]
    for p in (d_rot, d_t, d_e):
        ctx.ck(L.mmo_dev_free(p))
    lookups = n_poses * lig.n
    lps = lookups / (out["l2_warm"] * 1e-3)
    lookup = {"workload": "C3 lookup: 1e6 random rigid poses of ligdecs.mol2 (48 atoms) inside the 22 x 81^3 maps, per GPU",
              "ms": out["l2_warm"], "ms_l2_flushed_before_launch": out["l2_flushed"], "poses_per_s": ctx.world * n_poses / (out["l2_warm"] * 1e-3),
              "atom_lookups_per_s": ctx.world * lps, "scaling": "weak (poses sharded, maps replicated)",
              "maps_mb": nvox * T * 4 / 1e6, "pose_stream_mb": n_poses * 104 / 1e6,
              "maps_zpair_mb": nvox * T * 8 / 1e6 * (dims[2] - 1) / dims[2],
              "roofline": {"bound": "l2_gather", "kernel": "strict_interp_kernel<zpair>", "achieved": lps * 32 / 1e9, "peak": pk["l2_gather_zpair_gbs"],
                           "unit": "GB/s", "frac": lps / pk["l2_gather_zpair_lookups_per_s"],
                           "frac_of_plain_layout_roof": lps / pk["l2_gather_lookups_per_s"],
                           "bytes_model": "32 B gathered per atom lookup (8 f32 corners); SURVEY 8(d)'s 48 B figure (+16 B atom record) gives "
                                          f"{lps * 48 / 1e9:.0f} GB/s",
                           "peak_source": "mmo_measure_l2_gather_zpair, this run: random cells of the same maps in the layout the kernel reads "
                                          "(frac_of_plain_layout_roof: against the 8 x 4 B gather of the reference's G3D layout)"},
              "clocks": ck.summary()}
    return build, lookup


def leg_c4(ctx, quick=False):
    import mmo_b200
    from mmo_b200 import pqrs, workloads
    L = ctx.L
    c2 = workloads.load_c2("ligdecs")
    rec_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
    rec = mmo_b200.Receptor.from_mol(rec_m)
    cc = np.array(c2["roi"][:3])
    gd = mmo_b200.Grid.from_box(0.5, *(cc + 23.0))
    gmask = sphere_mask_bits(0.5, gd, cc, 21.0)
    ta, tq = pqrs.assign_ff_types([c2["lig"]])
    ctx.ck(L.mmo_kernel_timing(1))
    grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 0.5, gd, ta, tq, mask_bits=gmask, want_host=False)
    ms_grid, _ = ctx.ktime(K_GRID_BUILD)
    lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    per_gpu, n_steps = (128, 1000) if quick else (512, 10000)

    def run(n_chains, first):
        seeds = np.arange(first, first + n_chains, dtype=np.uint64) + workloads.SEED
        Rm, tm = workloads.random_poses_in_sphere(first + n_chains, c2["roi"][:3], 3.0, seed=41)
        Rm, tm = Rm[first:], tm[first:]
        mmo_b200.Lds.simulate_lig(grid, lig, c2["roi"], min(n_steps, 200), seeds, Rm, tm)                # warm-up
        ctx.ck(L.mmo_kernel_timing(1))
        ctx.barrier()
        res, _, _ = mmo_b200.Lds.simulate_lig(grid, lig, c2["roi"], n_steps, seeds, Rm, tm)
        ms, _ = ctx.ktime(K_MC)
        return ms, sum(r["frames_done"] for r in res), res

    with Clocks(ctx.rank == 0) as ck:
        ms, done, res = run(per_gpu, ctx.rank * per_gpu)
    ms_max = ctx.reduce([ms])[0]
    done_all = ctx.reduce([done], "sum")[0]
    leg = {"workload": f"C4: {per_gpu} Monte-Carlo chains per GPU x {n_steps} frames (ligdecs.mol2, 9 rotatable bonds, --hard-ROI --intra-NB, "
                       f"interpolated E_inter on 0.5 A maps); at 8 GPUs these are the 4096 chains of BASELINE configs[3]",
           "chains": per_gpu * ctx.world, "frames_per_chain": n_steps, "ms": ms_max, "chain_steps_per_s": done_all / (ms_max * 1e-3),
           "scaling": "weak (chains sharded, no collective: one process per (ligand, start) in the reference, lds.ml:1997-2000)",
           "grid_build_ms": ms_grid, "median_best_E": float(np.median([r["best_E"] for r in res])),
           "too_long_runs": int(sum(r["too_long"] for r in res)),
           "roofline": {"bound": "latency", "kernel": "mc_chain kernel", "frac": None,
                        "note": "a frame is a chain of dependent double operations summed in the reference's order (bit-identical "
                                "trajectories); neither ALU nor memory bound"},
           "clocks": ck.summary()}
    if ctx.world == 1 and not quick:
        ms1, done1, _ = run(4096, 0)
        leg["one_gpu_4096_chains"] = {"ms": ms1, "chain_steps_per_s": done1 / (ms1 * 1e-3)}
        ms1, done1, _ = run(1, 0)
        leg["one_chain"] = {"ms": ms1, "frames_per_s": done1 / (ms1 * 1e-3)}
    return leg


def leg_c5(ctx, pk, quick=False):
    import mmo_b200
    from mmo_b200 import workloads
    L = ctx.L
    total = 100_000 if quick else 1_000_000
    topk = 100
    rec_m = workloads.synthetic_receptor(10000, "sphere", 34.0, seed=workloads.SEED + 1, origin=(60.0, 60.0, 60.0))
    lig_m = workloads.c5_ligand()
    rec = mmo_b200.Receptor.from_mol(rec_m)
    lig = mmo_b200.Ligand.from_mol(lig_m, centered=False)
    base, rem = divmod(total, ctx.world)
    first = ctx.rank * base + min(ctx.rank, rem)
    n = base + (1 if ctx.rank < rem else 0)
    Ln = lig_m.n
    dxyz = [C.c_void_p() for _ in range(3)]
    d_e = C.c_void_p()
    for dp in dxyz:
        ctx.ck(L.mmo_dev_alloc(C.c_size_t(n * Ln * 8), C.byref(dp)))
    ctx.ck(L.mmo_dev_alloc(C.c_size_t(n * 8), C.byref(d_e)))
    chunk = 125_000
    for c0 in range(0, n, chunk):                      # conformer c of the screen = seeded block (first + c0) // chunk
        m = min(chunk, n - c0)
        X = workloads.c5_conformers(lig_m, m, (60.0, 60.0, 60.0), seed=workloads.SEED + 1000 * ((first + c0) // chunk) + (first + c0) % chunk)
        for dp, arr in zip(dxyz, X):
            ctx.ck(L.mmo_h2d(C.c_void_p(dp.value + c0 * Ln * 8), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)))
    # pair accounting: the instrumented kernel build on this rank's whole shard (untimed)
    ns = n
    ctx.ck(L.mmo_set_collect_stats(1))
    ctx.ck(L.mmo_score_coords_dev(rec.h, lig.h, 1, 0, C.c_int64(ns), dxyz[0], dxyz[1], dxyz[2], d_e))
    pe, pi_, pf = C.c_int64(), C.c_int64(), C.c_int64()
    ctx.ck(L.mmo_last_pair_stats(C.byref(pe), C.byref(pi_), C.byref(pf)))
    ctx.ck(L.mmo_set_collect_stats(0))
    flops_per_conf = (27.0 * pi_.value + 8.0 * (pe.value - pi_.value)) / ns
    ts, tf = np.empty(topk), np.empty(topk, np.int64)
    ms_, mf_ = np.empty(topk), np.empty(topk, np.int64)
    nt, on = C.c_int32(), C.c_int32()

    def call():
        ctx.ck(L.mmo_score_coords_dev(rec.h, lig.h, 1, 0, C.c_int64(n), dxyz[0], dxyz[1], dxyz[2], d_e))
        ctx.ck(L.mmo_topk_select_dev(d_e, C.c_int64(n), C.c_int32(topk), C.c_int64(first), ts.ctypes.data_as(_dp), tf.ctypes.data_as(_lp), C.byref(nt)))
        if ctx.world > 1:     # the path's only collective: k x 16 B per rank over NCCL, merged with the reference's tie rule
            ctx.ck(L.mmo_topk_allgather_merge(C.c_int32(topk), nt, ts.ctypes.data_as(_dp), tf.ctypes.data_as(_lp),
                                              ms_.ctypes.data_as(_dp), mf_.ctypes.data_as(_lp), C.byref(on)))
        else:
            ms_[:nt.value] = ts[:nt.value]; mf_[:nt.value] = tf[:nt.value]; on.value = nt.value

    reps = 2 if quick else 3
    with Clocks(ctx.rank == 0) as ck:
        call()                                         # warm-up (scratch arena, NCCL buffers)
        ctx.ck(L.mmo_kernel_timing(1))
        ctx.barrier()
        tot = 0.0
        for _ in range(reps):
            ctx.ck(L.mmo_l2_flush())
            tot += ctx.timed(call)
        k_pair, n_pair = ctx.ktime(K_DIRECT_FP32); k_fix, _ = ctx.ktime(K_HARD_FIX); k_prep, _ = ctx.ktime(K_ITEM_PREP)
    ms_call = ctx.reduce([tot / reps])[0]
    for p in dxyz + [d_e]:
        ctx.ck(L.mmo_dev_free(p))
    cps = total / (ms_call * 1e-3)
    tf_call = cps * flops_per_conf / 1e12 / ctx.world          # per GPU, over the whole call (prepare + sort + pair + fp64 pass + top-k)
    tf_pair = n * flops_per_conf / (k_pair / reps * 1e-3) / 1e12
    return {"workload": f"C5 virtual screen: {total} conformers (70 atoms: 40 heavy + 30 H, explicit coordinates) x 10000-atom receptor sphere, "
                        f"shifted direct UFF path in item mode, top-{topk} conformer ids",
            "conformers": total, "ms": ms_call, "conformers_per_s": cps, "scaling": "strong (the 1e6 conformers are dealt to the GPUs)",
            "nominal_pair_interactions_per_s": cps * rec_m.n * Ln, "evaluated_pair_interactions_per_s": cps * pe.value / ns,
            "pairs_evaluated_per_conformer": pe.value / ns, "pairs_inside_cutoff_per_conformer": pi_.value / ns,
            "fp64_fix_pairs_per_conformer": pf.value / ns,
            "kernel_ms_rank0": {"prepare_and_sort": k_prep / reps, "direct_items_kernel": k_pair / reps, "fp64_close_contact_pass": k_fix / reps,
                                "pair_kernel_launches_per_call": n_pair / reps},
            "topk_merged": int(on.value), "best_E": float(ms_[0]) if on.value else None, "best_conformer": int(mf_[0]) if on.value else None,
            "roofline": {"bound": "fp32", "kernel": "direct_items_kernel (+ prepare, sort, fp64 pass, top-k: whole call)", "achieved": tf_call,
                         "peak": pk["fp32_fma_tflops"], "unit": "TFLOP/s", "frac": tf_call / pk["fp32_fma_tflops"],
                         "pair_kernel_alone_tflops": tf_pair, "pair_kernel_alone_frac": tf_pair / pk["fp32_fma_tflops"],
                         "flop_model": "27 per evaluated pair inside 12 A, 8 outside (SURVEY 8d), counted by the instrumented build on the same conformers"},
            "clocks": ck.summary()}


def leg_fp64_scan(ctx, job_params, n_active, pps, n_rot, quick=False):
    """C2 in MMO_PREC_FP64: fp32 sweep + strict re-scoring of every pose that could be in the exact top-k"""
    import mmo_b200
    from mmo_b200 import ScanResult
    L = ctx.L
    P = job_params
    P.prec = mmo_b200.PREC_FP64
    job = C.c_void_p()
    try:
        ctx.ck(L.mmo_scan_create(C.byref(P), 0, C.byref(job)))
    except Exception:
        P.prec = mmo_b200.PREC_FP32
        raise
    n_slabs = n_active // pps
    steps = 3 if quick else 8
    sl = [int((s * ctx.world + ctx.rank + 0.5) * n_slabs / ((steps + 1) * ctx.world)) for s in range(steps + 1)]
    SR = ScanResult()
    ctx.ck(L.mmo_scan_run(job, sl[0] * pps, pps))
    ctx.barrier()
    tot = 0.0
    with Clocks(ctx.rank == 0) as ck:
        for s in sl[1:]:
            ctx.ck(L.mmo_l2_flush())
            tot += ctx.timed(lambda: ctx.ck(L.mmo_scan_run(job, s * pps, pps)))
        t_fin = ctx.timed(lambda: ctx.ck(L.mmo_scan_result_get(job, None, None, C.byref(SR))))
    ctx.ck(L.mmo_scan_destroy(job))
    P.prec = mmo_b200.PREC_FP32
    ms = ctx.reduce([tot + t_fin])[0]
    return {"workload": "C2 scan in MMO_PREC_FP64 (two-stage: fp32 sweep keeps every pose within 2 delta of the running k-th best, the strict fp64 "
                        "kernel re-scores them: argmin index and top-k order of exhaustive strict scoring)",
            "ms": ms, "steps": steps, "poses_per_s": ctx.world * steps * pps * n_rot / (ms * 1e-3), "strict_rescoring_ms": t_fin,
            "scaling": "weak", "best_score": SR.best_score, "best_frame": SR.best_frame, "clocks": ck.summary()}


def leg_grid_scan(ctx, job_params, n_active, pps, n_rot, quick=False):
    """C2's lattice and rotation set scanned through the interpolated energies (lds with its energy maps: the reference's
    grid mode of the same exhaustive search): prefilter / frame generation, strict_interp_kernel on the z-pair maps,
    reduce and top-k bookkeeping per slab"""
    import mmo_b200
    from mmo_b200 import ScanResult, pqrs, workloads
    L = ctx.L
    P = job_params
    c2 = workloads.load_c2("docked")
    rec_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
    rec = mmo_b200.Receptor.from_mol(rec_m)
    cc = np.array(c2["roi"][:3])
    gd = mmo_b200.Grid.from_box(1.0, *(cc + 23.0))
    ta, tq = pqrs.assign_ff_types([c2["lig"]])
    grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 1.0, gd, ta, tq, mask_bits=sphere_mask_bits(1.0, gd, cc, 21.0), want_host=False)
    old_rec, old_grid = P.rec, P.grid
    P.rec, P.grid = None, grid.h
    job = C.c_void_p()
    try:
        ctx.ck(L.mmo_scan_create(C.byref(P), 0, C.byref(job)))
    except Exception:
        P.rec, P.grid = old_rec, old_grid
        raise
    n_slabs = n_active // pps
    steps = 4 if quick else 16
    sl = [int((s * ctx.world + ctx.rank + 0.5) * n_slabs / ((steps + 1) * ctx.world)) for s in range(steps + 1)]
    SR = ScanResult()
    ctx.ck(L.mmo_scan_run(job, sl[0] * pps, pps))
    ctx.barrier()
    tot = 0.0
    ctx.ck(L.mmo_kernel_timing(1))
    with Clocks(ctx.rank == 0) as ck:
        for s in sl[1:]:
            ctx.ck(L.mmo_l2_flush())
            tot += ctx.timed(lambda: ctx.ck(L.mmo_scan_run(job, s * pps, pps)))
        t_fin = ctx.timed(lambda: ctx.ck(L.mmo_scan_result_get(job, None, None, C.byref(SR))))
    k_ms, k_n = ctx.ktime(K_INTERP)
    pf_ms, pf_n = ctx.ktime(K_PREFILTER)
    rd_ms, rd_n = ctx.ktime(K_REDUCE)
    ctx.ck(L.mmo_scan_destroy(job))
    P.rec, P.grid = old_rec, old_grid
    ms = ctx.reduce([tot + t_fin])[0]
    return {"workload": "C2 lattice x rotation set scanned through the interpolated energies (22 maps, 1 A, z-pair copy), top-1000",
            "ms": ms, "steps": steps, "poses_per_s": ctx.world * steps * pps * n_rot / (ms * 1e-3),
            "lookup_kernel_ms_per_slab": k_ms / max(1, k_n), "frame_generation_kernel_ms_per_slab": pf_ms / steps,
            "reduce_kernels_ms_per_slab": rd_ms / steps, "slab_ms": tot / steps, "scaling": "weak",
            "best_score": SR.best_score, "best_frame": SR.best_frame, "clocks": ck.summary()}


def leg_closure(ctx, quick=False):
    """the literal drop-in shape: `ene_inter : Mol.t -> float` (lds.ml:1952-1979) = ONE pose per call through the C ABI,
    host coordinates in, one double out; what an unmodified frame loop on the host would pay per evaluation"""
    import mmo_b200
    from mmo_b200 import pqrs, workloads
    c2 = workloads.load_c2("docked")
    rec_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
    rec = mmo_b200.Receptor.from_mol(rec_m)
    lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    R, t = workloads.random_poses_in_sphere(64, c2["roi"][:3], 3.0, seed=2)
    xs = (R[:, 0:1] * lig.xs + R[:, 1:2] * lig.ys + R[:, 2:3] * lig.zs) + t[:, 0:1]
    ys = (R[:, 3:4] * lig.xs + R[:, 4:5] * lig.ys + R[:, 5:6] * lig.zs) + t[:, 1:2]
    zs = (R[:, 6:7] * lig.xs + R[:, 7:8] * lig.ys + R[:, 8:9] * lig.zs) + t[:, 2:3]
    cc = np.array(c2["roi"][:3])
    gd = mmo_b200.Grid.from_box(1.0, *(cc + 23.0))
    ta, tq = pqrs.assign_ff_types([c2["lig"]])
    grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 1.0, gd, ta, tq, mask_bits=sphere_mask_bits(1.0, gd, cc, 21.0), want_host=False)
    n = 100 if quick else 400
    out = {}
    for name, fn in (("direct_fp32", lambda k: mmo_b200.Mol.ene_inter_UFF_shifted_brute(rec, lig, xs[k], ys[k], zs[k])),
                     ("direct_fp64", lambda k: mmo_b200.Mol.ene_inter_UFF_shifted_brute(rec, lig, xs[k], ys[k], zs[k], prec=mmo_b200.PREC_FP64)),
                     ("interp", lambda k: mmo_b200.Mol.ene_inter_UFF_interp(grid, lig, xs[k], ys[k], zs[k])),
                     ("intra_nb", lambda k: mmo_b200.Mol.ene_intra_UFFNB_brute(lig, xs[k], ys[k], zs[k]))):
        for k in range(8):
            fn(k)
        t0 = time.perf_counter()
        for k in range(n):
            fn(k % 64)
        out[name + "_us_per_call"] = 1e6 * (time.perf_counter() - t0) / n
    out["workload"] = "one pose per call (docked.mol2, 48 atoms, 1837-atom ROI receptor) through mmo_score_coords / mmo_score_interp_coords / mmo_intra_nb, ctypes overhead included"
    return out


def leg_n4(ctx, quick=False):
    """N4: the Majeux-Caflisch desolvation sums on the reference's own 0.5 A grid over the simulation box (wall clock,
    one GPU; bit-identical to the oracle's loops in tests/test_desolv.py)"""
    import mmo_b200
    from mmo_b200 import workloads
    c2 = workloads.load_c2("ligdecs")
    rm = c2["rec"]
    sdims = mmo_b200.Grid.from_box(0.5, *c2["sim_dims"])
    lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    shell = mmo_b200.Lds.first_solvent_shell(rm.xs, rm.ys, rm.zs, rm.r, 0.5, sdims)
    rec_all = mmo_b200.Receptor.from_mol(rm)
    ctx.ck(ctx.L.mmo_sync())
    t0 = time.perf_counter()
    dh, _ = mmo_b200.Lds.protein_desolv(c2["roi"], rec_all, shell, want_host=False)
    t_prot = time.perf_counter() - t0
    n_pen = 2000 if quick else 20000
    Rp, tp = workloads.random_poses_in_sphere(n_pen, c2["roi"][:3], 6.0, seed=43)
    mmo_b200.Lds.desolvation_penalty(dh, lig, rot9=Rp[:64], trans3=tp[:64])
    t0 = time.perf_counter()
    dp, dl = mmo_b200.Lds.desolvation_penalty(dh, lig, rot9=Rp, trans3=tp)
    t_pen = time.perf_counter() - t0
    return {"workload": "N4 desolvation: Lds.protein_desolv on the 0.5 A simulation grid, then Lds.desolvation_penalty of random poses",
            "grid_dims": list(sdims), "receptor_atoms": rm.n, "protein_desolv_wall_ms": t_prot * 1e3, "penalty_poses": n_pen,
            "penalty_wall_ms": t_pen * 1e3, "penalty_poses_per_s": n_pen / t_pen, "median_prot": float(np.median(dp)),
            "median_lig": float(np.median(dl))}


def run_all(ctx, quick=False, scan_params=None, n_active=0, pps=8, n_rot=0, only=None):
    out = {}
    t0 = time.perf_counter()
    pk = peaks(ctx)
    out["peaks"] = pk

    def leg(names, fn):
        # a leg that fails must not take the headline line (or the other legs) with it: its error is reported in its place
        try:
            res = fn()
        except Exception as e:
            res = tuple({"error": repr(e)} for _ in names) if len(names) > 1 else {"error": repr(e)}
        if len(names) > 1:
            for n, r in zip(names, res):
                out[n] = r
        else:
            out[names[0]] = res

    if only in (None, "c3"):
        leg(("c3_grid_build", "c3_lookup"), lambda: leg_c3(ctx, pk, quick))
    if only in (None, "c4"):
        leg(("c4_mc",), lambda: leg_c4(ctx, quick))
    if only in (None, "c5"):
        leg(("c5_screen",), lambda: leg_c5(ctx, pk, quick))
    if only == "n4":
        leg(("n4_desolvation",), lambda: leg_n4(ctx, quick))
    if only in (None, "closure") and ctx.rank == 0:
        leg(("single_pose_calls",), lambda: leg_closure(ctx, quick))
    if scan_params is not None:
        leg(("c2_fp64_scan",), lambda: leg_fp64_scan(ctx, scan_params, n_active, pps, n_rot, quick))
        leg(("c2_grid_scan",), lambda: leg_grid_scan(ctx, scan_params, n_active, pps, n_rot, quick))
    ctx.ck(ctx.L.mmo_kernel_timing(0))
    out["wall_s"] = time.perf_counter() - t0
    return out


if __name__ == "__main__":
    import mmo_b200
    mmo_b200.init(0)
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    print(json.dumps(run_all(Ctx(mmo_b200.lib()), quick="--quick" in sys.argv, only=only)))
