#!/usr/bin/env python
"""bench.py -- headline benchmark of the MMO scoring hot path on B200.

Workload (BASELINE.json configs[1], "C2"): rigid-body exhaustive scan of data/docked.mol2 against the
3A2J receptor, 100 000 SO(3) rotations x the ROI translation lattice (dx = 1 A), direct pair path
(UFF LJ + Coulomb, shifted, fp32 pair arithmetic with fp64 accumulation and fp64 close-contact
correction), top-1000 kept.  One "step" = one slab of POINTS_PER_STEP in-ROI lattice points x all
rotations through Lds.exhaustive_rigid_ligand_docking's replacement (mmo_scan_*).

  value : poses scored / s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e   : the same slabs through the one-shot host-buffer call mmo_scan() (rotations uploaded,
          top-k + argmin read back inside the timed region)
  roofline : the dominant kernel (direct_fp32_kernel) against the FP32 FMA peak measured on this box
  cpu_baseline / --impl reference : the CPU restatement of the OCaml reference (oracle/), all host
          cores, on a bounded sample of the same poses (OCaml itself is not installable here)

N > 1 (torchrun, one rank per GPU): the active lattice points are sharded over the ranks, no data
path collective; the per-rank top-k lists are merged with one small NCCL all-gather at the end.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_ROT = 100_000
TRANS_STEP = 1.0
TOPK = 1000
POINTS_PER_STEP = 8
WORKLOAD = ("C2 rigid exhaustive scan: docked.mol2 (48 atoms) x 3A2J ROI receptor, 100k SO3 rotations x dx=1.0 A "
            "ROI lattice, direct shifted UFF pair path, top-1000")


def clocks_sampler(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", dev, f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            return
        stop.wait(0.2)


def summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": reasons, "samples": len(sm)}


def setup_workload():
    from mmo_b200 import workloads
    c2 = workloads.load_c2("docked")
    rl = workloads.lig_radius(c2["centered"])
    rec_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + rl + 12.0)
    return c2, rec_m


def host_threads():
    """every host thread this process may run on -- not OMP_NUM_THREADS, which torchrun sets to 1 for N > 1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(c2, rec_m, rot, points_xyz, n_poses, nthreads):
    """Times the oracle (C restatement of mol.ml:822-849 + mol.ml:669-672) on n_poses of the workload."""
    import oracle
    rng = np.random.default_rng(1)
    ri = rng.integers(0, len(rot), n_poses)
    pi = rng.integers(0, len(points_xyz), n_poses)
    cx, cy, cz = c2["centered"]
    t0 = time.perf_counter()
    oracle.score_poses_mt(rec_m, cx, cy, cz, c2["lig"].q, c2["lig"].anum, rot[ri], points_xyz[pi], nthreads)
    return time.perf_counter() - t0


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one direct_fp32_kernel launch of this workload, from
    the committed ncu --set full capture (profiles/traffic.json; never measured under the timed run)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)["direct_fp32_kernel"]
        return t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
    except Exception:
        return None, None


def poses_last_slab(SR, pps):
    return pps * N_ROT


def lattice_points(roi, step):
    """in-ROI lattice nodes in the reference's loop order (lds.ml:1065-1091)"""
    import oracle
    lo = [roi[d] - roi[3] for d in range(3)]
    hi = [roi[d] + roi[3] for d in range(3)]
    dims = oracle.grid_from_box(step, hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2])
    pts = []
    for k in range(dims[2]):
        for j in range(dims[1]):
            for i in range(dims[0]):
                p = (lo[0] + oracle.grid_node(step, dims[0], i), lo[1] + oracle.grid_node(step, dims[1], j),
                     lo[2] + oracle.grid_node(step, dims[2], k))
                if (roi[0] - p[0]) ** 2 + (roi[1] - p[1]) ** 2 + (roi[2] - p[2]) ** 2 < roi[3] ** 2:
                    pts.append(p)
    return np.array(pts)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    import oracle
    c2, rec_m = setup_workload()
    rot = oracle.so3_rotations(N_ROT)
    pts = lattice_points(c2["roi"], TRANS_STEP)
    nthreads = host_threads()
    # size a step to ~2.5 s of CPU work
    dt = cpu_sample(c2, rec_m, rot, pts, 64 * nthreads, nthreads)
    per_step = max(nthreads, int(64 * nthreads * 2.5 / max(dt, 1e-6)))
    for _ in range(args.warmup):
        cpu_sample(c2, rec_m, rot, pts, max(nthreads, per_step // 8), nthreads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_sample(c2, rec_m, rot, pts, per_step, nthreads)
    value = per_step * args.steps / t
    pairs = value * rec_m.n * c2["lig"].n
    sample = f"{per_step} random (rotation, lattice point) poses of the workload per step"
    line = {"impl": "reference", "metric": "ligand poses scored/s", "value": value, "unit": "poses/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "receptor_atoms": rec_m.n, "ligand_atoms": c2["lig"].n},
            "pair_interactions_per_s": pairs,
            "cpu_baseline": {"value": value, "unit": "poses/s", "cores": nthreads, "kind": "port", "sample": sample,
                             "note": "C restatement of the OCaml reference (oracle/mmo_oracle.c, gcc -O2 "
                                     "-ffp-contract=off, OpenMP over poses); OCaml is not installable here"},
            "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_JSON_FD = None


def emit(line):
    """the ONE line of stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries the one JSON line and nothing else: whatever a library prints on fd 1 (NCCL's version
    # banner under NCCL_DEBUG=VERSION, for one) is sent to stderr; the JSON goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points-per-step", type=int, default=POINTS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefilter-run", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import mmo_b200
    from mmo_b200 import ScanParams, ScanResult
    L = mmo_b200.lib()

    def ck(rc):
        if rc != 0:
            raise RuntimeError(L.mmo_last_error().decode())

    dist = None
    mmo_b200.init(local_rank)
    if world > 1:
        # control plane (barrier, max over ranks, hand-out of the NCCL id): torch.distributed/gloo on the host.
        # data plane: the library's own ncclAllGather of the per-GPU top-k lists (mmo_topk_allgather_merge).
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")
        nid = np.zeros(128, np.uint8)
        if rank == 0:
            ck(L.mmo_nccl_unique_id(nid.ctypes.data_as(C.POINTER(C.c_uint8))))
        tid = torch.from_numpy(nid)
        dist.broadcast(tid, 0)
        ck(L.mmo_nccl_init(rank, world, tid.numpy().ctypes.data_as(C.POINTER(C.c_uint8))))
        # first collective = connection set-up (hundreds of ms): done here, so that the merge timed at the end is the merge
        w_s, w_f, w_n = np.empty(1), np.empty(1, np.int64), C.c_int32()
        ck(L.mmo_topk_allgather_merge(1, 0, None, None, w_s.ctypes.data_as(C.POINTER(C.c_double)),
                                      w_f.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(w_n)))

    c2, rec_m = setup_workload()
    rec = mmo_b200.Receptor.from_mol(rec_m)
    lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    rot = mmo_b200.SO3.rotations(N_ROT)
    import oracle   # checker + CPU arm only
    e_intra = float(oracle.ene_intra(c2["lig"], lig.xs, lig.ys, lig.zs)[0])

    P = ScanParams()
    P.rec, P.grid, P.lig, P.vdw_mask = rec.h, None, lig.h, None
    P.variant, P.prec = mmo_b200.VARIANT_SHIFTED, mmo_b200.PREC_FP32
    P.roi_c = (C.c_double * 3)(*c2["roi"][:3])
    P.roi_r, P.trans_step = c2["roi"][3], TRANS_STEP
    P.n_rot, P.rot9 = N_ROT, rot.ctypes.data_as(C.POINTER(C.c_double))
    P.e_intra_const, P.topk = e_intra, TOPK
    P.first_point, P.n_points = 0, -1

    job = C.c_void_p()
    ck(L.mmo_scan_create(C.byref(P), 0, C.byref(job)))
    n_active = C.c_int64()
    ck(L.mmo_scan_num_points(job, C.byref(n_active)))
    n_active = n_active.value
    pps = args.points_per_step
    total_steps = args.warmup + args.steps
    # rank r owns a contiguous block of the active points (weak scaling: the same slab size per GPU)
    # slabs are dealt round-robin: step s of rank r takes slab s*world + r of the active points, so that
    # every GPU sweeps a statistically similar part of the pocket (weak scaling, same slab size per GPU)
    assert (total_steps * world + 2 * world) * pps <= n_active, "not enough lattice points for this many steps"

    def slab_first(s):
        return (s * world + rank) * pps

    def barrier():
        ck(L.mmo_sync())
        if dist is not None:
            dist.barrier()

    # ---- pair accounting for the timed slabs (untimed pass of the instrumented kernel build) ------
    statjob = C.c_void_p()
    ck(L.mmo_scan_create(C.byref(P), 1, C.byref(statjob)))
    for s_ in range(args.warmup, args.warmup + min(2, args.steps)):
        ck(L.mmo_scan_run(statjob, slab_first(s_), pps))
    SR = ScanResult()
    ck(L.mmo_scan_result_get(statjob, None, None, C.byref(SR)))
    stat_poses = SR.n_scored
    pairs_eval_per_pose = SR.pairs_evaluated / max(1, stat_poses)
    pairs_in_per_pose = SR.pairs_inside / max(1, stat_poses)
    a_, b_, c_, f_ = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    ck(L.mmo_last_pair_stats(C.byref(a_), C.byref(b_), C.byref(c_)))      # the last slab of the accounting pass
    ck(L.mmo_last_fix_stats(C.byref(f_)))
    fix_pairs_per_pose = c_.value / max(1, poses_last_slab(SR, pps))
    flagged_per_pose = f_.value / max(1, poses_last_slab(SR, pps))
    ck(L.mmo_scan_destroy(statjob))

    # ---- device-resident timed region --------------------------------------------------------------
    for s in range(args.warmup):
        ck(L.mmo_scan_run(job, slab_first(s), pps))
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples), daemon=True)
    if rank == 0:
        th.start()
    ck(L.mmo_kernel_timing(1))
    launches0 = mmo_b200.launch_count()
    barrier()
    dev_ms = 0.0
    ms = C.c_float()
    for s in range(args.steps):
        ck(L.mmo_l2_flush())           # between timed iterations, outside the timed events
        ck(L.mmo_sync())
        ck(L.mmo_timer_start())
        ck(L.mmo_scan_run(job, slab_first(args.warmup + s), pps))
        ck(L.mmo_timer_stop(C.byref(ms)))
        dev_ms += ms.value
    barrier()
    launches = mmo_b200.launch_count() - launches0
    kms, kn = C.c_double(), C.c_int64()
    ck(L.mmo_kernel_time_get(0, C.byref(kms), C.byref(kn)))
    fix_ms = C.c_double()
    ck(L.mmo_kernel_time_get(1, C.byref(fix_ms), None))
    ck(L.mmo_kernel_timing(0))
    stop.set()
    if rank == 0:
        th.join(timeout=10)      # a concurrent nvidia-smi query stalls cudaMalloc/cudaFree in the e2e calls below
    ck(L.mmo_scan_result_get(job, None, None, C.byref(SR)))
    poses_per_step = pps * N_ROT

    # ---- the same scan with the reference's vdW bitmask prefilter on (lds.ml:1094-1097: clashing poses are
    #      rejected before scoring); reported beside the headline, not instead of it ----------------------
    prefilter = None
    if world == 1 and not args.no_prefilter_run:
        from mmo_b200 import workloads
        dims = mmo_b200.Grid.from_box(workloads.GRID_STEP, *c2["sim_dims"])
        m = c2["rec"]
        mask = mmo_b200.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, workloads.GRID_STEP, dims)
        P.vdw_mask = mask.h
        mjob = C.c_void_p()
        ck(L.mmo_scan_create(C.byref(P), 0, C.byref(mjob)))
        nm = C.c_int64()
        ck(L.mmo_scan_num_points(mjob, C.byref(nm)))
        msteps = max(1, min(args.steps, nm.value // pps - 1))
        ck(L.mmo_scan_run(mjob, 0, pps))                       # warm-up slab
        MR0 = ScanResult()
        ck(L.mmo_scan_result_get(mjob, None, None, C.byref(MR0)))
        pf_ms = 0.0
        for s_ in range(msteps):
            ck(L.mmo_l2_flush())
            ck(L.mmo_sync())
            ck(L.mmo_timer_start())
            # slabs spread over the whole ROI sphere (pocket, surface and buried lattice points alike)
            ck(L.mmo_scan_run(mjob, ((1 + s_) * (nm.value // pps - 1) // (msteps + 1)) * pps, pps))
            ck(L.mmo_timer_stop(C.byref(ms)))
            pf_ms += ms.value
        MR = ScanResult()
        ck(L.mmo_scan_result_get(mjob, None, None, C.byref(MR)))
        cand = msteps * pps * N_ROT
        scored = MR.n_scored - MR0.n_scored
        prefilter = {"candidate_poses_per_s": cand / (pf_ms * 1e-3), "scored_poses_per_s": scored / (pf_ms * 1e-3),
                     "surviving_fraction": scored / cand, "steps": msteps, "ms_per_step": pf_ms / msteps,
                     "active_lattice_points": nm.value,
                     "note": "vdW_clash_OR bitmask (0.5 A grid, all receptor atoms) before scoring, as lds --ext does; the 3A2J "
                             "pocket is buried: almost no pose of a blind 48-atom scan survives, the prefilter is the whole cost"}
        ck(L.mmo_scan_destroy(mjob))
        P.vdw_mask = None

    # ---- end-to-end: host buffers through the one-shot call, copies inside the timed region ---------
    ts = np.empty(TOPK); tf = np.empty(TOPK, np.int64)
    e2e_steps = max(2, min(args.steps, 20))
    R2 = ScanResult()
    # translate active-point slabs to raw lattice-point sub-ranges for mmo_scan()
    lat = SR.lattice_dims
    nvox = lat[0] * lat[1] * lat[2]
    act = []   # raw ids of the active points, recomputed on the host as scan_setup does
    lo = [c2["roi"][d] - c2["roi"][3] for d in range(3)]
    for p in range(nvox):
        k = p // (lat[0] * lat[1]); j = (p - k * lat[0] * lat[1]) // lat[0]; i = p - (k * lat[0] * lat[1] + j * lat[0])
        pos = (lo[0] + i * TRANS_STEP, lo[1] + j * TRANS_STEP, lo[2] + k * TRANS_STEP)
        if sum((c2["roi"][d] - pos[d]) ** 2 for d in range(3)) < c2["roi"][3] ** 2:
            act.append(p)
    assert len(act) == n_active
    # warm-up of the one-shot path (untimed, like the W warm-up steps of the resident path): the first call after the
    # resident jobs were destroyed pays for fresh device allocations (up to 0.3 s once per process)
    for s in range(args.warmup):
        a0 = slab_first(s)
        P.first_point = act[a0]
        P.n_points = act[a0 + pps - 1] - act[a0] + 1
        ck(L.mmo_scan(C.byref(P), ts.ctypes.data_as(C.POINTER(C.c_double)), tf.ctypes.data_as(C.POINTER(C.c_int64)),
                      C.byref(R2)))
    barrier()
    t0 = time.perf_counter()
    e2e_poses = 0
    for s in range(e2e_steps):
        a0 = slab_first(args.warmup + s)
        P.first_point = act[a0]
        P.n_points = act[a0 + pps - 1] - act[a0] + 1
        tc = time.perf_counter()
        ck(L.mmo_scan(C.byref(P), ts.ctypes.data_as(C.POINTER(C.c_double)), tf.ctypes.data_as(C.POINTER(C.c_int64)),
                      C.byref(R2)))
        if os.environ.get("MMO_BENCH_DEBUG"):
            print(f"e2e call {s}: raw points {P.first_point}+{P.n_points} scored {R2.n_scored} "
                  f"wall {1e3 * (time.perf_counter() - tc):.1f} ms device {R2.device_ms:.1f} ms", file=sys.stderr)
        e2e_poses += R2.n_scored
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = N_ROT * 9 * 8 + pps * 8 + 16
    d2h = TOPK * 16 + 64

    # ---- reduce over ranks -------------------------------------------------------------------------
    top_n = SR.n_top
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = tt[0].item(), tt[1].item()
        cnt = torch.tensor([float(e2e_poses), float(launches)], dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        e2e_poses, launches = int(cnt[0].item()), int(cnt[1].item())
        # the only collective of the path: NCCL all-gather of the per-GPU top-k (k x 16 B per rank) + merge
        ck(L.mmo_scan_result_get(job, ts.ctypes.data_as(C.POINTER(C.c_double)),
                                 tf.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(SR)))
        ms_s, ms_f = np.empty(TOPK), np.empty(TOPK, np.int64)
        out_n = C.c_int32()
        t_ag = time.perf_counter()
        ck(L.mmo_topk_allgather_merge(TOPK, SR.n_top, ts.ctypes.data_as(C.POINTER(C.c_double)),
                                      tf.ctypes.data_as(C.POINTER(C.c_int64)), ms_s.ctypes.data_as(C.POINTER(C.c_double)),
                                      ms_f.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(out_n)))
        t_ag = time.perf_counter() - t_ag
        top_n = out_n.value
        best = torch.tensor([SR.best_score, float(SR.best_frame)], dtype=torch.float64)
        allb = [torch.empty_like(best) for _ in range(world)]
        dist.all_gather(allb, best)
        gb = min((float(b[0]), int(b[1])) for b in allb if int(b[1]) >= 0)
        SR.best_score, SR.best_frame = gb[0], gb[1]
        assert top_n == 0 or ms_s[0] == gb[0], "merged top-1 disagrees with the global argmin"

    if rank == 0:
        total_poses = poses_per_step * args.steps * world
        value = total_poses / (dev_ms * 1e-3)
        pairs_nominal = value * rec_m.n * c2["lig"].n
        # roofline of the dominant kernel: algorithmic flops (27 inside / 8 outside the cut-off per
        # evaluated pair) / its own CUDA-event time, against the measured FP32 FMA peak
        fp32_peak = C.c_double()
        ck(L.mmo_measure_fp32_peak(C.byref(fp32_peak)))
        k_ms = kms.value / max(1, kn.value)
        flops_per_launch = poses_per_step * (27.0 * pairs_in_per_pose + 8.0 * (pairs_eval_per_pose - pairs_in_per_pose))
        achieved = flops_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        line = {
            "metric": "ligand poses scored/s", "value": value, "unit": "poses/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "receptor_atoms": rec_m.n, "ligand_atoms": c2["lig"].n,
                       "poses_per_step_per_gpu": poses_per_step, "timing": "L2 flushed (256 MB memset) between timed steps",
                       "parallelism": f"lattice-point slabs dealt round-robin to {world} GPU(s), no data-path collective, top-k merged by one NCCL all-gather"},
            "pair_interactions_per_s": pairs_nominal,
            "pairs_evaluated_per_pose": pairs_eval_per_pose, "pairs_inside_cutoff_per_pose": pairs_in_per_pose,
            "fp64_fix_pairs_per_pose": fix_pairs_per_pose, "atoms_flagged_for_fix_per_pose": flagged_per_pose,
            "gpu_launches": launches,
            "e2e": {"value": e2e_poses / e2e_s, "unit": "poses/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "warmup": args.warmup, "api": "mmo_scan() one-shot, host buffers"},
            "roofline": {"bound": "fp32", "kernel": "direct_fp32_kernel", "achieved": achieved, "peak": fp32_peak.value,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak.value if fp32_peak.value else None,
                         "traffic": ncu_traffic()[0], "traffic_unit": "bytes of DRAM per launch", "traffic_source": ncu_traffic()[1],
                         "algorithmic_bytes_per_launch": rec_m.n * 16 + poses_per_step * 16,
                         "peak_source": "measured on this box: FP32 FMA chain (mmo_measure_fp32_peak); "
                         "MEASURED_PEAKS.json has no FP32 ALU figure", "kernel_ms_per_launch": k_ms,
                         "kernel_share_of_step": kms.value / dev_ms if dev_ms else None,
                         "hard_fix_ms_per_launch": fix_ms.value / max(1, kn.value),
                         "flops_per_launch": flops_per_launch},
            "clocks": summarise_clocks(samples),
            "result": {"best_score": SR.best_score, "best_frame": SR.best_frame, "topk_merged": top_n,
                       "topk_allgather_ms": (1e3 * t_ag if dist is not None else None)},
        }
        if prefilter is not None:
            line["with_vdw_prefilter"] = prefilter
        if not args.no_cpu_baseline and world == 1:
            pts = lattice_points(c2["roi"], TRANS_STEP)
            nthreads = host_threads()
            dt = cpu_sample(c2, rec_m, rot, pts, 32 * nthreads, nthreads)
            n_s = max(nthreads, int(32 * nthreads * 12.0 / max(dt, 1e-6)))
            dt = cpu_sample(c2, rec_m, rot, pts, n_s, nthreads)
            line["cpu_baseline"] = {"value": n_s / dt, "unit": "poses/s", "cores": nthreads, "kind": "port",
                                    "sample": f"{n_s} random (rotation, lattice point) poses of the same workload, "
                                              f"{dt:.1f} s, C restatement of the OCaml reference (oracle/), OpenMP"}
        emit(line)
    ck(L.mmo_scan_destroy(job))
    if dist is not None:
        dist.barrier()
        ck(L.mmo_nccl_finalize())
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
