#!/usr/bin/env python
"""bench.py -- headline benchmark of the MMO scoring hot path on B200.

Workload (BASELINE.json configs[1], "C2"): rigid-body exhaustive scan of data/docked.mol2 against the
3A2J receptor, 100 000 SO(3) rotations x the ROI translation lattice (dx = 1 A), direct pair path
(UFF LJ + Coulomb, shifted, fp32 pair arithmetic with fp64 accumulation and fp64 close-contact
correction), top-1000 kept.  One "step" = one slab of POINTS_PER_STEP in-ROI lattice points x all
rotations through Lds.exhaustive_rigid_ligand_docking's replacement (mmo_scan_*).

  value : poses scored / s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks; the timed
          slabs are strided over ALL in-ROI lattice points (every step samples another part of the sphere) and
          the top-k read-out + NCCL all-gather + merge at the end is inside the timed total
  e2e   : the same slabs through the one-shot host-buffer call mmo_scan() (the 7.2 MB of rotations copied from
          pinned host memory on every call, top-k + argmin read back and merged, NCCL merge at the end, all
          inside the wall-clock region); the figure with the rotation set left resident is reported beside it
  aux   : the other BASELINE.json configs (C3 grid build + lookup, C4 MC chains, C5 conformer screen sharded
          over the N GPUs with the NCCL top-k merge, C2 in fp64 mode), tools/bench_legs.py
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

sys.path.insert(0, os.path.join(ROOT, "tools"))

N_ROT = 100_000
TRANS_STEP = 1.0
TOPK = 1000
POINTS_PER_STEP = 8
WORKLOAD = ("C2 rigid exhaustive scan: docked.mol2 (48 atoms) x 3A2J ROI receptor, 100k SO3 rotations x dx=1.0 A "
            "ROI lattice, direct shifted UFF pair path, top-1000")


def config_dict():
    """the same object in both arms (the driver compares them key by key)"""
    return {"workload": WORKLOAD, "receptor_atoms": 1837, "ligand_atoms": 48, "rotations": N_ROT, "trans_step_A": TRANS_STEP,
            "topk": TOPK, "lattice": "21 x 21 x 22 nodes, 4147 inside the ROI sphere (strict <)",
            "sampling": "every step = one slab of lattice points x all rotations; the slabs of a run are strided over the whole sphere",
            "timing": "L2 flushed (256 MB memset) between timed steps; CUDA events on the launching stream"}


def reference_toolchain():
    """is the reference's own toolchain on this box?  (it never was: recorded, not assumed)"""
    import shutil
    found = {t: shutil.which(t) for t in ("ocamlfind", "dune", "opam", "ocaml")}
    return {"found": found, "status": "present" if all(found[t] for t in ("ocamlfind", "dune")) else "absent",
            "consequence": "the CPU arm is the C restatement of the OCaml source (oracle/, kind 'port'), never presented as OCaml timing"}


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
            "config": config_dict(), "reference_toolchain": reference_toolchain(),
            "pair_interactions_per_s_nominal": pairs,
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
    ap.add_argument("--no-aux", action="store_true", help="skip the C3/C4/C5/fp64 legs (kernel tuning runs)")
    ap.add_argument("--quick-aux", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import mmo_b200
    import bench_legs
    from mmo_b200 import ScanParams, ScanResult
    L = mmo_b200.lib()
    dp, lp = C.POINTER(C.c_double), C.POINTER(C.c_int64)

    def ck(rc):
        if rc != 0:
            raise RuntimeError(L.mmo_last_error().decode())

    dist = torch = None
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
        # first collective = connection set-up (hundreds of ms): done here, so that the merges timed below are merges
        w_s, w_f, w_n = np.empty(1), np.empty(1, np.int64), C.c_int32()
        ck(L.mmo_topk_allgather_merge(1, 0, None, None, w_s.ctypes.data_as(dp), w_f.ctypes.data_as(lp), C.byref(w_n)))
    ctx = bench_legs.Ctx(L, rank, world, dist, torch)

    c2, rec_m = setup_workload()
    rec = mmo_b200.Receptor.from_mol(rec_m)
    lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    # SO3.rotations once per run (lds.ml:1748-1752), in page-locked host memory: the e2e calls copy it from there
    h_rot = C.c_void_p()
    ck(L.mmo_host_alloc(C.c_size_t(N_ROT * 72), C.byref(h_rot)))
    rot = np.ctypeslib.as_array(C.cast(h_rot, dp), shape=(N_ROT, 9))
    ck(L.mmo_so3_rotations(C.c_int32(N_ROT), C.cast(h_rot, dp)))
    # the rigid ligand's constant intra-ligand energy (lds.ml:706-712, 1318-1319) from the library's own kernel
    e_intra = float(mmo_b200.Mol.ene_intra_UFFNB_brute(lig, lig.xs, lig.ys, lig.zs)[0])

    P = ScanParams()
    P.rec, P.grid, P.lig, P.vdw_mask = rec.h, None, lig.h, None
    P.variant, P.prec = mmo_b200.VARIANT_SHIFTED, mmo_b200.PREC_FP32
    P.roi_c = (C.c_double * 3)(*c2["roi"][:3])
    P.roi_r, P.trans_step = c2["roi"][3], TRANS_STEP
    P.n_rot, P.rot9 = N_ROT, C.cast(h_rot, dp)
    P.e_intra_const, P.topk = e_intra, TOPK
    P.first_point, P.n_points = 0, -1
    ts = np.empty(TOPK); tf = np.empty(TOPK, np.int64)
    R2 = ScanResult()

    # ---- cold first call of the process through the one-shot API: fresh device allocations, the upload and k-d sort of
    #      the rotation set, module load -- what a caller pays once per run -------------------------------------------
    P.first_point, P.n_points = 0, 2 * 21 * 21    # the two lowest z planes of the lattice: a few dozen active points
    t0 = time.perf_counter()
    ck(L.mmo_scan(C.byref(P), ts.ctypes.data_as(dp), tf.ctypes.data_as(lp), C.byref(R2)))
    cold_ms = 1e3 * (time.perf_counter() - t0)
    cold_poses = R2.n_scored
    P.first_point, P.n_points = 0, -1

    job = C.c_void_p()
    ck(L.mmo_scan_create(C.byref(P), 0, C.byref(job)))
    n_active = C.c_int64()
    ck(L.mmo_scan_num_points(job, C.byref(n_active)))
    n_active = n_active.value
    total_steps = args.warmup + args.steps
    # Slabs of pps active lattice points; the run's (warm-up + timed) x world slabs are strided over ALL of them, so that
    # every step -- and every rank -- samples another part of the ROI sphere (pocket, surface, buried points alike).
    pps = max(1, min(args.points_per_step, n_active // (total_steps * world)))
    n_slabs = n_active // pps
    stride = n_slabs / float(total_steps * world)

    def slab_first(s):
        return int((s * world + rank + 0.5) * stride) * pps

    timed_slabs = [slab_first(args.warmup + s) for s in range(args.steps)]

    def barrier():
        ctx.barrier()

    # ---- pair accounting for the timed slabs (untimed pass of the instrumented kernel build, the same slabs) ------
    statjob = C.c_void_p()
    ck(L.mmo_scan_create(C.byref(P), 1, C.byref(statjob)))
    fix_pairs = flagged = 0
    a_, b_, c_, f_ = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    for a0 in timed_slabs:
        ck(L.mmo_scan_run(statjob, a0, pps))
        ck(L.mmo_last_pair_stats(C.byref(a_), C.byref(b_), C.byref(c_)))
        ck(L.mmo_last_fix_stats(C.byref(f_)))
        fix_pairs += c_.value; flagged += f_.value
    SR = ScanResult()
    ck(L.mmo_scan_result_get(statjob, None, None, C.byref(SR)))
    stat_poses = max(1, SR.n_scored)
    pairs_eval_per_pose = SR.pairs_evaluated / stat_poses
    pairs_in_per_pose = SR.pairs_inside / stat_poses
    fix_pairs_per_pose = fix_pairs / stat_poses
    flagged_per_pose = flagged / stat_poses
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
    for a0 in timed_slabs:
        ck(L.mmo_l2_flush())           # between timed iterations, outside the timed events
        ck(L.mmo_sync())
        ck(L.mmo_timer_start())
        ck(L.mmo_scan_run(job, a0, pps))
        ck(L.mmo_timer_stop(C.byref(ms)))
        dev_ms += ms.value
    # the end of the job, inside the timed total: read the rank's top-k, all-gather the lists over NCCL, merge
    ms_s, ms_f = np.empty(TOPK), np.empty(TOPK, np.int64)
    out_n = C.c_int32()

    def finish(jobh):
        ck(L.mmo_scan_result_get(jobh, ts.ctypes.data_as(dp), tf.ctypes.data_as(lp), C.byref(SR)))
        if world > 1:
            ck(L.mmo_topk_allgather_merge(TOPK, SR.n_top, ts.ctypes.data_as(dp), tf.ctypes.data_as(lp),
                                          ms_s.ctypes.data_as(dp), ms_f.ctypes.data_as(lp), C.byref(out_n)))
        else:
            ms_s[:SR.n_top] = ts[:SR.n_top]; ms_f[:SR.n_top] = tf[:SR.n_top]; out_n.value = SR.n_top

    t_merge = time.perf_counter()
    ck(L.mmo_timer_start())
    finish(job)
    ck(L.mmo_timer_stop(C.byref(ms)))
    merge_ms = max(ms.value, 1e3 * (time.perf_counter() - t_merge))      # host part of the merge included
    dev_ms += merge_ms
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
    top_n = out_n.value
    best_score, best_frame = SR.best_score, SR.best_frame
    if dist is not None:
        best = torch.tensor([SR.best_score, float(SR.best_frame)], dtype=torch.float64)
        allb = [torch.empty_like(best) for _ in range(world)]
        dist.all_gather(allb, best)
        best_score, best_frame = min((float(b[0]), int(b[1])) for b in allb if int(b[1]) >= 0)
        assert top_n == 0 or ms_s[0] == best_score, "merged top-1 disagrees with the global argmin"
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
    e2e_steps = max(2, min(args.steps, 20))
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
    run_s, run_f = np.empty(2 * TOPK), np.empty(2 * TOPK, np.int64)
    cnt2 = (C.c_int32 * 2)()
    mg_s, mg_f, mg_n = np.empty(TOPK), np.empty(TOPK, np.int64), C.c_int32()

    def e2e_pass(mode, n_steps, first_step):
        """n_steps one-shot calls + the running host merge + the NCCL merge at the end; returns (seconds, poses)"""
        ck(L.mmo_scan_set_rot_cache(mode))
        n_run = 0
        poses = 0
        barrier()
        t0 = time.perf_counter()
        for s in range(n_steps):
            a0 = slab_first(first_step + s)
            P.first_point = act[a0]
            P.n_points = act[a0 + pps - 1] - act[a0] + 1
            tc = time.perf_counter()
            ck(L.mmo_scan(C.byref(P), ts.ctypes.data_as(dp), tf.ctypes.data_as(lp), C.byref(R2)))
            if os.environ.get("MMO_BENCH_DEBUG"):
                print(f"e2e call {s}: raw points {P.first_point}+{P.n_points} scored {R2.n_scored} "
                      f"wall {1e3 * (time.perf_counter() - tc):.1f} ms device {R2.device_ms:.1f} ms", file=sys.stderr)
            poses += R2.n_scored
            # running top-k over the calls (Cpm.TopKeeper keeps k over the whole scan, lds.ml:1055-1064)
            run_s[TOPK:TOPK + R2.n_top] = ts[:R2.n_top]; run_f[TOPK:TOPK + R2.n_top] = tf[:R2.n_top]
            cnt2[0], cnt2[1] = n_run, R2.n_top
            ck(L.mmo_topk_merge(2, TOPK, run_s.ctypes.data_as(dp), run_f.ctypes.data_as(lp), cnt2, mg_s.ctypes.data_as(dp),
                                mg_f.ctypes.data_as(lp), C.byref(mg_n)))
            n_run = mg_n.value
            run_s[:n_run] = mg_s[:n_run]; run_f[:n_run] = mg_f[:n_run]
        if world > 1:
            ck(L.mmo_topk_allgather_merge(TOPK, n_run, run_s.ctypes.data_as(dp), run_f.ctypes.data_as(lp),
                                          mg_s.ctypes.data_as(dp), mg_f.ctypes.data_as(lp), C.byref(mg_n)))
        barrier()
        P.first_point, P.n_points = 0, -1
        return time.perf_counter() - t0, poses

    # warm-up of the one-shot path (untimed, like the W warm-up steps of the resident path)
    e2e_pass(1, args.warmup, 0)
    e2e_s, e2e_poses = e2e_pass(1, e2e_steps, args.warmup)            # headline: the rotation bytes move on every call
    e2e_s0, e2e_poses0 = e2e_pass(0, e2e_steps, args.warmup)          # rotation set left resident (memcmp'ed, not moved)
    # the same calls with the rotation set in ordinary (pageable) host memory, as an OCaml float array would be: the
    # driver stages the copy while the call blocks, but the call is issued behind the slab's kernels
    rot_pageable = np.array(rot, copy=True)
    P.rot9 = rot_pageable.ctypes.data_as(dp)
    e2e_sp, e2e_posesp = e2e_pass(1, e2e_steps, args.warmup)
    P.rot9 = C.cast(h_rot, dp)
    ck(L.mmo_scan_set_rot_cache(1))
    h2d = N_ROT * 72 + pps * 8 + 16          # rotations + the slab's lattice points + the threshold
    h2d_resident = pps * 8 + 16
    d2h = TOPK * 16 + 64

    # ---- reduce over ranks -------------------------------------------------------------------------
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_s, e2e_s0, merge_ms], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_s0, merge_ms = (float(x) for x in tt)
        cnt = torch.tensor([float(e2e_poses), float(launches), float(e2e_poses0)], dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        e2e_poses, launches, e2e_poses0 = int(cnt[0].item()), int(cnt[1].item()), int(cnt[2].item())

    # ---- the other BASELINE configs (every rank takes part: C4 / C5 shard over the ranks) ------------
    aux = None
    if not args.no_aux:
        aux = bench_legs.run_all(ctx, quick=args.quick_aux, scan_params=P, n_active=n_active, pps=pps, n_rot=N_ROT)

    if rank == 0:
        total_poses = poses_per_step * args.steps * world
        value = total_poses / (dev_ms * 1e-3)
        # roofline of the dominant kernel: algorithmic flops (27 inside / 8 outside the cut-off per
        # evaluated pair) / its own CUDA-event time, against the measured FP32 FMA peak
        fp32_peak = C.c_double()
        ck(L.mmo_measure_fp32_peak(C.byref(fp32_peak)))
        k_ms = kms.value / max(1, kn.value)
        flops_per_launch = poses_per_step * (27.0 * pairs_in_per_pose + 8.0 * (pairs_eval_per_pose - pairs_in_per_pose))
        achieved = flops_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        step_ms = (dev_ms - merge_ms) / args.steps
        cfg = config_dict()
        assert cfg["receptor_atoms"] == rec_m.n and cfg["ligand_atoms"] == c2["lig"].n and n_active == 4147
        line = {
            "metric": "ligand poses scored/s", "value": value, "unit": "poses/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "run": {"poses_per_step_per_gpu": poses_per_step, "lattice_points_per_step": pps, "active_lattice_points": n_active,
                    "first_active_point_of_timed_slabs_rank0": timed_slabs,
                    "parallelism": f"slabs dealt round-robin to {world} GPU(s), no data-path collective; the per-GPU top-{TOPK} lists "
                                   "are merged by one NCCL all-gather, inside the timed total",
                    "topk_readout_allgather_merge_ms": merge_ms},
            "pair_interactions_per_s_nominal": value * rec_m.n * c2["lig"].n,
            "pair_interactions_per_s_evaluated": value * pairs_eval_per_pose,
            "pair_interactions_note": "nominal = every (receptor ROI atom, ligand atom) pair of a pose, the count the reference's brute loop "
                                      "visits; evaluated = pairs whose distance the kernel computes after culling (SURVEY 8d's pair)",
            "pairs_evaluated_per_pose": pairs_eval_per_pose, "pairs_inside_cutoff_per_pose": pairs_in_per_pose,
            "fp64_fix_pairs_per_pose": fix_pairs_per_pose, "atoms_flagged_for_fix_per_pose": flagged_per_pose,
            "gpu_launches": launches,
            "e2e": {"value": e2e_poses / e2e_s, "unit": "poses/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "warmup": args.warmup,
                    "api": "mmo_scan() one-shot per step, host buffers (rotations in pinned memory, copied to the device on every call "
                           "and compared there with the resident set: the library default, mmo_scan_set_rot_cache(1)), running top-k "
                           "merged on the host, NCCL merge at the end",
                    "rotations_left_resident": {"value": e2e_poses0 / e2e_s0, "h2d_bytes_per_step": h2d_resident,
                                                "note": "mmo_scan_set_rot_cache(0): the same rotation bytes handed over again are recognised by "
                                                        "one memcmp on the host and not uploaded; slower here than moving them"},
                    "rotations_in_pageable_memory": {"value_per_gpu_rank0": e2e_posesp / e2e_sp,
                                                     "note": "the same one-shot calls with the rotation set in ordinary host memory "
                                                             "(what an OCaml float array is): rank 0's own rate"},
                    "cold_first_call": {"ms": cold_ms, "poses": cold_poses,
                                        "note": "first mmo_scan of the process: device allocations, rotation upload + k-d visiting order, module load"}},
            "roofline": {"bound": "fp32", "kernel": "direct_fp32_kernel", "achieved": achieved, "peak": fp32_peak.value,
                         "unit": "TFLOP/s", "frac": achieved / fp32_peak.value if fp32_peak.value else None,
                         "traffic": ncu_traffic()[0], "traffic_unit": "bytes of DRAM per launch", "traffic_source": ncu_traffic()[1],
                         "algorithmic_bytes_per_launch": rec_m.n * 16 + poses_per_step * 16,
                         "peak_source": "measured on this box: FP32 FMA chain (mmo_measure_fp32_peak); "
                         "MEASURED_PEAKS.json has no FP32 ALU figure", "kernel_ms_per_launch": k_ms,
                         "kernel_share_of_step": kms.value / (dev_ms - merge_ms) if dev_ms else None,
                         "hard_fix_ms_per_launch": fix_ms.value / max(1, kn.value),
                         "step_level_frac": (flops_per_launch / (step_ms * 1e-3) / 1e12) / fp32_peak.value if step_ms > 0 else None,
                         "flops_per_launch": flops_per_launch,
                         "sampled": "average over the timed slabs, which are strided over the whole ROI sphere"},
            "clocks": summarise_clocks(samples),
            "result": {"best_score": best_score, "best_frame": best_frame, "topk_merged": top_n},
            "reference_toolchain": reference_toolchain(),
        }
        if prefilter is not None:
            line["with_vdw_prefilter"] = prefilter
        if aux is not None:
            line["aux"] = aux
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
    ck(L.mmo_host_free(h_rot))
    if dist is not None:
        dist.barrier()
        ck(L.mmo_nccl_finalize())
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
