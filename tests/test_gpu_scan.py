"""Exhaustive rigid scan (lds.ml:1040-1114) on the GPU vs the oracle's literal loop nest."""
import numpy as np
import pytest

from conftest import tol_ok
from mmo_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(gpu, orc, c2, c2_roi_rec):
    rec = gpu.Receptor.from_mol(c2_roi_rec)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    dims = gpu.Grid.from_box(workloads.GRID_STEP, *c2["sim_dims"])
    m = c2["rec"]
    mask = gpu.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, workloads.GRID_STEP, dims)
    e_intra = orc.ene_intra(c2["lig"], lig.xs, lig.ys, lig.zs)[0]
    return rec, lig, mask, dims, e_intra


@pytest.mark.parametrize("use_mask", [False, True])
def test_scan_fp64_reproduces_argmin_and_topk_order(gpu, orc, c2, setup, use_mask):
    _, lig, mask, dims, e_intra = setup
    rot = gpu.SO3.rotations(48)
    # an ROI straddling the protein surface (the crystallographic pocket is buried: there every pose
    # of a 48-atom ligand trips the bitmask prefilter), coarse lattice: pocket, surface, solvent points
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2] + 26.0, 10.0)
    rec_m = c2["rec"]                       # whole receptor: the carved one is centred on the real ROI
    rec = gpu.Receptor.from_mol(rec_m)
    kw = dict(vdw_mask=mask.bits, m_step=workloads.GRID_STEP, m_dims=dims) if use_mask else {}
    want = orc.scan(rec_m, c2["lig"], lig.xs, lig.ys, lig.zs, roi, 4.0, rot, 25, scorer=0,
                    e_intra_const=e_intra, **kw)
    got = gpu.Lds.exhaustive_rigid_ligand_docking(25, roi, 4.0, rot, lig, rec=rec, vdw_mask=mask if use_mask else None,
                                                  e_intra_const=e_intra, prec=gpu.PREC_FP64)
    assert got["lattice_dims"] == want["lattice_dims"]
    assert got["n_candidates"] == want["n_candidates"] and got["n_scored"] == want["n_scored"]
    assert got["best_frame"] == want["best_frame"] and got["best_score"] == want["best_score"]
    assert np.array_equal(got["top_frames"], want["top_frames"])
    assert np.array_equal(got["top_scores"], want["top_scores"])
    if use_mask:
        assert 0 < want["n_scored"] < want["n_candidates"]
        # same scan, fp32 pair path: identical survivors, energies within tolerance
        got32 = gpu.Lds.exhaustive_rigid_ligand_docking(25, roi, 4.0, rot, lig, rec=rec, vdw_mask=mask,
                                                        e_intra_const=e_intra, prec=gpu.PREC_FP32)
        assert got32["n_scored"] == want["n_scored"]
        assert tol_ok(got32["top_scores"], want["top_scores"]).all()


def test_scan_fp32_within_tolerance_and_same_winner(gpu, orc, c2, c2_roi_rec, setup):
    rec, lig, mask, dims, e_intra = setup
    rot = gpu.SO3.rotations(200)
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 3.0)
    want = orc.scan(c2_roi_rec, c2["lig"], lig.xs, lig.ys, lig.zs, roi, 1.0, rot, 50, scorer=0, e_intra_const=e_intra)
    got = gpu.Lds.exhaustive_rigid_ligand_docking(50, roi, 1.0, rot, lig, rec=rec, e_intra_const=e_intra,
                                                  prec=gpu.PREC_FP32)
    assert got["n_scored"] == want["n_scored"]
    assert tol_ok(got["top_scores"], want["top_scores"]).all()
    # ordering can only differ between poses whose reference scores are closer than the tolerance
    for a, b in zip(got["top_frames"], want["top_frames"]):
        if a != b:
            sa = orc.scan(c2_roi_rec, c2["lig"], lig.xs, lig.ys, lig.zs, roi, 1.0, rot, 0, e_intra_const=e_intra,
                          score_frames=[a, b])
            assert abs(sa[0] - sa[1]) <= 2 * max(1e-6 * abs(sa[0]), 1e-4)
    assert got["best_frame"] == want["best_frame"] or abs(got["best_score"] - want["best_score"]) < 2e-4


def test_scan_sharded_ranges_merge_to_the_full_scan(gpu, c2, setup):
    """lattice-point sub-ranges (the multi-GPU sharding unit) + mmo_topk_merge == one full scan"""
    import ctypes as C
    rec, lig, mask, dims, e_intra = setup
    rot = gpu.SO3.rotations(64)
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 4.0)
    k = 20
    full = gpu.Lds.exhaustive_rigid_ligand_docking(k, roi, 2.0, rot, lig, rec=rec, prec=gpu.PREC_FP64)
    nvox = int(np.prod(full["lattice_dims"]))
    cuts = [0, nvox // 3, 2 * nvox // 3, nvox]
    parts = [gpu.Lds.exhaustive_rigid_ligand_docking(k, roi, 2.0, rot, lig, rec=rec, prec=gpu.PREC_FP64,
                                                     first_point=cuts[i], n_points=cuts[i + 1] - cuts[i]) for i in range(3)]
    assert sum(p["n_scored"] for p in parts) == full["n_scored"]
    S = np.full((3, k), np.inf); F = np.zeros((3, k), np.int64); cnt = np.zeros(3, np.int32)
    for i, p in enumerate(parts):
        n = len(p["top_scores"]); S[i, :n] = p["top_scores"]; F[i, :n] = p["top_frames"]; cnt[i] = n
    os_, of_ = np.empty(k), np.empty(k, np.int64)
    n = C.c_int32()
    dp, lp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    assert gpu.lib().mmo_topk_merge(3, k, S.ctypes.data_as(dp), F.ctypes.data_as(lp), cnt.ctypes.data_as(ip),
                                    os_.ctypes.data_as(dp), of_.ctypes.data_as(lp), C.byref(n)) == 0
    assert np.array_equal(os_[:n.value], full["top_scores"]) and np.array_equal(of_[:n.value], full["top_frames"])
    best = min(parts, key=lambda p: (p["best_score"], p["best_frame"]))
    assert best["best_frame"] == full["best_frame"]


def test_scan_interpolated_scorer(gpu, orc, c2, c2_roi_rec, setup):
    from mmo_b200 import pqrs
    rec, lig, mask, dims, e_intra = setup
    # small energy grid around the ROI centre: origin-anchored grid covering the pocket
    c = np.array(c2["roi"][:3])
    gd = gpu.Grid.from_box(1.0, *(c + 16.0))
    ta, tq = pqrs.assign_ff_types([c2["lig"]])
    gmask = orc.bitmask_sphere(1.0, gd, c, 14.0)
    g, maps = gpu.Lds.pre_calculate_FF_components_grid(rec, 1.0, gd, ta, tq, mask_bits=gmask)
    rot = gpu.SO3.rotations(32)
    roi = (c[0], c[1], c[2], 2.5)
    want = orc.scan(None, c2["lig"], lig.xs, lig.ys, lig.zs, roi, 1.0, rot, 10, scorer=2, maps=maps, g_step=1.0, g_dims=gd)
    got = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, grid=g)
    assert np.array_equal(got["top_frames"], want["top_frames"]) and np.array_equal(got["top_scores"], want["top_scores"])
    assert got["best_frame"] == want["best_frame"]


def test_scan_fp64_two_stage_equals_exhaustive_strict_scoring(gpu, orc, c2, setup):
    """MMO_PREC_FP64 scans sweep in fp32 and re-score only the poses within the error margin of the k-th
    best; the outcome must be what strict-fp64 scoring of EVERY pose gives (argmin frame, top-k order)."""
    rec, lig, mask, dims, e_intra = setup
    n_rot, k = 3000, 200
    rot = gpu.SO3.rotations(n_rot)
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 2.6)
    got = gpu.Lds.exhaustive_rigid_ligand_docking(k, roi, 1.0, rot, lig, rec=rec, e_intra_const=e_intra, prec=gpu.PREC_FP64)
    # every pose of the same scan through the strict kernel (bit-identical to the oracle, see test_gpu_direct)
    ld = got["lattice_dims"]
    lo = [roi[d] - roi[3] for d in range(3)]
    frames, R, T = [], [], []
    for kk in range(ld[2]):
        for jj in range(ld[1]):
            for ii in range(ld[0]):
                pos = np.array([lo[0] + orc.grid_node(1.0, ld[0], ii), lo[1] + orc.grid_node(1.0, ld[1], jj),
                                lo[2] + orc.grid_node(1.0, ld[2], kk)])
                if ((np.array(roi[:3]) - pos) ** 2).sum() < roi[3] ** 2:
                    pt = ii + jj * ld[0] + kk * ld[0] * ld[1]
                    frames.append(np.arange(n_rot, dtype=np.int64) + n_rot * pt)
                    R.append(rot); T.append(np.tile(pos, (n_rot, 1)))
    frames = np.concatenate(frames); R = np.concatenate(R); T = np.concatenate(T)
    assert len(frames) == got["n_scored"]
    e = e_intra + gpu.Mol.score_poses(rec, lig, R, T, prec=gpu.PREC_FP64)
    order = np.lexsort((frames, e))[:k]
    assert np.array_equal(got["top_frames"], frames[order])
    assert np.array_equal(got["top_scores"], e[order])
    assert got["best_frame"] == frames[order[0]] and got["best_score"] == e[order[0]]


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_scan_topk_equals_sorting_every_pose(gpu, orc, c2, setup, prec):
    """size-independent property at a slab large enough for the first-slab threshold estimate (k-th smallest
    per-block minimum): the scan's top-k and argmin == scoring every pose of the loop nest and sorting by
    (score, frame).  fp64: two-stage == strict scoring, bit for bit; fp32: within the accuracy contract (the fp32
    sum of a pose depends on its warp's companions at the 1e-7 level: x' = x - c is centred per warp)."""
    rec, lig, mask, dims, e_intra = setup
    n_rot, k = 1500, 7
    rot = gpu.SO3.rotations(n_rot)
    roi = (c2["roi"][0] + 3.0, c2["roi"][1], c2["roi"][2], 2.2)
    p = gpu.PREC_FP32 if prec == "fp32" else gpu.PREC_FP64
    got = gpu.Lds.exhaustive_rigid_ligand_docking(k, roi, 1.0, rot, lig, rec=rec, e_intra_const=e_intra, prec=p)
    assert got["n_scored"] == got["n_candidates"] >= 256 * k          # the estimate path is taken
    # the loop nest of lds.ml:1071-1105 on the host: frame = rot_i + n_rot * (i + j*x_dim + k*xy_dim)
    ld = got["lattice_dims"]
    lo = [roi[d] - roi[3] for d in range(3)]
    R, T, F = [], [], []
    for kk in range(ld[2]):
        for j in range(ld[1]):
            for i in range(ld[0]):
                pos = (lo[0] + orc.grid_node(1.0, ld[0], i), lo[1] + orc.grid_node(1.0, ld[1], j),
                       lo[2] + orc.grid_node(1.0, ld[2], kk))
                if sum((roi[d] - pos[d]) ** 2 for d in range(3)) < roi[3] ** 2:
                    pt = i + j * ld[0] + kk * ld[0] * ld[1]
                    R.append(rot); T.append(np.tile(pos, (n_rot, 1))); F.append(np.arange(n_rot) + n_rot * pt)
    R, T, F = np.concatenate(R), np.concatenate(T), np.concatenate(F)
    assert len(F) == got["n_candidates"]
    e = e_intra + gpu.Mol.score_poses(rec, lig, R, T, prec=p)
    order = np.lexsort((F, e))[:k]
    if prec == "fp64":
        assert np.array_equal(got["top_frames"], F[order])
        assert np.array_equal(got["top_scores"], e[order])
        assert got["best_frame"] == F[order[0]] and got["best_score"] == e[order[0]]
    else:
        assert tol_ok(got["top_scores"], e[order]).all()
        by_frame = dict(zip(F.tolist(), e.tolist()))
        for f, s_ in zip(got["top_frames"], got["top_scores"]):      # every reported pose carries its own score
            assert tol_ok([s_], [by_frame[int(f)]]).all()
        assert tol_ok([got["best_score"]], [e[order[0]]]).all()


def test_lds_ext_c_program_matches_the_python_path(gpu, orc, c2, c2_roi_rec, setup):
    """tools/lds_ext.c: `lds --ext` as a plain C program on the C ABI (no Python in that process), fed with the
    committed .pqrs / .bild files; same lattice, same survivors, same top-k as the ctypes path on the same inputs"""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(gpu.LIB_PATH), "lds_ext")
    assert os.path.exists(exe), "build it with `make -C mmo_b200/csrc lds_ext` (done by __graft_entry__.build())"
    G = workloads.GOLDEN
    for extra in ([], ["--no-prefilter"], ["--no-prefilter", "--fp64"]):
        out = subprocess.run([exe, "-lig", os.path.join(G, "docked.pqrs"), "-rec", os.path.join(G, "xtal_rec.pqrs"),
                              "-roi", os.path.join(G, "ROI.bild"), "--ext", "2.0,96", "-top", "12"] + extra,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.strip().split("\n")
        rows = [l.split("\t") for l in lines if l and l[0].isdigit()]
        best = [l for l in lines if l.startswith("best")][0].split("\t")
        rec, lig, mask, dims, e_intra = setup
        rot = gpu.SO3.rotations(96)
        want = gpu.Lds.exhaustive_rigid_ligand_docking(
            12, c2["roi"], 2.0, rot, lig, rec=rec, vdw_mask=None if "--no-prefilter" in extra else mask,
            e_intra_const=e_intra, prec=gpu.PREC_FP64 if "--fp64" in extra else gpu.PREC_FP32)
        assert f"candidates {want['n_candidates']}, scored {want['n_scored']}" in lines[0]
        assert [int(r[2]) for r in rows] == list(want["top_frames"])
        got_s = np.array([float(r[1]) for r in rows])
        if "--fp64" in extra:
            assert np.array_equal(got_s, want["top_scores"])
            assert int(best[2]) == want["best_frame"] and float(best[1]) == want["best_score"]
        else:
            assert tol_ok(got_s, want["top_scores"]).all()


def test_c2_full_size_scan_properties(gpu, orc, c2, setup):
    """BASELINE configs[1] at full size (100 000 rotations x 4139 in-ROI lattice points = 4.1e8 poses, ~7 s on a
    B200), checked through size-independent properties: bookkeeping, ordering, and the three scorers against each
    other -- the scan's pose kernel (fp32), the item kernel on an independent random sample of the same frames
    (fp32) and the strict fp64 kernel (bit-identical to the oracle on small cases) on the reported top-k."""
    rec, lig, mask, dims, e_intra = setup
    n_rot, k = 100_000, 1000
    rot = gpu.SO3.rotations(n_rot)
    got = gpu.Lds.exhaustive_rigid_ligand_docking(k, c2["roi"], 1.0, rot, lig, rec=rec, e_intra_const=e_intra)
    ld = got["lattice_dims"]
    assert tuple(ld) == (21, 21, 22)                                   # (c+r)-(c-r) is not exactly 2r: SURVEY 8a21
    lo0 = [c2["roi"][d] - c2["roi"][3] for d in range(3)]
    gi, gj, gk = np.meshgrid(np.arange(ld[0]), np.arange(ld[1]), np.arange(ld[2]), indexing="ij")
    n_in = int((((lo0[0] + gi * 1.0 - c2["roi"][0]) ** 2 + (lo0[1] + gj * 1.0 - c2["roi"][1]) ** 2 +
                 (lo0[2] + gk * 1.0 - c2["roi"][2]) ** 2) < c2["roi"][3] ** 2).sum())
    assert 4100 < n_in < 4200                                          # ~ (4/3) pi 10^3
    assert got["n_candidates"] == got["n_scored"] == n_in * n_rot
    s, f = got["top_scores"], got["top_frames"]
    assert len(s) == k and np.all(np.diff(s) >= 0) and len(set(f.tolist())) == k
    assert got["best_frame"] == f[0] and got["best_score"] == s[0]

    def poses_of(frames):
        pt = frames // n_rot
        ri = frames - pt * n_rot
        kk = pt // (ld[0] * ld[1]); j = (pt - kk * ld[0] * ld[1]) // ld[0]; i = pt - kk * ld[0] * ld[1] - j * ld[0]
        lo = [c2["roi"][d] - c2["roi"][3] for d in range(3)]
        T = np.stack([lo[0] + i * 1.0, lo[1] + j * 1.0, lo[2] + kk * 1.0], axis=1)
        inside = ((T - np.asarray(c2["roi"][:3])) ** 2).sum(1) < c2["roi"][3] ** 2
        return rot[ri], T, inside

    R, T, inside = poses_of(f)
    assert inside.all()
    strict = e_intra + gpu.Mol.score_poses(rec, lig, R, T, prec=gpu.PREC_FP64)
    assert tol_ok(s, strict).all()                                     # every reported pose carries its true energy
    # an independent sample of the loop nest through the other fp32 kernel: nothing better than the k-th was missed
    rng = np.random.default_rng(99)
    samp = np.unique(rng.integers(0, n_rot * ld[0] * ld[1] * ld[2], 400_000))
    Rs, Ts, ins = poses_of(samp)
    samp, Rs, Ts = samp[ins], Rs[ins], Ts[ins]
    es = e_intra + gpu.Mol.score_poses(rec, lig, Rs, Ts, prec=gpu.PREC_FP32)
    better = samp[es < s[-1] - 2e-4]
    assert set(better.tolist()) <= set(f.tolist())
    assert es.min() >= s[0] - 2e-4


def test_topk_select_dev_matches_a_host_sort(gpu):
    """mmo_topk_select_dev (a conformer screen's top-k): ascending, ties to the smaller id, NaN never selected"""
    import ctypes as C
    L = gpu.lib()
    rng = np.random.default_rng(5)
    for n, k in ((1, 3), (200, 50), (100_000, 100), (300_001, 1000)):
        e = rng.normal(size=n).round(2)                 # many exact ties
        e[rng.integers(0, n, max(1, n // 50))] = np.nan
        if n > 10:
            e[7] = e[3] = np.nanmin(e) - 1.0            # a tie for the best: the smaller id must come first
        d = C.c_void_p()
        assert L.mmo_dev_alloc(C.c_size_t(n * 8), C.byref(d)) == 0
        assert L.mmo_h2d(d, e.ctypes.data_as(C.c_void_p), C.c_size_t(n * 8)) == 0
        s, f, m = np.empty(k), np.empty(k, np.int64), C.c_int32()
        rc = L.mmo_topk_select_dev(d, C.c_int64(n), C.c_int32(k), C.c_int64(1000), s.ctypes.data_as(C.POINTER(C.c_double)),
                                   f.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(m))
        assert rc == 0, L.mmo_last_error()
        L.mmo_dev_free(d)
        ok = ~np.isnan(e)
        ids = np.arange(n)[ok]
        order = np.lexsort((ids, e[ok]))[:k]
        assert m.value == len(order)
        assert np.array_equal(s[:m.value], e[ok][order]) and np.array_equal(f[:m.value], ids[order] + 1000)


def test_rot_cache_mode_and_gather_microbenchmark(gpu, c2, c2_roi_rec):
    import ctypes as C
    L = gpu.lib()
    rec = gpu.Receptor.from_mol(c2_roi_rec)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    rot = gpu.SO3.rotations(64)
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 2.0)
    a = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, rec=rec)
    assert L.mmo_scan_set_rot_cache(1) == 0
    n0, n1 = C.c_int64(), C.c_int64()
    assert L.mmo_scan_rot_rescans(C.byref(n0)) == 0
    b = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, rec=rec)
    b2 = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, np.array(rot, copy=True), lig, rec=rec)     # same bytes, another buffer
    assert L.mmo_scan_rot_rescans(C.byref(n1)) == 0
    assert n1.value == n0.value          # the same set again: checked behind the kernels, never scanned twice
    assert np.array_equal(b2["top_scores"], b["top_scores"]) and np.array_equal(b2["top_frames"], b["top_frames"])
    assert L.mmo_scan_set_rot_cache(7) != 0 and L.mmo_scan_set_rot_cache(0) == 0
    c = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot, lig, rec=rec)
    assert L.mmo_scan_set_rot_cache(1) == 0
    assert np.array_equal(a["top_scores"], c["top_scores"])
    assert np.array_equal(a["top_scores"], b["top_scores"]) and np.array_equal(a["top_frames"], b["top_frames"])
    # another set of the same size: the device-side comparison must notice and rebuild the visiting order
    rot2 = np.ascontiguousarray(rot[::-1])
    assert L.mmo_scan_rot_rescans(C.byref(n0)) == 0
    d1 = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot2, lig, rec=rec)
    assert L.mmo_scan_rot_rescans(C.byref(n1)) == 0
    assert n1.value == n0.value + 1      # scanned on the resident set, found to be another one, scanned again
    assert L.mmo_scan_set_rot_cache(0) == 0
    d0 = gpu.Lds.exhaustive_rigid_ligand_docking(10, roi, 1.0, rot2, lig, rec=rec)
    assert L.mmo_scan_set_rot_cache(1) == 0
    assert np.array_equal(d1["top_scores"], d0["top_scores"]) and np.array_equal(d1["top_frames"], d0["top_frames"])
    n_rot = len(rot)
    assert tol_ok(np.sort(d1["top_scores"]), np.sort(a["top_scores"])).all()              # the same poses under other frame ids
    fb = int(d1["best_frame"])
    assert (fb // n_rot, n_rot - 1 - fb % n_rot) == (int(a["best_frame"]) // n_rot, int(a["best_frame"]) % n_rot)
    lps = C.c_double()
    assert L.mmo_measure_l2_gather((C.c_int32 * 3)(81, 81, 81), C.c_int32(22), C.byref(lps)) == 0
    assert lps.value > 1e9
    lz = C.c_double()
    assert L.mmo_measure_l2_gather_zpair((C.c_int32 * 3)(81, 81, 81), C.c_int32(22), C.byref(lz)) == 0
    assert lz.value > lps.value          # four 8-byte reads in two rows against eight 4-byte reads in four


def test_scan_against_a_receptor_larger_than_one_tile(gpu, orc, c2):
    """a ROI receptor of more than 2048 atoms (a protonated pocket): the pose kernel reads the groups through L1 in one
    launch (untiled variant) instead of one launch per shared-memory tile; the fp64 scan (fp32 sweep + strict re-scoring)
    must still be the oracle's scan, and the fp32 scores must honour the contract"""
    from mmo_b200 import workloads
    rec_m = workloads.synthetic_receptor(2900, "sphere", 22.0, seed=77, origin=tuple(c2["roi"][:3]))
    # carve a pocket so that some poses do not clash
    d = np.sqrt((rec_m.xs - c2["roi"][0]) ** 2 + (rec_m.ys - c2["roi"][1]) ** 2 + (rec_m.zs - c2["roi"][2]) ** 2)
    keep = d > 7.0
    from mmo_b200 import pqrs
    rec_m = pqrs.Mol("pocket", rec_m.xs[keep], rec_m.ys[keep], rec_m.zs[keep], rec_m.q[keep], rec_m.r[keep], rec_m.anum[keep])
    assert rec_m.n > 2048 + 200
    rec = gpu.Receptor.from_mol(rec_m)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    rot = gpu.SO3.rotations(24)
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 2.5)
    want = orc.scan(rec_m, c2["lig"], lig.xs, lig.ys, lig.zs, roi, 1.0, rot, 20)
    got = gpu.Lds.exhaustive_rigid_ligand_docking(20, roi, 1.0, rot, lig, rec=rec, prec=gpu.PREC_FP64)
    assert got["best_frame"] == want["best_frame"] and np.array_equal(got["top_frames"], want["top_frames"])
    assert np.array_equal(got["top_scores"], want["top_scores"])
    g32 = gpu.Lds.exhaustive_rigid_ligand_docking(20, roi, 1.0, rot, lig, rec=rec, prec=gpu.PREC_FP32)
    assert tol_ok(g32["top_scores"], want["top_scores"]).all()
