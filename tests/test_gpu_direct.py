"""K1/K2 parity on the GPU through the C ABI: CUDA direct pair path vs the CPU oracle.

fp64 mode: bit-identical (same operations, same order, no FMA).  fp32 mode: within the north-star
tolerance max(1e-6*|E|, 1e-4 kcal/mol) -- including clashing poses, which the close-contact fp64
correction pass exists for."""
import os

import numpy as np
import pytest

from conftest import tol_ok
from mmo_b200 import pqrs, workloads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _poses(c2, n, seed, radius=8.0):
    R, t = workloads.random_poses_in_sphere(n, c2["roi"][:3], radius, seed=seed)
    R[0] = np.eye(3).reshape(9)
    t[0] = c2["start_pos"]          # the docked pose itself
    return R, t


@pytest.fixture(scope="module")
def handles(gpu, c2, c2_roi_rec):
    rec = gpu.Receptor.from_mol(c2_roi_rec)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    return rec, lig


def test_fp64_bit_identical_shifted_and_global(gpu, orc, c2, c2_roi_rec, handles):
    rec, lig = handles
    m = c2["lig"]
    R, t = _poses(c2, 40, seed=11)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    for shifted, variant in ((True, gpu.VARIANT_SHIFTED), (False, gpu.VARIANT_GLOBAL)):
        want = orc.ene_inter(c2_roi_rec, m.q, m.anum, X, Y, Z, shifted=shifted)
        got_c = gpu.Mol._score(rec, lig, variant, gpu.PREC_FP64, X, Y, Z)
        got_p = gpu.Mol.score_poses(rec, lig, R, t, variant=variant, prec=gpu.PREC_FP64)
        assert np.array_equal(got_c, want)
        assert np.array_equal(got_p, want)      # on-device pose transform == Rot.rotate + translate_by


@pytest.fixture(params=[1, 2], ids=["pose_kernel", "item_kernel"])
def direct_mode(gpu, request):
    """both FP32 kernels must honour the contract on every kind of pose list"""
    assert gpu.lib().mmo_direct_set_mode(request.param) == 0
    yield request.param
    gpu.lib().mmo_direct_set_mode(0)


def test_fp32_within_tolerance_including_clashes(gpu, orc, c2, c2_roi_rec, handles, direct_mode):
    rec, lig = handles
    m = c2["lig"]
    R, t = _poses(c2, 600, seed=12)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    for shifted, variant in ((True, gpu.VARIANT_SHIFTED), (False, gpu.VARIANT_GLOBAL)):
        want = orc.ene_inter(c2_roi_rec, m.q, m.anum, X, Y, Z, shifted=shifted)
        got = gpu.Mol.score_poses(rec, lig, R, t, variant=variant, prec=gpu.PREC_FP32)
        ok = tol_ok(got, want)
        ratio = np.abs(got - want) / np.maximum(1e-6 * np.abs(want), 1e-4)
        print(f"random poses shifted={shifted}: worst |err|/tol = {ratio.max():.3f}, median {np.median(ratio):.4f}")
        assert ok.all(), f"worst: {np.abs(got - want)[~ok].max()} at E={want[~ok]}"
        got_c = gpu.Mol._score(rec, lig, variant, gpu.PREC_FP32, X, Y, Z)
        assert tol_ok(got_c, want).all()
    assert (want > 1e3).sum() > 50 and (want < 0).sum() > 0      # the sample holds clashes and good poses


def _coherent_poses(c2, n, seed, dtheta):
    """n poses per warp-sized family: one base rotation composed with rotations by < dtheta rad, shared
    translation -- what a warp of the scan driver sees (k-d ordered SO(3) sample on one lattice point)"""
    rng = np.random.default_rng(seed)
    base = workloads.random_rotations(n // 64 + 1, rng)
    R = np.empty((n, 9))
    t = np.empty((n, 3))
    for i in range(n):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        th = rng.uniform(-dtheta, dtheta)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        dR = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        R[i] = (dR @ base[i // 64].reshape(3, 3)).reshape(9)
        fam = np.random.default_rng(seed * 1000 + i // 64)
        v = fam.normal(size=3); v *= fam.uniform(0, 7.0) / np.linalg.norm(v)
        t[i] = np.asarray(c2["roi"][:3]) + v
    return R, t


@pytest.mark.parametrize("dtheta", [0.02, 0.15, 0.6])
def test_fp32_coherent_warps_within_tolerance(gpu, orc, c2, c2_roi_rec, handles, dtheta, direct_mode):
    """warps of similar poses take the cloud-centred expanded form of r^2 (small dtheta) or the
    difference form (large dtheta): both must honour the accuracy contract"""
    rec, lig = handles
    m = c2["lig"]
    R, t = _coherent_poses(c2, 64 * 12 + 7, seed=21, dtheta=dtheta)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    for shifted, variant in ((True, gpu.VARIANT_SHIFTED), (False, gpu.VARIANT_GLOBAL)):
        want = orc.ene_inter(c2_roi_rec, m.q, m.anum, X, Y, Z, shifted=shifted)
        got = gpu.Mol.score_poses(rec, lig, R, t, variant=variant, prec=gpu.PREC_FP32)
        ratio = np.abs(got - want) / np.maximum(1e-6 * np.abs(want), 1e-4)
        print(f"dtheta={dtheta} shifted={shifted}: worst |err|/tol = {ratio.max():.3f}, median {np.median(ratio):.4f}")
        assert (ratio <= 1.0).all(), f"worst ratio {ratio.max()} at E={want[ratio.argmax()]}"


@pytest.mark.parametrize("n", [1, 2, 63, 65, 700, 4096, 4097])
def test_fp32_small_batches_receptor_slices_and_block_fix(gpu, orc, c2, c2_roi_rec, handles, n):
    """single-pose calls (the reference's closure shape) and other small batches split the receptor into slices and run
    the fp64 pass with block = pose: same contract on both sides of the two thresholds, coordinates and rot/trans input,
    and the result of a pose does not depend on where in the batch it sits (the parts are added in a fixed order)"""
    rec, lig = handles
    m = c2["lig"]
    R, t = _poses(c2, n, seed=100 + n)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    for shifted, variant in ((True, gpu.VARIANT_SHIFTED), (False, gpu.VARIANT_GLOBAL)):
        want = orc.ene_inter(c2_roi_rec, m.q, m.anum, X, Y, Z, shifted=shifted)
        got = gpu.Mol.score_poses(rec, lig, R, t, variant=variant, prec=gpu.PREC_FP32)
        ratio = np.abs(got - want) / np.maximum(1e-6 * np.abs(want), 1e-4)
        assert (ratio <= 1.0).all(), f"n={n} shifted={shifted}: worst ratio {ratio.max()} at E={want[ratio.argmax()]}"
        got_c = gpu.Mol._score(rec, lig, variant, gpu.PREC_FP32, X, Y, Z)
        assert tol_ok(got_c, want).all()
        again = gpu.Mol.score_poses(rec, lig, R, t, variant=variant, prec=gpu.PREC_FP32)
        assert np.array_equal(got, again)                       # deterministic
    one = np.atleast_1d(gpu.Mol.ene_inter_UFF_shifted_brute(rec, lig, X[0], Y[0], Z[0]))
    assert tol_ok(one, orc.ene_inter(c2_roi_rec, m.q, m.anum, X[:1], Y[:1], Z[:1], shifted=True)).all()


def test_fp32_far_away_pose_is_exactly_zero(gpu, c2, handles, direct_mode):
    rec, lig = handles
    t = np.array([[c2["roi"][0] + 80.0, c2["roi"][1], c2["roi"][2]]])
    e = gpu.Mol.score_poses(rec, lig, np.eye(3).reshape(1, 9), t)
    assert e[0] == 0.0          # lds.ml:920: E_inter = 0.0 exactly is a reset signal


def test_empty_batch_and_ragged_sizes(gpu, orc, direct_mode):
    rng = np.random.default_rng(5)
    # receptor size not a multiple of the blob, ligand size not a multiple of the register chunk
    for P, L in ((1, 1), (17, 3), (33, 9), (250, 13)):
        rec_m = workloads.synthetic_receptor(P, "cube", 12.0, seed=P)
        lx, ly, lz = rng.uniform(2, 10, (3, L))
        lq = rng.uniform(-0.5, 0.5, L)
        la = rng.choice([1, 6, 7, 8, 16, 17], L).astype(np.int32)
        rec = gpu.Receptor.from_mol(rec_m)
        lig = gpu.Ligand(lx, ly, lz, lq, la)
        X = lx[None, :] + rng.uniform(-1, 1, (5, 1))
        Y, Z = np.tile(ly, (5, 1)), np.tile(lz, (5, 1))
        want = orc.ene_inter(rec_m, lq, la, X, Y, Z, shifted=True)
        assert np.array_equal(gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, gpu.PREC_FP64, X, Y, Z), want)
        assert tol_ok(gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, gpu.PREC_FP32, X, Y, Z), want).all()
        assert gpu.Mol._score(rec, lig, 1, 0, np.empty((0, L)), np.empty((0, L)), np.empty((0, L))).shape == (0,)


def test_empty_receptor_gives_zero(gpu):
    rec = gpu.Receptor(np.empty(0), np.empty(0), np.empty(0), np.empty(0), np.empty(0, np.int32))
    lig = gpu.Ligand([0.0, 1.0], [0.0, 0.0], [0.0, 0.0], [0.1, -0.1], [6, 8])
    for prec in (gpu.PREC_FP32, gpu.PREC_FP64):
        e = gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, prec, [[0.0, 1.0]], [[0.0, 0.0]], [[0.0, 0.0]])
        assert e[0] == 0.0


def test_coincident_atoms_use_the_001_clamp(gpu, orc, direct_mode):
    # Math.non_zero_dist (math.ml:58-62): r < 0.01 -> 0.01, energies of order 1e30
    rec_m = pqrs.Mol("one", np.array([5.0]), np.array([5.0]), np.array([5.0]), np.array([0.5]), np.array([1.7]),
                     np.array([6], np.int32))
    rec = gpu.Receptor.from_mol(rec_m)
    lig = gpu.Ligand([0.0], [0.0], [0.0], [0.5], [6])
    for dx in (0.0, 0.005, 0.0101):
        X, Y, Z = [[5.0 + dx]], [[5.0]], [[5.0]]
        want = orc.ene_inter(rec_m, [0.5], [6], X, Y, Z, shifted=True)
        assert np.array_equal(gpu.Mol._score(rec, lig, 1, gpu.PREC_FP64, X, Y, Z), want)
        assert tol_ok(gpu.Mol._score(rec, lig, 1, gpu.PREC_FP32, X, Y, Z), want).all()
    assert want[0] > 1e29


def test_unsupported_element_gives_nan_like_the_reference(gpu):
    rec = gpu.Receptor([0.0], [0.0], [0.0], [0.1], [30])          # Zn is not in UFF.ml:10-22
    lig = gpu.Ligand([0.0], [0.0], [0.0], [0.1], [6])
    for prec in (gpu.PREC_FP32, gpu.PREC_FP64):
        assert np.isnan(gpu.Mol._score(rec, lig, 1, prec, [[3.0]], [[0.0]], [[0.0]])[0])


def test_components_and_intra_bit_identical(gpu, orc, c2, c2_roi_rec, handles):
    rec, lig = handles
    m = c2["lig"]
    R, t = _poses(c2, 12, seed=13)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    e, v = gpu.Mol.ene_inter_UFF_shifted_bst_components(rec, lig, X, Y, Z)
    we, wv = orc.ene_inter_components(c2_roi_rec, m.q, m.anum, X, Y, Z)
    assert np.array_equal(e, we) and np.array_equal(v, wv)
    # intra: the docked conformer, rigid copies of it and perturbed conformers
    rng = np.random.default_rng(14)
    Xc = np.concatenate([X, X[:4] + rng.normal(0, 0.3, (4, lig.n))])
    Yc = np.concatenate([Y, Y[:4] + rng.normal(0, 0.3, (4, lig.n))])
    Zc = np.concatenate([Z, Z[:4] + rng.normal(0, 0.3, (4, lig.n))])
    assert np.array_equal(gpu.Mol.ene_intra_UFFNB_brute(lig, Xc, Yc, Zc), orc.ene_intra(m, Xc, Yc, Zc))


def test_linearity_in_receptor_charges(gpu, c2, c2_roi_rec):
    """size-independent property: E(q_rec) is affine in the receptor charges (Coulomb term)"""
    m = c2_roi_rec
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    R, t = _poses(c2, 64, seed=15, radius=3.0)
    e = []
    for s in (0.0, 1.0, 2.0):
        rec = gpu.Receptor(m.xs, m.ys, m.zs, m.q * s, m.anum)
        e.append(gpu.Mol.score_poses(rec, lig, R, t, prec=gpu.PREC_FP64))
    assert np.allclose(e[2] - e[1], e[1] - e[0], rtol=1e-9, atol=1e-7 * np.abs(e[1]).max())


def test_pose_tools_bit_identical(gpu, orc, c2):
    """place_ligand (Optim.apply_config, optim.ml:64-80) and lig_rot_sample (lig_rot_sample.ml:23-45)"""
    m = c2["lig"]
    lig = gpu.Ligand.from_mol(m, centered=True)
    rng = np.random.default_rng(77)
    for with_bonds in (False, True):
        cfg = list(rng.uniform(30, 60, 3)) + list(rng.uniform(-3, 3, 3))
        if with_bonds:
            cfg += list(rng.uniform(-3.1, 3.1, m.n_rbonds))
        x, y, z, tl = gpu.Optim.apply_config(lig, cfg)
        wx, wy, wz, wtl = orc.apply_config(m, lig.xs, lig.ys, lig.zs, cfg)
        assert np.array_equal(x, wx) and np.array_equal(y, wy) and np.array_equal(z, wz) and tl == wtl
    # rotated copies about the original centre of the (uncentred) ligand
    raw = gpu.Ligand.from_mol(m, centered=False)
    center = [orc.favg(m.xs), orc.favg(m.ys), orc.favg(m.zs)]
    rot = gpu.SO3.rotations(25)
    X, Y, Z = gpu.Optim.rotated_copies(raw, center, rot)
    for r in (0, 7, 24):
        ox = np.empty(m.n); oy = np.empty(m.n); oz = np.empty(m.n)
        import ctypes as C
        dp = C.POINTER(C.c_double)
        orc.lib().orc_center_rotate_translate(C.c_int(m.n), orc.d(m.xs)[1], orc.d(m.ys)[1], orc.d(m.zs)[1],
                                              orc.d(center)[1], orc.d(rot[r])[1], orc.d(center)[1],
                                              ox.ctypes.data_as(dp), oy.ctypes.data_as(dp), oz.ctypes.data_as(dp))
        assert np.array_equal(X[r], ox) and np.array_equal(Y[r], oy) and np.array_equal(Z[r], oz)
    # the rotated copies keep every inter-atomic distance (rigid motion)
    d0 = np.hypot(np.hypot(m.xs[0] - m.xs[5], m.ys[0] - m.ys[5]), m.zs[0] - m.zs[5])
    d1 = np.hypot(np.hypot(X[:, 0] - X[:, 5], Y[:, 0] - Y[:, 5]), Z[:, 0] - Z[:, 5])
    assert np.allclose(d1, d0, rtol=1e-13)


def test_fp32_conformer_screen_multi_tile(gpu, orc, direct_mode):
    """C5 shape: explicit conformers of a 70-atom ligand (n_fast = 72: padded chunk) against a receptor of
    more than one shared-memory tile (3000 atoms > 2048); poses scattered over the whole receptor."""
    rec_m = workloads.synthetic_receptor(3000, "sphere", 24.0, seed=3, origin=(40.0, 40.0, 40.0))
    lig_m = workloads.c5_ligand()
    rec = gpu.Receptor.from_mol(rec_m)
    lig = gpu.Ligand.from_mol(lig_m, centered=False)
    X, Y, Z = workloads.c5_conformers(lig_m, 520, (40.0, 40.0, 40.0), radius=36.0, seed=9)
    for shifted, variant in ((True, gpu.VARIANT_SHIFTED), (False, gpu.VARIANT_GLOBAL)):
        want = orc.ene_inter(rec_m, lig_m.q, lig_m.anum, X, Y, Z, shifted=shifted)
        got = gpu.Mol._score(rec, lig, variant, gpu.PREC_FP32, X, Y, Z)
        ratio = np.abs(got - want) / np.maximum(1e-6 * np.abs(want), 1e-4)
        print(f"conformers shifted={shifted}: worst |err|/tol = {ratio.max():.3f}")
        assert (ratio <= 1.0).all(), f"worst ratio {ratio.max()} at E={want[ratio.argmax()]}"
        assert np.array_equal(gpu.Mol._score(rec, lig, variant, gpu.PREC_FP64, X[:40], Y[:40], Z[:40]), want[:40])
    assert (np.abs(want) < 50.0).sum() > 20 and (want > 1e3).sum() > 20      # surface poses and clashes alike


def test_shutdown_and_reinit_rebuild_every_cache(gpu, orc, c2, c2_roi_rec):
    """mmo_shutdown drops the library-lifetime device buffers (tables, work counters, scratch, rotation set);
    a new mmo_init starts from nothing and gives the same numbers.  Runs last in this file on purpose: handles
    created before the shutdown belong to the old pool and are not used afterwards."""
    m = c2["lig"]
    R, t = _poses(c2, 700, seed=31)

    def run():
        rec = gpu.Receptor.from_mol(c2_roi_rec)
        lig = gpu.Ligand.from_mol(m, centered=True)
        e32 = gpu.Mol.score_poses(rec, lig, R, t, prec=gpu.PREC_FP32)          # item kernel (33.6 k items)
        gpu.lib().mmo_direct_set_mode(1)
        e32p = gpu.Mol.score_poses(rec, lig, R[:64], t[:64], prec=gpu.PREC_FP32)  # pose kernel
        gpu.lib().mmo_direct_set_mode(0)
        rot = gpu.SO3.rotations(64)
        sc = gpu.Lds.exhaustive_rigid_ligand_docking(5, (c2["roi"][0], c2["roi"][1], c2["roi"][2], 2.5), 2.0, rot, lig, rec=rec)
        del rec, lig
        return e32, e32p, sc["top_scores"], sc["top_frames"]

    a = run()
    assert gpu.lib().mmo_shutdown() == 0
    gpu.init(0)
    b = run()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_c5_large_conformer_screen_properties(gpu, orc):
    """BASELINE configs[4] at one GPU's share of the 2-GPU run (500 000 explicit conformers of the 70-atom ligand
    against the 10 000-atom receptor sphere), through size-independent properties: shard + top-k merge == whole,
    order independence within the contract, far conformers exactly 0, and a sample against the strict fp64 kernel
    and the oracle."""
    from mmo_b200 import sharding
    rec_m = workloads.synthetic_receptor(10000, "sphere", 34.0, seed=workloads.SEED + 1, origin=(60.0, 60.0, 60.0))
    lig_m = workloads.c5_ligand()
    rec = gpu.Receptor.from_mol(rec_m)
    lig = gpu.Ligand.from_mol(lig_m, centered=False)
    n, k = 500_000, 100
    X, Y, Z = workloads.c5_conformers(lig_m, n, (60.0, 60.0, 60.0), radius=44.0, seed=77)    # pocket, surface and solvent
    X[-1000:] += 200.0                                            # the last thousand: far beyond every cut-off
    e = gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, gpu.PREC_FP32, X, Y, Z)
    assert (e[-1000:] == 0.0).all() and np.isfinite(e).all()
    # (1) two shards + merge == top-k of the whole list (ties to the smaller conformer id)
    ids = np.arange(n, dtype=np.int64)
    tops = []
    for r in range(2):
        a, c = sharding.shard_range(n, r, 2)
        er = gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, gpu.PREC_FP32, X[a:a + c], Y[a:a + c], Z[a:a + c])
        assert tol_ok(er, e[a:a + c]).all()                       # same conformers, other warp companions
        o = np.lexsort((ids[a:a + c], er))[:k]
        tops.append((er[o], ids[a:a + c][o]))
    ms, mf = sharding.merge_lists(gpu.lib(), k, [t[0] for t in tops], [t[1] for t in tops])
    whole = np.lexsort((ids, e))[:k]
    assert tol_ok(ms, e[whole]).all()
    assert len(set(mf.tolist()) ^ set(ids[whole].tolist())) <= 4   # only near-ties at the k-th place may differ
    # (2) order independence: a shuffled subset gets the same energies within the contract
    rng = np.random.default_rng(8)
    sub = rng.choice(n - 1000, 60_000, replace=False)
    es = gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, gpu.PREC_FP32, X[sub], Y[sub], Z[sub])
    assert tol_ok(es, e[sub]).all()
    # (3) the best conformers and a random sample against the strict kernel; part of the sample against the oracle
    chk = np.concatenate([whole[:50], sub[:150]])
    strict = gpu.Mol._score(rec, lig, gpu.VARIANT_SHIFTED, gpu.PREC_FP64, X[chk], Y[chk], Z[chk])
    assert tol_ok(e[chk], strict).all()
    assert np.array_equal(strict[:40], orc.ene_inter(rec_m, lig_m.q, lig_m.anum, X[chk[:40]], Y[chk[:40]], Z[chk[:40]], shifted=True))


def test_shared_reciprocal_division_is_the_ieee_division(gpu):
    """strict kernels: a / r for many a through RN(1/r) and two fused corrections must equal a / r in every bit"""
    import ctypes as C
    bad = C.c_int64(-1)
    for seed in (1, 20231017):
        assert gpu.lib().mmo_selftest_division(C.c_uint64(seed), C.c_int64(1 << 28), C.byref(bad)) == 0
        assert bad.value == 0


def test_strict_tables_survive_a_reinit(tmp_path):
    """mmo_shutdown + mmo_init start a new epoch: the __constant__ UFF tables of the strict kernels and the pooled
    blocks belong to the old one.  Run in a subprocess so that the session's handles stay valid."""
    import subprocess
    import sys
    code = r'''
import sys, os
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, mmo_b200, oracle
from mmo_b200 import workloads
c2 = workloads.load_c2("docked")
rec_m = workloads.carve(c2["rec"], c2["roi"][:3], 25.0)
R, t = workloads.random_poses_in_sphere(64, c2["roi"][:3], 5.0, seed=3)
def run():
    rec = mmo_b200.Receptor.from_mol(rec_m); lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
    e = mmo_b200.Mol.score_poses(rec, lig, R, t, prec=mmo_b200.PREC_FP64)
    X, Y, Z = oracle.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    want = oracle.ene_inter(rec_m, c2["lig"].q, c2["lig"].anum, X, Y, Z, shifted=True)
    assert np.array_equal(e, want), "strict energies differ from the oracle"
    i = mmo_b200.Mol.ene_intra_UFFNB_brute(lig, X[:4], Y[:4], Z[:4])
    assert np.array_equal(i, oracle.ene_intra(c2["lig"], X[:4], Y[:4], Z[:4]))
    del rec, lig
mmo_b200.init(0); run()
assert mmo_b200.lib().mmo_shutdown() == 0
mmo_b200.init(0); run()
print("reinit ok")
'''
    script = tmp_path / "reinit.py"
    script.write_text(code)
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "reinit ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_strict_kernels_block_per_pose_and_thread_per_pose_agree(gpu, orc, c2):
    """MMO_PREC_FP64 has two kernels behind it: block per pose (up to 32 k poses: single evaluations, the re-scoring stage
    of the fp64 scan) and thread per pose (beyond).  Same doubles in the same order: both bit-identical to the oracle,
    hence to each other; likewise the two intra-ligand kernels (block per conformer for a few, thread per conformer)."""
    rec_m = workloads.carve(c2["rec"], c2["roi"][:3], 9.0)
    assert 100 < rec_m.n < 400
    rec = gpu.Receptor.from_mol(rec_m)
    m = c2["lig"]
    lig = gpu.Ligand.from_mol(m, centered=True)
    n = 33000
    R, t = workloads.random_poses_in_sphere(n, c2["roi"][:3], 7.0, seed=71)
    for variant, shifted in ((gpu.VARIANT_SHIFTED, True), (gpu.VARIANT_GLOBAL, False)):
        big = gpu.Mol.score_poses(rec, lig, R, t, variant=variant, prec=gpu.PREC_FP64)              # thread per pose
        small = gpu.Mol.score_poses(rec, lig, R[:700], t[:700], variant=variant, prec=gpu.PREC_FP64)  # block per pose
        one = gpu.Mol.score_poses(rec, lig, R[5:6], t[5:6], variant=variant, prec=gpu.PREC_FP64)
        assert np.array_equal(big[:700], small) and one[0] == big[5]
        X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R[:3000], t[:3000])
        assert np.array_equal(big[:3000], orc.ene_inter(rec_m, m.q, m.anum, X, Y, Z, shifted=shifted))
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R[:1200], t[:1200])
    few = gpu.Mol.ene_intra_UFFNB_brute(lig, X[:7], Y[:7], Z[:7])               # block per conformer
    many = gpu.Mol.ene_intra_UFFNB_brute(lig, X, Y, Z)                          # thread per conformer
    assert np.array_equal(few, many[:7]) and np.array_equal(many, orc.ene_intra(m, X, Y, Z))
