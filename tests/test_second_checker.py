"""The C oracle (oracle/mmo_oracle*.c) against a second restatement written independently from the OCaml text in pure
Python (oracle/pycheck/mmo_ref.py): whole-pose energies on C2- and C5-shaped inputs, one map voxel and one trilinear
cell on a C3-shaped input, intra-ligand energies, 200 Monte-Carlo frames, SO3.rotations, the vdW / solvent-shell bitmasks
and the clash test -- all BIT FOR BIT.  Plus a 40-digit mpmath
evaluation of the same whole-pose sums, which bounds the rounding error of either.  CPU only.

VERDICT r1 'weak #1': the oracle and the kernels came from one reading of the source; this removes the transcription
risk (two languages, two data layouts).  A shared misreading of the OCaml text would survive: DESIGN.md section 2 lists
the readings that stay single-source."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pycheck"))
import mmo_ref as ref  # noqa: E402

from mmo_b200 import pqrs, workloads  # noqa: E402


def _ref_prot(m):
    return ref.Mol(m.xs, m.ys, m.zs, m.q, m.anum)


def _ref_lig(lig, xs, ys, zs, center=None):
    groups = [list(g) for g in lig.groups] if hasattr(lig, "groups") else []
    pi = 4.0 * np.arctan(1.0)
    return ref.Mol(xs, ys, zs, lig.q, lig.anum, t_a=lig.typ, dists_a=lig.dists,
                   rbonds=list(zip(lig.rb_left, lig.rb_right)), rgroups=groups, max_rbond_rot=5.0 * (pi / 180.0), center=center)


def _groups(lig):
    off, idx = lig.rgroup_csr()
    return [list(idx[off[b]:off[b + 1]]) for b in range(lig.n_rbonds)]


@pytest.fixture(scope="module")
def c2lig(c2):
    lig = c2["lig"]
    lig.groups = _groups(lig)
    return lig


def test_whole_pose_energies_c2_shape_bit_for_bit(orc, c2, c2_roi_rec, c2lig):
    """docked.mol2 against the 1837-atom ROI receptor: docked pose, a clashing pose, a far pose; shifted and global"""
    cx, cy, cz = c2["centered"]
    R, t = workloads.random_poses_in_sphere(3, c2["roi"][:3], 6.0, seed=11)
    R[0] = np.eye(3).reshape(9); t[0] = c2["start_pos"]
    t[2] = np.array(c2["roi"][:3]) + np.array([0.0, 0.0, 27.0])
    X, Y, Z = orc.pose_coords(cx, cy, cz, R, t)
    prot = _ref_prot(c2_roi_rec)
    for shifted, fn in ((True, ref.ene_inter_UFF_shifted_brute), (False, ref.ene_inter_UFF_global_brute)):
        want = orc.ene_inter(c2_roi_rec, c2lig.q, c2lig.anum, X, Y, Z, shifted=shifted)
        for p in range(3):
            got = fn(prot, _ref_lig(c2lig, X[p], Y[p], Z[p]))
            assert got == want[p], (shifted, p, got, want[p])
    # the pose builder itself: Mol.rotate_then_translate_copy (mol.ml:664-672)
    m = ref.rotate_then_translate_copy(_ref_lig(c2lig, cx, cy, cz, center=(0, 0, 0)), tuple(R[1]), tuple(t[1]))
    assert m.xs == list(X[1]) and m.ys == list(Y[1]) and m.zs == list(Z[1])


def test_whole_pose_energy_c5_shape_bit_for_bit(orc):
    """a 70-atom conformer against a 10 000-atom synthetic receptor sphere (BASELINE configs[4])"""
    rec = workloads.synthetic_receptor(10000, "sphere", 34.0, seed=workloads.SEED + 1, origin=(60.0, 60.0, 60.0))
    lig = workloads.c5_ligand()
    X, Y, Z = workloads.c5_conformers(lig, 2, (60.0, 60.0, 60.0))
    want = orc.ene_inter(rec, lig.q, lig.anum, X, Y, Z, shifted=True)
    prot = _ref_prot(rec)
    for p in range(2):
        got = ref.ene_inter_UFF_shifted_brute(prot, ref.Mol(X[p], Y[p], Z[p], lig.q, lig.anum))
        assert got == want[p]


def test_whole_pose_sum_against_40_digit_arithmetic(orc, c2, c2_roi_rec, c2lig):
    """the same sums in 40-digit arithmetic: the double-precision result of the reference's evaluation order is within
    1e-11 relative (+ 1e-9 absolute) of the exact value of the formula -- a bound on what summation order can move"""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    cx, cy, cz = c2["centered"]
    X, Y, Z = orc.pose_coords(cx, cy, cz, np.eye(3).reshape(1, 9), np.array([c2["start_pos"]]))
    want = orc.ene_inter(c2_roi_rec, c2lig.q, c2lig.anum, X, Y, Z, shifted=True)[0]
    se = mp.mpf(0); sv = mp.mpf(0)
    P = c2_roi_rec
    for i in range(P.n):
        for j in range(c2lig.n):
            dx, dy, dz = mp.mpf(P.xs[i]) - mp.mpf(X[0][j]), mp.mpf(P.ys[i]) - mp.mpf(Y[0][j]), mp.mpf(P.zs[i]) - mp.mpf(Z[0][j])
            r2 = dx * dx + dy * dy + dz * dz
            if r2 < 144:
                r = mp.sqrt(r2)
                if r < mp.mpf("0.01"):
                    r = mp.mpf("0.01")
                w = (1 - (r / 12) ** 2) ** 2
                xi = dict(ref.ANUM_XI_DI)
                x_ij = mp.sqrt(mp.mpf(xi[int(P.anum[i])][0]) * mp.mpf(xi[int(c2lig.anum[j])][0]))
                d_ij = mp.sqrt(mp.mpf(xi[int(P.anum[i])][1]) * mp.mpf(xi[int(c2lig.anum[j])][1]))
                p6 = (x_ij / r) ** 6
                se += w * (mp.mpf(P.q[i]) * mp.mpf(c2lig.q[j]) / r)
                sv += w * d_ij * (p6 * p6 - 2 * p6)
    exact = mp.mpf(332.0637) / 4 * se + sv
    assert abs(mp.mpf(want) - exact) <= mp.mpf("1e-11") * abs(exact) + mp.mpf("1e-9")


def test_intra_energy_bit_for_bit(orc, c2, c2lig):
    cx, cy, cz = c2["centered"]
    want = orc.ene_intra(c2lig, cx, cy, cz)[0]
    assert ref.ene_intra_UFFNB_brute(_ref_lig(c2lig, cx, cy, cz)) == want
    rng = np.random.default_rng(3)
    x2, y2, z2 = cx + rng.normal(0, 0.2, len(cx)), cy + rng.normal(0, 0.2, len(cx)), cz + rng.normal(0, 0.2, len(cx))
    assert ref.ene_intra_UFFNB_brute(_ref_lig(c2lig, x2, y2, z2)) == orc.ene_intra(c2lig, x2, y2, z2)[0]


def test_map_voxels_and_trilinear_cell_c3_shape_bit_for_bit(orc):
    """C3 shape: 5000-atom receptor, 0.375 A grid, the 22 FF types of ligdecs.mol2: the 22 values of three voxels
    (Mol.ene_inter_UFF_shifted_grid + clamp + float32 store) and look-ups inside one cell of a small map set"""
    rec = workloads.synthetic_receptor(5000, "cube", 60.0, seed=workloads.SEED)
    rec.xs -= 15.0; rec.ys -= 15.0; rec.zs -= 15.0
    lig = pqrs.read_ligands_pqrs(os.path.join(workloads.GOLDEN, "ligdecs.pqrs"))[0]
    ta, tq = pqrs.assign_ff_types([lig])
    dims = orc.grid_from_box(0.375, 30.0, 30.0, 30.0)
    g = ref.Grid(0.375, 30.0, 30.0, 30.0)
    assert (g.x_dim, g.y_dim, g.z_dim) == tuple(dims) == (81, 81, 81)
    assert all(g.xs[i] == orc.grid_node(0.375, dims[0], i) for i in range(dims[0]))
    vox = [(40, 40, 40), (0, 80, 13), (17, 3, 66)]
    nvox = dims[0] * dims[1] * dims[2]
    mask = np.zeros((nvox + 7) // 8, np.uint8)
    for i, j, k in vox:
        idx = i + j * dims[0] + k * dims[0] * dims[1]
        mask[idx >> 3] |= 1 << (idx & 7)
    maps = orc.grid_build(rec, 0.375, dims, ta, tq, mask=mask)
    prot = _ref_prot(rec)
    probes = list(zip([int(a) for a in ta], [float(q) for q in tq]))
    for i, j, k in vox:
        idx = i + j * dims[0] + k * dims[0] * dims[1]
        e = ref.ene_inter_UFF_shifted_grid(prot, (g.xs[i], g.ys[j], g.zs[k]), probes)
        got = [ref.map_value(v) for v in e]
        assert got == [float(maps[t][idx]) for t in range(len(ta))]
        assert any(v != 0.0 for v in got)
    # trilinear: a small dense map set from the oracle, look-ups compared point by point
    sd = orc.grid_from_box(0.375, 3.0, 2.25, 2.625)
    small = orc.grid_build(rec, 0.375, sd, ta[:2], tq[:2])
    sg = ref.Grid(0.375, 3.0, 2.25, 2.625)
    rng = np.random.default_rng(8)
    pts = rng.uniform(0.0, 1.0, (40, 3)) * (np.array(sd) - 1.001) * 0.375
    pts[0] = (0.375 * 2, 0.375 * 3, 0.375 * 4)                   # exactly on a node
    for t in range(2):
        for p in pts:
            assert ref.trilin(sg, small[t], tuple(p)) == orc.trilin(0.375, sd, small[t], p[0], p[1], p[2])


@pytest.mark.parametrize("flags", [dict(), dict(tweak_rbonds=False), dict(hard_roi=False), dict(no_flip=True, temperature_K=450.0)])
def test_200_monte_carlo_frames_bit_for_bit(orc, c2, c2_roi_rec, c2lig, flags):
    """Lds.simulate_lig (lds.ml:741-1000) re-read from the OCaml text: molecule copies, shared acceptance windows,
    right-to-left draws, the dangling else -- 200 frames (600 with the default flags, so that adaptive step sizes and a
    bond flip are reached), every energy of every frame equal to the C oracle's"""
    c = np.array(c2["roi"][:3])
    dims = orc.grid_from_box(1.0, *(c + 23.0))
    mask = orc.bitmask_sphere(1.0, dims, c, 21.0)
    ta, tq = pqrs.assign_ff_types([c2lig])
    maps = orc.grid_build(c2_roi_rec, 1.0, dims, ta, tq, mask=mask)
    cx, cy, cz = c2["centered"]
    n_steps = 600 if not flags else 200
    rot0 = np.eye(3).reshape(9)
    want, wxyz, wtr = orc.mc_run(c2lig, cx, cy, cz, c2["roi"], n_steps, 1234, rot0, c2["start_pos"], maps=maps, g_step=1.0,
                                 g_dims=dims, **flags)
    g = ref.Grid(1.0, *(c + 23.0))
    assert (g.x_dim, g.y_dim, g.z_dim) == tuple(dims)
    comps = [maps[t] for t in range(len(ta))]
    centered = _ref_lig(c2lig, cx, cy, cz, center=(0.0, 0.0, 0.0))
    kw = dict(tweak_rbonds=flags.get("tweak_rbonds", True), enforce_ROI=flags.get("hard_roi", True),
              no_flip=flags.get("no_flip", False), temperature_K=flags.get("temperature_K", 293.15))
    trace, best_E, prev_E, cnt = ref.simulate_lig(centered, lambda m: ref.ene_inter_UFF_interp(g, comps, m), n_steps, 1234,
                                                  tuple(rot0), tuple(c2["start_pos"]), c2["roi"], **kw)
    assert cnt["frames"] == want["frames_done"] == len(trace)
    assert np.array_equal(np.array(trace), wtr)
    assert best_E == want["best_E"] and prev_E == want["prev_E"]
    assert (cnt["acc_rigid"], cnt["rej_rigid"], cnt["acc_conf"], cnt["rej_conf"], cnt["ooroi"], cnt["ezero"]) == \
        (want["n_accept_rigid"], want["n_reject_rigid"], want["n_accept_conf"], want["n_reject_conf"], want["n_ooroi"], want["n_ezero"])
    if not flags:
        assert cnt["acc_rigid"] > 0 and cnt["rej_rigid"] > 0 and cnt["acc_conf"] + cnt["rej_conf"] > 0


def test_so3_rotations_bit_for_bit(orc):
    """SO3.rotations (SO3.ml:18-39, quat.ml:33-37, rot.ml:136-146) with the platform's libm on both sides"""
    for n in (1, 7, 64):
        want = orc.so3_rotations(n)
        got = ref.so3_rotations(n)
        assert len(got) == n
        for i in range(n):
            assert tuple(want[i]) == got[i], (n, i)
    q = ref.super_fibonacci(10.0, 3)
    assert q == orc.so3_quat(10, 3)


def _bits(mask_bytes, nvox):
    """Bitv order of the C oracle's byte masks: bit idx = (byte[idx >> 3] >> (idx & 7)) & 1"""
    b = np.unpackbits(np.asarray(mask_bytes, np.uint8), bitorder="little")[:nvox]
    return [bool(x) for x in b]


def test_vdw_volume_solvent_shell_and_clash_bit_for_bit(orc, c2):
    """Lds.vdW_volume / first_solvent_shell (lds.ml:148-196) and G3D.vdW_clash_OR / Mol.protein_ligand_clash
    (G3D.ml:162-186, mol.ml:1195-1203) on a 1.0 A grid over a slice of the receptor"""
    m = c2["rec"]
    c = np.array(c2["roi"][:3])
    near = np.where((m.xs - c[0]) ** 2 + (m.ys - c[1]) ** 2 + (m.zs - c[2]) ** 2 < 9.0 ** 2)[0][:160]
    lo = np.array([m.xs[near].min(), m.ys[near].min(), m.zs[near].min()]) - 5.0
    xs, ys, zs, rr = m.xs[near] - lo[0], m.ys[near] - lo[1], m.zs[near] - lo[2], m.r[near]
    box = (float(xs.max() + 5.0), float(ys.max() + 5.0), float(zs.max() + 5.0))
    step = 1.0
    dims = orc.grid_from_box(step, *box)
    g = ref.Grid(step, *box)
    assert (g.x_dim, g.y_dim, g.z_dim) == tuple(dims)
    nvox = dims[0] * dims[1] * dims[2]
    atoms = list(zip(xs.tolist(), ys.tolist(), zs.tolist()))
    vol = ref.vdW_volume(g, atoms, rr.tolist())
    assert vol == _bits(orc.vdw_volume(xs, ys, zs, rr, step, dims), nvox)
    assert 0 < sum(vol) < nvox
    shell = ref.first_solvent_shell(g, atoms, rr.tolist())
    assert shell == _bits(orc.first_solvent_shell(xs, ys, zs, rr, step, dims), nvox)
    assert sum(shell) > 0 and not any(a and b for a, b in zip(vol, shell))      # the shell excludes the vdW volume
    # clash test of ligand poses: some inside the protein slice, some in the empty margin
    lig = c2["lig"]
    cx, cy, cz = c2["centered"]
    rng = np.random.default_rng(5)
    mask_bytes = orc.vdw_volume(xs, ys, zs, rr, step, dims)
    seen = set()
    for trial in range(14):
        t = np.array([rng.uniform(6.0, box[0] - 6.0), rng.uniform(6.0, box[1] - 6.0), rng.uniform(6.0, box[2] - 6.0)])
        s = 0.3                                              # shrink the ligand so that every atom stays inside the box
        if trial >= 12:                                      # two poses in the empty margin at the box corner
            t, s = np.array([1.5, 1.5, 1.5]) + 0.2 * (trial - 12), 0.08
        px, py, pz = cx * s + t[0], cy * s + t[1], cz * s + t[2]
        assert px.min() > 0 and px.max() < box[0] - 1 and py.min() > 0 and py.max() < box[1] - 1 and pz.min() > 0 and pz.max() < box[2] - 1
        want = orc.protein_ligand_clash(step, dims, mask_bytes, px, py, pz)
        got = ref.protein_ligand_clash(g, vol, list(zip(px.tolist(), py.tolist(), pz.tolist())))
        assert got == want
        seen.add(got)
    assert seen == {True, False}


def test_exhaustive_scan_bit_for_bit(orc, c2, c2_roi_rec, c2lig):
    """Lds.exhaustive_rigid_ligand_docking (lds.ml:1040-1114) on a small lattice and rotation set: the lattice of
    Grid.from_box over ROI.get_bounds, the strict in-ROI test, score = E_intra + E_inter per pose, first-pose-wins argmin
    and its frame number, the k best scores -- the C oracle's scan against the Python restatement, bit for bit"""
    cx, cy, cz = c2["centered"]
    roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 2.2)
    rot9 = orc.so3_rotations(5)
    e_intra = 1.25
    want = orc.scan(c2_roi_rec, c2lig, cx, cy, cz, roi, 1.5, rot9, 12, scorer=0, e_intra_const=e_intra)
    prot = _ref_prot(c2_roi_rec)
    lig0 = _ref_lig(c2lig, cx, cy, cz, center=(0.0, 0.0, 0.0))
    top, best, frame, n_scored = ref.exhaustive_rigid_ligand_docking(
        12, roi, 1.5, [tuple(r) for r in rot9], lig0, lambda m: e_intra + ref.ene_inter_UFF_shifted_brute(prot, m))
    assert n_scored == want["n_scored"] and n_scored >= 5 * 7
    assert best == want["best_score"] and frame == want["best_frame"]
    assert top == list(want["top_scores"])


def test_desolvation_sums_bit_for_bit(orc, c2):
    """Lds.protein_desolv and Lds.desolvation_penalty (lds.ml:204-267) on a 1 A grid over a slice of the receptor"""
    m = c2["rec"]
    c = np.array(c2["roi"][:3])
    near = np.where((m.xs - c[0]) ** 2 + (m.ys - c[1]) ** 2 + (m.zs - c[2]) ** 2 < 8.0 ** 2)[0][:120]
    lo = np.array([m.xs[near].min(), m.ys[near].min(), m.zs[near].min()]) - 6.0
    xs, ys, zs, rr, qq = m.xs[near] - lo[0], m.ys[near] - lo[1], m.zs[near] - lo[2], m.r[near], m.q[near]
    box = (float(xs.max() + 6.0), float(ys.max() + 6.0), float(zs.max() + 6.0))
    step = 1.0
    dims = orc.grid_from_box(step, *box)
    g = ref.Grid(step, *box)
    nvox = dims[0] * dims[1] * dims[2]
    atoms = list(zip(xs.tolist(), ys.tolist(), zs.tolist()))
    shell_bytes = orc.first_solvent_shell(xs, ys, zs, rr, step, dims)
    shell = _bits(shell_bytes, nvox)

    class _R:
        pass
    rec = _R(); rec.xs, rec.ys, rec.zs, rec.q, rec.n = xs, ys, zs, qq, len(xs)
    roi = (float(c[0] - lo[0]), float(c[1] - lo[1]), float(c[2] - lo[2]), 6.5)
    want = orc.protein_desolv(rec, step, dims, shell_bytes, roi)
    prot = ref.Mol(xs, ys, zs, qq, m.anum[near])
    got = ref.protein_desolv(roi, g, shell, prot)
    assert got == list(want)
    assert np.count_nonzero(want) > 50
    # a shrunken ligand pose next to the protein surface: both terms of the penalty
    lig = c2["lig"]
    cx, cy, cz = c2["centered"]
    nonzero = 0
    for t, s in ((np.array(roi[:3]) + np.array([0.0, 0.0, 3.0]), 0.35), (np.array(roi[:3]) + np.array([2.0, -1.0, 0.0]), 0.25),
                 (np.array([4.5, 4.5, 4.5]), 0.05)):
        px, py, pz = cx * s + t[0], cy * s + t[1], cz * s + t[2]
        assert min(px.min(), py.min(), pz.min()) > 4.0          # every atom's shell cube stays inside the grid (OCaml would raise)
        assert px.max() < box[0] - 4.0 and py.max() < box[1] - 4.0 and pz.max() < box[2] - 4.0
        wp, wl = orc.desolvation_penalty(step, dims, shell_bytes, want, px, py, pz, lig.q, lig.r)
        gp, gl = ref.desolvation_penalty(g, got, shell, list(zip(px.tolist(), py.tolist(), pz.tolist())), lig.q.tolist(), lig.r.tolist())
        assert (gp, gl) == (wp, wl), (t, gp, wp, gl, wl)
        nonzero += (wp != 0.0) + (wl != 0.0)
    assert nonzero >= 2


def test_n3_bitmasks_bit_for_bit(orc, c2):
    """Lds.bitmask_whole_protein (lds.ml:97-145) and Lds.bitmask_ROI_only (lds.ml:269-305) on a coarse grid"""
    m = c2["rec"]
    c = np.array(c2["roi"][:3])
    near = np.where((m.xs - c[0]) ** 2 + (m.ys - c[1]) ** 2 + (m.zs - c[2]) ** 2 < 7.0 ** 2)[0][:60]
    lo = np.array([m.xs[near].min(), m.ys[near].min(), m.zs[near].min()]) - 16.0
    xs, ys, zs = m.xs[near] - lo[0], m.ys[near] - lo[1], m.zs[near] - lo[2]
    box = (float(xs.max() + 16.0), float(ys.max() + 16.0), float(zs.max() + 16.0))
    step = 2.0
    dims = orc.grid_from_box(step, *box)
    g = ref.Grid(step, *box)
    nvox = dims[0] * dims[1] * dims[2]
    whole = ref.bitmask_whole_protein(g, list(zip(xs.tolist(), ys.tolist(), zs.tolist())))
    assert whole == _bits(orc.bitmask_whole_protein(xs, ys, zs, step, dims), nvox)
    assert 0 < sum(whole) < nvox
    roi = (float(c[0] - lo[0]), float(c[1] - lo[1]), float(c[2] - lo[2]), 3.0)
    only = ref.bitmask_ROI_only((roi[0], roi[1], roi[2], 3.0 - 24.0 + 9.0), g)        # out radius chosen so that the sphere cuts the grid
    want = _bits(orc.bitmask_sphere(step, dims, roi[:3], (3.0 - 24.0 + 9.0) + 12.0 * 2.0), nvox)
    assert only == want and 0 < sum(only) < nvox
