"""Pins the CPU oracle: hand-derived known answers, an independent mpmath
evaluation of the same formulas, and structural properties.  The reference ships no golden vectors
for this path and cannot be built here (F3): parity is "unpinned" beyond these."""
import json
import math
import os

import mpmath as mp
import numpy as np
import pytest

from mmo_b200 import pqrs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.mark.parametrize("case", KAT["pairs"])
def test_pair_known_answers(orc, case):
    a1, q1, a2, q2, r = case["anum1"], case["q1"], case["anum2"], case["q2"], case["r"]
    g = orc.pair_energy(a1, q1, a2, q2, r, shifted=False)
    s = orc.pair_energy(a1, q1, a2, q2, r, shifted=True)
    assert g == pytest.approx(case["global"], rel=1e-13, abs=1e-300)
    assert s == pytest.approx(case["shifted"], rel=1e-13, abs=1e-300)
    rr = max(r, 0.01)
    assert orc.shift_12A(rr) == pytest.approx(case["w"], rel=1e-14, abs=0.0)


def _mp_pair(a1, q1, a2, q2, r, shifted):
    mp.mp.dps = 40
    xi = {0: (0, 0), 1: (2.886, 0.044), 6: (3.851, 0.105), 7: (3.660, 0.069), 8: (3.500, 0.060), 9: (3.364, 0.050),
          12: (3.021, 0.111), 15: (4.147, 0.305), 16: (4.035, 0.274), 17: (3.947, 0.227), 35: (4.189, 0.251),
          53: (4.500, 0.339)}
    x = mp.sqrt(mp.mpf(xi[a1][0]) * mp.mpf(xi[a2][0]))
    d = mp.sqrt(mp.mpf(xi[a1][1]) * mp.mpf(xi[a2][1]))
    r = mp.mpf(r)
    if r < mp.mpf("0.01"):
        r = mp.mpf(0.01)
    p6 = (x / r) ** 6
    e = mp.mpf(332.0637) / 4 * mp.mpf(q1) * mp.mpf(q2) / r + d * (p6 * p6 - 2 * p6)
    if shifted:
        w = (1 - (r / 12) ** 2) ** 2 if r < 12 else mp.mpf(0)
        e = e * w
    return float(e)


def test_pair_against_mpmath(orc):
    rng = np.random.default_rng(1)
    anums = [1, 6, 7, 8, 9, 12, 15, 16, 17, 35, 53, 0]
    for _ in range(300):
        a1, a2 = rng.choice(anums, 2)
        q1, q2 = rng.uniform(-1, 1, 2)
        r = float(rng.choice([rng.uniform(0.0, 0.02), rng.uniform(0.5, 4.0), rng.uniform(4.0, 13.0)]))
        for shifted in (False, True):
            got = orc.pair_energy(int(a1), q1, int(a2), q2, r, shifted)
            want = _mp_pair(int(a1), q1, int(a2), q2, r, shifted)
            assert got == pytest.approx(want, rel=5e-14, abs=1e-18)


def test_unsupported_element_is_nan(orc):
    # UFF.ml:37: the table is initialised with NaN
    assert math.isnan(orc.pair_energy(6, 0.1, 30, 0.1, 3.0, True))


def test_trilinear_known_answer(orc):
    t = KAT["trilinear"]
    dims = (8, 8, 8)
    arr = np.zeros(dims[0] * dims[1] * dims[2], np.float32)
    i0, j0, k0 = t["ijk0"]
    x_dim, xy = dims[0], dims[0] * dims[1]
    idx = lambda i, j, k: i + j * x_dim + k * xy
    a, b, c, d, e, f, g, h = t["corners"]
    arr[idx(i0, j0, k0)] = a; arr[idx(i0 + 1, j0, k0)] = b; arr[idx(i0 + 1, j0 + 1, k0)] = c
    arr[idx(i0, j0 + 1, k0)] = d; arr[idx(i0, j0, k0 + 1)] = e; arr[idx(i0 + 1, j0, k0 + 1)] = f
    arr[idx(i0 + 1, j0 + 1, k0 + 1)] = g; arr[idx(i0, j0 + 1, k0 + 1)] = h
    got = orc.trilin(t["step"], dims, arr, *t["point"])
    assert got == pytest.approx(t["E"], rel=1e-14)


def test_so3_known_answers(orc):
    for i, q in enumerate(KAT["so3_n4"]):
        assert orc.so3_quat(4, i) == pytest.approx(q, rel=1e-14)
    R = orc.so3_rotations(1000).reshape(-1, 3, 3)
    eye = np.einsum("nij,nkj->nik", R, R)
    assert np.abs(eye - np.eye(3)).max() < 1e-14
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-14)


def test_beta(orc):
    assert orc.lib().orc_beta(__import__("ctypes").c_double(293.15)) == pytest.approx(KAT["beta_293_15"], rel=1e-15)


def test_fixture_topology():
    lig = pqrs.read_ligands_pqrs(os.path.join(GOLDEN, "docked.pqrs"))[0]
    t = KAT["docked_topology"]
    assert lig.n == 48 and lig.n_rbonds == 9
    got = sorted(tuple(sorted(p)) for p in zip(lig.rb_left.tolist(), lig.rb_right.tolist()))
    assert got == sorted(tuple(p) for p in t["rbonds"])
    n = lig.n
    npairs = sum(1 for i in range(n) for j in range(i + 1, n) if lig.dists[i + j * n] >= 3)
    assert npairs == t["pairs_ge3"] and int(lig.dists.max()) == t["max_topo_dist"]
    types_a, types_q = pqrs.assign_ff_types([lig])
    assert len(types_a) == t["n_types"]
    assert abs(lig.q.sum()) < 1e-9


def test_shifted_is_global_times_w_pairwise(orc):
    rng = np.random.default_rng(2)
    for _ in range(100):
        r = rng.uniform(0.3, 11.99)
        g = orc.pair_energy(6, 0.3, 8, -0.4, r, False)
        s = orc.pair_energy(6, 0.3, 8, -0.4, r, True)
        assert s == pytest.approx(g * orc.shift_12A(r), rel=1e-13)
    assert orc.pair_energy(6, 0.3, 8, -0.4, 12.0, True) == 0.0
    assert orc.pair_energy(6, 0.3, 8, -0.4, 15.0, True) == 0.0


def test_brl_equals_bst_up_to_summation_order(orc, c2, c2_roi_rec):
    lig = c2["lig"]
    cx, cy, cz = c2["centered"]
    rng = np.random.default_rng(3)
    from mmo_b200 import workloads
    R, t = workloads.random_poses_in_sphere(6, c2["roi"][:3], 6.0, seed=5)
    R[0] = np.eye(3).reshape(9); t[0] = c2["start_pos"]
    X, Y, Z = orc.pose_coords(cx, cy, cz, R, t)
    brl = orc.ene_inter(c2_roi_rec, lig.q, lig.anum, X, Y, Z, shifted=True)
    e, v = orc.ene_inter_components(c2_roi_rec, lig.q, lig.anum, X, Y, Z)
    assert np.allclose(e + v, brl, rtol=1e-11, atol=1e-9)
    # carving the receptor to the ROI neighbourhood does not change a shifted energy (w = 0 beyond 12 A)
    full = orc.ene_inter(c2["rec"], lig.q, lig.anum, X, Y, Z, shifted=True)
    assert np.allclose(full, brl, rtol=1e-12, atol=1e-10)


def test_ten_rotations_of_36_degrees_return_to_start(orc, c2):
    # the reference's own manual check (src/test_rot.ml:28-32), here asserted
    cx, cy, cz = c2["centered"]
    for axis in range(3):
        r = orc.rot_axis(axis, math.radians(36.0))
        x, y, z = np.array(cx), np.array(cy), np.array(cz)
        for _ in range(10):
            x, y, z = orc.rotate_then_translate(x, y, z, r, np.zeros(3))
        rmsd = math.sqrt(((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2).mean())
        assert rmsd < 1e-12


def test_rxyz_decompose_roundtrip(orc):
    rng = np.random.default_rng(4)
    for _ in range(50):
        a, g = rng.uniform(-3.1, 3.1, 2)
        b = rng.uniform(-1.5, 1.5)
        abg = orc.rot_decompose(orc.rot_r_xyz(a, b, g))
        assert abg == pytest.approx([a, b, g], abs=1e-12)


def test_intra_energy_is_rigid_motion_invariant(orc, c2):
    lig = c2["lig"]
    e0 = orc.ene_intra(lig, lig.xs, lig.ys, lig.zs)[0]
    cx, cy, cz = c2["centered"]
    R = orc.so3_rotations(7)
    X, Y, Z = orc.pose_coords(cx, cy, cz, R, np.tile([40.0, 50.0, 60.0], (7, 1)))
    e = orc.ene_intra(lig, X, Y, Z)
    assert np.allclose(e, e0, rtol=1e-10)


def test_interp_matches_direct_at_lattice_nodes(orc):
    from mmo_b200 import workloads
    rec = workloads.synthetic_receptor(200, "cube", 14.0, seed=9, origin=(3.0, 3.0, 3.0))
    dims = orc.grid_from_box(0.5, 10.0, 10.0, 10.0)
    assert dims == (21, 21, 21)
    ta, tq = np.array([6, 8, 1], np.int32), np.array([0.1, -0.5, 0.2])
    maps = orc.grid_build(rec, 0.5, dims, ta, tq)
    rng = np.random.default_rng(10)
    for _ in range(20):
        i, j, k = rng.integers(1, 19, 3)
        p = (0.5 * i, 0.5 * j, 0.5 * k)
        for t in range(3):
            direct = orc.ene_inter(rec, [tq[t]], [ta[t]], [[p[0]]], [[p[1]]], [[p[2]]], shifted=True)[0]
            direct = min(direct, 1e5)
            got = orc.trilin(0.5, dims, maps[t], *p)
            assert got == pytest.approx(np.float32(direct), rel=2e-7, abs=1e-30)
