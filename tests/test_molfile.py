"""N2: the library's C++ mol2 / pqrs readers (mmo_molfile_*, restating src/mol2pqrs.ml, src/mol_graph.ml,
src/pqrs.ml) against the Python restatement that produced the committed fixtures, on the fixtures themselves,
on a synthetic multi-molecule mol2 written here, and -- where the reference's data directory is present (build
container only) -- on data/*.mol2."""
import filecmp
import os

import numpy as np
import pytest

import mmo_b200
from mmo_b200 import pqrs, workloads

G = workloads.GOLDEN


def _same_mol(a, b):
    assert a.name == b.name and a.n == b.n
    for f in ("xs", "ys", "zs", "q", "r"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.anum, b.anum)
    if b.dists is not None:
        assert np.array_equal(a.dists, b.dists)
        assert np.array_equal(a.rb_left, b.rb_left) and np.array_equal(a.rb_right, b.rb_right)
        assert len(a.rgroups) == len(b.rgroups)
        for ga, gb in zip(a.rgroups, b.rgroups):
            assert np.array_equal(ga, gb)


@pytest.mark.parametrize("name", ["docked", "ligdecs", "minimized"])
def test_pqrs_reader_and_writer_round_trip(tmp_path, name):
    fn = os.path.join(G, name + ".pqrs")
    want = pqrs.read_ligands_pqrs(fn)
    f = mmo_b200.MolFile(fn, kind="ligand_pqrs")
    assert f.n_mols == len(want) and f.n_skipped == 0
    for k, w in enumerate(want):
        _same_mol(f.mol(k), w)
    ta, tq = pqrs.assign_ff_types(want)              # mol.ml:280-293: first-seen order, exact float equality
    ga, gq = f.types()
    assert np.array_equal(ga, ta) and np.array_equal(gq, tq)
    for k, w in enumerate(want):
        assert np.array_equal(f.mol(k).typ, w.typ)
    out = tmp_path / "again.pqrs"
    f.write_pqrs(str(out))
    assert filecmp.cmp(str(out), fn, shallow=False)  # byte for byte the fixture (values through %g)


def test_receptor_pqrs_reader(tmp_path):
    fn = os.path.join(G, "xtal_rec.pqrs")
    want = pqrs.read_receptor_pqrs(fn)
    f = mmo_b200.MolFile(fn, kind="receptor_pqrs")
    assert f.n_mols == 1
    _same_mol(f.mol(0), want)
    out = tmp_path / "rec.pqrs"
    f.write_pqrs(str(out))
    assert filecmp.cmp(str(out), fn, shallow=False)


MOL2 = """@<TRIPOS>MOLECULE
chain_ring
 13 13 1 0 0
SMALL
USER_CHARGES

@<TRIPOS>ATOM
      1 C1          0.0000    0.0000    0.0000 C.3       1 LIG        -0.1800
      2 C2          1.5000    0.1000    0.0000 C.3       1 LIG        -0.1200
      3 C3          2.1000    1.5000    0.2000 C.ar      1 LIG         0.0300
      4 C4          3.5000    1.7000    0.3000 C.ar      1 LIG        -0.0600
      5 C5          4.1000    3.0000    0.5000 C.ar      1 LIG        -0.0600
      6 C6          3.3000    4.1000    0.6000 C.ar      1 LIG        -0.0600
      7 C7          1.9000    3.9000    0.5000 C.ar      1 LIG        -0.0600
      8 C8          1.3000    2.6000    0.3000 C.ar      1 LIG        -0.0600
      9 O1         -0.6000   -1.2000    0.1000 O.3       1 LIG        -0.3900
     10 H1         -1.5500   -1.1000    0.1000 H         1 LIG         0.2100
     11 N1          5.5000    3.2000    0.6000 N.am      1 LIG        -0.3000
     12 LP1        -0.7000   -1.6000    0.9000 LP        1 LIG         0.0000
     13 Cl1         6.2000    4.8000    0.9000 Cl        1 LIG         0.1000
@<TRIPOS>BOND
     1     1     2    1
     2     2     3    1
     3     3     4   ar
     4     4     5   ar
     5     5     6   ar
     6     6     7   ar
     7     7     8   ar
     8     8     3   ar
     9     1     9    1
    10     9    10    1
    11     5    11   am
    12     9    12    1
    13    11    13    1
@<TRIPOS>MOLECULE
broken_in_two
 3 1 1 0 0
SMALL
USER_CHARGES

@<TRIPOS>ATOM
      1 C1          0.0000    0.0000    0.0000 C.3       1 LIG        -0.1000
      2 C2          1.5000    0.0000    0.0000 C.3       1 LIG        -0.1000
      3 C3          9.0000    0.0000    0.0000 C.3       1 LIG         0.2000
@<TRIPOS>BOND
     1     1     2    1
@<TRIPOS>MOLECULE
ethanol_like
 4 3 1 0 0
SMALL
USER_CHARGES

@<TRIPOS>ATOM
      1 C1          0.0000    0.0000    0.0000 C.3       1 LIG        -0.1800
      2 C2          1.5000    0.0000    0.0000 C.3       1 LIG         0.1500
      3 O1          2.0000    1.3000    0.0000 O.3       1 LIG        -0.3900
      4 H1          2.9000    1.3000    0.1000 H         1 LIG         0.2100
@<TRIPOS>BOND
     1     1     2    1
     2     2     3    1
     3     3     4    1
"""


def test_mol2_reader_multi_molecule_rotatable_bonds_and_lone_pairs(tmp_path):
    fn = tmp_path / "three.mol2"
    fn.write_text(MOL2)
    f = mmo_b200.MolFile(str(fn))
    assert f.n_mols == 2 and f.n_skipped == 1          # Mol_graph.Disconnected_atom: the reference emits nothing
    m = f.mol(0)
    assert m.name == "chain_ring" and m.n == 12        # the lone pair is gone (mol2.ml:54-56)
    assert list(m.anum) == [6, 6, 6, 6, 6, 6, 6, 6, 8, 1, 7, 17]
    # rotatable = single, not in a ring, neither end terminal: C1-C2, C2-C3 (ring substituent), C1-O1, C5-N1 ('am' = 1.0)
    bonds = sorted((int(l), int(r)) for l, r in zip(m.rb_left, m.rb_right))
    assert bonds == [(0, 8), (1, 0), (2, 1), (4, 10)]
    # C2-C3: the movable side is the smaller one, here the chain {C1, C2, O1, H1}; the axis tip C2 is not listed
    k = [i for i in range(m.n_rbonds) if (m.rb_left[i], m.rb_right[i]) == (2, 1)][0]
    assert sorted(m.rgroups[k]) == [0, 8, 9]
    assert m.dists[0 + 5 * m.n] == 5 and m.dists[9 + 11 * m.n] == 8      # C1..C6 around the ring; H1-O1-C1-C2-C3-C4-C5-N1-Cl1
    # same file, one molecule at a time, through the Python restatement
    single = tmp_path / "one.mol2"
    blocks = MOL2.split("@<TRIPOS>MOLECULE\n")[1:]
    for k, b in zip((0, 1), (blocks[0], blocks[2])):
        single.write_text("@<TRIPOS>MOLECULE\n" + "\n".join(l for l in b.split("\n") if " LP " not in l))
        if k == 0:
            continue      # the Python reader has no lone-pair handling; compared through the C++ writer below
        _same_mol(f.mol(k), pqrs.mol2_to_ligand(str(single)))
    out = tmp_path / "three.pqrs"
    f.write_pqrs(str(out))
    back = pqrs.read_ligands_pqrs(str(out))
    assert [b.name for b in back] == ["chain_ring", "ethanol_like"]
    for k, b in enumerate(back):
        _same_mol(f.mol(k), b)


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference data only in the build container")
@pytest.mark.parametrize("name", ["docked", "ligdecs", "minimized"])
def test_mol2_reader_on_the_reference_ligands(name):
    f = mmo_b200.MolFile(f"/root/reference/data/{name}.mol2")
    want = pqrs.read_ligands_pqrs(os.path.join(G, name + ".pqrs"))
    assert f.n_mols == len(want) == 1
    got = f.mol(0)
    # the fixture went through %g: compare the topology exactly and the numbers through the same formatting
    assert got.name == want[0].name and np.array_equal(got.anum, want[0].anum)
    assert np.array_equal(got.dists, want[0].dists)
    assert np.array_equal(got.rb_left, want[0].rb_left) and np.array_equal(got.rb_right, want[0].rb_right)
    for ga, gb in zip(got.rgroups, want[0].rgroups):
        assert np.array_equal(ga, gb)
    for fld in ("xs", "ys", "zs", "q", "r"):
        assert [float("%g" % v) for v in getattr(got, fld)] == list(getattr(want[0], fld))


@pytest.mark.gpu
def test_molfile_ligand_scores_like_the_python_path(gpu, orc, c2, c2_roi_rec):
    """the handle built by mmo_molfile_ligand is the ligand Ligand.from_mol builds: identical fp64 energies"""
    f = mmo_b200.MolFile(os.path.join(G, "docked.pqrs"), kind="ligand_pqrs")
    lig_c = f.ligand(0, centered=True)
    lig_p = gpu.Ligand.from_mol(c2["lig"], centered=True)
    rec = gpu.Receptor.from_mol(c2_roi_rec)
    R, t = workloads.random_poses_in_sphere(50, c2["roi"][:3], 6.0, seed=3)
    a = gpu.Mol.score_poses(rec, lig_c, R, t, prec=gpu.PREC_FP64)
    b = gpu.Mol.score_poses(rec, lig_p, R, t, prec=gpu.PREC_FP64)
    assert np.array_equal(a, b)
    Xc = np.tile(lig_c.xs, (3, 1)) + 40.0
    assert np.array_equal(gpu.Mol.ene_intra_UFFNB_brute(lig_c, Xc, np.tile(lig_c.ys, (3, 1)), np.tile(lig_c.zs, (3, 1))),
                          gpu.Mol.ene_intra_UFFNB_brute(lig_p, Xc, np.tile(lig_p.ys, (3, 1)), np.tile(lig_p.zs, (3, 1))))


@pytest.mark.parametrize("name", ["docked", "ligdecs", "minimized", "xtal_rec"])
def test_reference_pqrs_center_script_agrees(name):
    """golden vectors made by the reference's own bin/pqrs_center.py run on our fixtures
    (tools/make_refpy_fixture.py): same header line, same geometric centre through '%g'"""
    import json
    want = json.load(open(os.path.join(G, "pqrs_center.json")))[name]
    f = mmo_b200.MolFile(os.path.join(G, name + ".pqrs"), kind="receptor_pqrs" if name == "xtal_rec" else "ligand_pqrs")
    m = f.mol(0)
    header = f"{m.n}:{m.name}" if name == "xtal_rec" else f"{m.n}:{m.n_rbonds}:{m.name}"
    assert header == want[0]
    sx = sy = sz = 0.0
    for i in range(m.n):                         # the script's plain left-to-right sums
        sx += m.xs[i]; sy += m.ys[i]; sz += m.zs[i]
    assert "%g %g %g" % (sx / m.n, sy / m.n, sz / m.n) == want[1]


def test_less_charges_rounds_away_from_zero_and_merges_types():
    """lds --less-charges (lds.ml:1887-1894; Utls.reduce_precision, utls.ml:127-132)"""
    f = mmo_b200.MolFile(os.path.join(G, "ligdecs.pqrs"), kind="ligand_pqrs")
    before = f.mol(0)
    ta0, tq0 = f.types()
    f.reduce_charges()
    after = f.mol(0)
    want = np.array([float(int(q * 100.0 + (0.5 if q >= 0.0 else -0.5))) / 100.0 for q in before.q])
    assert np.array_equal(after.q, want) and np.array_equal(after.xs, before.xs)
    ta1, tq1 = f.types()
    assert len(ta1) <= len(ta0) and len(set(zip(ta1.tolist(), tq1.tolist()))) == len(ta1)
    assert all(tq1[t] == q and ta1[t] == a for t, q, a in zip(after.typ, after.q, after.anum))
