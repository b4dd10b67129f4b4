"""The reference's two pose-feed tools restated as C programs on the C ABI (tools/lig_rot_sample.c,
tools/place_ligand.c): same argv as src/lig_rot_sample.ml:11-21 and src/place_ligand.ml:9-33, host only (no GPU, no
mmo_init).  Coordinates are checked against the oracle's restatement of Mol.center_rotate_translate_copy
(mol.ml:705-710) and Optim.apply_config (optim.ml:64-80) through the writer's own %10.4f, the mol2 text against
Mol2.output_one's formats (mol2.ml:184-190, 209-210, 326-343)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mmo_b200
from mmo_b200 import pqrs
from test_molfile import MOL2

BIN = os.path.dirname(mmo_b200.LIB_PATH)
CHAIN_RING = "@<TRIPOS>MOLECULE\n" + MOL2.split("@<TRIPOS>MOLECULE\n")[1]      # 13 atoms, one of them a lone pair


def _exe(name):
    p = os.path.join(BIN, name)
    assert os.path.exists(p), f"build it with `make -C mmo_b200/csrc tools` (done by __graft_entry__.build())"
    return p


def _blocks(path):
    """[(name, counts line, atom lines, bond lines)] of a mol2 file written by Mol2.output_one"""
    out = []
    for b in open(path).read().split("@<TRIPOS>MOLECULE\n")[1:]:
        l = b.split("\n")
        assert l[2:5] == ["SMALL", "USER_CHARGES", ""] and l[5] == "@<TRIPOS>ATOM"
        k = l.index("@<TRIPOS>BOND")
        out.append((l[0], l[1], l[6:k], [x for x in l[k + 1:] if x]))
    return out


def _atom_line(i, name, x, y, z, typ, q):
    return "%7d %-8s%10.4f%10.4f%10.4f %-8s  1 <0>     %10.4f" % (i, name, x, y, z, typ, q)


NAMES = ["C1", "C2", "C3", "C4", "C5", "C6", "C7", "C8", "O1", "H1", "N1", "Cl1"]
TYPES = ["C.3", "C.3", "C.ar", "C.ar", "C.ar", "C.ar", "C.ar", "C.ar", "O.3", "H", "N.am", "Cl"]
# bond 12 (O1-LP1) goes with the lone pair; Cl1 moves from atom 13 to atom 12
BONDS = ["%6d%5d%5d %s" % (i + 1, s, d, t) for i, (s, d, t) in enumerate(
    [(1, 2, "1"), (2, 3, "1"), (3, 4, "ar"), (4, 5, "ar"), (5, 6, "ar"), (6, 7, "ar"), (7, 8, "ar"), (8, 3, "ar"),
     (1, 9, "1"), (9, 10, "1"), (5, 11, "am"), (11, 12, "1")])]


@pytest.fixture()
def mol2_in(tmp_path):
    fn = tmp_path / "in.mol2"
    fn.write_text(CHAIN_RING)
    return str(fn)


def test_lig_rot_sample_tool(tmp_path, orc, mol2_in):
    out = str(tmp_path / "rot.mol2")
    n = 9
    r = subprocess.run([_exe("lig_rot_sample"), str(n), mol2_in, out], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    m = mmo_b200.MolFile(mol2_in).mol(0)
    assert m.n == 12
    center = [orc.favg(m.xs), orc.favg(m.ys), orc.favg(m.zs)]
    rot = orc.so3_rotations(n)
    blocks = _blocks(out)
    assert len(blocks) == n
    dp = C.POINTER(C.c_double)
    for k, (name, counts, atoms, bonds) in enumerate(blocks):
        assert name == "chain_ring" and counts == "%5d%6d%6d%6d%6d" % (12, 12, 0, 0, 0) and bonds == BONDS
        ox, oy, oz = (np.empty(m.n) for _ in range(3))
        orc.lib().orc_center_rotate_translate(C.c_int(m.n), orc.d(m.xs)[1], orc.d(m.ys)[1], orc.d(m.zs)[1],
                                              orc.d(center)[1], orc.d(rot[k])[1], orc.d(center)[1],
                                              ox.ctypes.data_as(dp), oy.ctypes.data_as(dp), oz.ctypes.data_as(dp))
        assert atoms == [_atom_line(i + 1, NAMES[i], ox[i], oy[i], oz[i], TYPES[i], m.q[i]) for i in range(m.n)]
    # the output is itself a mol2 file of n molecules with the input's topology
    back = mmo_b200.MolFile(out)
    assert back.n_mols == n and back.n_skipped == 0
    b3 = back.mol(3)
    assert np.array_equal(b3.dists, m.dists) and np.array_equal(b3.rb_left, m.rb_left)
    # rigid motion about the centre: the centre stays, distances stay (to the 4 decimals of the text)
    assert abs(orc.favg(b3.xs) - center[0]) < 1e-4 and abs(orc.favg(b3.zs) - center[2]) < 1e-4
    d = lambda a, i, j: np.sqrt((a.xs[i] - a.xs[j]) ** 2 + (a.ys[i] - a.ys[j]) ** 2 + (a.zs[i] - a.zs[j]) ** 2)
    assert abs(d(b3, 0, 11) - d(m, 0, 11)) < 5e-4
    # the same through the library calls, without the text
    X, Y, Z = mmo_b200.MolFile(mol2_in).rotated_copies(0, rot)
    assert ["%10.4f" % v for v in X[3]] == [a[16:26] for a in blocks[3][2]]


def test_lig_rot_sample_usage_and_errors(tmp_path, mol2_in):
    exe = _exe("lig_rot_sample")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("usage:\n") and "num_samples input.mol2 output.mol2" in r.stderr
    r = subprocess.run([exe, "3", str(tmp_path / "missing.mol2"), str(tmp_path / "o.mol2")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr
    two = tmp_path / "two.mol2"
    two.write_text(CHAIN_RING + CHAIN_RING.replace("chain_ring", "again"))
    r = subprocess.run([exe, "3", str(two), str(tmp_path / "o.mol2")], capture_output=True, text=True)
    assert r.returncode != 0 and "several ligands" in r.stderr            # lig_rot_sample.ml:36
    r = subprocess.run([exe, "0", mol2_in, str(tmp_path / "o.mol2")], capture_output=True, text=True)
    assert r.returncode == 0 and open(tmp_path / "o.mol2").read() == ""   # SO3.sample 0: an empty file


@pytest.mark.parametrize("with_bonds", [False, True])
def test_place_ligand_tool(tmp_path, orc, mol2_in, with_bonds):
    out = str(tmp_path / "placed.mol2")
    f = mmo_b200.MolFile(mol2_in)
    m = f.mol(0)
    assert m.n_rbonds == 4
    cfg = [31.25, -4.5, 17.125, 0.3, -1.1, 2.9] + ([0.7, -2.2, 3.0, -0.05] if with_bonds else [])
    r = subprocess.run([_exe("place_ligand"), mol2_in, out] + [repr(v) for v in cfg], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert f"rbonds given: {len(cfg) - 6}" in r.stderr                    # place_ligand.ml:33
    c = [orc.favg(m.xs), orc.favg(m.ys), orc.favg(m.zs)]
    cx, cy, cz = m.xs + (-c[0]), m.ys + (-c[1]), m.zs + (-c[2])           # Mol.center (mol.ml:699-702)
    wx, wy, wz, too_long = orc.apply_config(m, cx, cy, cz, cfg)
    assert not too_long
    (name, counts, atoms, bonds), = _blocks(out)
    assert name == "chain_ring" and bonds == BONDS
    assert atoms == [_atom_line(i + 1, NAMES[i], wx[i], wy[i], wz[i], TYPES[i], m.q[i]) for i in range(m.n)]
    x, y, z, tl = f.apply_config(0, cfg)
    assert np.array_equal(x, wx) and np.array_equal(y, wy) and np.array_equal(z, wz) and not tl
    # a rigid-body placement puts the centre at (x, y, z)
    if not with_bonds:
        assert abs(orc.favg(x) - cfg[0]) < 1e-9 and abs(orc.favg(y) - cfg[1]) < 1e-9


def test_place_ligand_usage_and_errors(tmp_path, mol2_in):
    exe = _exe("place_ligand")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "input.mol2 output.mol2 x y z a b g [rbonds]" in r.stderr
    r = subprocess.run([exe, mol2_in, str(tmp_path / "o.mol2"), "0", "0", "0", "0", "0", "0", "0.5", "0.5"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "2 rbond values but mol has 4" in r.stderr       # place_ligand.ml:49-55
    r = subprocess.run([exe, mol2_in, str(tmp_path / "o.mol2"), "0", "0", "zero", "0", "0", "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "not a number" in r.stderr


def test_write_mol2_refuses_pqrs_molecules(tmp_path):
    from mmo_b200 import workloads
    f = mmo_b200.MolFile(os.path.join(workloads.GOLDEN, "docked.pqrs"), kind="ligand_pqrs")
    with pytest.raises(RuntimeError, match="not read from a mol2 file"):
        f.write_mol2(str(tmp_path / "x.mol2"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference data only in the build container")
@pytest.mark.parametrize("name", ["docked", "ligdecs", "minimized"])
def test_identity_placement_reproduces_the_reference_mol2_atom_lines(tmp_path, orc, name):
    """data/*.mol2 carry the reference writer's own line formats: placing the ligand back on its own centre with
    zero angles must give back every ATOM and BOND line of the input, character for character"""
    src = f"/root/reference/data/{name}.mol2"
    m = mmo_b200.MolFile(src).mol(0)
    c = [orc.favg(m.xs), orc.favg(m.ys), orc.favg(m.zs)]
    out = str(tmp_path / "back.mol2")
    r = subprocess.run([_exe("place_ligand"), src, out] + [repr(float(v)) for v in c] + ["0", "0", "0"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = open(src).read().split("\n")
    a0, b0 = lines.index("@<TRIPOS>ATOM"), lines.index("@<TRIPOS>BOND")
    (nm, counts, atoms, bonds), = _blocks(out)
    assert nm == lines[1].strip() and atoms == lines[a0 + 1:b0]
    assert bonds == [l for l in lines[b0 + 1:b0 + 1 + len(bonds)]]


def test_mol2_atoms_reader_takes_disconnected_molecules(tmp_path):
    """Mol2.read_one_from_file as scissors uses it (scissors.ml:48-55): first molecule only, no graph analysis"""
    import ctypes as C
    fn = tmp_path / "three.mol2"
    fn.write_text(MOL2)
    L = mmo_b200.lib()
    h = C.c_void_p()
    assert L.mmo_molfile_read_mol2_atoms(os.fsencode(str(fn)), C.byref(h)) == 0
    n, sk = C.c_int32(), C.c_int32()
    assert L.mmo_molfile_count(h, C.byref(n), C.byref(sk)) == 0 and (n.value, sk.value) == (1, 0)
    na = C.c_int32()
    assert L.mmo_molfile_shape(h, 0, C.byref(na), None, None, None, 0) == 0 and na.value == 12   # lone pair dropped
    L.mmo_molfile_destroy(h)
    # a molecule in two pieces is skipped by the mol2pqrs reader but is fine for the atoms-only one
    broken = tmp_path / "broken.mol2"
    broken.write_text("@<TRIPOS>MOLECULE\n" + MOL2.split("@<TRIPOS>MOLECULE\n")[2])
    assert mmo_b200.MolFile(str(broken)).n_mols == 0
    assert L.mmo_molfile_read_mol2_atoms(os.fsencode(str(broken)), C.byref(h)) == 0
    assert L.mmo_molfile_count(h, C.byref(n), C.byref(sk)) == 0 and (n.value, sk.value) == (1, 0)
    L.mmo_molfile_destroy(h)


@pytest.mark.gpu
def test_scissors_tool_and_carve_kernel(tmp_path, gpu, c2):
    """scissors (scissors.ml:24-66): protein atoms whose nearest ligand atom is within the cut-off, as pqrs lines"""
    import ctypes as C
    m, lm = c2["rec_orig"], c2["lig"]
    # the carve itself against numpy (sqrt of the nearest squared distance <= cutoff)
    keep = np.zeros(m.n, np.uint8)
    nk = C.c_int32()
    dp = C.POINTER(C.c_double)
    arr = [np.ascontiguousarray(a, np.float64) for a in (m.xs, m.ys, m.zs, lm.xs, lm.ys, lm.zs)]
    rc = gpu.lib().mmo_carve_near_ligand(C.c_int32(m.n), arr[0].ctypes.data_as(dp), arr[1].ctypes.data_as(dp), arr[2].ctypes.data_as(dp),
                                         C.c_int32(lm.n), arr[3].ctypes.data_as(dp), arr[4].ctypes.data_as(dp), arr[5].ctypes.data_as(dp),
                                         C.c_double(5.0), keep.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nk))
    assert rc == 0
    d2 = ((m.xs[:, None] - lm.xs[None, :]) ** 2 + (m.ys[:, None] - lm.ys[None, :]) ** 2) + (m.zs[:, None] - lm.zs[None, :]) ** 2
    want = np.sqrt(d2.min(axis=1)) <= 5.0
    assert np.array_equal(keep.astype(bool), want) and nk.value == want.sum() and 20 < nk.value < m.n
    # the tool: protein and ligand as mol2 files (written here from the fixtures), output = "%g %g %g %g %g %s" lines
    sym = {1: "H", 6: "C", 7: "N", 8: "O", 9: "F", 12: "Mg", 15: "P", 16: "S", 17: "Cl", 35: "Br", 53: "I"}

    def write_mol2(path, mol):
        with open(path, "w") as f:
            f.write("@<TRIPOS>MOLECULE\n%s\n%5d%6d%6d%6d%6d\nSMALL\nUSER_CHARGES\n\n@<TRIPOS>ATOM\n" % (mol.name, mol.n, 0, 0, 0, 0))
            for i in range(mol.n):
                f.write("%7d %-8s%10.4f%10.4f%10.4f %-8s  1 <0>     %10.4f\n" %
                        (i + 1, "A%d" % i, mol.xs[i], mol.ys[i], mol.zs[i], sym[int(mol.anum[i])], mol.q[i]))
            f.write("@<TRIPOS>BOND\n")
    write_mol2(tmp_path / "prot.mol2", m)
    write_mol2(tmp_path / "lig.mol2", lm)
    out = tmp_path / "site.pqrs"
    r = subprocess.run([_exe("scissors"), "-l", str(tmp_path / "lig.mol2"), "-p", str(tmp_path / "prot.mol2"), "-o", str(out), "-d", "6.5"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    # what the tool saw are the 4-decimal coordinates of the mol2 text
    rx, ry, rz = (np.array([float("%.4f" % v) for v in a]) for a in (m.xs, m.ys, m.zs))
    qx, qy, qz = (np.array([float("%.4f" % v) for v in a]) for a in (lm.xs, lm.ys, lm.zs))
    d2 = ((rx[:, None] - qx[None, :]) ** 2 + (ry[:, None] - qy[None, :]) ** 2) + (rz[:, None] - qz[None, :]) ** 2
    k = np.sqrt(d2.min(axis=1)) <= 6.5
    lines = open(out).read().strip().split("\n")
    want_lines = ["%g %g %g %g %g %s" % (rx[i], ry[i], rz[i], float("%.4f" % m.q[i]), m.r[i], sym[int(m.anum[i])])
                  for i in range(m.n) if k[i]]
    assert lines == want_lines
    r = subprocess.run([_exe("scissors")], capture_output=True, text=True)
    assert r.returncode == 1 and "-l <ligand.mol2>: xtal ligand input file" in r.stderr
