"""K3/K4/K5 parity: grid-map build, trilinear lookup, vdW bitmask -- bit-identical to the oracle."""
import os

import numpy as np
import pytest

from mmo_b200 import pqrs, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(gpu, orc):
    rec_m = workloads.synthetic_receptor(400, "cube", 20.0, seed=21, origin=(2.0, 2.0, 2.0))
    dims = orc.grid_from_box(0.375, 9.0, 8.0, 7.0)
    lig = pqrs.read_ligands_pqrs(os.path.join(workloads.GOLDEN, "ligdecs.pqrs"))[0]
    ta, tq = pqrs.assign_ff_types([lig])
    return rec_m, dims, lig, ta, tq


def test_grid_build_bit_identical(gpu, orc, small):
    rec_m, dims, lig, ta, tq = small
    rec = gpu.Receptor.from_mol(rec_m)
    want = orc.grid_build(rec_m, 0.375, dims, ta, tq)
    g, got = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta, tq)
    assert got.shape == want.shape == (22, dims[0] * dims[1] * dims[2])
    assert np.array_equal(got, want)
    assert (got == np.float32(1e5)).any()          # the max_E clamp is exercised
    assert np.array_equal(g.download(), want)


def test_grid_build_masked(gpu, orc, small):
    rec_m, dims, lig, ta, tq = small
    rec = gpu.Receptor.from_mol(rec_m)
    mask = orc.bitmask_sphere(0.375, dims, (4.0, 4.0, 3.0), 3.0)
    want = orc.grid_build(rec_m, 0.375, dims, ta[:5], tq[:5], mask=mask)
    g, got = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta[:5], tq[:5], mask_bits=mask)
    assert np.array_equal(got, want)
    assert (got == 0.0).sum() > got.size // 3      # unmasked voxels stay exactly 0.0 (G3D.create)


def test_trilin_and_interp_bit_identical(gpu, orc, small):
    rec_m, dims, lig, ta, tq = small
    rec = gpu.Receptor.from_mol(rec_m)
    g, maps = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta, tq)
    rng = np.random.default_rng(22)
    hi = 0.375 * (np.array(dims) - 1)
    P = rng.uniform(0.0, 0.999, (500, 3)) * hi
    for t in (0, 7, 21):
        got = gpu.G3D.trilin(g, t, P[:, 0], P[:, 1], P[:, 2])
        want = np.array([orc.trilin(0.375, dims, maps[t], *p) for p in P])
        assert np.array_equal(got, want)
    # whole-ligand interpolated energy: small rigid poses that stay inside the grid
    L = gpu.Ligand.from_mol(lig, centered=True)
    scale = 0.25                                    # shrink the template so that it fits the 9x8x7 box
    Ls = gpu.Ligand(L.xs * scale, L.ys * scale, L.zs * scale, lig.q, lig.anum, typ=lig.typ)
    R, t = workloads.random_poses_in_sphere(200, hi / 2, 0.8, seed=23)
    X, Y, Z = orc.pose_coords(L.xs * scale, L.ys * scale, L.zs * scale, R, t)
    assert X.min() > 0 and (X.max(), Y.max(), Z.max()) < tuple(hi)
    want = orc.ene_inter_interp(0.375, dims, maps, lig.typ, X, Y, Z)
    assert np.array_equal(gpu.Mol.ene_inter_UFF_interp(g, Ls, X, Y, Z), want)
    assert np.array_equal(gpu.Mol.interp_poses(g, Ls, R, t), want)


def test_trilinear_known_answer_on_device(gpu):
    dims = (8, 8, 8)
    arr = np.zeros((1, 512), np.float32)
    idx = lambda i, j, k: i + j * 8 + k * 64
    for v, (i, j, k) in zip(range(1, 9), [(2, 3, 4), (3, 3, 4), (3, 4, 4), (2, 4, 4), (2, 3, 5), (3, 3, 5), (3, 4, 5), (2, 4, 5)]):
        arr[0, idx(i, j, k)] = v
    g = gpu.G3D.upload(0.5, dims, arr)
    assert gpu.G3D.trilin(g, 0, [1.1], [1.7], [2.3])[0] == pytest.approx(4.639999999999999, rel=1e-15)


def test_ba1_cache_files_roundtrip(gpu, small, tmp_path):
    rec_m, dims, lig, ta, tq = small
    rec = gpu.Receptor.from_mol(rec_m)
    g, maps = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta[:3], tq[:3])
    paths = []
    for t in range(3):
        p = str(tmp_path / f"lig.t{t}.ba1")
        g.write_ba1(t, p)
        paths.append(p)
        raw = np.fromfile(p, "<f4")                 # G3D.to_ba1_file: raw little-endian f32, x fastest
        assert np.array_equal(raw, maps[t])
    txt = open(paths[0] + ".dims").read().split("\n")
    assert txt[:4] == ["step: 0.375", f"x_dim: {dims[0]}", f"y_dim: {dims[1]}", f"z_dim: {dims[2]}"]
    g2 = gpu.G3D.of_ba1_files(paths)
    assert np.array_equal(g2.download(), maps)


def test_vdw_mask_and_clash_bit_identical(gpu, orc, c2, c2_roi_rec):
    dims = gpu.Grid.from_box(workloads.GRID_STEP, *c2["sim_dims"])
    m = c2["rec"]
    mask = gpu.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, workloads.GRID_STEP, dims)
    want = orc.vdw_volume(m.xs, m.ys, m.zs, m.r, workloads.GRID_STEP, dims)
    assert np.array_equal(mask.bits, want[:len(mask.bits)])
    assert 0 < np.unpackbits(mask.bits).sum() < 0.2 * mask.bits.size * 8
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    R, t = workloads.random_poses_in_sphere(400, c2["roi"][:3], 9.0, seed=24)
    t[::3] += np.array([0.0, 0.0, 34.0])           # a third of the poses out in the solvent: no clash
    t[1::3] += np.array([0.0, 0.0, 24.0])          # and a third grazing the surface
    got = gpu.Mol.protein_ligand_clash(mask, lig, R, t)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    ref = np.array([orc.protein_ligand_clash(workloads.GRID_STEP, dims, want, X[p], Y[p], Z[p]) for p in range(len(R))])
    assert np.array_equal(got, ref)
    assert 0 < ref.sum() < len(ref)


@pytest.mark.gpu
def test_clash_poses_partly_outside_the_mask_box(gpu, orc, c2):
    """poses whose atoms leave the mask box (negative coordinates, beyond the last voxel): Bitv.get would raise in the
    reference; kernel, host prefilter and oracle agree that outside voxels read as not occupied -- no out-of-bounds
    read, same flags"""
    dims = gpu.Grid.from_box(workloads.GRID_STEP, *c2["sim_dims"])
    m = c2["rec"]
    mask = gpu.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, workloads.GRID_STEP, dims)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    rng = np.random.default_rng(77)
    R, _ = workloads.random_poses_in_sphere(600, (0, 0, 0), 1.0, seed=25)
    ext = np.array(dims) * workloads.GRID_STEP
    t = rng.uniform(-30.0, 1.0, (600, 3)) * (rng.random((600, 3)) < 0.5) + rng.uniform(0.0, 1.0, (600, 3)) * ext
    t[::4] = rng.uniform(-1e4, 1e4, (150, 3))                   # far outside, both signs
    t[1::4] = ext + rng.uniform(-3.0, 20.0, (150, 3))           # around the upper corner
    got = gpu.Mol.protein_ligand_clash(mask, lig, R, t)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    ref = np.array([orc.protein_ligand_clash(workloads.GRID_STEP, dims, mask.bits, X[p], Y[p], Z[p]) for p in range(len(R))])
    assert np.array_equal(got, ref)
    assert not got[::4].any()
    # the scan's host-side centre prefilter (vdW_clash_AND) with a ROI that pokes out of the mask box
    rot = gpu.SO3.rotations(4)
    roi = (1.0, 1.0, 1.0, 3.0)
    res = gpu.Lds.exhaustive_rigid_ligand_docking(3, roi, 1.0, rot, lig, rec=gpu.Receptor.from_mol(workloads.carve(m, roi[:3], 30.0)),
                                                  vdw_mask=mask)
    assert res["n_scored"] >= 0


def test_n3_masks_bit_identical(gpu, orc, c2):
    """first solvent shell, whole-protein and ROI-only bitmasks (lds.ml:97-145, 172-184, 269-305) on the device
    against the oracle's literal loops; 1 A grid over the simulation box keeps the CPU side short"""
    m = c2["rec"]
    dims = gpu.Grid.from_box(1.0, *c2["sim_dims"])
    got = gpu.Lds.first_solvent_shell(m.xs, m.ys, m.zs, m.r, 1.0, dims)
    want = orc.first_solvent_shell(m.xs, m.ys, m.zs, m.r, 1.0, dims)
    assert np.array_equal(got.bits, want) and 0 < np.unpackbits(want).sum()
    vdw = gpu.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, 1.0, dims)
    assert not (got.bits & vdw.bits).any()                 # the shell lies outside every vdW sphere
    sel = slice(0, 400)                                    # brute-force nearest neighbour on the CPU: a part of the protein
    gotw = gpu.Lds.bitmask_whole_protein(m.xs[sel], m.ys[sel], m.zs[sel], 1.0, dims)
    wantw = orc.bitmask_whole_protein(m.xs[sel], m.ys[sel], m.zs[sel], 1.0, dims)
    assert np.array_equal(gotw.bits, wantw) and np.unpackbits(wantw).sum() > 1000
    gotr = gpu.Lds.bitmask_ROI_only(c2["roi"], 1.0, dims)
    wantr = orc.bitmask_sphere(1.0, dims, c2["roi"][:3], c2["roi"][3] + 24.0)
    assert np.array_equal(gotr.bits, wantr)


def test_c3_full_size_grid_and_lookup(gpu, orc):
    """BASELINE configs[2] at full size: 22 maps of 81^3 voxels (0.375 A over 30 A) from a 5000-atom receptor,
    then a million interpolated poses.  The oracle builds the same maps on a random 4000-voxel bitmask (seconds on
    the CPU); the full device build must carry exactly those values there, and the lookup must reproduce the
    oracle's trilinear sums bit for bit on a sample of the poses."""
    rec_m = workloads.synthetic_receptor(5000, "cube", 60.0, seed=workloads.SEED)
    rec_m.xs -= 15.0; rec_m.ys -= 15.0; rec_m.zs -= 15.0      # the 30 A box centred in the receptor starts at the origin
    lig_m = pqrs.read_ligands_pqrs(os.path.join(workloads.GOLDEN, "ligdecs.pqrs"))[0]
    ta, tq = pqrs.assign_ff_types([lig_m])
    dims = gpu.Grid.from_box(0.375, 30.0, 30.0, 30.0)
    assert tuple(dims) == (81, 81, 81) and len(ta) == 22
    rec = gpu.Receptor.from_mol(rec_m)
    g, maps = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta, tq)
    nvox = 81 ** 3
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(nvox, 4000, replace=False))
    bits = np.zeros(nvox, np.uint8); bits[pick] = 1
    mask = np.packbits(bits, bitorder="little")
    want = orc.grid_build(rec_m, 0.375, dims, ta, tq, mask=mask)
    assert np.array_equal(maps[:, pick], want[:, pick])
    assert (maps[:, pick] != 0.0).all() and (maps == np.float32(1e5)).any()
    # a million rigid poses whose atoms stay inside the grid; 300 of them through the oracle
    lig = gpu.Ligand.from_mol(lig_m, centered=True)
    rl = workloads.lig_radius((lig.xs, lig.ys, lig.zs))
    n = 1_000_000
    R = workloads.random_rotations(n, rng)
    t = rng.uniform(rl + 0.4, 30.0 - rl - 0.4, (n, 3))
    e = gpu.Mol.interp_poses(g, lig, R, t)
    assert np.isfinite(e).all()
    sel = rng.choice(n, 300, replace=False)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R[sel], t[sel])
    assert np.array_equal(e[sel], orc.ene_inter_interp(0.375, dims, maps, lig_m.typ, X, Y, Z))
    # atoms that sit exactly on lattice nodes read the stored value: single-atom "ligands" of three types
    for typ in (0, 9, 21):
        one = gpu.Ligand([0.0], [0.0], [0.0], [float(tq[typ])], [int(ta[typ])], typ=np.array([typ], np.int32))
        nodes = rng.integers(1, 80, (64, 3))
        tt = nodes * 0.375
        idx = nodes[:, 0] + nodes[:, 1] * 81 + nodes[:, 2] * 81 * 81
        got = gpu.Mol.interp_poses(g, one, np.tile(np.eye(3).reshape(9), (64, 1)), tt)
        assert np.array_equal(got, maps[typ, idx].astype(np.float64))


def test_bitmask_text_file_round_trip(gpu, c2, tmp_path):
    """N1: `<rec>.bitmask` (Utls.bitmask_to_file, utls.ml:12-20): one line of n '0'/'1' characters, Bitv.M order
    (first character = bit n-1; unpinned library), read back to the same bits"""
    import ctypes as C
    L = gpu.lib()
    dims = gpu.Grid.from_box(2.0, *c2["sim_dims"])
    m = c2["rec"]
    mask = gpu.Lds.vdW_volume(m.xs, m.ys, m.zs, m.r, 2.0, dims)
    n = dims[0] * dims[1] * dims[2]
    want = np.unpackbits(mask.bits, bitorder="little")[:n]
    fn = str(tmp_path / "rec.pqrs.bitmask")
    for msb in (1, 0):
        assert L.mmo_mask_write_bitmask(mask.h, fn.encode(), C.c_int(msb)) == 0, L.mmo_last_error()
        txt = open(fn).read()
        assert txt.endswith("\n") and len(txt) == n + 1 and set(txt[:-1]) <= {"0", "1"}
        chars = np.frombuffer(txt[:-1].encode(), np.uint8) - ord("0")
        assert np.array_equal(chars[::-1] if msb else chars, want)
        h = C.c_void_p()
        assert L.mmo_mask_read_bitmask(fn.encode(), C.c_double(2.0), (C.c_int32 * 3)(*dims), C.c_int(msb), C.byref(h)) == 0, L.mmo_last_error()
        got = np.zeros_like(mask.bits)
        assert L.mmo_mask_download(h, got.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        L.mmo_mask_destroy(h)
        assert np.array_equal(got, mask.bits)
    # wrong length / wrong character are refused (Bitv.M.of_string raises)
    open(fn, "w").write("0101\n")
    h = C.c_void_p()
    assert L.mmo_mask_read_bitmask(fn.encode(), C.c_double(2.0), (C.c_int32 * 3)(*dims), C.c_int(1), C.byref(h)) != 0


def test_ba1_zst_cache_through_the_zstd_program(gpu, small, tmp_path, monkeypatch):
    """N1: `.ba1.zst` (utls.ml:22-45, lds.ml:516-553).  The reference shells out to `zstd`; this image has none, so a
    stand-in script with the same command line (--rm -qf / -dqfk) is put on PATH: the test pins the commands the library
    issues and the uncompress-read-remove sequence, not the compression format"""
    import ctypes as C
    import shutil
    import stat
    L = gpu.lib()
    rec_m, dims, lig, ta, tq = small
    rec = gpu.Receptor.from_mol(rec_m)
    g, maps = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.375, dims, ta[:2], tq[:2])
    paths = [str(tmp_path / f"lig.mol2.t{t}.ba1") for t in range(2)]
    if shutil.which("zstd") is None:
        h = C.c_void_p()
        g.write_ba1(0, paths[0])
        assert L.mmo_zstd_compress_file(paths[0].encode()) != 0 and b"zstd" in L.mmo_last_error()
        fake = tmp_path / "bin"
        fake.mkdir()
        (fake / "zstd").write_text('#!/bin/sh\n'
                                   'case "$1" in\n'
                                   '  --rm) [ "$2" = "-qf" ] || exit 2; cp "$3" "$3.zst" && rm "$3";;\n'
                                   '  -dqfk) cp "$2" "${2%.zst}";;\n'
                                   '  *) exit 2;;\n'
                                   'esac\n')
        os.chmod(fake / "zstd", os.stat(fake / "zstd").st_mode | stat.S_IEXEC)
        monkeypatch.setenv("PATH", str(fake) + os.pathsep + os.environ["PATH"])
    for t in range(2):
        g.write_ba1(t, paths[t])
        assert L.mmo_zstd_compress_file(paths[t].encode()) == 0, L.mmo_last_error()
        assert not os.path.exists(paths[t]) and os.path.exists(paths[t] + ".zst") and os.path.exists(paths[t] + ".dims")
    g2 = gpu.G3D.of_ba1_files([p + ".zst" for p in paths])
    assert np.array_equal(g2.download(), maps)
    assert not os.path.exists(paths[0])                     # the transient uncompressed copy is gone (lds.ml:551-552)
    assert L.mmo_zstd_uncompress_file(b"/tmp/x; rm -rf y.zst") != 0      # nothing but plain path characters reaches the shell
