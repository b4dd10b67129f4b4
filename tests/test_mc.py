"""Monte-Carlo chains (lds.ml:741-1000): oracle behaviours on the CPU, and CUDA chains against the
oracle frame by frame.  Both sides use include/mmo_detmath.h for the random stream and sin/cos/exp,
so trajectories -- not just distributions -- must agree bit for bit."""
import numpy as np
import pytest

from mmo_b200 import pqrs, workloads


def _pocket_grid(orc, c2, rec_m, step=1.0, reach=21.0):
    c = np.array(c2["roi"][:3])
    dims = orc.grid_from_box(step, *(c + reach + 2.0))
    mask = orc.bitmask_sphere(step, dims, c, reach)
    ta, tq = pqrs.assign_ff_types([c2["lig"]])
    return dims, mask, ta, tq


@pytest.fixture(scope="module")
def mc_setup(orc, c2, c2_roi_rec):
    dims, mask, ta, tq = _pocket_grid(orc, c2, c2_roi_rec)
    maps = orc.grid_build(c2_roi_rec, 1.0, dims, ta, tq, mask=mask)
    return dims, mask, ta, tq, maps


def test_oracle_chain_mirrors_the_dangling_else(orc, c2, mc_setup):
    dims, mask, ta, tq, maps = mc_setup
    cx, cy, cz = c2["centered"]
    rot0 = np.eye(3).reshape(9)
    # without --hard-ROI nothing is ever accepted or rejected
    r, xyz, tr = orc.mc_run(c2["lig"], cx, cy, cz, c2["roi"], 300, 1234, rot0, c2["start_pos"], maps=maps, g_step=1.0,
                            g_dims=dims, hard_roi=False)
    assert r["n_accept_rigid"] + r["n_reject_rigid"] + r["n_accept_conf"] + r["n_reject_conf"] == 0
    assert (tr[:, 3] == -1).all() and r["frames_done"] == 300
    # with --hard-ROI the loop is a Metropolis chain
    r, xyz, tr = orc.mc_run(c2["lig"], cx, cy, cz, c2["roi"], 2000, 1234, rot0, c2["start_pos"], maps=maps, g_step=1.0,
                            g_dims=dims, hard_roi=True)
    tested = r["n_accept_rigid"] + r["n_reject_rigid"] + r["n_accept_conf"] + r["n_reject_conf"]
    assert tested == r["frames_done"] - r["n_ooroi"] - r["n_ezero"]
    assert 0 < r["n_accept_rigid"] and 0 < r["n_reject_rigid"]
    assert r["best_E"] <= tr[0, 0] + 1e-9 or r["n_ooroi"] + r["n_ezero"] > 0
    # D3: best is the minimum over every tested trial since the last reset
    if r["n_ooroi"] + r["n_ezero"] == 0:
        start_E = orc.ene_inter_interp(1.0, dims, maps, c2["lig"].typ, *[np.atleast_2d(v) for v in xyz])  # noqa: F841
        assert r["best_E"] == min(tr[tr[:, 3] >= 0, 0].min(), r["best_E"])


def test_oracle_rigid_ligand_keeps_intra_constant(orc, c2, mc_setup):
    dims, mask, ta, tq, maps = mc_setup
    cx, cy, cz = c2["centered"]
    r, xyz, tr = orc.mc_run(c2["lig"], cx, cy, cz, c2["roi"], 200, 7, np.eye(3).reshape(9), c2["start_pos"], maps=maps,
                            g_step=1.0, g_dims=dims, tweak_rbonds=False)
    e0 = orc.ene_intra(c2["lig"], cx, cy, cz)[0]
    assert (tr[:, 2] == e0).all()        # lds.ml:706-712: const_ene_intra of the centred ligand


def test_rng_and_detmath_are_platform_independent(orc):
    import ctypes as C
    f = orc.lib().orc_rng_uniform
    assert f(C.c_uint64(1234), C.c_uint64(0)) == 0.73066652454062397
    assert f(C.c_uint64(1234), C.c_uint64(1)) == 0.59288985801498617


@pytest.mark.gpu
@pytest.mark.parametrize("threads", ["128", "32", "64", "256"])      # block per chain (default), warp per chain, other block sizes
@pytest.mark.parametrize("flags", [dict(), dict(tweak_rbonds=False), dict(no_flip=True, temperature_K=600.0),
                                   dict(hard_roi=False), dict(intra_nb=False)])
def test_cuda_chains_match_the_oracle_bit_for_bit(gpu, orc, c2, c2_roi_rec, mc_setup, flags, threads, monkeypatch):
    monkeypatch.setenv("MMO_MC_THREADS", threads)
    dims, mask, ta, tq, maps = mc_setup
    rec = gpu.Receptor.from_mol(c2_roi_rec)
    g, gmaps = gpu.Lds.pre_calculate_FF_components_grid(rec, 1.0, dims, ta, tq, mask_bits=mask)
    assert np.array_equal(gmaps, maps)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    n_steps, seeds = 1200, np.array([1234, 99, 20231017, 5], np.uint64)
    R, t = workloads.random_poses_in_sphere(len(seeds), c2["roi"][:3], 4.0, seed=31)
    R[0] = np.eye(3).reshape(9); t[0] = c2["start_pos"]
    res, xyz, trace = gpu.Lds.simulate_lig(g, lig, c2["roi"], n_steps, seeds, R, t, want_xyz=True, want_trace=True, **flags)
    for c, seed in enumerate(seeds):
        want, wxyz, wtr = orc.mc_run(c2["lig"], lig.xs, lig.ys, lig.zs, c2["roi"], n_steps, int(seed), R[c], t[c],
                                     maps=maps, g_step=1.0, g_dims=dims, **flags)
        got = res[c]
        if c == 0:
            assert np.array_equal(trace[:want["frames_done"]], wtr)
        for k in ("best_E", "prev_E", "max_rot", "max_trans", "n_accept_rigid", "n_reject_rigid", "n_accept_conf",
                  "n_reject_conf", "n_ooroi", "n_ezero", "too_long", "frames_done"):
            assert got[k] == want[k], (c, k, got[k], want[k])
        assert np.array_equal(got["best_rot"], want["best_rot"]) and np.array_equal(got["best_pos"], want["best_pos"])
        assert np.array_equal(xyz[c], wxyz)


@pytest.mark.gpu
def test_many_chains_statistics(gpu, orc, c2, c2_roi_rec, mc_setup):
    """512 chains: every chain is the oracle's chain for its seed (spot-checked); chains that start from
    random (clashing) poses reject most moves, so the adaptive scheme (lds.ml:586-600) must have shrunk
    their step sizes below the params.ml defaults."""
    dims, mask, ta, tq, maps = mc_setup
    g = gpu.G3D.upload(1.0, dims, maps)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    n = 512
    seeds = np.arange(n, dtype=np.uint64) + 20231017
    R, t = workloads.random_poses_in_sphere(n, c2["roi"][:3], 3.0, seed=32)
    res, _, _ = gpu.Lds.simulate_lig(g, lig, c2["roi"], 4000, seeds, R, t)
    for c in (0, 17, 511):
        want, _, _ = orc.mc_run(c2["lig"], lig.xs, lig.ys, lig.zs, c2["roi"], 4000, int(seeds[c]), R[c], t[c], maps=maps,
                                g_step=1.0, g_dims=dims)
        assert res[c]["best_E"] == want["best_E"] and res[c]["n_accept_rigid"] == want["n_accept_rigid"]
    ar = np.array([r["n_accept_rigid"] / max(1, r["n_accept_rigid"] + r["n_reject_rigid"]) for r in res])
    assert 0.0 < np.median(ar) < 0.45
    shrunk = np.array([r["max_rot"] < np.radians(15.0) and r["max_trans"] < 0.15 for r in res])
    assert shrunk.mean() > 0.9
    assert all(r["frames_done"] == 4000 or r["too_long"] for r in res)


@pytest.mark.gpu
def test_cuda_chain_on_the_direct_scorer_follows_the_oracle(gpu, orc, c2, c2_roi_rec, monkeypatch):
    """--no-interp: E_inter = Mol.ene_inter_UFF_shifted_brute (mol.ml:822-849).  The kernel uses the
    reference's fp64 pair terms with a lane-strided summation, so energies agree to ~1e-13 relative and
    the Metropolis decisions -- hence the trajectory -- are the oracle's."""
    rec = gpu.Receptor.from_mol(c2_roi_rec)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    n_steps, seeds = 400, np.array([4242, 7], np.uint64)
    R = np.tile(np.eye(3).reshape(9), (2, 1)); t = np.tile(c2["start_pos"], (2, 1))
    t[1] += 0.3
    res, xyz, trace = gpu.Lds.simulate_lig(None, lig, c2["roi"], n_steps, seeds, R, t, want_xyz=True, want_trace=True, rec=rec)
    monkeypatch.setenv("MMO_MC_THREADS", "32")         # the warp-per-chain kernel runs the same arithmetic
    res32, xyz32, trace32 = gpu.Lds.simulate_lig(None, lig, c2["roi"], n_steps, seeds, R, t, want_xyz=True, want_trace=True, rec=rec)
    assert np.array_equal(trace, trace32) and np.array_equal(xyz, xyz32)
    for c in range(2):
        want, wxyz, wtr = orc.mc_run(c2["lig"], lig.xs, lig.ys, lig.zs, c2["roi"], n_steps, int(seeds[c]), R[c], t[c],
                                     rec=c2_roi_rec)
        if c == 0:
            assert np.array_equal(trace[:, 3], wtr[:, 3])                       # same accept/reject sequence
            assert np.allclose(trace[:, :3], wtr[:, :3], rtol=1e-10, atol=1e-9)
        assert res[c]["n_accept_rigid"] == want["n_accept_rigid"] and res[c]["n_accept_conf"] == want["n_accept_conf"]
        assert res[c]["best_E"] == pytest.approx(want["best_E"], rel=1e-10, abs=1e-9)
        assert np.allclose(xyz[c], wxyz, rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_c4_full_size_chains_shard_and_match(gpu, orc, c2, c2_roi_rec, mc_setup):
    """BASELINE configs[3] at full size: 4096 chains x 10 000 frames (interpolated E_inter + intra NB, --hard-ROI).
    Chains are independent (one RNG per chain, lds.ml:1997-2000), so (1) an eighth of them run alone -- one GPU's
    share when the job is sharded over 8 -- must reproduce its slice of the full launch bit for bit, and (2) any
    chain must be the oracle's chain for its seed (three spot checks, 10 000 frames each on the CPU)."""
    dims, mask, ta, tq, maps = mc_setup
    g = gpu.G3D.upload(1.0, dims, maps)
    lig = gpu.Ligand.from_mol(c2["lig"], centered=True)
    n, steps = 4096, 10_000
    seeds = np.arange(n, dtype=np.uint64) + 20231017
    R, t = workloads.random_poses_in_sphere(n, c2["roi"][:3], 3.0, seed=33)
    res, _, _ = gpu.Lds.simulate_lig(g, lig, c2["roi"], steps, seeds, R, t)
    assert all(r["frames_done"] == steps or r["too_long"] for r in res)
    lo, hi = 5 * 512, 6 * 512                                   # rank 5 of 8
    part, _, _ = gpu.Lds.simulate_lig(g, lig, c2["roi"], steps, seeds[lo:hi], R[lo:hi], t[lo:hi])
    keys = ("best_E", "prev_E", "max_rot", "max_trans", "n_accept_rigid", "n_reject_rigid", "n_accept_conf",
            "n_reject_conf", "n_ooroi", "n_ezero", "too_long", "frames_done")
    for a, b in zip(part, res[lo:hi]):
        assert all(a[k] == b[k] for k in keys)
        assert np.array_equal(a["best_rot"], b["best_rot"]) and np.array_equal(a["best_pos"], b["best_pos"])
    for c in (3, 2049, 4095):
        want, _, _ = orc.mc_run(c2["lig"], lig.xs, lig.ys, lig.zs, c2["roi"], steps, int(seeds[c]), R[c], t[c], maps=maps,
                                g_step=1.0, g_dims=dims)
        assert all(res[c][k] == want[k] for k in keys), c
    best = np.array([r["best_E"] for r in res])
    assert np.isfinite(best).all() and best.min() < np.median(best)


@pytest.mark.gpu
def test_place_ligand_in_roi_properties(gpu, orc, c2):
    """Lds.place_ligand_in_ROI (lds.ml:308-345): every start lies strictly inside the ROI, no ligand heavy atom comes
    closer than 0.8 (r_i + r_j) to a protein heavy atom (Mol.heavy_atom_clash), the draws are a pure function of the
    seed, and without the clash test the poses are the first in-ROI draws of the same stream"""
    import mmo_b200
    m, lm = c2["rec"], c2["lig"]
    lig = gpu.Ligand.from_mol(lm, centered=True)
    # a ROI at the protein surface: some random poses touch the protein and are rejected, others pass
    roi = (float(m.xs.max()) + 5.0, float(np.median(m.ys)), float(np.median(m.zs)), 8.0)
    rot, pos, trials = mmo_b200.place_ligand_in_ROI(m, lig, roi, 1234, 12)
    rot2, pos2, trials2 = mmo_b200.place_ligand_in_ROI(m, lig, roi, 1234, 12)
    assert np.array_equal(rot, rot2) and np.array_equal(pos, pos2) and trials == trials2 >= 12
    c, r = np.array(roi[:3]), roi[3]
    assert (((pos - c) ** 2).sum(1) < r * r).all()
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, rot, pos)
    heavy_l, heavy_p = lm.anum > 1, m.anum > 1
    lim = 0.8 * (lm.r[heavy_l][:, None] + m.r[heavy_p][None, :])
    for p in range(12):
        d2 = ((X[p][heavy_l][:, None] - m.xs[heavy_p][None, :]) ** 2 + (Y[p][heavy_l][:, None] - m.ys[heavy_p][None, :]) ** 2 +
              (Z[p][heavy_l][:, None] - m.zs[heavy_p][None, :]) ** 2)
        assert (d2 >= lim * lim).all()
        R = rot[p].reshape(3, 3)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1.0) < 1e-12
    # without the rejection every in-ROI draw is taken: fewer trials, and the first pose is the first in-ROI point
    rot0, pos0, trials0 = mmo_b200.place_ligand_in_ROI(m, lig, roi, 1234, 12, clash_check=False)
    assert 12 <= trials0 < trials
    u = [orc.lib().orc_rng_uniform for _ in range(1)][0]
    import ctypes as C
    u.restype = C.c_double
    ctr, first = 0, None
    while first is None:
        z = (c[2] - r) + u(C.c_uint64(1234), C.c_uint64(ctr)) * ((c[2] + r) - (c[2] - r))
        y = (c[1] - r) + u(C.c_uint64(1234), C.c_uint64(ctr + 1)) * ((c[1] + r) - (c[1] - r))
        x = (c[0] - r) + u(C.c_uint64(1234), C.c_uint64(ctr + 2)) * ((c[0] + r) - (c[0] - r))
        ctr += 3
        if (c[0] - x) ** 2 + (c[1] - y) ** 2 + (c[2] - z) ** 2 < r * r:
            first = (x, y, z)
    assert tuple(pos0[0]) == first
    # the buried 3A2J pocket itself cannot be populated by blind rigid placement when the clash test is on: the
    # reference's message after 100 000 draws (with its nan radius the reference never gets there, see the header)
    with pytest.raises(RuntimeError, match="100k trials"):
        mmo_b200.place_ligand_in_ROI(m, lig, c2["roi"], 5, 1)
    _, pos_b, _ = mmo_b200.place_ligand_in_ROI(m, lig, c2["roi"], 5, 3, clash_check=False)
    assert (((pos_b - np.array(c2["roi"][:3])) ** 2).sum(1) < c2["roi"][3] ** 2).all()


@pytest.mark.gpu
def test_lds_mc_c_program_matches_the_python_path(gpu, orc, c2):
    """tools/lds_mc.c: `lds` in its Monte-Carlo mode as a plain C program on the C ABI, fed with the committed .pqrs /
    .bild files: same start poses, same chains, same best energies as the ctypes path on the same inputs"""
    import os
    import subprocess
    import mmo_b200
    exe = os.path.join(os.path.dirname(gpu.LIB_PATH), "lds_mc")
    assert os.path.exists(exe), "build it with `make -C mmo_b200/csrc tools` (done by __graft_entry__.build())"
    G = workloads.GOLDEN
    n_steps, starts, seed = 400, 5, 77
    out = subprocess.run([exe, "-lig", os.path.join(G, "docked.pqrs"), "-rec", os.path.join(G, "xtal_rec.pqrs"), "-roi",
                          os.path.join(G, "ROI.bild"), "-steps", str(n_steps), "-starts", str(starts), "-s", str(seed),
                          "--intra-NB", "--hard-ROI"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    rows = [l.split("\t") for l in lines if l and not l.startswith("#") and "\t" in l]
    assert len(rows) == starts and lines[-1].startswith(f"{n_steps} frames in ")
    # the same run through the ctypes mirror: maps on the 0.5 A simulation grid behind the ROI-only bitmask
    m, lm = c2["rec"], c2["lig"]
    dims = gpu.Grid.from_box(0.5, *c2["sim_dims"])
    mask = gpu.Lds.bitmask_ROI_only(c2["roi"], 0.5, dims)
    ta, tq = pqrs.assign_ff_types([lm])
    rec = gpu.Receptor.from_mol(m)
    g, _ = gpu.Lds.pre_calculate_FF_components_grid(rec, 0.5, dims, ta, tq, mask_bits=mask.bits, want_host=False)
    lig = gpu.Ligand.from_mol(lm, centered=True)
    rot, pos, _ = mmo_b200.place_ligand_in_ROI(m, lig, c2["roi"], seed, starts, clash_check=False)
    seeds = np.array([seed + 1 + s for s in range(starts)], np.uint64)
    res, _, _ = gpu.Lds.simulate_lig(g, lig, c2["roi"], n_steps, seeds, rot, pos)
    for s, row in enumerate(rows):
        assert int(row[1]) == s and float(row[2]) == res[s]["best_E"] and int(row[3]) == res[s]["frames_done"]
        assert int(row[4]) == res[s]["n_accept_rigid"] and int(row[7]) == res[s]["n_reject_conf"]
    # flag handling of the OCaml main (lds.ml:1759-1771): exactly one E_intra flag
    bad = subprocess.run([exe, "-lig", "a", "-rec", "b", "-roi", "c", "-steps", "10k"], capture_output=True, text=True)
    assert bad.returncode != 0 and "which lig_E_intra FF to use?" in bad.stderr
