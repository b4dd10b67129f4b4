"""N4: the Majeux-Scarsi-Caflisch desolvation sums (Lds.protein_desolv, src/lds.ml:204-236; Lds.desolvation_penalty,
src/lds.ml:239-267).  CPU: the oracle's restatement against a hand-derived case.  GPU: the device kernels against
the oracle, bit for bit (same doubles, same summation order)."""
import math

import numpy as np
import pytest

from mmo_b200 import workloads

K_DESOLV = (1.0 / 4.0 - 1.0 / 78.5) / (8.0 * (4.0 * math.atan(1.0)))      # src/const.ml:17-31


class _Rec:
    def __init__(self, xs, ys, zs, q):
        self.xs, self.ys, self.zs, self.q = (np.asarray(a, np.float64) for a in (xs, ys, zs, q))
        self.n = len(self.xs)


def test_oracle_desolvation_known_answers(orc):
    """one protein atom (q = 0.5) 3 A from the only shell voxel; one ligand atom (q = -0.4, r = 1.7) 2.2 A from it:
    contribution = K * 1 A^3 * (0.5/9)^2, ligand term = K * 1 A^3 * (0.4/4.84)^2"""
    dims = (12, 12, 12)
    nvox = 12 ** 3
    idx = 8 + 5 * 12 + 5 * 144
    bits = np.zeros(nvox, np.uint8); bits[idx] = 1
    far = 2 + 2 * 12 + 2 * 144                                            # a second shell voxel, outside the ROI
    bits[far] = 1
    shell = np.packbits(bits, bitorder="little")
    rec = _Rec([5.0], [5.0], [5.0], [0.5])
    res = orc.protein_desolv(rec, 1.0, dims, shell, (6.0, 5.0, 5.0, 5.0))
    assert res[idx] == K_DESOLV * (1.0 * ((0.5 / 9.0) * (0.5 / 9.0)))
    assert res[far] == 0.0 and np.count_nonzero(res) == 1                 # ROI.is_inside is strict and required
    # 12.5 A away: outside Const.charged_cutoff, nothing to add
    res_far = orc.protein_desolv(_Rec([8.0 - 12.5], [5.0], [5.0], [0.5]), 1.0, dims, shell, (6.0, 5.0, 5.0, 5.0))
    assert not res_far.any()
    prot, lig = orc.desolvation_penalty(1.0, dims, shell, res, [10.2], [5.0], [5.0], [-0.4], [1.7])
    assert prot == res[idx]
    d2 = (8.0 - 10.2) * (8.0 - 10.2)
    assert lig == K_DESOLV * (1.0 * ((-0.4 / d2) * (-0.4 / d2)))
    # the ligand's vdW sphere covers the voxel: no longer part of its solvent shell, nothing is desolvated
    prot, lig = orc.desolvation_penalty(1.0, dims, shell, res, [9.0], [5.0], [5.0], [-0.4], [1.7])
    assert prot == 0.0 and lig == 0.0
    # too far for the shell (r + 1.4 = 3.1 A): nothing either
    prot, lig = orc.desolvation_penalty(1.0, dims, shell, res, [11.2], [5.0], [5.0], [-0.4], [1.7])
    assert prot == 0.0 and lig == 0.0


@pytest.fixture(scope="module")
def shell_setup(gpu, orc, c2):
    m = c2["rec"]
    step = 1.0
    dims = gpu.Grid.from_box(step, *c2["sim_dims"])
    shell = gpu.Lds.first_solvent_shell(m.xs, m.ys, m.zs, m.r, step, dims)
    rec = gpu.Receptor.from_mol(m)
    return m, rec, step, dims, shell


@pytest.mark.gpu
def test_protein_desolv_bit_identical(gpu, orc, c2, shell_setup):
    m, rec, step, dims, shell = shell_setup
    h, got = gpu.Lds.protein_desolv(c2["roi"], rec, shell)
    want = orc.protein_desolv(m, step, dims, shell.bits, c2["roi"])
    assert np.array_equal(got, want)
    nz = np.count_nonzero(want)
    assert 50 < nz < np.unpackbits(shell.bits).sum()                      # the ROI keeps a part of the shell only
    # an ROI that holds no shell voxel: all zeros, no launch of the sum kernel
    h0, got0 = gpu.Lds.protein_desolv((1.0, 1.0, 1.0, 0.5), rec, shell)
    assert not got0.any()


@pytest.mark.gpu
def test_desolvation_penalty_bit_identical(gpu, orc, c2, shell_setup):
    m, rec, step, dims, shell = shell_setup
    lm = c2["lig"]
    lig = gpu.Ligand.from_mol(lm, centered=True)
    h, contribs = gpu.Lds.protein_desolv(c2["roi"], rec, shell)
    R, t = workloads.random_poses_in_sphere(24, c2["roi"][:3], 8.0, seed=11)
    R[0] = np.eye(3).reshape(9); t[0] = c2["start_pos"]                   # the docked pose itself
    t[1] = [3.0, 3.0, 3.0]                                                # far from the protein: nothing desolvated
    gp, gl = gpu.Lds.desolvation_penalty(h, lig, rot9=R, trans3=t)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    want = [orc.desolvation_penalty(step, dims, shell.bits, contribs, X[p], Y[p], Z[p], lm.q, lm.r) for p in range(len(R))]
    wp, wl = np.array([w[0] for w in want]), np.array([w[1] for w in want])
    assert np.array_equal(gp, wp) and np.array_equal(gl, wl)
    assert gp[0] > 0.0 and gl[0] > 0.0 and gp[1] == 0.0 and gl[1] == 0.0
    # explicit coordinates take the same path
    cp, cl = gpu.Lds.desolvation_penalty(h, lig, xs=X, ys=Y, zs=Z)
    assert np.array_equal(cp, wp) and np.array_equal(cl, wl)


@pytest.mark.gpu
def test_desolvation_on_the_default_half_angstrom_grid(gpu, orc, c2):
    """the reference's own grid (0.5 A over the simulation box, 16 M voxels): device only, checked through
    properties -- contributions live on shell voxels inside the ROI, and the docked pose's penalty is the sum over
    its desolvated voxels computed here with numpy from the device's own per-voxel array"""
    m = c2["rec"]
    step = 0.5
    dims = gpu.Grid.from_box(step, *c2["sim_dims"])
    shell = gpu.Lds.first_solvent_shell(m.xs, m.ys, m.zs, m.r, step, dims)
    rec = gpu.Receptor.from_mol(m)
    h, contribs = gpu.Lds.protein_desolv(c2["roi"], rec, shell)
    nvox = dims[0] * dims[1] * dims[2]
    bits = np.unpackbits(shell.bits, bitorder="little")[:nvox].astype(bool)
    assert not contribs[~bits].any() and (contribs >= 0.0).all() and contribs.any()
    idx = np.nonzero(contribs)[0]
    k = idx // (dims[0] * dims[1]); j = (idx - k * dims[0] * dims[1]) // dims[0]; i = idx - k * dims[0] * dims[1] - j * dims[0]
    d2 = (i * step - c2["roi"][0]) ** 2 + (j * step - c2["roi"][1]) ** 2 + (k * step - c2["roi"][2]) ** 2
    assert (d2 < c2["roi"][3] ** 2 + 1e-9).all()
    lm = c2["lig"]
    lig = gpu.Ligand.from_mol(lm, centered=True)
    R = np.eye(3).reshape(1, 9); t = np.array([c2["start_pos"]])
    gp, gl = gpu.Lds.desolvation_penalty(h, lig, rot9=R, trans3=t)
    X, Y, Z = orc.pose_coords(lig.xs, lig.ys, lig.zs, R, t)
    lshell = gpu.Lds.first_solvent_shell(X[0], Y[0], Z[0], lm.r, step, dims)
    both = np.unpackbits(shell.bits & lshell.bits, bitorder="little")[:nvox].astype(bool)
    vox = np.nonzero(both)[0]
    assert len(vox) > 100
    prot = 0.0
    for v in vox:
        prot += contribs[v]
    assert gp[0] == prot
    assert gl[0] > 0.0
