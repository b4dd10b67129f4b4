"""Host-side set-up of the BASELINE.json workloads (deterministic, seed 20231017 by default).

Mirrors what `lds` does before it reaches the scoring closures (src/lds.ml:20-41, 1832-1946):
the protein is moved into a positive-octant simulation box with a 36 A margin, the ROI follows,
the ligand is centred.  Also the synthetic receptors / pose sets of SURVEY.md section 8(d).
numpy only; no energies are computed here.
"""
from __future__ import annotations

import os

import numpy as np

from . import pqrs

SEED = 20231017
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
MARGIN = 12.0 * 3.0          # src/lds.ml:1832  margin = charged_cutoff * 3
GRID_STEP = 0.5              # src/params.ml:22


def _favg(a):
    s = 0.0
    c = 0.0
    for v in a:
        y = float(v) - c
        t = s + y
        c = (t - s) - y
        s = t
    return s / len(a)


def preprocess_protein(rec: pqrs.Mol, margin: float = MARGIN):
    """src/lds.ml:20-41: returns (moved receptor, sim box dims, init_to_prod translation)."""
    rmax = float(rec.r.max())
    lo = np.array([rec.xs.min() - rmax, rec.ys.min() - rmax, rec.zs.min() - rmax])     # mol.ml:713-722
    hi = np.array([rec.xs.max() + rmax, rec.ys.max() + rmax, rec.zs.max() + rmax])
    box_dims = hi - lo
    sim_dims = box_dims + 2.0 * margin
    center_new = sim_dims * 0.5
    center_old = np.array([_favg(rec.xs), _favg(rec.ys), _favg(rec.zs)])
    delta = center_new - center_old               # translate_to: V3.diff p m.center (mol.ml:687-690)
    out = rec.copy()
    out.xs = rec.xs + delta[0]
    out.ys = rec.ys + delta[1]
    out.zs = rec.zs + delta[2]
    return out, sim_dims, delta


def carve(rec: pqrs.Mol, center, radius: float) -> pqrs.Mol:
    """Receptor atoms within `radius` of `center` (exact for the shifted scorers when
    radius >= R_roi + R_lig + 12, since w = 0 beyond 12 A; cf. src/scissors.ml)."""
    d2 = (rec.xs - center[0]) ** 2 + (rec.ys - center[1]) ** 2 + (rec.zs - center[2]) ** 2
    k = d2 < radius * radius
    return pqrs.Mol(rec.name + "_roi", rec.xs[k].copy(), rec.ys[k].copy(), rec.zs[k].copy(), rec.q[k].copy(),
                    rec.r[k].copy(), rec.anum[k].copy())


def lig_radius(lig_xyz_centered) -> float:
    x, y, z = lig_xyz_centered
    return float(0.01 + np.sqrt(x * x + y * y + z * z).max())      # mol.ml:576-583


def load_c2(ligand="docked"):
    """Config C1/C2 inputs: xtal receptor stand-in in the simulation box, ROI, centred ligand."""
    rec0 = pqrs.read_receptor_pqrs(os.path.join(GOLDEN, "xtal_rec.pqrs"))
    lig = pqrs.read_ligands_pqrs(os.path.join(GOLDEN, ligand + ".pqrs"))[0]
    pqrs.assign_ff_types([lig])
    roi0 = pqrs.read_roi_bild(os.path.join(GOLDEN, "ROI.bild"))
    rec, sim_dims, delta = preprocess_protein(rec0)
    roi = (roi0[0] + delta[0], roi0[1] + delta[1], roi0[2] + delta[2], roi0[3])
    c = [_favg(lig.xs), _favg(lig.ys), _favg(lig.zs)]
    centered = (lig.xs + (0.0 - c[0]), lig.ys + (0.0 - c[1]), lig.zs + (0.0 - c[2]))
    start_pos = np.array(c) + delta          # where the docked ligand sits in the simulation box
    return dict(rec=rec, rec_orig=rec0, lig=lig, centered=centered, roi=roi, sim_dims=sim_dims, delta=delta,
                start_pos=start_pos)


# ---------------------------------------------------------------------------------------------
_ELTS = np.array([1, 6, 7, 8, 16], np.int32)
_ELT_P = np.array([0.50, 0.32, 0.08, 0.09, 0.01])


def synthetic_receptor(n_atoms: int, shape: str, size: float, seed: int = SEED, min_sep: float = 1.0,
                       origin=(0.0, 0.0, 0.0)) -> pqrs.Mol:
    """Random receptor: uniform in a cube (edge `size`) or sphere (radius `size`) with minimum
    separation, elements {H 50, C 32, N 8, O 9, S 1}%, charges U[-0.8,0.8] shifted to net 0."""
    rng = np.random.default_rng(seed)
    pts = np.empty((0, 3))
    cell = min_sep
    occupied = {}
    out = []
    while len(out) < n_atoms:
        cand = rng.uniform(0.0, 1.0, (4 * n_atoms, 3))
        if shape == "cube":
            cand = cand * size
        else:
            cand = (cand * 2.0 - 1.0) * size
            cand = cand[(cand ** 2).sum(1) < size * size]
        for p in cand:
            key = tuple((p // cell).astype(int))
            ok = True
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        for qq in occupied.get((key[0] + dx, key[1] + dy, key[2] + dz), ()):
                            if ((p - qq) ** 2).sum() < min_sep * min_sep:
                                ok = False
            if ok:
                occupied.setdefault(key, []).append(p)
                out.append(p)
                if len(out) == n_atoms:
                    break
    pts = np.array(out) + np.array(origin)
    anum = rng.choice(_ELTS, n_atoms, p=_ELT_P).astype(np.int32)
    q = rng.uniform(-0.8, 0.8, n_atoms)
    q -= q.mean()
    rad = np.array([pqrs.VDW_RADII[int(a)] for a in anum])
    return pqrs.Mol(f"synth{n_atoms}", pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), q, rad, anum)


def random_rotations(n: int, rng) -> np.ndarray:
    """n proper rotation matrices (row-major 9) from normalised Gaussian quaternions."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 9))
    R[:, 0] = 1 - 2 * (y * y + z * z); R[:, 1] = 2 * (x * y - z * w); R[:, 2] = 2 * (x * z + y * w)
    R[:, 3] = 2 * (x * y + z * w); R[:, 4] = 1 - 2 * (x * x + z * z); R[:, 5] = 2 * (y * z - x * w)
    R[:, 6] = 2 * (x * z - y * w); R[:, 7] = 2 * (y * z + x * w); R[:, 8] = 1 - 2 * (x * x + y * y)
    return R


def random_poses_in_sphere(n: int, center, radius: float, seed: int = SEED):
    rng = np.random.default_rng(seed)
    R = random_rotations(n, rng)
    t = np.empty((0, 3))
    while len(t) < n:
        c = rng.uniform(-1, 1, (2 * n, 3))
        c = c[(c ** 2).sum(1) < 1.0]
        t = np.concatenate([t, c])
    t = t[:n] * radius + np.asarray(center)
    return R, t


def c5_ligand(seed: int = SEED) -> pqrs.Mol:
    """C5 template: 40 heavy atoms as a self-avoiding 1.5 A random walk + 30 hydrogens 1.09 A off
    random heavy atoms; elements and charges recycled from docked.mol2 (tests/golden/docked.pqrs)."""
    rng = np.random.default_rng(seed + 5)
    src = pqrs.read_ligands_pqrs(os.path.join(GOLDEN, "docked.pqrs"))[0]
    heavy_src = [i for i in range(src.n) if src.anum[i] != 1]
    h_src = [i for i in range(src.n) if src.anum[i] == 1]
    pts = [np.zeros(3)]
    while len(pts) < 40:
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        c = pts[rng.integers(max(0, len(pts) - 3), len(pts))] + 1.5 * d
        if min(np.linalg.norm(c - p) for p in pts) > 1.3:
            pts.append(c)
    anum = [int(src.anum[heavy_src[i % len(heavy_src)]]) for i in range(40)]
    q = [float(src.q[heavy_src[i % len(heavy_src)]]) for i in range(40)]
    while len(pts) < 70:
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        c = pts[rng.integers(0, 40)] + 1.09 * d
        if min(np.linalg.norm(c - p) for p in pts) > 0.95:
            k = len(pts) - 40
            pts.append(c); anum.append(1); q.append(float(src.q[h_src[k % len(h_src)]]))
    pts = np.array(pts); pts -= pts.mean(0)
    q = np.array(q); q -= q.mean()
    anum = np.array(anum, np.int32)
    rad = np.array([pqrs.VDW_RADII[int(a)] for a in anum])
    return pqrs.Mol("c5lig", pts[:, 0].copy(), pts[:, 1].copy(), pts[:, 2].copy(), q, rad, anum)


def c5_conformers(lig: pqrs.Mol, n: int, center, radius: float = 10.0, seed: int = SEED, jitter: float = 0.35):
    """n conformers as explicit coordinates (pose-major): random rotation, centre uniform in a sphere, and a
    per-atom Gaussian displacement standing in for the random dihedrals of the C5 recipe."""
    rng = np.random.default_rng(seed + 55)
    R, t = random_poses_in_sphere(n, center, radius, seed=seed + 56)
    P = np.stack([lig.xs, lig.ys, lig.zs])                        # 3 x L
    Rm = R.reshape(n, 3, 3)
    X = np.einsum("nij,jl->nil", Rm, P) + t[:, :, None]           # n x 3 x L
    X += rng.normal(0.0, jitter, X.shape)
    return np.ascontiguousarray(X[:, 0, :]), np.ascontiguousarray(X[:, 1, :]), np.ascontiguousarray(X[:, 2, :])
