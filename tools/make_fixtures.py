#!/usr/bin/env python
"""Build the committed input fixtures under tests/golden/ from the reference's data files.

Runs ONLY in the build container (needs /root/reference/data); the outputs are committed so that
nothing on the GPU box reads /root/reference.

  docked.pqrs / ligdecs.pqrs / minimized.pqrs : data/*.mol2 through the restated mol2pqrs
                                                (mmo_b200/pqrs.py, src/mol2pqrs.ml:10-43)
  xtal_rec.pqrs : data/xtal.pdb ATOM records through a deterministic stand-in for the external
                  OpenBabel protonation/charging step: element from cols 77-78, radius
                  from ptable.ml:41-54, charge ~ U[-0.6,0.6] (seed 20231017) recentred to net 0 and
                  rounded to 4 decimals like a mol2 charge column.  The energy math does not depend
                  on how the charges were made.
  ROI.bild      : copy of data/ROI.bild (one `.sphere` line)
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmo_b200 import pqrs  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")


def receptor_from_pdb(fn):
    xs, ys, zs, an = [], [], [], []
    for l in open(fn):
        if l.startswith("ATOM  "):
            xs.append(float(l[30:38])); ys.append(float(l[38:46])); zs.append(float(l[46:54]))
            an.append(pqrs.SYM2ANUM[l[76:78].strip().capitalize()])
    n = len(xs)
    rng = np.random.default_rng(20231017)
    q = rng.uniform(-0.6, 0.6, n)
    q -= q.mean()
    q = np.round(q, 4)
    q[0] = np.round(q[0] - q.sum(), 4)          # net charge exactly 0 after rounding
    an = np.array(an, np.int32)
    rad = np.array([pqrs.VDW_RADII[int(a)] for a in an])
    return pqrs.Mol("3A2J_heavy", np.array(xs), np.array(ys), np.array(zs), q, rad, an)


def main():
    os.makedirs(OUT, exist_ok=True)
    for base in ("docked", "ligdecs", "minimized"):
        m = pqrs.mol2_to_ligand(os.path.join(REF, base + ".mol2"))
        pqrs.write_ligand_pqrs(os.path.join(OUT, base + ".pqrs"), m)
        back = pqrs.read_ligands_pqrs(os.path.join(OUT, base + ".pqrs"))[0]
        npairs = sum(1 for i in range(m.n) for j in range(i + 1, m.n) if m.dists[i + j * m.n] >= 3)
        print(base, "atoms", m.n, "rbonds", m.n_rbonds, list(zip(back.rb_left.tolist(), back.rb_right.tolist())),
              "pairs>=3:", npairs, "maxdist", int(m.dists.max()))
    rec = receptor_from_pdb(os.path.join(REF, "xtal.pdb"))
    pqrs.write_receptor_pqrs(os.path.join(OUT, "xtal_rec.pqrs"), rec)
    print("receptor atoms", rec.n, "net q", rec.q.sum())
    with open(os.path.join(OUT, "ROI.bild"), "w") as f:
        f.write(open(os.path.join(REF, "ROI.bild")).read())


if __name__ == "__main__":
    main()
