import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, mmo_b200, bench_legs
from mmo_b200 import pqrs, workloads
mmo_b200.init(0)
L = mmo_b200.lib()
c2 = workloads.load_c2("ligdecs")
rec_m = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
rec = mmo_b200.Receptor.from_mol(rec_m)
cc = np.array(c2["roi"][:3])
gd = mmo_b200.Grid.from_box(0.5, *(cc + 23.0))
gmask = bench_legs.sphere_mask_bits(0.5, gd, cc, 21.0)
ta, tq = pqrs.assign_ff_types([c2["lig"]])
grid, _ = mmo_b200.Lds.pre_calculate_FF_components_grid(rec, 0.5, gd, ta, tq, mask_bits=gmask, want_host=False)
lig = mmo_b200.Ligand.from_mol(c2["lig"], centered=True)
for n in (1, 512, 4096):
    seeds = np.arange(n, dtype=np.uint64) + workloads.SEED
    Rm, tm = workloads.random_poses_in_sphere(n, c2["roi"][:3], 3.0, seed=41)
    mmo_b200.Lds.simulate_lig(grid, lig, c2["roi"], 2000, seeds, Rm, tm)
