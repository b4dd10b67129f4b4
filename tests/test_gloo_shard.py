"""N > 1 path on the CPU: two gloo ranks shard the lattice points of a scan, score their shard with the
CPU oracle (no GPU here), all-gather their top-k lists and merge them with libmmo_b200's host-side
mmo_topk_merge.  The merged result must be the single-process scan."""
import os
import subprocess
import sys

import numpy as np

from mmo_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle, mmo_b200
from mmo_b200 import workloads, sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
c2 = workloads.load_c2()
rec = workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + workloads.lig_radius(c2["centered"]) + 12.0)
cx, cy, cz = c2["centered"]
rot = oracle.so3_rotations(12)
roi = (c2["roi"][0], c2["roi"][1], c2["roi"][2], 3.0)
k = 15
full = oracle.scan(rec, c2["lig"], cx, cy, cz, roi, 2.0, rot, k)
nvox = int(np.prod(full["lattice_dims"]))
first, count = sharding.shard_range(nvox, rank, world)
mine = oracle.scan(rec, c2["lig"], cx, cy, cz, roi, 2.0, rot, k, first_point=first, n_points=count)
s, f = sharding.allgather_topk(dist, mmo_b200.lib(), k, mine["top_scores"], mine["top_frames"])
best = torch.tensor([mine["best_score"], float(mine["best_frame"])], dtype=torch.float64)
allb = [torch.empty_like(best) for _ in range(world)]
dist.all_gather(allb, best)
bs, bf = min((float(b[0]), int(b[1])) for b in allb if int(b[1]) >= 0)
ok = np.array_equal(s, full["top_scores"]) and np.array_equal(f, full["top_frames"]) and bf == full["best_frame"]
tot = torch.tensor([mine["n_scored"]]); dist.all_reduce(tot)
ok = ok and int(tot.item()) == full["n_scored"]
print(json.dumps({"rank": rank, "ok": bool(ok), "n": len(s)}))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_shard_range_partitions():
    for n in (0, 1, 7, 4139, 9702):
        for w in (1, 2, 3, 8):
            blocks = [sharding.shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in blocks) == n
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_two_gloo_ranks_merge_to_the_single_process_scan(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('"ok": true') == 2


WORKER_C5 = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle, mmo_b200
from mmo_b200 import workloads, sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# C5 shape, small: conformers of the 70-atom template against a synthetic receptor sphere; the conformer id is the "frame"
rec = workloads.synthetic_receptor(600, "sphere", 14.0, seed=7, origin=(20.0, 20.0, 20.0))
lig = workloads.c5_ligand()
n, k = 301, 25
X, Y, Z = workloads.c5_conformers(lig, n, (20.0, 20.0, 20.0), radius=4.0)
X[17] = X[3]; Y[17] = Y[3]; Z[17] = Z[3]                     # an exact tie: the smaller id must win
def topk(e, first):
    ids = np.arange(first, first + len(e))
    o = np.lexsort((ids, e))[:k]
    return e[o], ids[o]
full_s, full_f = topk(oracle.ene_inter(rec, lig.q, lig.anum, X, Y, Z, shifted=True), 0)
first, count = sharding.shard_range(n, rank, world)
sl = slice(first, first + count)
s_loc, f_loc = topk(oracle.ene_inter(rec, lig.q, lig.anum, X[sl], Y[sl], Z[sl], shifted=True), first)
s, f = sharding.allgather_topk(dist, mmo_b200.lib(), k, s_loc, f_loc)
ok = np.array_equal(s, full_s) and np.array_equal(f, full_f) and (list(f).index(3) < list(f).index(17) if 3 in f and 17 in f else True)
print(json.dumps({"rank": rank, "ok": bool(ok), "n": len(s)}))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_two_gloo_ranks_merge_conformer_ids_of_a_sharded_screen(tmp_path):
    """C5 at N > 1: conformers are dealt to the ranks, every rank keeps its top-k (energy, conformer id), the lists are
    all-gathered and merged by mmo_topk_merge: ids and order of the single-process screen, ties to the smaller id"""
    script = tmp_path / "worker_c5.py"
    script.write_text(WORKER_C5)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('"ok": true') == 2
